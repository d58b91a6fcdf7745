"""one small pass through every idl_bam_* entry point (for compute-sanitizer runs): python tools/bam_tiny.py"""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from indelope_b200 import abi, cuda, host  # noqa: E402

cfg = dict(host.CONFIGS["pr1"]); cfg.update(chrom_len=60_000, n_events=12, max_indel=40, qual_levels=8)
ds = host.Dataset(**cfg)
d = tempfile.mkdtemp()
p = os.path.join(d, "t.bam"); ds.write_bam(p, level=6); data = open(p, "rb").read()
b = cuda.Bam(data)
rois = ds.sweep(min_reads=5); a = rois.arrays()
s = b.sweep(0, min_event_support=3, min_read_coverage=5)
assert np.array_equal(s["roi_start"], a["roi_start"])
r = b.fetch(s["read_idx"][:50])
b.set_reference(0, a["chrom_seqs"][0])
params = abi.default_params(min_reads=5, min_ctg_len=73, min_event_len=5)
ctx = cuda.Context(0, params)
t = ctx.bam_submit(b, np.zeros(len(s["roi_start"]), np.int32), s["roi_start"], s["roi_end"], s["roi_n_reads"], s["read_idx"])
res = ctx.wait(t); print("regions", res.contents.n_regions, "events", res.contents.n_events); ctx.release(t)
ctx.close(); b.close()
