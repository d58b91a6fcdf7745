"""Host-side mirror of the reference's calling loop over the two native libraries.

    for r in gen_roi(...): for v in callsemble(r, ...): dedup; echo v        (src/indelope.nim:601-608)

becomes: pack regions into pinned batches -> idl_submit (async, one lane per stream) -> idl_wait -> filter cascade +
VCF text on the host.  Regions keep their emission order (batches are contiguous slices, records are written in
ticket order), so the order-dependent dedup sees the same sequence as the reference's single loop.
"""
import ctypes as C

import numpy as np

from . import cuda, host
from .abi import default_params


def plan_batches(rois, lo, hi, max_reads=400_000, max_regions=20_000):
    """contiguous slices [a, b) of the region list, bounded by reads and regions per batch"""
    n = np.ctypeslib.as_array(rois.c.roi_n_reads, shape=(rois.n_rois,)) if rois.n_rois else np.zeros(0, np.int32)
    out, a, acc = [], lo, 0
    for k in range(lo, hi):
        if k > a and (acc + int(n[k]) > max_reads or k - a >= max_regions):
            out.append((a, k)); a, acc = k, 0
        acc += int(n[k])
    if hi > a:
        out.append((a, hi))
    return out


def gpu_rois(ds, min_reads=3, max_read_coverage=600, device=0, timings=None):
    """gen_roi for every target of a dataset ON THE GPU (idl_sweep): the same regions and record lists as
    ds.sweep(min_reads) -- src/indelope.nim:515-545,601-602 with min_event_support = max(3, min_reads - 2) -- as a host.Rois"""
    base = host.Rois(host.lib().idlh_sweep(ds.h, 255, 1 << 30, 1 << 30), ds)  # no regions: just the record arrays of the dataset
    a = dict(base.arrays())
    chrom, rs, re, nr, idx = [], [], [], [], []
    for c in range(ds.n_chroms):
        cr = ds.chrom_reads(c)
        r = cuda.sweep(cr["chrom_len"], cr["start"], cr["stop"], cr["flag"], cr["cigar"], cr["cig_off"], min_event_support=max(3, min_reads - 2),
                       min_read_coverage=min_reads, max_read_coverage=max_read_coverage, device=device)
        chrom.append(np.full(len(r["roi_start"]), c, np.int32)); rs.append(r["roi_start"]); re.append(r["roi_end"]); nr.append(r["roi_n_reads"])
        idx.append(r["read_idx"] + cr["first_read"])
        if timings is not None:
            timings.append({k: r[k] for k in ("ms_h2d", "ms_kernels", "ms_d2h", "algorithmic_bytes", "streamed_bytes", "n_runs")})
    a["roi_chrom"] = np.concatenate(chrom); a["roi_start"] = np.concatenate(rs); a["roi_stop"] = np.concatenate(re); a["roi_n_reads"] = np.concatenate(nr)
    a["read_idx"] = np.concatenate(idx)
    a["roi_read_begin"] = np.concatenate([[0], np.cumsum(a["roi_n_reads"])[:-1]]).astype(np.int64) if len(a["roi_n_reads"]) else np.zeros(0, np.int64)
    out = host.Rois(arrays=a)
    out._keep = (base, ds)  # the arrays are views of the dataset's memory
    return out


class Caller:
    """`indelope --min-reads M --min-contig-len C --min-event-len E` over regions, on one GPU"""

    def __init__(self, device=0, min_reads=3, min_ctg_len=73, min_event_len=4, **kw):
        self.params = default_params(min_reads=min_reads, min_ctg_len=min_ctg_len, min_event_len=min_event_len, **kw)
        self.ctx = cuda.Context(device, self.params)
        self._batches = {}

    def _batch_for(self, lane, sizes):
        """pinned batch per lane, grown on demand"""
        cur = self._batches.get(lane)
        need = (sizes[0], sizes[1], sizes[2], sizes[3])
        if cur is None or any(n > c for n, c in zip(need, cur[1])):
            if cur is not None:
                self.ctx.batch_free(cur[0])
            cap = tuple(int(x * 1.25) + 64 for x in need)
            cur = (self.ctx.batch_alloc(*cap), cap)
            self._batches[lane] = cur
        return cur[0]

    def call(self, rois, lo=0, hi=None, dump_level=0, max_reads=400_000, timings=None, dedup=True):
        """returns (vcf record text, dump text). `timings` (list) receives one dict per batch.  dedup=False leaves the
        order-dependent dedup to the caller (interval shards are merged first, see indelope_b200.shard)."""
        hi = rois.n_rois if hi is None else hi
        writer = host.VcfWriter(dedup=dedup)
        plans = plan_batches(rois, lo, hi, max_reads=max_reads)
        n_lanes = self.params.n_streams
        vcf, dump = [], []
        inflight = []

        def drain():
            a, b, t = inflight.pop(0)
            res = self.ctx.wait(t)
            v, d = writer.records(rois, a, self.params, res, dump_level)
            if timings is not None:
                r = res.contents
                timings.append({k: getattr(r, k) for k in ("ms_h2d", "ms_assemble", "ms_align", "ms_genotype", "ms_al", "ms_d2h", "ms_total", "n_regions",
                                                           "n_contigs", "n_alns", "n_events", "offsets_tested", "dp_cells_a", "dp_cells_b", "dp_a", "dp_b",
                                                           "kmer_reads", "kmer_bytes", "al_events", "kernel_launches", "pool_retries")})
            self.ctx.release(t)
            vcf.append(v); dump.append(d)

        for i, (a, b) in enumerate(plans):
            if len(inflight) >= n_lanes:
                drain()
            nr, sb, rb = rois.pack_size(a, b, self.params)
            batch = self._batch_for(i % n_lanes, (b - a, nr, sb, rb))
            rois.pack(a, b, self.params, batch)
            inflight.append((a, b, self.ctx.submit(batch)))
        while inflight:
            drain()
        self.status_counts = writer.status_counts()  # regions per IDL_RS_* bit of the last call (warnings went to stderr)
        return "".join(vcf), "".join(dump)

    def close(self):
        for b, _ in self._batches.values():
            self.ctx.batch_free(b)
        self._batches = {}
        self.ctx.close()


def call_bam(fasta, bam, device=0, min_reads=3, min_ctg_len=73, min_event_len=4, max_reads=400_000, by_target=False, timings=None, **kw):
    """`indelope --gpu-decode [options] <fasta> <bam>` in process: the VCF (header + records) with everything in front of callsemble on the
    device as well -- BGZF inflate and record parse (idl_bam_open), gen_roi per target (idl_bam_sweep, src/indelope.nim:515-545,601-602), the
    batches built from the resident records (idl_bam_submit).  The host reads the two files and writes the VCF text.
    by_target: the file is not decoded as a whole but target by target, each from the run of BGZF members <bam>.bai points to
    (idl_bam_open_slice) -- for files that do not fit in device memory; needs the index."""
    if by_target:
        st = host.Stream(fasta, bam)          # parses the BAM header and orders the FASTA's sequences after it; no record is read through it
        tg = st.targets()
        ta = tg.arrays()
        names, seqs = ta["chrom_names"], ta["chrom_seqs"]
        ref_len = [len(x) for x in seqs]
        keep = (st, tg)
    else:
        ref = host.Dataset.load_fasta(fasta)
        data = open(bam, "rb").read()
        whole = cuda.Bam(data, device=device)
        ref.set_targets(whole.ref_names, whole.ref_len)
        names, seqs = ref.sequences()
        keep = (ref,)
    caller = Caller(device, min_reads=min_reads, min_ctg_len=min_ctg_len, min_event_len=min_event_len, **kw)
    writer = host.VcfWriter()
    empty = dict(start=[], stop=[], mapq=[], flag=[], len=[], seq_off=[], bases=[], quals=[])
    out = [host.Rois(arrays=dict(empty, roi_chrom=[], roi_start=[], roi_stop=[], roi_read_begin=[], roi_n_reads=[], read_idx=[], chrom_names=names, chrom_seqs=seqs)).header()]
    n_regions_done = 0
    try:
        for c, name in enumerate(names):
            if name == "hs37d5" or name.startswith("GL"):   # skippable targets, src/indelope.nim:41-42
                continue
            if by_target:
                sp = host.bai_target_span(bam, c)
                if sp is None:
                    continue
                with open(bam, "rb") as f:
                    f.seek(sp["file_begin"]); run = f.read(sp["file_end"] - sp["file_begin"])
                b = cuda.Bam(run, device=device, slice=dict(ref_names=names, ref_len=ref_len, first_record=sp["first_record"], end_member=sp["end_member"],
                                                            end_offset=sp["end_offset"]))
            else:
                b = whole
            try:
                r = b.sweep(c, min_event_support=max(3, min_reads - 2), min_read_coverage=min_reads, max_read_coverage=600)
                n = len(r["roi_start"])
                if n == 0:
                    continue
                chrom = np.full(n, c, np.int32); rs, re, nr, idx = r["roi_start"], r["roi_end"], r["roi_n_reads"], r["read_idx"]
                begin = np.concatenate([[0], np.cumsum(nr)]).astype(np.int64)
                rois = host.Rois(arrays=dict(empty, roi_chrom=chrom, roi_start=rs, roi_stop=re, roi_read_begin=begin[:-1], roi_n_reads=nr, read_idx=idx, chrom_names=names, chrom_seqs=seqs))
                b.set_reference(c, seqs[c])
                inflight = []

                def drain():
                    lo, t = inflight.pop(0)
                    res = caller.ctx.wait(t)
                    v, _ = writer.records(rois, lo, caller.params, res, 0)
                    if timings is not None:
                        timings.append({k: getattr(res.contents, k) for k in ("ms_assemble", "ms_align", "ms_genotype", "ms_al", "n_regions", "n_events")})
                    caller.ctx.release(t); out.append(v)
                for lo, hi in plan_batches(rois, 0, n, max_reads=max_reads):
                    if len(inflight) >= caller.params.n_streams:
                        drain()
                    inflight.append((lo, caller.ctx.bam_submit(b, chrom[lo:hi], rs[lo:hi], re[lo:hi], nr[lo:hi], idx[begin[lo]:begin[hi]], ordinal_base=n_regions_done + lo)))
                while inflight:
                    drain()
                n_regions_done += n
            finally:
                if by_target:
                    b.close()
        caller.status_counts = writer.status_counts()
    finally:
        caller.close()
        if not by_target:
            whole.close()
    del keep
    return "".join(out)
