"""helpers shared by the parity tests"""
import os

import numpy as np

from indelope_b200 import host

HAS_GPU = None


def has_gpu():
    global HAS_GPU
    if HAS_GPU is None:
        try:
            import torch
            HAS_GPU = bool(torch.cuda.is_available())
        except Exception:
            HAS_GPU = False
    return HAS_GPU


def small_dataset(name="pr1", **over):
    cfg = dict(host.CONFIGS[name])
    cfg.update(over)
    return host.Dataset(**cfg)


def diff_lines(a, b, limit=12):
    """first differing lines of two dumps, grouped by record type"""
    la, lb = a.splitlines(), b.splitlines()
    out = []
    if len(la) != len(lb):
        out.append("line counts differ: %d vs %d" % (len(la), len(lb)))
    n = 0
    for i, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            out.append("line %d:\n  oracle: %s\n  gpu   : %s" % (i, x[:600], y[:600]))
            n += 1
            if n >= limit:
                break
    return "\n".join(out)


def by_type(dump):
    d = {}
    for l in dump.splitlines():
        d.setdefault(l[:1], []).append(l)
    return d


def rois_from_reads(reads, chrom_seq, roi_start=None, roi_stop=None, chrom_name="chrT"):
    """build a one-region Rois from a list of dicts(start, seq, qual=None, mapq=60, flag=0, stop=None)"""
    starts, stops, mapq, flag, lens, offs = [], [], [], [], [], []
    bases, quals = [], []
    off = 0
    for r in reads:
        s = r["seq"]
        starts.append(r["start"]); stops.append(r.get("stop", r["start"] + len(s))); mapq.append(r.get("mapq", 60)); flag.append(r.get("flag", 0))
        lens.append(len(s)); offs.append(off); off += len(s)
        bases.append(np.frombuffer(s.encode(), dtype=np.uint8))
        q = r.get("qual")
        quals.append(np.full(len(s), 30, dtype=np.uint8) if q is None else np.asarray(q, dtype=np.uint8))
    n = len(reads)
    arrays = dict(
        start=np.array(starts, np.int32), stop=np.array(stops, np.int32), mapq=np.array(mapq, np.uint8), flag=np.array(flag, np.uint16),
        len=np.array(lens, np.int32), seq_off=np.array(offs, np.int64),
        bases=np.concatenate(bases) if bases else np.zeros(0, np.uint8), quals=np.concatenate(quals) if quals else np.zeros(0, np.uint8),
        roi_chrom=np.zeros(1, np.int32), roi_start=np.array([roi_start if roi_start is not None else min(starts)], np.int32),
        roi_stop=np.array([roi_stop if roi_stop is not None else max(stops)], np.int32), roi_read_begin=np.zeros(1, np.int64),
        roi_n_reads=np.array([n], np.int32), read_idx=np.arange(n, dtype=np.int64),
        chrom_names=[chrom_name], chrom_seqs=[np.frombuffer(chrom_seq.encode(), dtype=np.uint8) if isinstance(chrom_seq, str) else chrom_seq],
    )
    return host.Rois(arrays=arrays), arrays


def merge_rois(list_of_arrays):
    """concatenate several one-chromosome region sets that share the same chromosome into one"""
    out = {k: [] for k in ("start", "stop", "mapq", "flag", "len", "seq_off", "bases", "quals", "roi_chrom", "roi_start", "roi_stop", "roi_read_begin",
                           "roi_n_reads", "read_idx")}
    nread = nbase = nidx = 0
    for a in list_of_arrays:
        for k in ("start", "stop", "mapq", "flag", "len", "bases", "quals", "roi_chrom", "roi_start", "roi_stop", "roi_n_reads"):
            out[k].append(a[k])
        out["seq_off"].append(a["seq_off"] + nbase)
        out["roi_read_begin"].append(a["roi_read_begin"] + nidx)
        out["read_idx"].append(a["read_idx"] + nread)
        nread += len(a["start"]); nbase += len(a["bases"]); nidx += len(a["read_idx"])
    m = {k: np.concatenate(v) for k, v in out.items()}
    m["chrom_names"] = list_of_arrays[0]["chrom_names"]; m["chrom_seqs"] = list_of_arrays[0]["chrom_seqs"]
    return m


def out_dir():
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    return d
