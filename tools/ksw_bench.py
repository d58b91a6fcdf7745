"""Microbenchmark of kernel 2 through idl_ksw2_batch at fixed shapes (GCUPS = exact in-band cells / kernel time).
  python tools/ksw_bench.py [n]
"""
import os
import sys
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from indelope_b200 import cuda


def make(rng, n, ql, tl, site):
    qs, ts = [], []
    for _ in range(n):
        base = rng.integers(0, 4, tl + ql + 64).astype(np.uint8)
        if site == "A":  # contig vs window with one indel
            q = base[:ql].copy(); pos = ql // 2
            q = np.concatenate([q[:pos], q[pos + 12:ql], base[tl:tl + 12]])
            t = base[:tl].copy()
        else:            # read vs suffix
            o = int(rng.integers(0, max(1, tl - ql)))
            q = base[o:o + ql].copy(); t = base[:tl].copy()
        qs.append(q); ts.append(t)
    return qs, ts


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 56832  # 4 full waves of 148 SMs x 3 CTAs x 32 groups
    ctx = cuda.Context(0)
    rng = np.random.default_rng(1)
    for site, ql, tl, go, w, z in [("A", 300, 420, 4, 50, 400), ("A", 600, 720, 4, 50, 400), ("B", 150, 400, 5, -1, -1), ("B", 150, 700, 5, -1, -1)]:
        qs, ts = make(rng, n, ql, tl, site)
        best = None
        for rep in range(3):
            f, c, extra, ms = ctx.ksw2_batch(qs, ts, gapo=go, gape=1, w=w, zdrop=z)
            best = ms if best is None else min(best, ms)
        cells = sum(e["cells"] for e in extra)
        print("site %s %dx%d n=%d: %.3f ms, %.1f GCUPS, %.2f us/alignment/warp-slot, zdropped %.2f, avg cigar ops %.1f" % (
            site, ql, tl, n, best, cells / best / 1e6, best * 1e3 / n, np.mean([x["zdropped"] for x in f]), np.mean([x["n_cigar"] for x in f])))
    ctx.close()


if __name__ == "__main__":
    main()
