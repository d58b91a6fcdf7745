"""Soak of kernel 2 through idl_ksw2_batch against the reference's own compiled ksw2_extz2_sse.c (oracle/_ref), or the lane model
when that is absent: many more random pairs than the test-suite runs, every ksw_extz_t field and the full CIGAR.
  python tools/ksw_soak.py [n_pairs] [seed]
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from indelope_b200 import cuda  # noqa: E402
from oracle import pyoracle as orc  # noqa: E402
from test_oracle_ksw2 import random_pair  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    seed = int(sys.argv[2]) if len(sys.argv) > 2 else 4242
    impl = "ref" if orc.have_ref() else "lane"
    rng = np.random.default_rng(seed)
    groups = {}
    for _ in range(n):
        q, t, go, w, z = random_pair(rng)
        groups.setdefault((go, w, z), []).append((q, t))
    ctx = cuda.Context(0)
    bad = 0; unb = 0; t0 = time.time()
    for (go, w, z), pairs in groups.items():
        f, c, extra, _ = ctx.ksw2_batch([p[0] for p in pairs], [p[1] for p in pairs], gapo=go, gape=1, w=w, zdrop=z)
        for i, (q, t) in enumerate(pairs):
            fo, co, _ = orc.ksw2(q, t, gapo=go, gape=1, w=w, zdrop=z, impl=impl)
            bad += int(extra[i]["status"] < 0 or fo != f[i] or co != c[i])
        if w < 0 and z < 0:
            unb += len(pairs)
    ctx.close()
    print("ksw soak: %d pairs (%d unbanded without z-drop: the row-owned variant for reads up to 160 bases), %d parameter sets, checked against '%s': %d mismatches, %.0f s" % (
        n, unb, len(groups), impl, bad, time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
