"""host-side timing of the order-dependent dedup over a genome's worth of record text (the serial tail of the strong-scaling run):
  python tools/dedup_bench.py [dup_every]     (no GPU needed)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from indelope_b200 import host  # noqa: E402

dup = int(sys.argv[1]) if len(sys.argv) > 1 else 29
lines = []
for i in range(151756):
    lines.append("chrS%d\t%d\t.\t%s\t%s\t%d\tPASS\t%s\tGT:AD:DP:GQ:GL\t0/1:10,12:22:45.0:-10,0,-20\n" % (
        1 + i // 6500, 1000 + (i - (dup > 0 and i % dup == 0)) * 37, "A" * 3, "ACGT" * 2, 30, "X" * 320))
txt = "".join(lines).encode()
a = np.concatenate([np.frombuffer(txt, dtype=np.uint8), np.zeros(8, np.uint8)])
for th in (1, 2, 4, 8, 16):
    host.set_threads(th)
    best = 1e9
    for _ in range(5):
        b = a.copy(); t = time.time(); n = host.dedup_inplace(b.ctypes.data, len(txt)); best = min(best, time.time() - t)
    print("%2d threads: %.1f ms  (%d -> %d bytes)" % (th, best * 1e3, len(txt), n))
