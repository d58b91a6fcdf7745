"""Per-source-line executed warp instructions of one ncu report, in source order (for a line range of one file).
  python tools/ncu_perline.py gpurun_out/prof_kswB.ncu-rep ksw2.cuh 180 420
"""
import csv
import subprocess
import sys


def main():
    rep, fn, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, fname, out = None, "", []
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            try:
                n = int(r[hdr.index("Instructions Executed")]); smp = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            out.append((fname, int(r[0]), n, smp, r[1].strip()[:90]))
    tot = sum(o[2] for o in out) or 1; ts = sum(o[3] for o in out) or 1
    print("total warp instructions %d, samples %d" % (tot, ts))
    byfile = {}
    for f, l, n, smp, s in out:
        byfile[f] = byfile.get(f, 0) + n
    print({k: "%.1f%%" % (100.0 * v / tot) for k, v in byfile.items() if v})
    for f, l, n, smp, s in sorted(out):
        if f == fn and lo <= l <= hi and n > 0:
            print("%4d %5.2f%% inst %5.2f%% smp %10d  %s" % (l, 100.0 * n / tot, 100.0 * smp / ts, n, s))


if __name__ == "__main__":
    main()
