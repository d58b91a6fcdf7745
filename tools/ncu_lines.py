"""Aggregate an ncu report's executed instructions and stall samples by CUDA source line.
  python tools/ncu_lines.py gpurun_out/prof_al_kernel.ncu-rep [top]
"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
            out.append((n, smp, fname, r[0], r[1].strip()[:100]))
    tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
    print("total warp instructions %d, samples %d" % (tot, ts))
    for n, smp, f, ln, src in sorted(out, key=lambda x: -x[0])[:top]:
        print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * n / tot, 100.0 * smp / ts, f, ln, src))


if __name__ == "__main__":
    main()
