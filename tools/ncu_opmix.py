"""Dynamic opcode mix of one ncu report (executed warp instructions per SASS opcode, with the issue pipe of each).
  python tools/ncu_opmix.py gpurun_out/prof_x.ncu-rep
"""
import csv
import re
import subprocess
import sys

FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMAD.WIDE"}
LSU = {"LDS", "STS", "LDG", "STG", "LDC", "LDCU", "ATOMG", "ATOMS", "SHFL", "LDL", "STL", "RED", "MATCH", "VOTE", "REDUX"}


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = None
    mix = {}
    for r in rows:
        if len(r) > 5 and r[0] == "Address":
            hdr = r; continue
        if not hdr or len(r) != len(hdr):
            continue
        src = r[1].strip()
        m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
        if not m:
            continue
        op = m.group(2)
        n = int(r[hdr.index("Instructions Executed")])
        mix[op] = mix.get(op, 0) + n
    tot = sum(mix.values()) or 1
    pipes = {"fma": 0, "lsu/other-mem": 0, "alu+rest": 0}
    for op, n in mix.items():
        pipes["fma" if op in FMA else "lsu/other-mem" if op in LSU else "alu+rest"] += n
    print("total warp instructions", tot, {k: "%.1f%%" % (100.0 * v / tot) for k, v in pipes.items()})
    for op, n in sorted(mix.items(), key=lambda kv: -kv[1])[:30]:
        print("%-12s %6.2f%%" % (op, 100.0 * n / tot))


if __name__ == "__main__":
    main()
