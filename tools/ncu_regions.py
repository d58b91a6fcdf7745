"""Executed warp instructions of an ncu report aggregated by source line ranges.
  python tools/ncu_regions.py report.ncu-rep file:lo-hi[:name] ...   (no ranges: per-file totals and the raw key metrics)
"""
import csv
import subprocess
import sys


def load(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, fname, agg, smp = None, "", {}, {}
    for r in rows:
        if len(r) == 2 and r[0] == "File Path":
            fname = r[1].split("/")[-1]; continue
        if len(r) > 5 and r[0] == "Line No":
            hdr = r; continue
        if hdr and len(r) == len(hdr) and r[0] != "":
            try:
                n = int(r[hdr.index("Instructions Executed")]); s = int(r[hdr.index("# Samples")])
            except ValueError:
                continue
            k = (fname, int(r[0]))
            agg[k] = agg.get(k, 0) + n; smp[k] = smp.get(k, 0) + s
    return agg, smp


def raw(rep, keys):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    return {k: d.get(k) for k in keys}


def main():
    rep = sys.argv[1]
    agg, smp = load(rep)
    tot = sum(agg.values()) or 1; ts = sum(smp.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, ts))
    for k, v in raw(rep, ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
                          "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
                          "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__icc_request_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
                          "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
                          "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
                          "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
                          "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]).items():
        print("  %-85s %s" % (k, v))
    if len(sys.argv) == 2:
        files = {}
        for (f, l), v in agg.items():
            files[f] = files.get(f, 0) + v
        for f, v in sorted(files.items(), key=lambda x: -x[1]):
            print("%-28s %5.1f%%" % (f, 100.0 * v / tot))
        return
    for spec in sys.argv[2:]:
        parts = spec.split(":")
        f = parts[0]; lo, hi = [int(x) for x in parts[1].split("-")]; name = parts[2] if len(parts) > 2 else spec
        v = sum(n for (ff, l), n in agg.items() if ff == f and lo <= l <= hi)
        s = sum(n for (ff, l), n in smp.items() if ff == f and lo <= l <= hi)
        print("%-34s %5.1f%% inst  %5.1f%% samples" % (name, 100.0 * v / tot, 100.0 * s / ts))


if __name__ == "__main__":
    main()
