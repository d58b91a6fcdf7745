"""Summarise ncu captures from gpurun_out/ into tracked text files under profiles/.

  python tools/summarize_profiles.py r01
writes profiles/<tag>_launches.csv (copy of the launch list), profiles/<tag>_launches_summary.txt (per-kernel shares) and
profiles/<tag>_<kernel>.txt (key raw metrics + the hottest source lines) for every gpurun_out/prof_<kernel>.ncu-rep.
"""
import csv
import glob
import os
import shutil
import subprocess
import sys
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")

WANT = """gpu__time_duration.sum launch__grid_size launch__block_size launch__registers_per_thread launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem launch__waves_per_multiprocessor sm__warps_active.avg.pct_of_peak_sustained_active
smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio smsp__issue_active.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__throughput.avg.pct_of_peak_sustained_elapsed
dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed lts__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__throughput.avg.pct_of_peak_sustained_elapsed l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum sm__cycles_active.avg
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio
smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
sm__icc_request_hit_rate.pct gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct
sm__sass_inst_executed_op_shared_ld.sum sm__sass_inst_executed_op_shared_st.sum sm__sass_inst_executed_op_global_ld.sum sm__sass_inst_executed_op_global_st.sum""".split()


def raw_metrics(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    if len(rows) < 3:
        return []
    hdr, units = rows[0], rows[1]
    out = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        out.append([(h, d[h], u) for h, u in zip(hdr, units) if h in WANT] + [("Kernel Name", d.get("Kernel Name", ""), "")])
    return out


def hot_lines(rep, top=25):
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
            out.append((n, smp, fname, r[0], r[1].strip()[:110]))
    tot = sum(o[0] for o in out) or 1; ts = sum(o[1] for o in out) or 1
    lines = ["total warp instructions %d, stall samples %d" % (tot, ts)]
    for n, smp, f, ln, src in sorted(out, key=lambda x: -x[0])[:top]:
        lines.append("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100.0 * n / tot, 100.0 * smp / ts, f, ln, src))
    return lines


def launches(tag):
    src = os.path.join(OUT, "launches.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, tag + "_launches.csv"))
    rows = [r for r in csv.reader(open(src)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(list)
    for r in rows:
        if r[0] == "ID":
            hdr = r; continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        if d.get("Metric Name") == "gpu__time_duration.sum":
            v = float(d["Metric Value"].replace(",", ""))
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d["Metric Unit"], 1.0)
            agg[d["Kernel Name"].split("(")[0]].append(v)
    tot = sum(sum(v) for v in agg.values()) or 1
    with open(os.path.join(PROF, tag + "_launches_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none, one short bench run (cold-cache, serialised: compare SHARES)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
            f.write("%-34s launches=%3d total_us=%12.1f share=%6.3f avg_us=%10.1f\n" % (k, len(v), sum(v), sum(v) / tot, sum(v) / len(v)))


def traffic(tag):
    """per-kernel DRAM bytes / instructions / pipe utilisation at the full default workload -> profiles/<tag>_traffic.json (bench.py reads it)"""
    import json
    src = os.path.join(OUT, "traffic_full.csv")
    if not os.path.exists(src):
        return
    shutil.copy(src, os.path.join(PROF, tag + "_traffic_full.csv"))
    hdr, per = None, collections.defaultdict(lambda: collections.defaultdict(list))
    for r in csv.reader(open(src)):
        if len(r) > 5 and r[0] == "ID":
            hdr = r; continue
        if hdr is None or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        v = float(d["Metric Value"].replace(",", ""))
        if d["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(d["Metric Unit"], 1.0)
        if d["Metric Unit"] in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[d["Metric Unit"]]
        per[d["Kernel Name"].split("(")[0].replace("void ", "")][d["Metric Name"]].append(v)
    out = {}
    for full, m in per.items():
        # the resident leg is the SECOND launch of each kernel (one warm-up first; the end-to-end leg's smaller launches follow);
        # the two assemble_kernel variants of a step are summed
        k = full.split("<")[0]
        def second(name):
            v = m.get(name, [])
            return v[1] if len(v) > 1 else (v[0] if v else None)
        rd, wr = second("dram__bytes_read.sum") or 0, second("dram__bytes_write.sum") or 0
        cur = {"dram_bytes_per_launch": rd + wr, "read": rd, "write": wr, "ncu_ms": second("gpu__time_duration.sum"),
               "thread_inst": second("smsp__thread_inst_executed.sum"), "warp_inst": second("smsp__inst_executed.sum"),
               "pipe_alu_pct": second("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
               "issue_active_pct": second("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "warps_active_pct": second("sm__warps_active.avg.pct_of_peak_sustained_active")}
        if k in out:
            for f in ("dram_bytes_per_launch", "read", "write", "ncu_ms", "thread_inst", "warp_inst"):
                out[k][f] = (out[k][f] or 0) + (cur[f] or 0)
            for f in ("pipe_alu_pct", "issue_active_pct", "warps_active_pct"):
                out[k][f] = max(out[k][f] or 0, cur[f] or 0)
        else:
            out[k] = cur
    out["_how"] = ("ncu --metrics dram__bytes_*,smsp__thread_inst_executed.sum,... --clock-control none on `python bench.py --steps 1 --warmup 1` "
                   "(default chr1 workload), second launch of each kernel = the device-resident leg; profiles/%s_traffic_full.csv" % tag)
    json.dump(out, open(os.path.join(PROF, tag + "_traffic.json"), "w"), indent=1)
    print("wrote traffic json")


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(PROF, exist_ok=True)
    launches(tag)
    traffic(tag)
    for rep in sorted(glob.glob(os.path.join(OUT, "prof_*.ncu-rep"))):
        name = os.path.basename(rep)[5:-8]
        with open(os.path.join(PROF, "%s_%s.txt" % (tag, name)), "w") as f:
            f.write("ncu --set full --clock-control none --import-source on  (%s)\n\n" % os.path.basename(rep))
            for launch in raw_metrics(rep):
                for h, v, u in launch:
                    f.write("%-78s %18s %s\n" % (h, v, u))
                f.write("\n")
            f.write("\n".join(hot_lines(rep)) + "\n")
        print("wrote", name)


if __name__ == "__main__":
    main()
