"""Throughput of gen_roi on the GPU (idl_sweep) on a BASELINE workload: kernel time by CUDA events, algorithmic and
streamed bytes against the measured HBM copy bandwidth, the host stand-in's sweep timed beside it.
  python tools/sweep_bench.py [workload] [reps]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from indelope_b200 import cuda, host  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "chr1"
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    ds = host.Dataset(**host.CONFIGS[name])
    t0 = time.time(); rois = ds.sweep(min_reads=5); host_s = time.time() - t0
    peak = 6650.0
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        peak = json.load(open(p)).get("hbm_gbs", peak)
    out = {"workload": name, "host_sweep_s": host_s, "regions": rois.n_rois, "chroms": []}
    for c in range(ds.n_chroms):
        cr = ds.chrom_reads(c)
        best = None
        for _ in range(reps):
            r = cuda.sweep(cr["chrom_len"], cr["start"], cr["stop"], cr["flag"], cr["cigar"], cr["cig_off"], min_event_support=3, min_read_coverage=5, max_read_coverage=600)
            if best is None or r["ms_kernels"] < best["ms_kernels"]:
                best = r
        ms = best["ms_kernels"]
        out["chroms"].append({"chrom_len": cr["chrom_len"], "records": len(cr["start"]), "regions": len(best["roi_start"]), "runs": best["n_runs"], "ms_kernels": ms,
                              "ms_h2d": best["ms_h2d"], "ms_d2h": best["ms_d2h"], "algorithmic_gbs": best["algorithmic_bytes"] / ms / 1e6,
                              "streamed_gbs": best["streamed_bytes"] / ms / 1e6, "hbm_peak_gbs": peak, "algorithmic_frac": best["algorithmic_bytes"] / ms / 1e6 / peak,
                              "streamed_frac": best["streamed_bytes"] / ms / 1e6 / peak, "positions_per_s": cr["chrom_len"] / ms * 1e3})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
