#!/bin/bash
# quick iteration loop for kernel 2 (run under gpurun): parity tests, micro-bench, one full capture of call-site B, DRAM traffic of al_kernel
OUT=gpurun_out; mkdir -p $OUT
(timeout 600 python -m pytest tests/test_gpu_ksw2.py -m gpu -x -q 2>&1 | tail -4) > $OUT/q_tests.log
timeout 300 python tools/ksw_bench.py > $OUT/q_kswbench.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 200 > $OUT/q_bench.json 2> $OUT/q_bench.err
if [ "${1:-}" != "noprof" ]; then
ncu --set full --clock-control none --import-source on -k regex:ksw2_batch_kernel -s 9 -c 1 -o $OUT/prof_ksw2_siteB_150x700 -f python tools/ksw_bench.py 14208 > $OUT/ncu_kswB.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'al_kernel' -s 1 -c 1 --csv --log-file $OUT/q_traffic.csv python bench.py --steps 1 --warmup 1 --cpu-sample 200 > $OUT/ncu_traffic.log 2>&1
fi
cat $OUT/q_tests.log $OUT/q_kswbench.log; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/q_bench.json")); print(d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("bench failed", e)
PY
tail -12 $OUT/q_traffic.csv 2>/dev/null | cut -c1-300
