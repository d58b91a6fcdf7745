#!/bin/bash
# one short bench run + the kernel-2 micro-benchmark (run under gpurun while iterating on a kernel)
OUT=gpurun_out; mkdir -p $OUT
timeout 100 python tools/ksw_bench.py > $OUT/q_kswbench4.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 200 > $OUT/q_bench.json 2> $OUT/q_bench.err
grep "site B" $OUT/q_kswbench4.log; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/q_bench.json")); print(d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("bench failed", e)
PY
tail -3 $OUT/q_bench.err
