OUT=gpurun_out; mkdir -p $OUT
(timeout 300 python -m pytest tests/test_gpu_ksw2.py -m gpu -x -q 2>&1 | tail -8) > $OUT/q_tests.log
timeout 100 python tools/ksw_bench.py 14208 > $OUT/q_kswbench.log 2>&1
timeout 100 python tools/ksw_bench.py > $OUT/q_kswbench4.log 2>&1
timeout 300 python bench.py --steps 3 --warmup 3 --cpu-sample 200 > $OUT/q_bench.json 2> $OUT/q_bench.err
cat $OUT/q_tests.log $OUT/q_kswbench.log $OUT/q_kswbench4.log; python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/q_bench.json")); print(d["value"], d["ms_per_step"], d["kernel_ms"], d["e2e"]["value"])
except Exception as e: print("bench failed", e)
PY
tail -3 $OUT/q_bench.err
