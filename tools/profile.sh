#!/bin/bash
# ncu evidence for the kernels of the calling path (run under gpurun, one GPU). Outputs land in gpurun_out/.
#   tools/profile.sh [scale]
set -u
SCALE=${1:-0.05}
OUT=gpurun_out
mkdir -p $OUT
# 1. launch list: every kernel launch of one short bench run with its device time (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --scale $SCALE --cpu-sample 200 > $OUT/ncu_bench.log 2>&1
# 2. full captures, one launch of each kernel (skip the warm-up launch)
for K in assemble_kernel align_kernel kmer_kernel al_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_$K -f \
      python bench.py --steps 1 --warmup 1 --scale $SCALE --cpu-sample 200 > $OUT/ncu_$K.log 2>&1
done
ls -la $OUT
