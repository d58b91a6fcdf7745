#!/bin/bash
# ncu evidence for the kernels of the calling path (run under gpurun, one GPU). Outputs land in gpurun_out/.
#   tools/profile.sh [scale]        then, back in the container:  python tools/summarize_profiles.py r01
set -u
SCALE=${1:-0.05}
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/prof_*.ncu-rep
# 1. launch list: every kernel launch of one short bench run with its device time (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --scale $SCALE --cpu-sample 200 > $OUT/ncu_bench.log 2>&1
# 2. full captures, one launch of each kernel of the resident leg (skip the warm-up launch)
for K in assemble_kernel align_kernel kmer_kernel al_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_$K -f \
      python bench.py --steps 1 --warmup 1 --scale $SCALE --cpu-sample 200 > $OUT/ncu_$K.log 2>&1
done
# 3. kernel 2 alone at fixed shapes (tools/ksw_bench.py, one full wave of 14208 alignments): call-site A 300x420 (launch 0) and call-site B 150x700 (launch 9)
ncu --set full --clock-control none --import-source on -k regex:ksw2_batch_kernel -s 0 -c 1 -o $OUT/prof_ksw2_siteA_300x420 -f python tools/ksw_bench.py 14208 > $OUT/ncu_kswA.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ksw2_batch_kernel -s 9 -c 1 -o $OUT/prof_ksw2_siteB_150x700 -f python tools/ksw_bench.py 14208 > $OUT/ncu_kswB.log 2>&1
# 4. DRAM traffic, executed thread instructions and pipe utilisation of every pipeline kernel at the FULL default workload (second launch of each)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:'assemble_kernel|align_kernel|kmer_kernel|al_kernel' --csv --log-file $OUT/traffic_full.csv \
    python bench.py --steps 1 --warmup 1 --cpu-sample 200 > $OUT/ncu_traffic.log 2>&1
# 5. the dominant kernel at the FULL default workload: one full capture of the resident leg's al_kernel launch
ncu --set full --clock-control none --import-source on -k regex:al_kernel -s 1 -c 1 -o $OUT/prof_al_kernel_full -f \
    python bench.py --steps 1 --warmup 1 --cpu-sample 200 > $OUT/ncu_al_full.log 2>&1
ls -la $OUT | head -40
