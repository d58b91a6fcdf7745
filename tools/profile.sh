#!/bin/bash
# ncu evidence for the kernels of the calling path at the FULL default workload (run under gpurun, one GPU). Outputs land in gpurun_out/.
#   tools/profile.sh        then, back in the container:  python tools/summarize_profiles.py r02
set -u
OUT=gpurun_out
mkdir -p $OUT
rm -f $OUT/prof_*.ncu-rep
B="python bench.py --steps 1 --warmup 1 --cpu-sample 200 --no-bam-leg"
# 1. launch list: every kernel launch of one short bench run with its device time (shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --cpu-sample 200 --no-bam-leg > $OUT/ncu_bench.log 2>&1
# 2. full captures: the resident leg's launch of each kernel (launch 0 is the warm-up step)
for K in assemble_kernel align4_kernel kmer_kernel al_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 1 -c 1 -o $OUT/prof_${K}_full -f $B > $OUT/ncu_$K.log 2>&1
done
# 3. DRAM traffic, executed thread instructions and pipe utilisation of every pipeline kernel (second launch of each)
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__thread_inst_executed.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_write.sum \
    --clock-control none -k regex:'assemble_kernel|align4_kernel|kmer_kernel|al_kernel' --csv --log-file $OUT/traffic_full.csv $B > $OUT/ncu_traffic.log 2>&1
# 4. gen_roi on the GPU: every pass of idl_sweep on the 248 Mb contig
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'sw_' -c 40 --csv \
    --log-file $OUT/sweep_kernels.csv python tools/sweep_bench.py chr1 1 > $OUT/ncu_sweep.log 2>&1
# 5. the BAM decoder: launch list of idl_bam_open + idl_bam_sweep-free path on a 10 Mb BAM with per-base qualities, full capture of the inflate kernel
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:'bgzf_|bam_|scan_' -c 60 --csv \
    --log-file $OUT/bam_kernels.csv python tools/bam_bench.py 10 8 1 1 --no-host > $OUT/ncu_bam.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bgzf_inflate -s 1 -c 1 -o $OUT/prof_bgzf_inflate_kernel -f python tools/bam_bench.py 10 8 1 2 --no-host > $OUT/ncu_inflate.log 2>&1
ls -la $OUT | head -40
