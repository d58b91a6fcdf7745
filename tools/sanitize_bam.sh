#!/bin/bash
# compute-sanitizer over the BAM decoder and the batch builder (run under gpurun): idl_bam_open / sweep / fetch / pack / submit through tests/test_gpu_bam.py
OUT=gpurun_out; mkdir -p $OUT
{
echo "compute-sanitizer (B200, round 2) on pytest tests/test_gpu_bam.py -m gpu"
for T in memcheck synccheck initcheck; do
  echo "== $T"
  timeout 1500 compute-sanitizer --tool $T --error-exitcode 0 python -m pytest tests/test_gpu_bam.py -m gpu -q 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Invalid|hazard|Uninit|error" | tail -8
done
echo "== racecheck (shared-memory hazards: the decode tables of the inflate kernel)"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 0 python -m pytest tests/test_gpu_bam.py -m gpu -q -k "members_of_every_kind or long_records or bait or empty or odd_letters" 2>&1 | grep -E "passed|failed|RACECHECK SUMMARY|hazard|error" | tail -12
} > $OUT/r02_sanitizer_bam.txt 2>&1
cat $OUT/r02_sanitizer_bam.txt
