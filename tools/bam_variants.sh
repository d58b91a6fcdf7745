#!/bin/bash
# compile-time variants of the BAM decoder kernels, timed with tools/bam_bench.py (run under gpurun; nvcc is on the box):
#   tools/bam_variants.sh "" "-DINF_CTAS_PER_SM_=6" ...
mkdir -p gpurun_out
for V in "$@"; do
  IDL_NVCC_EXTRA="$V" python -c "from indelope_b200 import build as b; b.build_cuda(force=True)" > /dev/null 2>&1 || { echo "$V: build failed"; continue; }
  echo "== $V"
  timeout 300 python tools/bam_bench.py 10 8 1 5 --no-host 2> gpurun_out/bamvar.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); i=d['idl_bam_open']
print('copy+inflate %.2f ms (%.1f GB/s out, %.1f in)  parse %.2f ms  h2d %.2f ms  wall %.1f ms' % (i['ms_inflate'], d['copy_inflate_gbs_out'], d['copy_inflate_gbs_in'], i['ms_parse'], i['ms_h2d'], i['wall_s']*1e3))"
done
python -c "from indelope_b200 import build as b; b.build_cuda(force=True)" > /dev/null 2>&1
