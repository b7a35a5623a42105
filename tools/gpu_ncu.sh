#!/bin/bash
# `ncu --set full` captures of a round, one launch each.  The reports travel back through gpurun_out/ (64 MiB at most per
# visit), so the headline kernel and the two tcgen05 kernels are separate visits:
#   gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh r02 decode'      gpurun --timeout 1500 -- 'bash tools/gpu_ncu.sh r02 umma'
set -u
TAG=${1:-r02}
WHAT=${2:-decode}
OUT=gpurun_out
mkdir -p $OUT
if [ "$WHAT" = decode ]; then
  echo "== decode kernel (bench state)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 60 -c 1 -f -o $OUT/${TAG}_decode \
    python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 --layers 4 > $OUT/${TAG}_decode_cmd.log 2>&1
else
  echo "== tcgen05 chunk kernel (Mistral stride 16, 8 sequences)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:chunk_umma_kernel -s 4 -c 1 -f -o $OUT/${TAG}_chunk_umma python tools/chunk_profile.py 8 32 8 8208 16 h2o_head > /dev/null 2>&1
  echo "== tcgen05 GQA decode kernel (70B layout, 8 sequences, n = 8256, bench state)"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_umma_kernel -s 20 -c 1 -f -o $OUT/${TAG}_decode_umma \
    python bench.py --workload c5 --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 --layers 4 > /dev/null 2>&1
fi
ls -la $OUT/${TAG}_*.ncu-rep
