#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 60 -c 1 -f -o $OUT/r02x_decode_c2 python bench.py --workload c2 --no-sweep --no-cpu-baseline --no-gpu-reference --steps 2 --warmup 3 --min-seconds 0.05 --layers 4 > $OUT/r02x_c2.log 2>&1
tail -3 $OUT/r02x_c2.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:decode_umma_kernel -s 60 -c 1 -f -o $OUT/r02x_decode_c5 python bench.py --workload c5 --no-sweep --no-cpu-baseline --no-gpu-reference --steps 2 --warmup 3 --min-seconds 0.05 --layers 4 > $OUT/r02x_c5.log 2>&1
tail -3 $OUT/r02x_c5.log | cut -c1-300
ls -la $OUT/r02x*
