#!/bin/bash
set -u
OUT=gpurun_out
for pol in recency h2o_head tova roco; do echo "-- c5 b8 $pol"; timeout 120 python tools/decode_umma_timeline.py 8 64 8 8256 0 $pol 2>&1 | tee -a $OUT/r02v_decode_umma_timeline.txt; done
for pol in recency roco; do echo "-- c5 b8 C=4 $pol"; timeout 120 python tools/decode_umma_timeline.py 8 64 8 8256 4 $pol 2>&1 | tee -a $OUT/r02v_decode_umma_timeline.txt; done
