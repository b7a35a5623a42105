#!/bin/bash
set -u
OUT=gpurun_out
for args in "1 64 8 8256 2" "8 64 8 8256 2" "8 32 32 1088 1" "32 32 8 1088 1" "8 64 8 8256 2 h2o_head"; do echo "-- $args"; timeout 120 python tools/decode_umma_timeline.py $args 2>&1 | tee -a $OUT/r02l_decode_umma_timeline.txt; done
