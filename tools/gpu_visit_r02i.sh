#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== timeline C3 b8"; timeout 120 python tools/umma_timeline.py 8 32 8 8208 16 h2o_head 0 2>&1 | tee $OUT/r02i_timeline.txt
echo "== sanitize targets natively first"; timeout 300 python tools/sanitize_targets.py 2>&1 | tail -14
echo "== sanitizer"; bash tools/sanitize.sh 2>&1 | tail -60
