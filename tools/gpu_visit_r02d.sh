#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
python -m easykv_b200.build > /dev/null
echo "== chunk debug nsplit4"; timeout 300 python tools/chunk_debug.py 0 2>&1 | grep -v "^   cache equal: True" | cut -c1-150 | tail -30 | tee $OUT/r02d_chunk_debug.txt
echo "== chunk debug nsplit2"; timeout 300 python tools/chunk_debug.py 3 2>&1 | grep -v "^   cache equal: True" | cut -c1-150 | tail -30 | tee -a $OUT/r02d_chunk_debug.txt
for v in 0 3; do
echo "== timeline C3 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 8 8208 16 h2o_head $v 2>&1 | tee -a $OUT/r02d_timeline.txt
echo "== timeline C2 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 32 1088 64 roco $v 2>&1 | tee -a $OUT/r02d_timeline.txt
done
echo "== chunk sweep nsplit4"; timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02d_sweep_chunk.jsonl | cut -c1-260
echo "== chunk sweep nsplit2"; EKV_CHUNK_VARIANT=3 timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02d_sweep_chunk_nsplit2.jsonl | cut -c1-260
