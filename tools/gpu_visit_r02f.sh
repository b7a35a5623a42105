#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== chunk debug"; timeout 300 python tools/chunk_debug.py 0 2>&1 | grep -v "^   cache equal: True" | cut -c1-130 | tail -28 | tee $OUT/r02f_chunk_debug.txt
for v in 0; do
echo "== timeline C3 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 8 8208 16 h2o_head $v 2>&1 | tee -a $OUT/r02f_timeline.txt
echo "== timeline C2 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 32 1088 64 roco $v 2>&1 | tee -a $OUT/r02f_timeline.txt
done
echo "== chunk sweep"; timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02f_sweep_chunk.jsonl | cut -c1-260
echo "== cluster decode timelines"
for args in "8 64 8 8256 0" "16 32 8 8208 0" "1 32 32 1088 0" "8 32 32 1088 0"; do echo "-- $args"; timeout 120 python tools/cluster_timeline.py $args 2>&1 | tee -a $OUT/r02f_cluster_timeline.txt; done
echo "== full gpu suite"; timeout 2400 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/r02f_pytest.txt
