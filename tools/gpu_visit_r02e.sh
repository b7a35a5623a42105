#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== chunk debug"; timeout 300 python tools/chunk_debug.py 0 2>&1 | grep -v "^   cache equal: True" | cut -c1-130 | tail -28 | tee $OUT/r02e_chunk_debug.txt
for v in 0 1536 2048 1024; do
echo "== timeline C3 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 8 8208 16 h2o_head $v 2>&1 | tee -a $OUT/r02e_timeline.txt
done
for v in 0 256 512; do
echo "== timeline C2 b8 variant $v"; timeout 120 python tools/umma_timeline.py 8 32 32 1088 64 roco $v 2>&1 | tee -a $OUT/r02e_timeline.txt
done
echo "== timeline C5 b1"; timeout 120 python tools/umma_timeline.py 1 64 8 8256 64 roco 0 2>&1 | tee -a $OUT/r02e_timeline.txt
echo "== chunk sweep"; timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02e_sweep_chunk.jsonl | cut -c1-260
echo "== bench (short)"; timeout 900 python bench.py --steps 5 --warmup 3 --min-seconds 0.3 2>&1 | tail -3 | tee $OUT/r02e_bench.json | cut -c1-6000
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 5 --warmup 2 2>&1 | tail -2 | tee $OUT/r02e_bench_ref.json | cut -c1-1500
