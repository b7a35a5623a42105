#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== chunk debug (tcgen05 path)"; timeout 300 python tools/chunk_debug.py 0 2>&1 | tail -60 | tee $OUT/r02b_chunk_debug.txt
echo "== reference on the B200: goldens"; timeout 900 python -m oracle.gen_golden_gpu 2>&1 | tail -12 | tee $OUT/r02b_gen_golden_gpu.txt
echo "== replay statistics (B200-recorded goldens)"; timeout 900 python tools/gpu_ref_report.py $OUT/golden_gpu 2>&1 | tee $OUT/r02b_ref_report_gpu.jsonl | cut -c1-600
ls -la $OUT/golden_gpu
