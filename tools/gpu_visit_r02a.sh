#!/bin/bash
# Round-2 visit A: tensor-core primitive probe, reference-on-B200 goldens (+ topk tie probe), replay statistics.
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/r02a_smi.txt 2>&1
echo "== umma probe"; timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -x -q 2>&1 | tail -15 | tee $OUT/r02a_umma.txt
echo "== reference on the B200: goldens"; timeout 900 python -m oracle.gen_golden_gpu 2>&1 | tail -30 | tee $OUT/r02a_gen_golden_gpu.txt
echo "== replay statistics (B200-recorded goldens)"; timeout 900 python tools/gpu_ref_report.py $OUT/golden_gpu 2>&1 | tee $OUT/r02a_ref_report_gpu.jsonl | cut -c1-400
echo "== replay statistics (CPU-recorded goldens, arith 0/1)"; timeout 900 python tools/gpu_ref_report.py tests/golden 2>&1 | tee $OUT/r02a_ref_report_cpu.jsonl | cut -c1-300 | tail -20
echo "== sweeps (round-1 kernels, today's box)"
timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02a_sweep_chunk.jsonl | cut -c1-250
ls -la $OUT/golden_gpu
