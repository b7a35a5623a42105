#!/bin/bash
# compute-sanitizer over small invocations of every kernel family (memcheck + racecheck + synccheck).
# Run as: gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck racecheck synccheck; do
  echo "== $tool: smoke (tensor-core chunk path + cluster decode kernel, GQA fp16)"
  timeout 600 $CS --tool $tool --error-exitcode 9 python __graft_entry__.py --smoke 2>&1 | tail -4
  echo "rc=$?"
done
echo "== memcheck: persistent decode kernel (MHA), general kernel, select / export / evict_explicit / rope, cluster C=2,8 and tensor-core decode"
timeout 900 $CS --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "(decode_random and (cluster-1 or cluster2 or cluster8) and (roco or h2o)) or rope_kernel or (select_matches and 144) or (chunk_random and general and recency)" 2>&1 | tail -4
echo "== racecheck: cluster decode C=2 + persistent decode"
timeout 900 $CS --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
  -k "decode_random and (cluster-1 or cluster2) and roco" 2>&1 | tail -4
