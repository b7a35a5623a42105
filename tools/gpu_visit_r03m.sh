#!/bin/bash
for args in "2 8 4 2304 float16 0 0" "1 16 2 333 bfloat16 3 0" "2 8 2 260 float16 0 0"; do
  timeout 120 python tools/stream_probe.py $args 2>&1 | grep -v "^$" | tail -1 | cut -c1-160
done
CUDA_LAUNCH_BLOCKING=1 timeout 120 python tools/stream_probe.py 2 8 4 2304 float16 0 0 2>&1 | grep -v "^$" | tail -1 | cut -c1-160
