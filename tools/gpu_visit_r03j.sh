#!/bin/bash
for args in "1 16 2 333 bfloat16 3 0" "1 16 2 333 float16 3 0" "1 16 2 333 bfloat16 0 0" "2 16 2 333 bfloat16 3 0" "1 16 2 333 bfloat16 3 1" "1 16 2 333 bfloat16 3 2" "1 16 2 640 bfloat16 3 0" "2 8 4 2304 float16 0 0" "2 8 4 2304 float16 0 2"; do
  timeout 120 python tools/stream_probe.py $args 2>&1 | grep -v "^$" | tail -2 | cut -c1-220
done
