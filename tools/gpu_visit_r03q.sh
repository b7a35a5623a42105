#!/bin/bash
for args in "2 4 2 260 float16 3 1" "2 4 2 2304 float16 3 1" "2 16 2 333 float16 3 1" "2 16 2 333 bfloat16 3 0" "2 8 4 2304 float16 0 0"; do
  r=$(timeout 120 python tools/stream_probe.py $args 2>&1 | grep -v "^$" | tail -1 | cut -c1-70); echo "$args: $r"
done
