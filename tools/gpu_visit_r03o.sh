#!/bin/bash
for args in "2 4 2 260 float16 3 1" "2 4 2 2304 float16 3 1" "2 16 2 333 float16 3 1"; do
  r=$(timeout 120 python tools/stream_probe.py $args 2>&1 | grep -v "^$" | tail -1 | cut -c1-70); echo "$args: $r"
done
timeout 300 compute-sanitizer --tool memcheck --print-limit 1 python tools/stream_probe.py 2 4 2 260 float16 3 1 2>&1 | grep -E "Invalid|at |by thread|Access|Device Frame|nearest|ok" | head -10 | cut -c1-220
