#!/bin/bash
timeout 300 compute-sanitizer --tool memcheck --print-limit 2 python tools/stream_probe.py 2 8 4 2304 float16 0 0 2>&1 | grep -E "Invalid|at |by thread|Access|Device Frame|nearest|ok" | head -24 | cut -c1-250
echo ---
timeout 300 compute-sanitizer --tool memcheck --print-limit 2 python tools/stream_probe.py 2 8 2 260 float16 0 0 2>&1 | grep -E "Invalid|at |by thread|Access|Device Frame|nearest|ok" | head -24 | cut -c1-250
