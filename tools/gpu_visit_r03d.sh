#!/bin/bash
set -u
for v in 0 8 7; do timeout 200 python tools/timeline_bench.py c2 32 $v 2>&1 | tail -3; done
EKV_BENCH_FIXED_INPUTS=1 timeout 200 python tools/timeline_bench.py c2 32 0 2>&1 | tail -3
