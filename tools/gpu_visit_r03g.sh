#!/bin/bash
set -u
timeout 200 python tools/timeline_bench.py c2 32 0 2>&1 | tail -3
timeout 200 python tools/timeline_bench.py c2_flat 32 0 2>&1 | tail -3
EKV_BENCH_FIXED_INPUTS=1 timeout 200 python tools/timeline_bench.py c2 32 0 2>&1 | tail -3
