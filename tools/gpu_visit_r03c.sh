#!/bin/bash
set -u
OUT=gpurun_out
echo "-- pool refresh"; timeout 300 python tools/decode_ab.py c2 2>&1 | grep auto
echo "-- fixed inputs"; EKV_BENCH_FIXED_INPUTS=1 timeout 300 python tools/decode_ab.py c2 2>&1 | grep auto
echo "-- sweep.py 7B b64"; timeout 300 python tools/sweep.py decode 2>&1 | grep '"7B b64"' | cut -c1-260
