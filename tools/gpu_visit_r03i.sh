#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gpu_stream_ragged.py -m gpu -q -x -k "1-16-2-333" 2>&1 | grep -v "^$" | head -60 | cut -c1-220 | tee $OUT/r03i_memcheck.txt
for k in "2-8-4-2304" "ragged"; do timeout 600 python -m pytest tests/test_gpu_stream_ragged.py -m gpu -q -k "$k" 2>&1 | tail -25 | cut -c1-220; done | tee $OUT/r03i_pytest.txt
