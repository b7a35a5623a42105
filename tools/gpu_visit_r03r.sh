#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_ragged.py -m gpu -q 2>&1 | tail -4 | cut -c1-200 | tee $OUT/r03r_pytest_new.txt
timeout 400 python tools/sweep.py decode 2>&1 | grep -i "stream\|7B b64\"\|7B b1\"\|7B b8\"" | cut -c1-330 | tee $OUT/r03r_sweep_stream.jsonl
timeout 600 python tools/decode_ab.py c2 c2_b1 c2_b8 c3_decode c5 2>&1 | grep auto | tee $OUT/r03r_decode_ab.jsonl
