#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_stream_ragged.py -m gpu -q 2>&1 | tail -30 | cut -c1-260 | tee $OUT/r03h_pytest_new.txt
timeout 1500 python -m pytest tests -m gpu -q -x -k "stream or golden or driver" 2>&1 | tail -8 | cut -c1-260 | tee $OUT/r03h_pytest.txt
timeout 300 python tools/sweep.py decode 2>&1 | grep -i "stream\|7B b64\"" | cut -c1-300 | tee $OUT/r03h_sweep_stream.jsonl
