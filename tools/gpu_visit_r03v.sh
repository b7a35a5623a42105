#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "umma or decode or stream or ragged" 2>&1 | tail -4 | cut -c1-200
timeout 200 python tools/decode_umma_timeline_bench.py c5 2>&1 | tail -16
timeout 200 python tools/decode_umma_timeline_bench.py c3_decode 2>&1 | tail -8
