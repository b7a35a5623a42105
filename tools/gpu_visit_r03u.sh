#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "decode or stream or ragged or golden or fullsize" 2>&1 | tail -4 | cut -c1-200
timeout 300 python tools/decode_ab.py c3_gen c3_decode c5 2>&1 | grep auto
timeout 2400 bash tools/sanitize.sh 2>&1 | tail -60 | cut -c1-220
