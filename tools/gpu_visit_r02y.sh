#!/bin/bash
set -u
OUT=gpurun_out
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -12 | cut -c1-300 | tee $OUT/r02y_pytest.txt
echo "== decode A/B"; timeout 900 python tools/decode_ab.py c2 c2_b1 c2_b8 c4_roco c3_decode c5 c5_b32 2>&1 | grep "auto" | tee $OUT/r02y_decode_ab.jsonl
echo "-- timeline persistent"; B=64 timeout 120 python tools/timeline.py 2>&1 | tail -12 | tee $OUT/r02y_timeline.txt
