#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== decode A/B"; timeout 900 python tools/decode_ab.py c5 c5_b32 c3_decode c3_gen m7b_n1088_b32 70b_n1088_b32 c5_b1 c3_decode_b1 c3_decode_b4 c2_b8 c2_b1 2>&1 | tee $OUT/r02k_decode_ab.jsonl
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -8 | tee $OUT/r02k_pytest.txt
