#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python tools/decode_tune_ab.py c2 c4_roco c2_b8 2>&1 | tee $OUT/r03a_tune_ab.jsonl
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -6 | cut -c1-300 | tee $OUT/r03a_pytest.txt
