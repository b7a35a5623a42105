#!/bin/bash
set -u
OUT=gpurun_out
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -6 | cut -c1-300 | tee $OUT/r02w_pytest.txt
echo "== decode A/B"; timeout 900 python tools/decode_ab.py c2 c2_b1 c2_b8 c3_decode c5 c5_b32 2>&1 | grep "auto" | tee $OUT/r02w_decode_ab.jsonl
for args in "8 64 8 8256 0"; do echo "-- $args"; timeout 120 python tools/decode_umma_timeline.py $args 2>&1 | tee -a $OUT/r02w_decode_umma_timeline.txt; done
echo "-- timeline persistent"; B=64 timeout 120 python tools/timeline.py 2>&1 | tail -30 | tee $OUT/r02w_timeline.txt
