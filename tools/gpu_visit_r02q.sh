#!/bin/bash
set -u
OUT=gpurun_out
echo "== decode debug"; timeout 300 python tools/decode_debug.py 5 0 2>&1 | cut -c1-200 | grep -v "victims_equal True$" | tail -20
echo "== forced cluster 2"; timeout 300 python tools/decode_debug.py 5 2 2>&1 | cut -c1-200 | grep -v "victims_equal True$" | tail -20
for args in "8 64 8 8256 0" "32 32 8 1088 0" "16 32 8 8208 0"; do echo "-- $args"; timeout 120 python tools/decode_umma_timeline.py $args 2>&1 | tee -a $OUT/r02q_decode_umma_timeline.txt; done
echo "== decode A/B"; timeout 900 python tools/decode_ab.py c5 c5_b32 c3_decode c3_decode_b4 2>&1 | grep -v "umma_c1\|umma_c4" | tee $OUT/r02q_decode_ab.jsonl
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -x -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -8 | tee $OUT/r02q_pytest.txt
