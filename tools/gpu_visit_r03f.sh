#!/bin/bash
set -u
OUT=gpurun_out
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -6 | cut -c1-300 | tee $OUT/r03f_pytest.txt
timeout 200 python tools/decode_umma_timeline_bench.py c5 2>&1 | tail -9 | tee $OUT/r03f_umma_tl.txt
timeout 600 python tools/decode_ab.py c3_decode c5 c5_b32 2>&1 | grep auto | tee $OUT/r03f_decode_ab.jsonl
