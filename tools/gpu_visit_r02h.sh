#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== chunk debug"; timeout 300 python tools/chunk_debug.py 0 2>&1 | grep -v "^   cache equal: True" | awk '{print $1,$2,$3,$4,$5,$6,$7,$11,$17,$18}' | tail -28 | tee $OUT/r02h_chunk_debug.txt
echo "== timeline C3 b8"; timeout 120 python tools/umma_timeline.py 8 32 8 8208 16 h2o_head 0 2>&1 | tee -a $OUT/r02h_timeline.txt
echo "== timeline C2 b8"; timeout 120 python tools/umma_timeline.py 8 32 32 1088 64 roco 0 2>&1 | tee -a $OUT/r02h_timeline.txt
echo "== chunk sweep"; timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02h_sweep_chunk.jsonl | cut -c1-260
echo "== chunk parity tests"; timeout 1200 python -m pytest tests -m gpu -q -x -k "chunk or golden or reference or fullsize" 2>&1 | tail -5 | tee $OUT/r02h_pytest.txt
