#!/bin/bash
set -u
OUT=gpurun_out
timeout 200 python tools/decode_umma_timeline_bench.py c5 2>&1 | tail -9 | tee $OUT/r03e_umma_tl.txt
timeout 200 python tools/decode_umma_timeline_bench.py c5_b32 2>&1 | tail -9 | tee -a $OUT/r03e_umma_tl.txt
timeout 200 python tools/decode_umma_timeline_bench.py c3_decode 2>&1 | tail -9 | tee -a $OUT/r03e_umma_tl.txt
timeout 600 python tools/decode_ab.py c2 c2_b1 c2_b8 c4_roco 2>&1 | grep auto | tee $OUT/r03e_decode_ab.jsonl
