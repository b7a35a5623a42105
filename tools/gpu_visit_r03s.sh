#!/bin/bash
set -u
OUT=gpurun_out
timeout 600 python tools/decode_variants.py c2 c2_b8 2>&1 | tee $OUT/r03s_decode_variants.jsonl
timeout 200 python tools/timeline_bench.py c2 32 0 2>&1 | tail -3 | tee $OUT/r03s_timeline_bench.txt
