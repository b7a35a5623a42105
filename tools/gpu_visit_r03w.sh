#!/bin/bash
timeout 200 python tools/decode_umma_timeline_bench.py c5 2>&1 | tail -10
timeout 600 python -m pytest tests -m gpu -q -x -k "umma or decode or select or golden or fullsize" 2>&1 | tail -3 | cut -c1-200
timeout 300 python tools/decode_ab.py c2_b1 c3_decode c5 c5_b32 2>&1 | grep auto
