#!/bin/bash
timeout 120 python tools/decode_umma_timeline_bench.py c5 2>&1 | tail -16
timeout 600 python -m pytest tests -m gpu -q -x -k "umma or decode or select or golden or fullsize or stream or ragged" --timeout 120 2>&1 | tail -3 | cut -c1-200
timeout 300 python tools/decode_ab.py c3_gen c3_decode c5 c5_b32 2>&1 | grep auto
