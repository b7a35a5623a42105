#!/bin/bash
timeout 900 python -m pytest tests -m gpu -q -x --timeout 180 2>&1 | tail -3 | cut -c1-200
timeout 600 python tools/decode_ab.py c2 c2_b1 c2_b8 c3_gen c3_decode c5 c5_b32 2>&1 | grep auto
