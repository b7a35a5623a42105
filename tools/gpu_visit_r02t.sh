#!/bin/bash
set -u
OUT=gpurun_out
echo "== decode parity tests"; timeout 1500 python -m pytest tests -m gpu -q -k "decode or golden or reference or fullsize or driver or select" 2>&1 | tail -40 | cut -c1-300 | tee $OUT/r02t_pytest.txt
