#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== umma probes"; timeout 300 python -m pytest tests/test_gpu_umma.py -m gpu -q 2>&1 | tail -4
echo "== decode debug (tcgen05 GQA decode kernel)"; timeout 300 python tools/decode_debug.py 5 0 2>&1 | cut -c1-220 | tail -70 | tee $OUT/r02j_decode_debug.txt
echo "== forced cluster 2"; timeout 300 python tools/decode_debug.py 5 2 2>&1 | cut -c1-200 | grep -v "victims_equal True$" | tail -20 | tee -a $OUT/r02j_decode_debug.txt
echo "== forced cluster 4"; timeout 300 python tools/decode_debug.py 5 4 2>&1 | cut -c1-200 | grep -v "victims_equal True$" | tail -20 | tee -a $OUT/r02j_decode_debug.txt
echo "== racecheck cluster leg details"; timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool racecheck --print-limit 12 python tools/sanitize_targets.py cluster 2>&1 | grep -v "^=========     at\|^=========         in\|^=========     Host Frame\|^=========                in" | head -80 | tee $OUT/r02j_racecheck_cluster.txt
