#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
echo "== new parity tests"; timeout 1500 python -m pytest tests/test_gpu_umma.py tests/test_gpu_fullsize_vs_restate.py tests/test_gpu_reference_live.py tests/test_gpu_reference_goldens.py -m gpu -q 2>&1 | tail -25 | tee $OUT/r02c_pytest_new.txt
echo "== chunk sweep, tcgen05 path"; timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/r02c_sweep_chunk_umma.jsonl | cut -c1-260
echo "== ncu launch list: chunk kernels (Mistral stride 16, 8 sequences)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chunk_ --csv --log-file $OUT/r02c_chunk_launches.csv python tools/chunk_profile.py 8 32 8 8208 16 h2o_head > /dev/null 2>&1
grep -E "chunk_" $OUT/r02c_chunk_launches.csv | awk -F'","' '{print $5, $NF}' | tail -8
echo "== ncu full: chunk_umma_kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chunk_umma_kernel -s 4 -c 2 -f -o $OUT/r02c_chunk_umma python tools/chunk_profile.py 8 32 8 8208 16 h2o_head > /dev/null 2>&1
ls -la $OUT | tail -8
