#!/bin/bash
set -u
OUT=gpurun_out
timeout 900 python -m pytest tests/test_driver.py tests/test_gpu_reference_live.py -m gpu -q -x 2>&1 | tail -4 | cut -c1-200
timeout 600 python tools/e2e_generate.py --layers 32 --prompt 4096 --new 256 2>&1 | tail -1 | tee $OUT/r03t_e2e_llama7b.json
timeout 600 python tools/e2e_generate.py --arch mistral --layers 32 --prompt 16384 --new 16 --mode encoding --budget 0.5 --stride 16 --policy h2o --keep-attention 2>&1 | tail -1 | tee $OUT/r03t_e2e_mistral7b.json
