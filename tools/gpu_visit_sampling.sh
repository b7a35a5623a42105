#!/bin/bash
# Short box visit for the sampling / perplexity tail: full GPU parity suite, the tail's timings, one end-to-end run.
set -u
TAG=${1:-r01g}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== sampling tail timings" ; timeout 200 python tools/time_sampling.py 2>&1 | tee $OUT/${TAG}_sampling.jsonl
echo "== end to end generate (7B shape)" ; timeout 400 python tools/e2e_generate.py --layers 32 --prompt 4096 --new 64 2>&1 | tail -1 | tee $OUT/${TAG}_e2e_llama7b.json
