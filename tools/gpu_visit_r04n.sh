#!/bin/bash
# the entry-limit query against the kernels' real ceilings; then the whole GPU suite once more on the final library
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "entry_limit" 2>&1 | tail -15 | cut -c1-250
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r04n_pytest.txt
