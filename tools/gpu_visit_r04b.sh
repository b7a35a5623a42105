#!/bin/bash
# chunk kernels with programmatic dependent launch: parity first, then the chunk sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_umma.py tests/test_gpu_parity.py tests/test_gpu_fullsize_vs_restate.py tests/test_gpu_reference_goldens.py -m gpu -q -x --timeout 180 2>&1 | tail -3 | cut -c1-200
timeout 400 python tools/sweep.py chunk > gpurun_out/sweep_chunk_pdl.jsonl 2> gpurun_out/sweep_chunk_pdl.err
cat gpurun_out/sweep_chunk_pdl.jsonl | cut -c1-260
