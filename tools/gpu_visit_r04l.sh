#!/bin/bash
# cluster kernel: slice prefetch into L2 ahead of the PDL wait — parity, then small-batch timings (graph and eager)
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream_ragged.py tests/test_gpu_reference_goldens.py tests/test_gpu_fullsize_vs_restate.py -m gpu -q -x --timeout 180 2>&1 | tail -2 | cut -c1-200
timeout 300 python tools/pdl_probe.py c2_b1 2>&1 | grep workload
EKV_NO_PDL=1 timeout 300 python tools/pdl_probe.py c2_b1 2>&1 | grep workload
timeout 600 python tools/decode_ab.py c2_b1 c3_decode_b1 c5_b1 c3_decode_b4 2>&1 | grep '"auto"'
