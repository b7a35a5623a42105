#!/bin/bash
# leaner classify pass / parallel boundary ranks in the bucket select: parity, then A/B on the decode workloads + timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_umma.py tests/test_gpu_stream_ragged.py tests/test_gpu_fullsize_vs_restate.py tests/test_gpu_reference_goldens.py -m gpu -q -x --timeout 180 2>&1 | tail -3 | cut -c1-200
timeout 600 python tools/decode_ab.py c5 c5_b32 c3_decode c2_b8 c2_b1 2>&1 | grep '"auto"'
timeout 200 python tools/decode_umma_timeline_bench.py c5 > gpurun_out/r04d_decode_umma_timeline.txt 2>&1
head -30 gpurun_out/r04d_decode_umma_timeline.txt
