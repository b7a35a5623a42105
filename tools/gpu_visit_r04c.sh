#!/bin/bash
# head_dim 64 / 96 through the exact kernel: restatement parity, the reference's new goldens, the driver end to end
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_driver.py tests/test_gpu_stream_ragged.py -m gpu -q -x --timeout 180 -k "head_dim or d64 or d96 or general or ragged" 2>&1 | tail -8 | cut -c1-300
