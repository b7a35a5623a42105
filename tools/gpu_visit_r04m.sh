#!/bin/bash
# final build: the whole GPU suite, smoke, and the default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r04m_pytest.txt
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2 | tee gpurun_out/r04m_smoke.txt
timeout 600 python bench.py 2>&1 | tail -1 > gpurun_out/r04m_bench.json; cut -c1-300 gpurun_out/r04m_bench.json
