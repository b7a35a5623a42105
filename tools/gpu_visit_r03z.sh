#!/bin/bash
OUT=gpurun_out
timeout 600 python bench.py 2>&1 | tail -1 | tee $OUT/r02z_bench.json | cut -c1-200
