#!/bin/bash
# where does an 8-sequence step of the persistent kernel spend its time?  (phase timeline at B = 8 and B = 1-cluster)
timeout 200 python tools/timeline_bench.py c2 8 0 2>&1 | tail -4
timeout 200 python tools/timeline_bench.py c2 8 1 2>&1 | tail -4
timeout 200 python tools/timeline_bench.py c2 16 0 2>&1 | tail -4
