#!/bin/bash
timeout 300 python tools/stream_probe.py 2 8 4 2304 float16 0 0 2>&1 | grep -E "DBG|ok" | sort | uniq -c | sort -rn | head -30 | cut -c1-200
