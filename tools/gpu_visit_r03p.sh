#!/bin/bash
timeout 400 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 200 python tools/stream_probe.py 2 4 2 260 float16 3 1 2>&1 > /tmp/race.txt
grep -n "Error:" /tmp/race.txt | head -5
grep -A6 "Error:" /tmp/race.txt | grep -E "Error:|Thread|Current" | head -24 | cut -c1-230
