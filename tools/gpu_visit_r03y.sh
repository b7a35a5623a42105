#!/bin/bash
timeout 300 python bench.py --no-sweep --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 --min-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline'], j['clocks'])"
timeout 300 python bench.py --workload c5 --no-sweep --no-cpu-baseline --no-gpu-reference --steps 5 --warmup 3 --min-seconds 0.3 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['roofline'], j['clocks'])"
