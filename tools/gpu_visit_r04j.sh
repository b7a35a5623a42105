#!/bin/bash
# kernel durations (ncu, serialised) against the per-layer time of the replayed step, small batches
mkdir -p gpurun_out
for w in c2_b8 c2_b1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_' -c 120 --csv --log-file gpurun_out/r04j_${w}_launches.csv \
    python bench.py --workload $w --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 > /dev/null 2>&1
  python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/r04j_${w}_launches.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
ki=rows[h].index('Kernel Name'); vi=rows[h].index('Metric Value')
v=[float(r[vi].replace(',','')) for r in rows[h+1:] if len(r)>vi and r[vi].replace(',','').replace('.','').isdigit()]
print('${w}', 'kernel', rows[h+1][ki][:60], 'n', len(v), 'mean us', sum(v)/len(v)/1e3, 'min', min(v)/1e3, 'max', max(v)/1e3)
PY
  timeout 300 python bench.py --workload $w --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w bench ms_per_step', d['ms_per_step'], 'avg_launch_us', d['roofline']['avg_launch_us'], d['roofline']['frac'])"
  EKV_BENCH_FIXED_INPUTS=1 timeout 300 python bench.py --workload $w --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w bench (fixed inputs, no refresh) avg_launch_us', d['roofline']['avg_launch_us'], d['roofline']['frac'])"
done
