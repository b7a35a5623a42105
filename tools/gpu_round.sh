#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list and one full capture of the decode
# kernel.  Run as:  gpurun --timeout 1500 -- 'bash tools/gpu_round.sh [tag]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench" ; timeout 600 python bench.py 2>&1 | tail -2 | tee $OUT/${TAG}_bench.json
echo "== bench 32 seqs" ; timeout 600 python bench.py --seqs-per-gpu 32 --no-cpu-baseline 2>&1 | tail -1 | tee $OUT/${TAG}_bench_b32.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref.json
echo "== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_kernel|general_kernel|select_kernel|export_kernel|evict_explicit|tova_' -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > $OUT/${TAG}_launches_cmd.log 2>&1
echo "== ncu full capture of the decode kernel"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 40 -c 3 -f -o $OUT/${TAG}_decode \
  python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --layers 8 > $OUT/${TAG}_decode_cmd.log 2>&1
ls -la $OUT
