#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list, full captures of the headline decode kernel
# and of the chunk kernels, and the kernel sweeps.  Run as:  gpurun --timeout 2400 -- 'bash tools/gpu_round.sh [tag]'
# Everything lands in gpurun_out/<tag>_*; copy what should be judged into profiles/.
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee $OUT/${TAG}_smoke.txt
echo "== bench" ; timeout 600 python bench.py 2>&1 | tail -2 | tee $OUT/${TAG}_bench.json
echo "== bench 32 seqs" ; timeout 600 python bench.py --seqs-per-gpu 32 --no-cpu-baseline --no-gpu-reference --no-sweep 2>&1 | tail -1 | tee $OUT/${TAG}_bench_b32.json
echo "== bench reference arm" ; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee $OUT/${TAG}_bench_ref.json
echo "== ncu launch list (our kernels; the step is 32 launches of the decode kernel)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'decode_|general_kernel|select_kernel|export_kernel|evict_explicit|tova_|chunk_' -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 > $OUT/${TAG}_launches_cmd.log 2>&1
echo "== ncu full capture of the decode kernel"
[ "${EKV_SKIP_NCU_FULL:-0}" = 1 ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_kernel -s 40 -c 3 -f -o $OUT/${TAG}_decode \
  python bench.py --steps 2 --warmup 3 --no-sweep --no-cpu-baseline --no-gpu-reference --min-seconds 0.05 --layers 8 > $OUT/${TAG}_decode_cmd.log 2>&1
echo "== sweeps"
timeout 600 python tools/sweep.py decode 2>&1 | tee $OUT/${TAG}_sweep_decode.jsonl | tail -3
timeout 600 python tools/sweep.py cluster 2>&1 | tee $OUT/${TAG}_sweep_cluster.jsonl | tail -3
timeout 600 python tools/sweep.py chunk 2>&1 | tee $OUT/${TAG}_sweep_chunk.jsonl | tail -6
echo "== ncu: chunk kernels (Mistral stride 16, 8 sequences)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:chunk_ --csv --log-file $OUT/${TAG}_chunk_launches.csv python tools/chunk_profile.py 8 32 8 8208 16 h2o_head > /dev/null 2>&1
[ "${EKV_SKIP_NCU_FULL:-0}" = 1 ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:chunk_umma_kernel -s 4 -c 2 -f -o $OUT/${TAG}_chunk_umma python tools/chunk_profile.py 8 32 8 8208 16 h2o_head > /dev/null 2>&1
echo "== ncu: tcgen05 GQA decode kernel (70B layout, 32 sequences, n = 8256)"
[ "${EKV_SKIP_NCU_FULL:-0}" = 1 ] || timeout 900 ncu --set full --clock-control none --import-source on -k regex:decode_umma_kernel -s 4 -c 2 -f -o $OUT/${TAG}_decode_umma python tools/decode_profile.py 32 64 8 8256 0 roco > /dev/null 2>&1
echo "== end to end generate (7B shape, installed transformers model classes)"
timeout 600 python tools/e2e_generate.py --layers 32 --prompt 4096 --new 256 2>&1 | tail -1 | tee $OUT/${TAG}_e2e_llama7b.json
timeout 600 python tools/e2e_generate.py --arch mistral --layers 32 --prompt 16384 --new 16 --mode encoding --budget 0.5 --stride 16 --policy h2o --keep-attention 2>&1 | tail -1 | tee $OUT/${TAG}_e2e_mistral7b.json
echo "== timelines under the bench state"
timeout 200 python tools/timeline_bench.py c2 32 0 2>&1 | tail -3 | tee $OUT/${TAG}_timeline_bench.txt
for w in c5 c5_b32 c3_decode; do timeout 200 python tools/decode_umma_timeline_bench.py $w 2>&1 | tail -9; done | tee $OUT/${TAG}_decode_umma_timeline.txt
ls -la $OUT | tail -30
