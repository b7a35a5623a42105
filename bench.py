#!/usr/bin/env python
"""Headline benchmark: decode tokens/s at KV budget 1024 on the Llama-2-7B head layout, and the fused
evict+attention kernel's achieved HBM bandwidth against the measured roofline.

A *step* is one decode step of the hot path for every sequence on the GPU: L = 32 fused
`ekv_attend_evict` launches (one per layer), each streaming the retained cache of all sequences once,
updating the RoCo statistics, selecting one victim per (sequence, layer, kv head), evicting it in
place and appending the new token.  Workload = BASELINE.json configs[1] in the form that actually
evicts ("2b"): `mode='auto'`, budget 1024, stride 64 -> retained cache 1088, decode phase
of easykv.py:670-748 at 1089 keys per step.  Projections / MLP / sampling are outside the path
 and are not timed; q, k_new, v_new are synthetic N(0,1) fp16.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--seqs-per-gpu B] [--impl reference]

Under torchrun every rank runs its own shard of sequences on its own GPU (no collective on the data
path; one barrier + max-over-ranks of the device time).  `--impl reference` times the reference's
algorithm on the host CPU (the oracle port of the reference's PyTorch path; /root/reference itself
cannot travel to the GPU box) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

L_LAYERS, H, HKV, D = 32, 32, 32, 128          # Llama-2-7B attention geometry
BUDGET, STRIDE = 1024, 64
RETAINED = 1088                                # plan('auto', 4096, 1024, 64).idx  (tests/test_budget.py)
METRIC = "decode tokens/sec @ KV budget=1024, Llama-2-7B shape (evict+attn hot path)"


def bytes_alg_per_launch(B, n_keys):
    """SURVEY §8(d) / BASELINE.md §2 per (layer x sequence x step), times the B sequences of a launch."""
    e = 2
    per_seq = (2 * HKV * n_keys * D * e      # K and V read once
               + 2 * HKV * 1 * D * e         # new K, V rows written
               + 2 * H * 1 * D * e           # q read, out written
               + 6 * HKV * n_keys * 4        # roco: S, SQ, C read + write
               + HKV * 1 * 4)                # victim ids
    return B * per_seq


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        if not self.samples:          # a timed region shorter than one sampling period: one direct query
            try:
                self.samples.append(subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                                    "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                                   timeout=10).stdout.strip())
            except (OSError, subprocess.SubprocessError):
                pass
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(seconds_budget=12.0, layers=2, max_steps=40):
    """The reference's algorithm for this path on the host CPU: the oracle port of its PyTorch code
    (oracle/restate.py: QK^T, fp32 softmax, PV, roco accumulate, topk/argmin select, order-preserving
    K/V + state compaction), fp32, one sequence, `layers` layers of the 7B head layout at 1088+1 keys.
    tokens/s is extrapolated to the 32-layer stack (layers are independent on this path)."""
    from oracle import restate
    torch.set_num_threads(os.cpu_count() or 1)
    g = torch.Generator().manual_seed(0)
    n = RETAINED
    layers_ = []
    for _ in range(layers):
        lo = restate.LayerOracle(HKV, D, torch.float32)
        lo.load_prefill(torch.randn(HKV, n, D, generator=g), torch.randn(HKV, n, D, generator=g), n,
                        torch.arange(n, 0, -1).float())
        lo.S = torch.rand(HKV, n, generator=g) * lo.C / n
        lo.SQ = lo.S * lo.S / lo.C * 1.5
        layers_.append(lo)
    st = restate.Step(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=RETAINED - int(RETAINED * 0.3))
    q = torch.randn(H, 1, D, generator=g) * 0.3
    k, v = torch.randn(HKV, 1, D, generator=g), torch.randn(HKV, 1, D, generator=g)
    for lo in layers_:
        lo.forward(st, q, k, v)             # warm-up
    t0 = time.perf_counter()
    steps = 0
    while steps < max_steps and time.perf_counter() - t0 < seconds_budget:
        for lo in layers_:
            lo.forward(st, q, k, v)
        steps += 1
    dt = time.perf_counter() - t0
    per_layer_step = dt / (steps * layers)
    return {"value": 1.0 / (per_layer_step * L_LAYERS), "unit": "tokens/s", "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"1 sequence x {layers} layers x {steps} evicting decode steps, fp32, 7B head layout, "
                      f"{RETAINED}+1 keys, extrapolated to {L_LAYERS} layers ({per_layer_step*1e3:.2f} ms per layer-step)"}


def dist_env():
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    return rank, world, local


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    steps = max(args.steps, 1)
    r = cpu_reference(seconds_budget=min(60.0, 3.0 * steps), layers=2, max_steps=steps + args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": "tokens/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Llama-2-7B head layout (L=32,H=32,Hkv=32,d=128), mode=auto budget={BUDGET} "
                                   f"stride={STRIDE} -> {RETAINED}+1 keys per decode step, roco, 1 sequence (reference batch is 1), CPU"},
            "cpu_baseline": r,
            "e2e": {"value": r["value"], "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--seqs-per-gpu", type=int, default=64)
    ap.add_argument("--layers", type=int, default=L_LAYERS)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from easykv_b200.cache import BudgetedKVCache, SteadyDecode
    from easykv_b200.plan import StepParams, resolve_plan
    plan = resolve_plan("auto", 4096, BUDGET, STRIDE)
    assert plan.idx == RETAINED
    W = max(args.warmup, 3)
    B, L = args.seqs_per_gpu, args.layers
    n = RETAINED
    torch.manual_seed(1234 + rank)
    cache = BudgetedKVCache(L, B, H, HKV, D, n + 1, dtype=torch.float16, device=dev, arith=1)   # ATen's CUDA flavour
    cinit = [float(n - i) for i in range(n)]
    for l in range(L):
        cache.load_prefill(l, torch.randn(B, HKV, n, D, device=dev, dtype=torch.float16),
                           torch.randn(B, HKV, n, D, device=dev, dtype=torch.float16), n, cinit)
        cache.S[l][:, :, :n] = torch.rand(B, HKV, n, device=dev) * cache.Cn[l][:, :, :n] / n
        cache.SQ[l][:, :, :n] = cache.S[l][:, :, :n] ** 2 / cache.Cn[l][:, :, :n] * 1.5
    sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=plan.budget - int(plan.budget * 0.3))
    q = torch.randn(L, B, H, 1, D, device=dev, dtype=torch.float16) * 0.3
    kn = torch.randn(L, B, HKV, 1, D, device=dev, dtype=torch.float16)
    vn = torch.randn(L, B, HKV, 1, D, device=dev, dtype=torch.float16)
    steady = SteadyDecode(cache, sp, q, kn, vn)
    use_graph = not args.no_graph
    if use_graph:
        steady.capture()
    step = steady.replay if use_graph else steady.run

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---------------- device-resident timing ----------------------------------------------------------
    for _ in range(W):
        step()
    clocks = ClockSampler(local)
    barrier()
    if rank == 0:
        clocks.start()
    launches0 = cache.lib.ekv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clk = clocks.stop() if rank == 0 else None
    host_launches = cache.lib.ekv_launch_count() - launches0
    gpu_launches = args.steps * L           # kernel nodes executed (graph replays do not pass the host counter)

    # ---------------- end to end through the public API with host buffers --------------------------------
    hq = torch.randn(L, B, H, 1, D, dtype=torch.float16).pin_memory()
    hk = torch.randn(L, B, HKV, 1, D, dtype=torch.float16).pin_memory()
    hv = torch.randn(L, B, HKV, 1, D, dtype=torch.float16).pin_memory()
    hout = torch.empty(L, B, H, 1, D, dtype=torch.float16).pin_memory()
    hvic = torch.empty(L, B, HKV, 1, dtype=torch.int32).pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    # Per-layer pipeline: the H2D copy of layer l+1's q / k_new / v_new and the D2H read-back of layer l-1's
    # out / victim ids run on their own streams while layer l's kernel streams the cache; every byte still
    # crosses the bus inside the timed region, every step.
    copy_in, copy_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    ev_in = [torch.cuda.Event() for _ in range(L)]
    ev_done = [torch.cuda.Event() for _ in range(L)]

    def h2d(l):
        with torch.cuda.stream(copy_in):
            q[l].copy_(hq[l], non_blocking=True); kn[l].copy_(hk[l], non_blocking=True); vn[l].copy_(hv[l], non_blocking=True)
            ev_in[l].record(copy_in)

    def e2e_step():
        main = torch.cuda.current_stream()
        copy_in.wait_stream(main)
        h2d(0)
        for l in range(L):
            if l + 1 < L:
                h2d(l + 1)                             # issued one layer ahead of the launch that needs it
            main.wait_event(ev_in[l])
            steady.run_layer(l)
            ev_done[l].record(main)
            copy_out.wait_event(ev_done[l])
            with torch.cuda.stream(copy_out):
                hout[l].copy_(steady.out[l], non_blocking=True); hvic[l].copy_(steady.victim_lidx[l], non_blocking=True)
        main.wait_stream(copy_out)
        main.synchronize()                             # the caller consumes out / victim ids every step
    for _ in range(2):
        e2e_step()
    barrier()
    e0.record()
    for _ in range(e2e_steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1)

    from easykv_b200.shard import reduce_job
    ms, tokens = reduce_job(ms, B * args.steps, device=dev)            # max over ranks, sum over ranks
    ms_e2e, tokens_e2e = reduce_job(ms_e2e, B * e2e_steps, device=dev)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = tokens / (ms / 1e3)
    per_launch_s = ms / 1e3 / (args.steps * L)
    balg = bytes_alg_per_launch(B, n + 1)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    achieved = balg / per_launch_s / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_r01.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("seqs_per_gpu") == B:
            traffic = tj.get("dram_bytes_per_launch")
    h2d = (hq.numel() + hk.numel() + hv.numel()) * 2
    d2h = hout.numel() * 2 + hvic.numel() * 4
    line = {
        "metric": METRIC, "value": value, "unit": "tokens/s", "n_gpus": world, "steps": args.steps, "warmup": W,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16", "data": "synthetic",
        "config": {"workload": f"Llama-2-7B head layout (L={L},H={H},Hkv={HKV},d={D}), 4K prompt already reduced by mode=auto "
                               f"budget={BUDGET} stride={STRIDE} to {RETAINED} retained slots; decode phase: {RETAINED}+1 keys per step, "
                               f"roco, one eviction per (sequence, layer, kv head) per step",
                   "arithmetic": "f16 K/V/q/probabilities, f32 accumulate and policy state",
                   "seqs_per_gpu": B, "global_seqs": B * world, "cuda_graph": use_graph,
                   "l2": f"inputs larger than L2: {L} layers x {B} seqs x {2*HKV*(n+1)*D*2/1e6:.1f} MB of K/V = "
                         f"{L*B*2*HKV*(n+1)*D*2/1e9:.1f} GB streamed per step, distinct buffers per layer",
                   "parallelism": f"sequences sharded over {world} GPU(s), no collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src,
                     "bytes_alg_per_launch": balg, "avg_launch_us": per_launch_s * 1e6,
                     "kernel": "ekv::decode_kernel<__half,1>"},
        "e2e": {"value": tokens_e2e / (ms_e2e / 1e3), "unit": "tokens/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                "note": "q,k_new,v_new from pinned host memory, out + victim ids back to host, every step, per-layer copies "
                        "pipelined on side streams around the layer launches; cache resident"},
        "gpu_launches": gpu_launches, "host_launch_calls_in_timed_region": host_launches,
        "clocks": clk,
    }
    if not args.no_cpu_baseline and world >= 1:
        line["cpu_baseline"] = cpu_reference()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
