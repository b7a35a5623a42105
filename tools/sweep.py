"""Kernel-level sweep over the BASELINE configs' geometries (development aid; bench.py is the contract).
Prints one JSON line per case: per-launch time, algorithmic GB/s (SURVEY §8d formula) and fraction of the
measured HBM roofline.

    python tools/sweep.py [decode|chunk|all]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache  # noqa: E402
from easykv_b200.plan import StepParams  # noqa: E402

PEAK = 6538.0
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    PEAK = float(json.load(open(pk))["hbm_gbs"])
A_POL = {"roco": 6, "h2o_head": 2, "tova": 1, "recency": 0, "full": 0}


def bytes_alg(B, H, Hkv, d, n, q, policy, evict, e=2):
    return B * (2 * Hkv * n * d * e + 2 * Hkv * q * d * e + 2 * H * q * d * e + A_POL[policy] * Hkv * n * 4 + Hkv * evict * 4)


def steady_state(cache, l, n, dev):
    B, Hkv = cache.B, cache.Hkv
    cache.S[l][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[l][:, :, :n] / n
    cache.SQ[l][:, :, :n] = cache.S[l][:, :, :n] ** 2 / cache.Cn[l][:, :, :n] * 1.5


def run_case(name, B, H, Hkv, n, q_len, policy, L=4, steps=10, dtype=torch.float16, kernel=0, cluster=0, variant=0, literal_prompt=0):
    d, dev = 128, "cuda"
    torch.manual_seed(0)
    grow = policy == "full" or literal_prompt
    cache = BudgetedKVCache(L, B, H, Hkv, d, n + q_len * (3 + steps + 1 if grow else 1), dtype=dtype, arith=1)
    cache.lib.ekv_debug_set_dispatch(variant, cluster)
    for l in range(L):
        cache.load_prefill(l, torch.randn(B, Hkv, n, d, device=dev, dtype=dtype), torch.randn(B, Hkv, n, d, device=dev, dtype=dtype),
                           n, [float(n - i) for i in range(n)])
        steady_state(cache, l, n, dev)
    budget = n
    if q_len == 1:
        recent = int(budget * 0.3)
        sp = StepParams(policy=policy, accumulate=policy in ("roco", "h2o_head", "tova"), evict=0 if policy == "full" else 1,
                        counter_add=1.0, k_feasible=budget - recent, win_recent=recent if policy == "h2o_head" else 0, range_start=4)
    else:
        recent = int(budget * 0.1)
        sp = StepParams(policy=policy, accumulate=policy in ("roco", "h2o_head", "tova"), evict=0 if policy == "full" else q_len,
                        counter_add=float(q_len), c_new_step=1.0, k_feasible=max(budget - recent - 4, q_len), sink_protect=4,
                        win_lo=4, win_recent=recent, range_start=4)
    if literal_prompt:
        # BASELINE configs[1] exactly as written: mode='decoding', the prompt's slots carry no state and are never
        # evicted, the generated ones (< budget) are only scored (easykv.py:294-303): attention over n keys,
        # roco accumulate over the n - prompt generated slots, no eviction
        sp = StepParams(policy=policy, accumulate=True, evict=0, score_offset=literal_prompt, c_new0=1.0, k_feasible=700)
    q = torch.randn(L, B, H, q_len, d, device=dev, dtype=dtype) * 0.3
    kn = torch.randn(L, B, Hkv, q_len, d, device=dev, dtype=dtype)
    vn = torch.randn(L, B, Hkv, q_len, d, device=dev, dtype=dtype)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if literal_prompt:
        cache2 = cache                              # append mode: the cache grows by one slot per step
        for _ in range(2):
            for l in range(L):
                cache2.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            for l in range(L):
                cache2.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
        e1.record()
    elif q_len == 1 and sp.evict == 1 and kernel == 0:
        # steady-state decode: one CUDA graph per step (L launches), no host work in the timed region
        from easykv_b200.cache import SteadyDecode
        sd = SteadyDecode(cache, sp, q, kn, vn).capture()
        for _ in range(3):
            sd.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            sd.replay()
        e1.record()
    else:
        for _ in range(3):
            for l in range(L):
                cache.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            for l in range(L):
                cache.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
        e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (steps * L)
    ba = bytes_alg(B, H, Hkv, d, n + q_len, q_len, policy, sp.evict, e=torch.empty(0, dtype=dtype).element_size())
    gbs = ba / us / 1e3
    print(json.dumps(dict(case=name, B=B, H=H, Hkv=Hkv, n=n, q_len=q_len, policy=policy, dtype=str(dtype).split(".")[1], kernel=kernel, cluster=cluster, variant=variant,
                          us_per_launch=round(us, 1), bytes_alg=ba, GBps=round(gbs, 1), frac_of_measured=round(gbs / PEAK, 3))), flush=True)
    cache.lib.ekv_debug_set_dispatch(0, 0)
    del cache
    torch.cuda.empty_cache()


def run_stream_case(name, B, H, Hkv, n, L=4, steps=10, dtype=torch.float16, fused="auto"):
    """Decode step of the streaming variant (generation_config['streaming']): ekv_rope_cache (re-rotate the whole
    cache at cache-relative positions) + ekv_rope_qk + ekv_attend_evict + the raw-row scatter, per layer.
    Algorithmic bytes = the non-streaming step's + one extra read and one write of K."""
    d, dev = 128, "cuda"
    torch.manual_seed(0)
    cache = BudgetedKVCache(L, B, H, Hkv, d, n + 1, dtype=dtype, arith=1)
    cache.enable_streaming()
    cache.fused_streaming = fused
    for l in range(L):
        cache.load_prefill(l, torch.randn(B, Hkv, n, d, device=dev, dtype=dtype), torch.randn(B, Hkv, n, d, device=dev, dtype=dtype),
                           n, [float(n - i) for i in range(n)])
        steady_state(cache, l, n, dev)
    recent = int(n * 0.3)
    sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - recent)
    inv = 1.0 / (10000.0 ** (torch.arange(0, d, 2, device=dev).float() / d))
    emb = torch.cat([torch.outer(torch.arange(n + 8, device=dev).float(), inv)] * 2, dim=-1)
    cos, sin = emb.cos().to(dtype), emb.sin().to(dtype)
    q = torch.randn(L, B, 1, H * d, device=dev, dtype=dtype) * 0.3
    kn = torch.randn(L, B, 1, Hkv * d, device=dev, dtype=dtype)
    vn = torch.randn(L, B, 1, Hkv * d, device=dev, dtype=dtype)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        for l in range(L):
            cache.step_stream(l, sp, q[l], kn[l], vn[l], cos, sin)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        for l in range(L):
            cache.step_stream(l, sp, q[l], kn[l], vn[l], cos, sin)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (steps * L)
    ba = bytes_alg(B, H, Hkv, d, n + 1, 1, "roco", 1) + 2 * B * Hkv * n * d * 2
    gbs = ba / us / 1e3
    print(json.dumps(dict(case=name, B=B, H=H, Hkv=Hkv, n=n, q_len=1, policy="roco", dtype=str(dtype).split(".")[1], kernel=0, cluster=0,
                          variant=0, streaming=True, us_per_launch=round(us, 1), bytes_alg=ba, GBps=round(gbs, 1),
                          frac_of_measured=round(gbs / PEAK, 3))), flush=True)
    del cache
    torch.cuda.empty_cache()


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what in ("decode", "all"):
        # (name, B, H, Hkv, n) — decode step, roco
        for name, B, H, Hkv, n in [("7B b1", 1, 32, 32, 1088), ("7B b8", 8, 32, 32, 1088), ("7B b32", 32, 32, 32, 1088),
                                   ("7B b64", 64, 32, 32, 1088), ("13B n2112 b32", 32, 40, 40, 2112),
                                   ("mistral n8208 b16", 16, 32, 8, 8208), ("70B n8256 b8", 8, 64, 8, 8256),
                                   ("70B n8256 b32", 32, 64, 8, 8256), ("7B literal n4352 b16", 16, 32, 32, 4352)]:
            run_case(name, B, H, Hkv, n, 1, "roco")
        for pol in ("h2o_head", "tova", "recency", "full"):
            run_case("13B sweep " + pol, 32, 40, 40, 2112, 1, pol)
        run_case("7B b64 bf16", 64, 32, 32, 1088, 1, "roco", dtype=torch.bfloat16)
        # the generation phase of BASELINE configs[2] / [4] ('encoding' mode: attention over the retained cache, no policy)
        run_case("C3 generation: mistral n8208 b16, no policy", 16, 32, 8, 8208, 1, "full")
        run_case("C5 generation: 70B n8256 b8, no policy", 8, 64, 8, 8256, 1, "full")
        run_case("C2 literal: decoding, 4096 prompt + 200 generated, no eviction", 16, 32, 32, 4296, 1, "roco", literal_prompt=4096)
        run_stream_case("7B b64 streaming variant (two passes: ekv_rope_cache + the step; 4 launches per layer-step)", 64, 32, 32, 1088)
        run_stream_case("7B b1 streaming variant, two passes", 1, 32, 32, 1088, L=32, fused=False)
        run_stream_case("7B b1 streaming variant, rotation fused into the decode kernel", 1, 32, 32, 1088, L=32, fused=True)
        run_stream_case("mistral n8208 b1 streaming variant, two passes", 1, 32, 8, 8208, L=32, fused=False)
        run_stream_case("mistral n8208 b1 streaming variant, rotation fused into the decode kernel", 1, 32, 8, 8208, L=32, fused=True)
        run_case("7B b32 general-kernel", 32, 32, 32, 1088, 1, "roco", kernel=1, steps=3)
    if what in ("cluster",):
        for B in (1, 2, 4, 8):
            for c in (-1, 0, 2, 4, 8):
                run_case(f"7B b{B}", B, 32, 32, 1088, 1, "roco", cluster=c)
        for B in (1, 4, 16):
            for c, v in ((-1, 0), (0, 0), (4, 0), (8, 0), (8, 3)):
                run_case(f"mistral n8208 b{B}", B, 32, 8, 8208, 1, "roco", cluster=c, variant=v)
        for B in (1, 8, 32):
            for c, v in ((0, 0), (4, 0), (8, 0), (8, 3)):
                run_case(f"70B n8256 b{B}", B, 64, 8, 8256, 1, "roco", cluster=c, variant=v)
        for c in (-1, 0):
            run_case("mistral n1088 b32", 32, 32, 8, 1088, 1, "roco", cluster=c)
            run_case("70B n1088 b32", 32, 64, 8, 1088, 1, "roco", cluster=c)
        for c in (-1, 2):
            run_case("7B literal n4352 b16", 16, 32, 32, 4352, 1, "roco", cluster=c)
    if what in ("chunk", "all"):
        for name, B, H, Hkv, n, q in [("C3 mistral stride16", 1, 32, 8, 8208, 16), ("C3 mistral stride16 b8", 8, 32, 8, 8208, 16),
                                      ("C2 7B stride64", 1, 32, 32, 1088, 64), ("C2 7B stride64 b8", 8, 32, 32, 1088, 64),
                                      ("C5 70B stride64", 1, 64, 8, 8256, 64)]:
            pol = "h2o_head" if "C3" in name else "roco"
            run_case(name, B, H, Hkv, n, q, pol, L=2, steps=3)


if __name__ == "__main__":
    main()
