"""BASELINE geometries at FULL size against the CPU restatement (oracle/restate.py) directly — not kernel against
kernel: the persistent ping-pong decode kernel with more units than SMs (decode_kernel<half,1,2>, the instantiation
bench.py times), the cluster-split decode kernel on the Mistral (n = 8208, g = 4) and 70B (n = 8256, g = 8) layouts,
and the strided-chunk kernels on the configs[1] / [2] / [4] chunk shapes.  Teacher-forced: every step is compared on
identical state; victims must be identical except where the oracle's own decision margin is a near-tie of 16-bit
probabilities (the kernels and torch's CPU GEMM sum a logit's 128 products in different orders)."""
import pytest
import torch

from oracle import restate

pytestmark = pytest.mark.gpu


def _setup(B, H, Hkv, n, cap, dtype, C_init, seed):
    from easykv_b200.cache import BudgetedKVCache
    d = 128
    g = torch.Generator().manual_seed(seed)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype)
    cache = BudgetedKVCache(1, B, H, Hkv, d, cap, dtype=dtype, arith=0)
    K0, V0 = rnd(B, Hkv, n, d), rnd(B, Hkv, n, d)
    cache.load_prefill(0, K0.cuda(), V0.cuda(), n, C_init)
    orcs = []
    for b in range(B):
        o = restate.LayerOracle(Hkv, d, dtype)
        o.load_prefill(K0[b], V0[b], n, torch.tensor(C_init, dtype=torch.float32))
        orcs.append(o)
    return cache, orcs, rnd


def _run(cache, orcs, rnd, st, B, H, Hkv, q_len, steps, tol=1e-3):
    from easykv_b200.plan import StepParams
    sp = StepParams.from_fields(st)
    near_ties = 0
    for t in range(steps):
        q, k, v = rnd(B, H, q_len, 128) * 0.3, rnd(B, Hkv, q_len, 128), rnd(B, Hkv, q_len, 128)
        out, vl = cache.step(0, sp, q.cuda(), k.cuda(), v.cuda(), apply=False)
        out, vl = out.cpu(), torch.sort(vl.cpu().long(), dim=-1)[0]
        ref_v = []
        for b, o in enumerate(orcs):
            o_ref, v_ref = o.forward(st, q[b], k[b], v[b])
            assert (out[b].float() - o_ref.float()).abs().max().item() <= tol * max(1.0, o_ref.float().abs().max().item())
            v_ref = torch.sort(v_ref, dim=-1)[0]
            if not torch.equal(vl[b], v_ref):
                assert min(o.last_margin) < 1e-5, (t, b, o.last_margin)
                near_ties += 1
            ref_v.append(v_ref)
        cache.evict(0, torch.stack(ref_v))            # both sides continue from the oracle's choice
    for b, o in enumerate(orcs):
        Ke, Ve = cache.export(0)
        assert torch.equal(Ke[b].cpu(), o.K) and torch.equal(Ve[b].cpu(), o.V)
    return near_ties


@pytest.fixture
def dispatch(ekv_lib):
    yield ekv_lib
    ekv_lib.ekv_debug_set_dispatch(0, 0)
    ekv_lib.ekv_debug_set_chunk_variant(0)


def _decode_step(n):
    return restate.Step(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))


def _chunk_step(policy, n, stride):
    recent = int(n * 0.1)
    return restate.Step(policy=policy, accumulate=True, evict=stride, counter_add=float(stride), c_new_step=1.0,
                        k_feasible=max(n - recent - 4, stride), sink_protect=4, win_lo=4, win_recent=recent, range_start=4)


@pytest.mark.parametrize("variant", [2, 0], ids=["forced_pingpong", "automatic"])
def test_persistent_decode_kernel_7b_layout_vs_restate(dispatch, variant):
    """configs[1]: Llama-2-7B head layout, 1088 retained slots, roco, 6 sequences x 32 kv heads = 192 units > 148 SMs."""
    B, H, Hkv, n = 6, 32, 32, 1088
    dispatch.ekv_debug_set_dispatch(variant, -1 if variant else 0)
    cache, orcs, rnd = _setup(B, H, Hkv, n, n + 1, torch.float16, [float(n - i) for i in range(n)], seed=21)
    assert _run(cache, orcs, rnd, _decode_step(n), B, H, Hkv, 1, steps=10) <= 2


@pytest.mark.parametrize("variant", [5, 6], ids=["tcgen05", "round1_kernels"])
@pytest.mark.parametrize("B,H,Hkv,n", [(2, 32, 8, 8208), (2, 64, 8, 8256), (9, 64, 8, 1088)],
                         ids=["mistral_n8208_g4", "70b_n8256_g8", "70b_n1088_g8_b9"])
def test_gqa_decode_kernels_long_caches_vs_restate(dispatch, B, H, Hkv, n, variant):
    """configs[2] / [4] decode geometries through the tcgen05 GQA decode kernel and through the round-1 cluster-split
    kernels."""
    dispatch.ekv_debug_set_dispatch(variant, 0)
    cache, orcs, rnd = _setup(B, H, Hkv, n, n + 1, torch.float16, [float(n - i) for i in range(n)], seed=22)
    assert _run(cache, orcs, rnd, _decode_step(n), B, H, Hkv, 1, steps=4) <= 1


@pytest.mark.parametrize("chunk_variant", [0, 2], ids=["tcgen05", "mma_sync"])
@pytest.mark.parametrize("B,H,Hkv,n,stride,policy", [(1, 32, 8, 8208, 16, "h2o_head"), (2, 32, 32, 1088, 64, "roco"),
                                                     (1, 64, 8, 8256, 64, "roco")],
                         ids=["c3_mistral_stride16", "c2_7b_stride64", "c5_70b_stride64"])
def test_chunk_kernels_full_size_vs_restate(dispatch, B, H, Hkv, n, stride, policy, chunk_variant):
    dispatch.ekv_debug_set_chunk_variant(chunk_variant)
    cache, orcs, rnd = _setup(B, H, Hkv, n, n + stride, torch.float16, [0.0] * n, seed=23)
    assert _run(cache, orcs, rnd, _chunk_step(policy, n, stride), B, H, Hkv, stride, steps=2) <= 1
