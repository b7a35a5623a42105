"""ABI v8 on the GPU: (a) the streaming variant fused into the decode kernels (io.rope_cos / rope_sin / k_new_raw: cached
rows rotated while they are read — llama_forward_stream, llama_patch.py:310-327) against the two-pass path
(ekv_rope_cache into a second buffer, then the unchanged step) that the reference-recorded streaming goldens pin; (b) ragged
batches (io.seq_n_before + step.budget_gate) against the same sequences run one by one."""
import dataclasses

import pytest
import torch

from easykv_b200.cache import BudgetedKVCache, RaggedDecode
from easykv_b200.plan import StepParams

D = 128


def _tables(rows, dtype, dev):
    inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
    f = torch.outer(torch.arange(rows, dtype=torch.float32), inv)
    emb = torch.cat([f, f], dim=-1)
    return emb.cos().to(dtype).to(dev), emb.sin().to(dtype).to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("B,H,Hkv,n,dtype,dispatch", [
    (2, 4, 4, 200, torch.float16, (0, 0)),        # MHA, few units: cluster kernel
    (10, 8, 8, 300, torch.float16, (0, -1)),      # MHA, single-CTA kernels only: the persistent kernel
    (40, 8, 8, 520, torch.bfloat16, (0, 0)),      # MHA, 320 units: the persistent ping-pong kernel, bf16
    (2, 8, 2, 260, torch.float16, (0, 0)),        # GQA g = 4: cluster kernel, FMA path
    (1, 16, 2, 333, torch.bfloat16, (3, 0)),      # GQA g = 8: the cluster kernel's FMA path on both sides (the tensor-core path is not fused)
    (2, 8, 4, 2304, torch.float16, (0, 0)),       # g = 2, long cache: the tcgen05 kernel declines, cluster kernel takes it
])
def test_fused_streaming_equals_two_pass(ekv_lib, B, H, Hkv, n, dtype, dispatch):
    dev = "cuda"
    torch.manual_seed(11)
    steps = 12
    cos, sin = _tables(n + steps + 4, dtype, dev)
    caches = [BudgetedKVCache(1, B, H, Hkv, D, n + 8, dtype=dtype, arith=1) for _ in range(2)]
    K0 = torch.randn(B, Hkv, n, D, device=dev).to(dtype)
    V0 = torch.randn(B, Hkv, n, D, device=dev).to(dtype)
    for c, fused in zip(caches, (True, False)):
        c.enable_streaming()
        c.fused_streaming = fused
        c.load_prefill(0, K0, V0, n, [float(n - i) for i in range(n)])
    sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))
    ekv_lib.ekv_debug_set_dispatch(*dispatch)
    try:
        launches = ekv_lib.ekv_launch_count()
        for t in range(steps):
            q = torch.randn(B, 1, H * D, device=dev).to(dtype) * 0.5
            k = torch.randn(B, 1, Hkv * D, device=dev).to(dtype)
            v = torch.randn(B, 1, Hkv * D, device=dev).to(dtype)
            if t == 0:
                launches = ekv_lib.ekv_launch_count()
            o0, v0 = caches[0].step_stream(0, sp, q, k, v, cos, sin)
            if t == 0:
                fused_launches = ekv_lib.ekv_launch_count() - launches
            o1, v1 = caches[1].step_stream(0, sp, q, k, v, cos, sin)
            assert torch.equal(v0, v1), f"step {t}: victims differ"
            assert torch.equal(o0, o1), f"step {t}: outputs differ by {(o0.float() - o1.float()).abs().max().item()}"
        assert fused_launches == 2, "fused step = ekv_rope_qk + one attention launch (no ekv_rope_cache pass)"
    finally:
        ekv_lib.ekv_debug_set_dispatch(0, 0)
    for a, b in zip(caches[0].export(0, with_state=True), caches[1].export(0, with_state=True)):
        assert torch.equal(a, b)


@pytest.mark.gpu
@pytest.mark.parametrize("H,Hkv,dtype,dispatch,kernel", [
    (4, 4, torch.float16, (0, 1), 0),             # cluster kernel (one CTA per unit on both sides: identical summation order)
    (4, 4, torch.float32, (0, -1), 0),            # persistent kernel, fp32
    (8, 2, torch.float16, (5, 1), 0),             # tcgen05 GQA kernel (forced)
    (8, 2, torch.bfloat16, (0, 1), 0),            # GQA: cluster kernel
    (4, 2, torch.float16, (0, 0), 1),             # the exact general kernel
])
def test_ragged_batch_equals_sequences_one_by_one(ekv_lib, H, Hkv, dtype, dispatch, kernel):
    """Three sequences of 70 / 96 / 83 slots, budget gate 90: the first grows for the whole run, the second evicts from
    the first step, the third starts evicting on the way.  Victims identical, outputs within rounding of the kernels'
    different dispatch, final caches identical to each sequence run alone."""
    dev = "cuda"
    torch.manual_seed(5)
    lengths, gate, steps = [70, 96, 83], 90, 14
    B = len(lengths)
    cap = 128
    Ks = [torch.randn(Hkv, n, D, device=dev).to(dtype) for n in lengths]
    Vs = [torch.randn(Hkv, n, D, device=dev).to(dtype) for n in lengths]
    Cs = [[float(n - i) for i in range(n)] for n in lengths]
    qs = [torch.randn(B, H, 1, D, device=dev).to(dtype) * 0.4 for _ in range(steps)]
    ks = [torch.randn(B, Hkv, 1, D, device=dev).to(dtype) for _ in range(steps)]
    vs = [torch.randn(B, Hkv, 1, D, device=dev).to(dtype) for _ in range(steps)]
    sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=gate - int(gate * 0.3))
    ekv_lib.ekv_debug_set_dispatch(*dispatch)
    try:
        batch = BudgetedKVCache(1, B, H, Hkv, D, cap, dtype=dtype, arith=1)
        rag = RaggedDecode(batch, lengths, gate)
        rag.load_prefill(0, Ks, Vs, Cs)
        outs, vics = [], []
        for t in range(steps):
            o, vl = rag.step(0, sp, qs[t], ks[t], vs[t], kernel=kernel)
            outs.append(o); vics.append(vl)
        tol = 2e-6 if dtype == torch.float32 else (1e-3 if dtype == torch.float16 else 8e-3)
        for b, n0 in enumerate(lengths):
            one = BudgetedKVCache(1, 1, H, Hkv, D, cap, dtype=dtype, arith=1)
            one.load_prefill(0, Ks[b], Vs[b], n0, Cs[b])
            n = n0
            for t in range(steps):
                evict = n + 1 > gate
                o, vl = one.step(0, sp if evict else dataclasses.replace(sp, evict=0), qs[t][b:b + 1], ks[t][b:b + 1], vs[t][b:b + 1],
                                 kernel=kernel)
                if evict:
                    assert torch.equal(vl[0], vics[t][b]), f"sequence {b} step {t}: victims {vl[0].flatten().tolist()} vs {vics[t][b].flatten().tolist()}"
                else:
                    assert bool((vics[t][b] == -1).all())
                    n += 1
                assert (o.float() - outs[t][b:b + 1].float()).abs().max().item() <= tol
            assert rag.n[0][b] == n
            Ko, Vo, S, SQ, Cn = one.export(0, with_state=True)
            # the batch's sequence b in logical order
            lidx = batch.lidx[0][b]                                        # [Hkv, cap]
            for h in range(Hkv):
                valid = (lidx[h] >= 0).nonzero().flatten()
                order = valid[torch.argsort(lidx[h][valid])]
                assert order.numel() == n
                assert torch.equal(batch.K[0][b, h, order], Ko[0, h]) and torch.equal(batch.V[0][b, h, order], Vo[0, h])
                if dtype == torch.float32:          # fp32 probabilities: the longer batch launch sums in another order
                    assert torch.allclose(batch.S[0][b, h, order], S[0, h], rtol=1e-5, atol=1e-7)
                    assert torch.allclose(batch.SQ[0][b, h, order], SQ[0, h], rtol=1e-5, atol=1e-9)
                else:
                    assert torch.equal(batch.S[0][b, h, order], S[0, h]) and torch.equal(batch.SQ[0][b, h, order], SQ[0, h])
                assert torch.equal(batch.Cn[0][b, h, order], Cn[0, h])
    finally:
        ekv_lib.ekv_debug_set_dispatch(0, 0)
