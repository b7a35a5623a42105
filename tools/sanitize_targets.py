"""Small invocations of every kernel family, sized for compute-sanitizer (racecheck / synccheck / memcheck are
10-100x slower than native): the persistent ping-pong decode kernel with more units than SMs (both consumer groups
and the shared mbarrier ring), its steady-state form where victim_slots aliases new_slots (SteadyDecode), the
cluster-split decode kernel (FMA and tensor-core variants), the tcgen05 cluster chunk kernel, the mma.sync chunk
kernels, the general kernel, select / explicit evict / export / rope.  Every result is checked against the CPU
restatement, and the script prints how many library launches ran so that an empty run cannot read as a clean one.

    compute-sanitizer --tool racecheck python tools/sanitize_targets.py [family ...]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import restate  # noqa: E402
from easykv_b200 import _lib, build  # noqa: E402
from easykv_b200.cache import BudgetedKVCache, SteadyDecode  # noqa: E402
from easykv_b200.plan import StepParams  # noqa: E402

build.build()
lib = _lib.load()
dev = "cuda"
want = set(sys.argv[1:])


def run(name):
    return not want or name in want


def oracle_step(orcs, st, q, k, v):
    outs, vics = [], []
    for b, o in enumerate(orcs):
        ob, vb = o.forward(st, q[b], k[b], v[b])
        outs.append(ob)
        vics.append(None if vb is None else torch.sort(vb, dim=-1)[0])
    return torch.stack(outs), (None if vics[0] is None else torch.stack(vics))


def decode_case(label, B, H, Hkv, n, steps, variant, cluster, dtype=torch.float16, steady=False, d=128, kernel=0, hard_select=False):
    g = torch.Generator().manual_seed(3)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype)
    lib.ekv_debug_set_dispatch(variant, cluster)
    cache = BudgetedKVCache(1, B, H, Hkv, d, n + 1, dtype=dtype, arith=0)
    K0, V0 = rnd(B, Hkv, n, d), rnd(B, Hkv, n, d)
    C0 = [float(n - i) for i in range(n)]
    cache.load_prefill(0, K0.cuda(), V0.cuda(), n, C0)
    orcs = []
    for b in range(B):
        o = restate.LayerOracle(Hkv, d, dtype)
        o.load_prefill(K0[b], V0[b], n, torch.tensor(C0))
        orcs.append(o)
    if hard_select:
        # the lowest-mean slots carry the LARGEST standard deviations: the candidate attempt is rejected and the victim comes
        # from the bucket select (histograms, classify pass, listed entries ranked exactly) — the path racecheck should see
        gs = torch.Generator().manual_seed(9)
        Cc = torch.tensor(C0)
        S0 = torch.rand(B, Hkv, n, generator=gs) * 0.5 + 0.5
        S0[:, :, ::3] *= 0.01                                       # a third of the slots: tiny mean ...
        SQ0 = S0 * S0 / Cc * (1.0 + torch.rand(B, Hkv, n, generator=gs) * 0.2)
        SQ0[:, :, ::3] = 4.0 + torch.rand(B, Hkv, (n + 2) // 3, generator=gs)    # ... and a huge second moment
        cache.S[0][:, :, :n] = S0.cuda(); cache.SQ[0][:, :, :n] = SQ0.cuda()
        for b, o in enumerate(orcs):
            o.S, o.SQ = S0[b].clone(), SQ0[b].clone()
    st = restate.Step(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3) if not hard_select else n // 2)
    sp = StepParams.from_fields(st)
    sd = None
    qb = torch.zeros(1, B, H, 1, d, dtype=dtype, device=dev)
    kb = torch.zeros(1, B, Hkv, 1, d, dtype=dtype, device=dev)
    vb = torch.zeros(1, B, Hkv, 1, d, dtype=dtype, device=dev)
    for t in range(steps):
        q, k, v = rnd(B, H, 1, d) * 0.3, rnd(B, Hkv, 1, d), rnd(B, Hkv, 1, d)
        o_ref, v_ref = oracle_step(orcs, st, q, k, v)
        check_victims = True
        if steady:
            qb[0].copy_(q.cuda()); kb[0].copy_(k.cuda()); vb[0].copy_(v.cuda())
            if sd is None:
                sd = SteadyDecode(cache, sp, qb, kb, vb)       # runs the first step itself (brings the layer to steady state)
                out, vl, check_victims = sd.out[0].clone(), None, False
            else:
                sd.run()                                       # victim_slots aliases new_slots from here on
                out, vl = sd.out[0].clone(), sd.victim_lidx[0].clone()
        else:
            out, vl = cache.step(0, sp, q.cuda(), k.cuda(), v.cuda(), kernel=kernel)
        torch.cuda.synchronize()
        assert (out.cpu().float() - o_ref.float()).abs().max().item() <= 1e-3, label
        if check_victims:
            assert torch.equal(vl.cpu().long(), v_ref), (label, t)
    lib.ekv_debug_set_dispatch(0, 0)
    print(f"ok {label}", flush=True)


def chunk_case(label, B, H, Hkv, n, stride, steps, variant, dtype=torch.float16, policy="roco", d=128, kernel=0):
    g = torch.Generator().manual_seed(4)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype)
    lib.ekv_debug_set_chunk_variant(variant)
    cache = BudgetedKVCache(1, B, H, Hkv, d, n + stride, dtype=dtype, arith=0)
    K0, V0 = rnd(B, Hkv, n, d), rnd(B, Hkv, n, d)
    cache.load_prefill(0, K0.cuda(), V0.cuda(), n, [0.0] * n)
    orcs = []
    for b in range(B):
        o = restate.LayerOracle(Hkv, d, dtype)
        o.load_prefill(K0[b], V0[b], n, torch.zeros(n))
        orcs.append(o)
    recent = int(n * 0.1)
    st = restate.Step(policy=policy, accumulate=True, evict=stride, counter_add=float(stride), c_new_step=1.0,
                      k_feasible=max(n - recent - 4, stride), sink_protect=4, win_lo=4, win_recent=recent, range_start=4)
    for t in range(steps):
        q, k, v = rnd(B, H, stride, d) * 0.3, rnd(B, Hkv, stride, d), rnd(B, Hkv, stride, d)
        o_ref, v_ref = oracle_step(orcs, st, q, k, v)
        out, vl = cache.step(0, StepParams.from_fields(st), q.cuda(), k.cuda(), v.cuda(), kernel=kernel)
        torch.cuda.synchronize()
        assert (out.cpu().float() - o_ref.float()).abs().max().item() <= 1e-3, label
        assert torch.equal(vl.cpu().long(), v_ref), (label, t)
    lib.ekv_debug_set_chunk_variant(0)
    print(f"ok {label}", flush=True)


n0 = lib.ekv_launch_count()
if run("persistent"):
    # 5 sequences x 32 kv heads = 160 units > 148 SMs: decode_kernel<half,1,2>, both consumer groups, ring shared
    decode_case("persistent ping-pong decode_kernel<half,1,2>, 160 units", 5, 32, 32, 96, 3, variant=2, cluster=-1)
    decode_case("persistent decode_kernel<half,1,1> (one group per CTA), 160 units", 5, 32, 32, 96, 2, variant=1, cluster=-1)
if run("steady"):
    decode_case("steady decode (victim_slots == new_slots), ping-pong", 5, 32, 32, 96, 4, variant=2, cluster=-1, steady=True)
if run("cluster"):
    decode_case("cluster decode C=2 (FMA)", 1, 8, 8, 200, 3, variant=0, cluster=2)
    decode_case("cluster decode C=4, g=4 (FMA)", 1, 8, 2, 300, 3, variant=3, cluster=4)
    decode_case("cluster decode C=2, g=8 (tensor-core variant)", 1, 16, 2, 300, 3, variant=4, cluster=2)
if run("bucket"):
    # the bucket select itself (ekv_bucket.cuh) in both kernels that use it, on a state built to reject the candidate walk
    decode_case("cluster decode C=2 (FMA), bucket select", 1, 8, 8, 400, 2, variant=0, cluster=2, hard_select=True)
    decode_case("cluster decode C=4, g=4 (FMA), bucket select", 1, 8, 2, 600, 2, variant=3, cluster=4, hard_select=True)
if run("bucket_umma"):
    decode_case("tcgen05 GQA decode, g=4, CTA pairs, bucket select", 1, 8, 2, 700, 2, variant=5, cluster=2, hard_select=True)
if run("decode_umma"):
    decode_case("tcgen05 GQA decode, g=4, one CTA per unit", 2, 8, 2, 300, 3, variant=5, cluster=1)
    decode_case("tcgen05 GQA decode, g=8, CTA pairs", 1, 16, 2, 700, 3, variant=5, cluster=2)
    decode_case("tcgen05 GQA decode, g=2, clusters of 4", 1, 4, 2, 1100, 2, variant=5, cluster=4)
if run("chunk_umma"):
    chunk_case("tcgen05 chunk, 1 CTA per unit", 1, 4, 4, 200, 16, 2, variant=0)
    chunk_case("tcgen05 chunk, cluster of CTAs per unit, g=4", 1, 8, 2, 1500, 16, 2, variant=0)
    chunk_case("tcgen05 chunk, two row blocks, g=2", 1, 4, 2, 300, 64, 2, variant=0)
if run("chunk_tc"):
    chunk_case("mma.sync chunk", 1, 8, 2, 300, 16, 2, variant=2)
if run("general"):
    # the exact kernel, a template over head_dim: 64 and 96 arrive here at the automatic dispatch, 128 with kernel = 1
    decode_case("general kernel decode, head_dim 64, g=4", 2, 8, 2, 203, 3, variant=0, cluster=0, d=64)
    decode_case("general kernel decode, head_dim 96", 1, 4, 4, 150, 2, variant=0, cluster=0, d=96)
    chunk_case("general kernel chunk, head_dim 64, g=2", 1, 4, 2, 260, 16, 2, variant=0, d=64)
    chunk_case("general kernel chunk, head_dim 96, stride 24", 1, 4, 4, 200, 24, 2, variant=0, d=96, policy="h2o_head")
    decode_case("general kernel decode, head_dim 128 (kernel = 1)", 1, 8, 2, 203, 2, variant=0, cluster=0, kernel=1)
print(f"library launches in this run: {lib.ekv_launch_count() - n0}")
assert lib.ekv_launch_count() - n0 > 0
