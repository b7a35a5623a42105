"""Small decode cases through a chosen decode kernel (ekv_debug_set_dispatch variant / cluster) against the CPU
restatement, with detailed differences — development aid.  variant 5 = the tcgen05 GQA decode kernel.

    python tools/decode_debug.py [variant] [cluster]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import replay, restate  # noqa: E402
import engines  # noqa: E402
from easykv_b200 import _lib, build  # noqa: E402

build.build()
lib = _lib.load()
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 5
cluster = int(sys.argv[2]) if len(sys.argv) > 2 else 0
lib.ekv_debug_set_dispatch(variant, cluster)

CASES = [  # dtype, H, Hkv, policy, n0, steps, score_offset
    (torch.float16, 8, 2, "roco", 203, 6, 0),
    (torch.float16, 16, 2, "roco", 300, 6, 0),
    (torch.float16, 8, 4, "roco", 1500, 4, 0),
    (torch.float16, 8, 8, "roco", 203, 4, 0),
    (torch.bfloat16, 8, 1, "h2o_head", 203, 4, 0),
    (torch.float16, 8, 2, "tova", 640, 4, 0),
    (torch.float16, 4, 2, "recency", 130, 3, 0),
    (torch.float16, 8, 2, "full", 200, 3, 0),
    (torch.float16, 8, 2, "roco", 4500, 3, 0),
    (torch.float16, 8, 1, "roco", 8256, 2, 0),
    (torch.float16, 8, 2, "roco", 260, 5, 200),
]
for dtype, H, Hkv, policy, n0, steps, P in CASES:
    d = 128
    g = torch.Generator().manual_seed(7)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype)
    eng = engines.CudaEngine(1, H, Hkv, d, dtype, capacity=n0 + steps + 8)
    orc = replay.OracleEngine(1, H, Hkv, d, dtype)
    K, V = rnd(Hkv, n0, d), rnd(Hkv, n0, d)
    ns = n0 - P
    C0 = torch.arange(ns, 0, -1).float()
    for e in (eng, orc):
        e.load_prefill(0, K, V, ns, C0)
    recent = int(ns * 0.3)
    if policy == "full":
        st = restate.Step()
    else:
        st = restate.Step(policy=policy, accumulate=True, evict=1, counter_add=1.0, k_feasible=ns - recent, score_offset=P,
                          win_recent=recent if policy == "h2o_head" else 0, range_start=4)
    for t in range(steps):
        q, k, v = rnd(H, 1, d) * 0.3, rnd(Hkv, 1, d), rnd(Hkv, 1, d)
        o_ref, v_ref = orc.forward(0, st, q, k, v)
        o, vic = eng.forward(0, st, q, k, v, force=v_ref)
        torch.cuda.synchronize()
        err = (o.float() - o_ref.float()).abs()
        same = True if v_ref is None else torch.equal(vic, v_ref)
        nan = int(torch.isnan(o.float()).sum())
        print(f"{str(dtype)[6:]:9s} H{H} Hkv{Hkv} {policy:8s} n{n0} P{P} step{t}: out err max {err.max().item():.3e} (ref max {o_ref.float().abs().max().item():.2f}) "
              f"nan {nan} victims_equal {same}" + ("" if same else f" got {vic.flatten().tolist()} ref {v_ref.flatten().tolist()} margin {orc.margin(0)}"), flush=True)
        if err.max().item() > 5e-3 or nan:
            print("   per-head max err", [round(float(x), 4) for x in err.amax(dim=(1, 2))])
    Kc, Vc = eng.export(0)
    print("   cache equal:", torch.equal(Kc, orc.export(0)[0]) and torch.equal(Vc, orc.export(0)[1]), flush=True)
lib.ekv_debug_set_dispatch(0, 0)
