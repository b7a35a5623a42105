"""Small strided-chunk cases through the tcgen05 cluster kernel (chunk variant 0) and the mma.sync two-pass kernels
(variant 2) against the CPU restatement, with detailed differences — development aid for the chunk kernels.

    python tools/chunk_debug.py [variant]
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
variant = int(sys.argv[1]) if len(sys.argv) > 1 else 0
lib.ekv_debug_set_chunk_variant(variant)

CASES = [  # dtype, H, Hkv, stride, policy, n0, steps
    (torch.float16, 4, 4, 16, "roco", 100, 3),
    (torch.float16, 8, 2, 16, "roco", 300, 3),
    (torch.float16, 8, 1, 8, "h2o_head", 200, 3),
    (torch.bfloat16, 4, 4, 64, "roco", 1100, 3),
    (torch.float16, 16, 2, 24, "tova", 260, 3),
    (torch.float16, 4, 2, 96, "roco", 700, 3),
    (torch.float16, 4, 4, 7, "recency", 150, 3),
    (torch.float16, 32, 8, 16, "h2o_head", 3000, 2),
    (torch.float16, 8, 8, 64, "roco", 1088, 2),
]
for dtype, H, Hkv, stride, policy, n0, steps in CASES:
    d = 128
    g = torch.Generator().manual_seed(11)
    rnd = lambda *s: torch.randn(*s, generator=g).to(dtype)
    eng = engines.CudaEngine(1, H, Hkv, d, dtype, kernel=0, capacity=n0 + stride)
    orc = replay.OracleEngine(1, H, Hkv, d, dtype)
    K, V = rnd(Hkv, n0, d), rnd(Hkv, n0, d)
    for e in (eng, orc):
        e.load_prefill(0, K, V, n0, torch.zeros(n0))
    recent, sink = int(n0 * 0.1), 4
    st = restate.Step(policy=policy, accumulate=True, evict=stride, counter_add=float(stride), c_new_step=1.0,
                      k_feasible=max(n0 - recent - sink, stride), sink_protect=sink, win_lo=sink, win_recent=recent,
                      range_start=sink)
    for t in range(steps):
        q, k, v = rnd(H, stride, d) * 0.3, rnd(Hkv, stride, d), rnd(Hkv, stride, d)
        S_before = orc.layers[0].S.clone()
        o_ref, v_ref = orc.forward(0, st, q, k, v)
        o, vic = eng.forward(0, st, q, k, v, force=v_ref)
        torch.cuda.synchronize()
        err = (o.float() - o_ref.float()).abs()
        same = torch.equal(torch.sort(vic, dim=-1)[0], torch.sort(v_ref, dim=-1)[0])
        nan = int(torch.isnan(o.float()).sum())
        print(f"{str(dtype)[6:]:9s} H{H} Hkv{Hkv} s{stride} {policy:8s} n{n0} step{t}: out err max {err.max().item():.3e} (ref max {o_ref.float().abs().max().item():.2f}) "
              f"nan {nan} victims_equal {same} margin {orc.margin(0)}", flush=True)
        if err.max().item() > 5e-3 or nan:
            hh, qq, dd = [int(x) for x in torch.nonzero(err == err.max())[0]] if not nan else (0, 0, 0)
            print("   worst at head", hh, "query", qq, "dim", dd, "| per-head max err", [round(float(x), 4) for x in err.amax(dim=(1, 2))][:16])
            print("   per-query max err", [round(float(x), 4) for x in err.amax(dim=(0, 2))][:32])
    Kc, Vc = eng.export(0)
    print("   cache equal:", torch.equal(Kc, orc.export(0)[0]) and torch.equal(Vc, orc.export(0)[1]), flush=True)
lib.ekv_debug_set_chunk_variant(0)
