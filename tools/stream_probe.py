"""Probe: one fused-streaming decode step for a given shape (development aid): python tools/stream_probe.py B H Hkv n dtype variant cluster"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B, H, Hkv, n = (int(x) for x in sys.argv[1:5])
dtype = getattr(torch, sys.argv[5]); variant, cluster = int(sys.argv[6]), int(sys.argv[7])
D, dev = 128, "cuda"
inv = 1.0 / (10000.0 ** (torch.arange(0, D, 2, dtype=torch.float32) / D))
f = torch.outer(torch.arange(n + 16, dtype=torch.float32), inv)
emb = torch.cat([f, f], dim=-1)
cos, sin = emb.cos().to(dtype).to(dev), emb.sin().to(dtype).to(dev)
c = BudgetedKVCache(1, B, H, Hkv, D, n + 8, dtype=dtype, arith=1)
c.enable_streaming()
c.load_prefill(0, torch.randn(B, Hkv, n, D, device=dev).to(dtype), torch.randn(B, Hkv, n, D, device=dev).to(dtype), n, [float(n - i) for i in range(n)])
sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))
c.lib.ekv_debug_set_dispatch(variant, cluster)
for t in range(3):
    q = torch.randn(B, 1, H * D, device=dev).to(dtype) * 0.5
    k = torch.randn(B, 1, Hkv * D, device=dev).to(dtype); v = torch.randn(B, 1, Hkv * D, device=dev).to(dtype)
    o, vl = c.step_stream(0, sp, q, k, v, cos, sin)
    torch.cuda.synchronize()
print("ok", sys.argv[1:], vl.flatten().tolist()[:4])
