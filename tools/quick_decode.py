"""Quick timing of the fused decode kernel (development aid; bench.py is the contract)."""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache  # noqa: E402
from easykv_b200.plan import StepParams  # noqa: E402

B = int(os.environ.get("B", 32)); L = int(os.environ.get("L", 4)); n = int(os.environ.get("N", 1088))
H = int(os.environ.get("H", 32)); Hkv = int(os.environ.get("HKV", 32)); d = 128
steps = int(os.environ.get("STEPS", 20))
kernel = int(os.environ.get("KERNEL", 0))
torch.manual_seed(0)
cache = BudgetedKVCache(L, B, H, Hkv, d, n + 1, dtype=torch.float16)
for l in range(L):
    K = torch.randn(B, Hkv, n, d, device="cuda", dtype=torch.float16)
    V = torch.randn(B, Hkv, n, d, device="cuda", dtype=torch.float16)
    cache.load_prefill(l, K, V, n, [float(n - i) for i in range(n)])
budget = n
sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=budget - int(budget * 0.3))
q = torch.randn(L, B, H, 1, d, device="cuda", dtype=torch.float16)
kn = torch.randn(L, B, Hkv, 1, d, device="cuda", dtype=torch.float16)
vn = torch.randn(L, B, Hkv, 1, d, device="cuda", dtype=torch.float16)
for _ in range(3):
    for l in range(L):
        cache.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    for l in range(L):
        out, vl = cache.step(l, sp, q[l], kn[l], vn[l], kernel=kernel)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / (steps * L)
bytes_alg = B * (2 * Hkv * (n + 1) * d * 2 + 2 * Hkv * d * 2 + 2 * H * d * 2 + 6 * Hkv * (n + 1) * 4 + Hkv * 4)
print(f"B={B} L={L} n={n} H={H} Hkv={Hkv} kernel={kernel}: {ms*1e3:.1f} us/layer-call  {bytes_alg/ms/1e6:.0f} GB/s  victims[0,0]={vl[0,0].tolist()}")
