"""Phase timeline of decode_umma_kernel (ekv_debug_set_timeline): python tools/decode_umma_timeline.py B H Hkv n [cluster] [policy]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B, H, Hkv, n = (int(x) for x in sys.argv[1:5])
cluster = int(sys.argv[5]) if len(sys.argv) > 5 else 0
policy = sys.argv[6] if len(sys.argv) > 6 else "roco"
d, dev = 128, "cuda"
cache = BudgetedKVCache(1, B, H, Hkv, d, n + 1, dtype=torch.float16, arith=1)
cache.lib.ekv_debug_set_dispatch(5, cluster)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device=dev).half(), torch.randn(B, Hkv, n, d, device=dev).half(), n,
                   [float(n - i) for i in range(n)])
cache.S[0][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[0][:, :, :n] / n
cache.SQ[0][:, :, :n] = cache.S[0][:, :, :n] ** 2 / cache.Cn[0][:, :, :n] * 1.5
recent = int(n * 0.3)
sp = StepParams(policy=policy, accumulate=policy != "recency", evict=1, counter_add=1.0, k_feasible=n - recent,
                win_recent=recent if policy == "h2o_head" else 0, range_start=4)
q = torch.randn(B, H, 1, d, device=dev).half() * 0.3; k = torch.randn(B, Hkv, 1, d, device=dev).half(); v = torch.randn_like(k)
for _ in range(5):
    cache.step(0, sp, q, k, v)
tl = torch.zeros(4096, 16, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
cache.lib.ekv_debug_set_dispatch(0, 0)
t = tl.cpu().double()
t = t[t[:, 0] > 0]
names = ["K phase", "row statistics + exchange", "V phase (p, P^T, state, keys)", "new token + output gather", "select walk", "apply"]
print(f"{t.shape[0]} CTAs; softmax-warp lifetime mean {(t[:, 6] - t[:, 0]).mean() / 1e3:.1f} us; span {(t[:, 6].max() - t[:, 0].min()) / 1e3:.1f} us; walk rounds mean {t[:, 8].mean():.1f} max {t[:, 8].max():.0f}")
for i, nm in enumerate(names):
    dlt = (t[:, i + 1] - t[:, i]) / 1e3
    print(f"  {nm:32s} mean {dlt.mean():8.2f} us   p10 {dlt.quantile(0.1):8.2f}   p90 {dlt.quantile(0.9):8.2f}")
