"""Phase timeline of the cluster decode kernel (ekv_debug_set_timeline): average cycles per phase."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B, H, Hkv, n, cluster = (int(x) for x in sys.argv[1:6])
variant = int(sys.argv[6]) if len(sys.argv) > 6 else 0
d, dev = 128, "cuda"
cache = BudgetedKVCache(1, B, H, Hkv, d, n + 1, dtype=torch.float16, arith=1)
cache.lib.ekv_debug_set_dispatch(variant, cluster)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device=dev).half(), torch.randn(B, Hkv, n, d, device=dev).half(), n,
                   [float(n - i) for i in range(n)])
cache.S[0][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[0][:, :, :n] / n
cache.SQ[0][:, :, :n] = cache.S[0][:, :, :n] ** 2 / cache.Cn[0][:, :, :n] * 1.5
sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))
q = torch.randn(B, H, 1, d, device=dev).half() * 0.3; k = torch.randn(B, Hkv, 1, d, device=dev).half(); v = torch.randn_like(k)
for _ in range(5):
    cache.step(0, sp, q, k, v)
ncta = B * Hkv * 8
tl = torch.zeros(ncta, 8, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
tl = tl.cpu()
used = tl[:, 0] != 0
t = tl[used].double()
names = ["header", "K phase", "softmax(2 barriers)", "V phase", "tail pass1+best", "barrier3 wait.. select", "-"]
print(f"CTAs {int(used.sum())}; attempts max {int(t[:,7].max())}")
for i in range(6):
    dcy = (t[:, i + 1] - t[:, i])
    print(f"  {names[i]:28s} mean {dcy.mean():10.0f} cyc  max {dcy.max():10.0f}")
print(f"  total mean {(t[:,6]-t[:,0]).mean():.0f} cyc = {(t[:,6]-t[:,0]).mean()/1.9e3:.1f} us")
