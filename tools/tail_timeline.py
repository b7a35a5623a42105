"""Stage timeline of chunk_tail_kernel (ekv_debug_set_timeline)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B, H, Hkv, n, stride = (int(x) for x in sys.argv[1:6]); policy = sys.argv[6]
d, dev = 128, "cuda"
cache = BudgetedKVCache(1, B, H, Hkv, d, n + stride, dtype=torch.float16, arith=1)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device=dev).half(), torch.randn(B, Hkv, n, d, device=dev).half(), n, [float(n - i) for i in range(n)])
cache.S[0][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[0][:, :, :n] / n
cache.SQ[0][:, :, :n] = cache.S[0][:, :, :n] ** 2 / cache.Cn[0][:, :, :n] * 1.5
recent = int(n * 0.1)
sp = StepParams(policy=policy, accumulate=True, evict=stride, counter_add=float(stride), c_new_step=1.0, k_feasible=max(n - recent - 4, stride),
                sink_protect=4, win_lo=4, win_recent=recent, range_start=4)
q = torch.randn(B, H, stride, d, device=dev).half() * 0.3; k = torch.randn(B, Hkv, stride, d, device=dev).half(); v = torch.randn_like(k)
for _ in range(3):
    cache.step(0, sp, q, k, v)
tl = torch.zeros(B * Hkv, 8, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
t = tl.cpu().double()
names = ["load lj + append -> entry", "pass 1 (state, keys)", "(single-victim path)", "select stages", "gather + sort victims", "renumber"]
seq = [t[:, 6], t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4], t[:, 5]]
for i in range(6):
    print(f"  {names[i]:28s} mean {(seq[i+1]-seq[i]).mean():9.0f} cyc")
print(f"  total {(seq[6]-seq[0]).mean():.0f} cyc = {(seq[6]-seq[0]).mean()/1.9e3:.1f} us")
