"""Phase timeline of chunk_umma_kernel (ekv_debug_set_timeline): per CTA, %globaltimer stamps at
start | setup done | K phase done | max exchange done | L pass done | sum exchange done | V phase done | epilogue done.

    python tools/umma_timeline.py B H Hkv n stride policy [chunk_variant]
"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B, H, Hkv, n, stride = (int(x) for x in sys.argv[1:6]); policy = sys.argv[6]
variant = int(sys.argv[7]) if len(sys.argv) > 7 else 0
d, dev = 128, "cuda"
cache = BudgetedKVCache(1, B, H, Hkv, d, n + stride, dtype=torch.float16, arith=1)
cache.lib.ekv_debug_set_chunk_variant(variant)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device=dev).half(), torch.randn(B, Hkv, n, d, device=dev).half(), n, [float(n - i) for i in range(n)])
cache.S[0][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[0][:, :, :n] / n
cache.SQ[0][:, :, :n] = cache.S[0][:, :, :n] ** 2 / cache.Cn[0][:, :, :n] * 1.5
recent = int(n * 0.1)
sp = StepParams(policy=policy, accumulate=True, evict=stride, counter_add=float(stride), c_new_step=1.0, k_feasible=max(n - recent - 4, stride),
                sink_protect=4, win_lo=4, win_recent=recent, range_start=4)
q = torch.randn(B, H, stride, d, device=dev).half() * 0.3; k = torch.randn(B, Hkv, stride, d, device=dev).half(); v = torch.randn_like(k)
for _ in range(3):
    cache.step(0, sp, q, k, v)
U = B * Hkv
NC = 8192
tl = torch.zeros(U * 8 + NC * 16, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
t = tl[U * 8:].view(NC, 16).cpu().double()
t = t[t[:, 0] > 0]
for i in (4, 5):                      # stamps 4 / 5 exist only when the exact L pass ran (slow path)
    t[:, i] = torch.where(t[:, i] > 0, t[:, i], t[:, i - 1])
print(f"{t.shape[0]} CTAs; kernel span {(t[:, 7].max() - t[:, 0].min()) / 1e3:.1f} us; CTA lifetime mean {(t[:, 7] - t[:, 0]).mean() / 1e3:.1f} us, max {(t[:, 7] - t[:, 0]).max() / 1e3:.1f} us")
names = ["setup", "K phase (logits, max, sums)", "row-statistics exchange", "exact L pass (slow path)", "its exchange", "V phase (p, stats, P^T)", "epilogue"]
for i, nm in enumerate(names):
    dlt = (t[:, i + 1] - t[:, i]) / 1e3
    print(f"  {nm:26s} mean {dlt.mean():7.2f} us   p10 {dlt.quantile(0.1):7.2f}   p90 {dlt.quantile(0.9):7.2f}")
print(f"  inside setup: barriers+first TMA issued {((t[:, 8] - t[:, 0]) / 1e3).mean():.2f} us | thread 0 done with vmask + Q fill {((t[:, 9] - t[:, 0]) / 1e3).mean():.2f} | "
      f"__syncthreads passed {((t[:, 10] - t[:, 0]) / 1e3).mean():.2f} | cluster sync passed {((t[:, 1] - t[:, 0]) / 1e3).mean():.2f}")
starts = torch.sort(t[:, 0] - t[:, 0].min())[0] / 1e3
print("  CTA start times (us), deciles:", [round(float(starts[int(i * (len(starts) - 1) / 10)]), 1) for i in range(11)])
