"""Phase timeline of the decode kernel (ekv_debug_set_timeline): prints per-phase cycle counts."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B = int(os.environ.get("B", 32)); n = 1088; H = Hkv = 32; d = 128
dev = "cuda"
cache = BudgetedKVCache(1, B, H, Hkv, d, n + 1, dtype=torch.float16)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device=dev).half(), torch.randn(B, Hkv, n, d, device=dev).half(), n,
                   [float(n - i) for i in range(n)])
cache.S[0][:, :, :n] = torch.rand(B, Hkv, n, device=dev) * cache.Cn[0][:, :, :n] / n     # bench.py's synthetic steady state
cache.SQ[0][:, :, :n] = cache.S[0][:, :, :n] ** 2 / cache.Cn[0][:, :, :n] * 1.5
sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))
q = torch.randn(B, H, 1, d, device=dev).half() * 0.3; k = torch.randn(B, Hkv, 1, d, device=dev).half(); v = torch.randn_like(k)
for _ in range(20):
    cache.step(0, sp, q, k, v)
tl = torch.zeros(296, 16, 8, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
tl = tl.cpu()
names = ["start", "hdr", "tile0", "Kend", "smax", "Vend", "out", "tail"]
ends = []
for cta in range(296):
    if tl[cta, 0, 0].item() == 0:
        continue
    base = tl[cta, 0, 0].item()
    last = max(tl[cta, ku, 7].item() for ku in range(8))
    ends.append(last - base)
    if cta not in (0, 73, 147):
        continue
    print(f"CTA {cta}: producer ring-full wait {tl[cta,15,0].item()} of {tl[cta,15,1].item()} cycles")
    for ku in range(8):
        row = tl[cta, ku]
        if row[0].item() == 0:
            continue
        rel = [(x.item() - base) for x in row]
        dif = [rel[0]] + [rel[i] - rel[i - 1] for i in range(1, 8)]
        print(f"  unit#{ku} g{ku%2}: start@{rel[0]:>7}  " + " ".join(f"{nm}+{dd}" for nm, dd in zip(names[1:], dif[1:])) + f"  end@{rel[7]}")
        if ku < 7:
            tr = [x.item() for x in tl[cta, 8 + ku][:7]]
            print("      tail: pre+%d pass1+%d select+%d apply+%d | tile-wait in K+V: %d" % (
                tr[0] - row[6].item(), tr[1] - tr[0], tr[2] - tr[1], tr[5] - tr[2], tr[6]))
import statistics
print("CTA lifetimes (cycles): min %d median %d max %d" % (min(ends), statistics.median(ends), max(ends)))
pw = [tl[c, 15, 0].item() / max(tl[c, 15, 1].item(), 1) for c in range(296) if tl[c, 15, 1].item() > 0]
print("producer ring-full fraction: mean %.3f" % (sum(pw) / len(pw)))
