"""Phase timeline of the decode kernel (ekv_debug_set_timeline): prints per-phase cycle counts."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams
B = int(os.environ.get("B", 32)); n = 1088; H = Hkv = 32; d = 128
cache = BudgetedKVCache(1, B, H, Hkv, d, n + 1, dtype=torch.float16)
cache.load_prefill(0, torch.randn(B, Hkv, n, d, device="cuda").half(), torch.randn(B, Hkv, n, d, device="cuda").half(), n,
                   [float(n - i) for i in range(n)])
sp = StepParams(policy="roco", accumulate=True, evict=1, counter_add=1.0, k_feasible=n - int(n * 0.3))
q = torch.randn(B, H, 1, d, device="cuda").half(); k = torch.randn(B, Hkv, 1, d, device="cuda").half(); v = torch.randn_like(k)
for _ in range(3):
    cache.step(0, sp, q, k, v)
tl = torch.zeros(148, 16, 8, dtype=torch.int64, device="cuda")
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
cache.step(0, sp, q, k, v)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
tl = tl.cpu()
names = ["start", "hdr", "tile0", "Kend", "smax", "Vend", "out", "tail"]
for cta in (0, 1, 73, 147):
    base = tl[cta, 0, 0].item()
    if base == 0:
        continue
    print(f"CTA {cta} (globaltimer start {tl[cta,15,7].item() - tl[:,15,7][tl[:,15,7]>0].min().item()} ns)")
    for ku in range(8):
        row = tl[cta, ku]
        if row[0].item() == 0:
            continue
        rel = [(x.item() - base) for x in row]
        dif = [rel[0]] + [rel[i] - rel[i - 1] for i in range(1, 8)]
        print(f"  unit#{ku} g{ku%2}: start@{rel[0]:>7}  " + " ".join(f"{nm}+{dd}" for nm, dd in zip(names[1:], dif[1:])) + f"  end@{rel[7]}")
        if ku < 7:
            tr = [x.item() for x in tl[cta, 8 + ku][:6]]
            print("      tail: pre+%d pass1+%d fast+%d radix+%d gather+%d apply+%d post+%d" % (
                tr[0] - row[6].item(), tr[1] - tr[0], tr[2] - tr[1], tr[3] - tr[2], tr[4] - tr[3], tr[5] - tr[4], row[7].item() - tr[5]))
