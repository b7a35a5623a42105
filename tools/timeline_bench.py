"""Phase timeline of the persistent decode kernel UNDER bench.py's steady state (pool-refreshed inputs): mean cycles per
phase over all CTAs and units.  python tools/timeline_bench.py [workload=c2] [B=32] [variant=0]
(B <= 32 on the 7B layout keeps <= 7 units per CTA, which is what the kernel's profiling rows hold.)"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "c2"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = torch.device("cuda", 0)
w = dict(bench.WORKLOADS[name]); w["B"] = B
cache, steady, L, q, kn, vn = bench.build_workload(w, B, dev, L=2)
cache.lib.ekv_debug_set_dispatch(variant, 0)
for _ in range(60):
    steady.replay()
torch.cuda.synchronize()
tl = torch.zeros(296, 16, 8, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
steady.run_layer(0)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
cache.lib.ekv_debug_set_dispatch(0, 0)
t = tl.cpu().double()
names = ["hdr", "tile0 wait", "K phase", "softmax", "V phase", "out", "tail"]
acc = {n: [] for n in names}
tail = {"pass1": [], "select": [], "apply": []}
life = []
for cta in range(296):
    if t[cta, 0, 0] == 0:
        continue
    last = 0
    for ku in range(7):
        row = t[cta, ku]
        if row[0] == 0:
            continue
        for i, n in enumerate(names):
            acc[n].append((row[i + 1] - row[i]).item())
        last = max(last, row[7].item())
        tr = t[cta, 8 + ku]
        if tr[0] > 0 and tr[5] >= tr[2] >= tr[1] >= tr[0]:
            tail["pass1"].append((tr[1] - tr[0]).item()); tail["select"].append((tr[2] - tr[1]).item()); tail["apply"].append((tr[5] - tr[2]).item())
    life.append(last - t[cta, 0, 0].item())
    pw = t[cta, 15, 0].item() / max(t[cta, 15, 1].item(), 1)
mean = lambda v: sum(v) / max(len(v), 1)
print(f"{name} B={B} variant={variant}: {len(life)} CTAs, lifetime mean {mean(life):.0f} cycles = {mean(life) / 1.92e3:.1f} us")
print("  " + "  ".join(f"{n} {mean(v):.0f}" for n, v in acc.items()))
print("  tail: " + "  ".join(f"{n} {mean(v):.0f} (max {max(v) if v else 0:.0f})" for n, v in tail.items()))
