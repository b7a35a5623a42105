"""Phase timeline of decode_umma_kernel UNDER bench.py's steady state: python tools/decode_umma_timeline_bench.py [workload=c5] [B]"""
import os, sys, collections
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
name = sys.argv[1] if len(sys.argv) > 1 else "c5"
w = dict(bench.WORKLOADS[name])
B = int(sys.argv[2]) if len(sys.argv) > 2 else w["B"]
w["B"] = B
dev = torch.device("cuda", 0)
cache, steady, L, q, kn, vn = bench.build_workload(w, B, dev, L=2)
for _ in range(40):
    steady.replay()
torch.cuda.synchronize()
tl = torch.zeros(4096, 16, dtype=torch.int64, device=dev)
cache.lib.ekv_debug_set_timeline(tl.data_ptr())
steady.run_layer(0)
torch.cuda.synchronize()
cache.lib.ekv_debug_set_timeline(None)
t = tl.cpu().double()
t = t[t[:, 0] > 0]
names = ["K phase", "row statistics + exchange", "V phase (p, P^T)", "new token + output gather", "select", "apply"]
print(f"{name} B={B}: {t.shape[0]} CTAs; softmax-warp lifetime mean {(t[:, 6] - t[:, 0]).mean() / 1e3:.1f} us; span {(t[:, 6].max() - t[:, 0].min()) / 1e3:.1f} us")
codes = collections.Counter(int(x) for x in t[:, 8].tolist())
print("  select path (1 + status + 4 * kind; status 0 ok 1 none 2 fallback; kind 0 direct 1 pure bucket 2 refined):", dict(codes))
for i, nm in enumerate(names):
    dlt = (t[:, i + 1] - t[:, i]) / 1e3
    print(f"  {nm:32s} mean {dlt.mean():8.2f} us   p10 {dlt.quantile(0.1):8.2f}   p90 {dlt.quantile(0.9):8.2f}")
if (t[:, 9] > 0).all():          # the select's pieces (roco): tail barrier + histogram exchange | scan | pass | exchange | final ranks + argmin
    parts = [("tail barrier + histogram exchange", 7, 9), ("scan", 9, 10), ("classify pass", 10, 11), ("exchange", 11, 12), ("ranks + argmin", 12, 5)]
    for nm, i0, i1 in parts:
        dlt = (t[:, i1] - t[:, i0]) / 1e3
        print(f"    select: {nm:34s} mean {dlt.mean():6.2f} us")
    print(f"    classify pass: its entry loop alone {((t[:, 13] - t[:, 10]) / 1e3).mean():.2f} us (thread 0)")
    print(f"    boundary entries ranked by brute force: mean {t[:, 14].mean():.0f} max {t[:, 14].max():.0f}")
if (t[:, 7] > 0).all():          # the tail runs on the helper warps beside the softmax warps' output gather
    print(f"    tail start (helper warps) relative to the end of the V phase {((t[:, 7] - t[:, 3]) / 1e3).mean():+.2f} us; tail end after the output gather's end {((t[:, 6] - t[:, 4]) / 1e3).mean():+.2f} us; CTA busy {((t[:, 6] - t[:, 0]) / 1e3).mean():.1f} us")

