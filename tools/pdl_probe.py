"""Does programmatic dependent launch survive CUDA-graph capture?  Times L back-to-back layer launches of a small decode
workload (a) replayed from a graph and (b) launched eagerly on the stream; run once with EKV_NO_PDL=1 and once without.
python tools/pdl_probe.py [workload ...]"""
import json, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["c2_b1", "c2_b8", "c5"]:
    w = dict(bench.WORKLOADS[name])
    cache, steady, L, q, kn, vn = bench.build_workload(w, w["B"], dev, L=8)
    for _ in range(20):
        steady.replay()
    torch.cuda.synchronize()
    res = {}
    for label, fn in (("graph", steady.replay), ("eager", lambda: [steady.run_layer(l) for l in range(L)])):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 40
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res[label] = round(e0.elapsed_time(e1) * 1e3 / (reps * L), 2)
    print(json.dumps(dict(workload=name, pdl=0 if os.environ.get("EKV_NO_PDL") else 1, us_per_layer=res)), flush=True)
