"""A/B of the decode kernels on bench.py's GQA workloads (L2-cold, graph replay): automatic dispatch (the tcgen05 GQA
kernel where it applies) against decode_variant 6 (round-1 kernels only) and forced cluster sizes."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from easykv_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
hbm, tf, _ = bench.peaks()


def barrier():
    torch.cuda.synchronize()


names = sys.argv[1:] or ["c5", "c5_b32", "c3_decode", "c3_gen", "c2_b8", "c2_b1"]
extra = {"m7b_n1088_b32": dict(kind="decode", model="Mistral-7B", L=32, H=32, Hkv=8, n=1088, B=32, policy="roco"),
         "70b_n1088_b32": dict(kind="decode", model="Llama-2-70B", L=80, H=64, Hkv=8, n=1088, B=32, policy="roco"),
         "c5_b1": dict(kind="decode", model="Llama-2-70B", L=80, H=64, Hkv=8, n=8256, B=1, policy="roco"),
         "c3_decode_b1": dict(kind="decode", model="Mistral-7B", L=32, H=32, Hkv=8, n=8208, B=1, policy="roco"),
         "c3_decode_b4": dict(kind="decode", model="Mistral-7B", L=32, H=32, Hkv=8, n=8208, B=4, policy="roco")}
bench.WORKLOADS.update(extra)
for name in names:
    for label, variant, cluster in (("auto", 0, 0), ("round1", 6, 0), ("umma_c1", 5, 1), ("umma_c2", 5, 2), ("umma_c4", 5, 4)):
        lib.ekv_debug_set_dispatch(variant, cluster)
        try:
            r = bench.run_sweep_item(name, dev, barrier, hbm, tf, target_s=0.15)
            print(json.dumps(dict(workload=name, dispatch=label, us=r["us_per_layer_forward"], frac=r["frac"])), flush=True)
        except Exception as exc:
            print(json.dumps(dict(workload=name, dispatch=label, error=f"{type(exc).__name__}: {exc}"[:160])), flush=True)
        finally:
            lib.ekv_debug_set_dispatch(0, 0)
