"""A/B of the MHA decode dispatch on bench.py's workloads (evolving state): automatic, one consumer group per CTA with two
CTAs per SM (decode_variant 1), ping-pong groups (2), and the cluster kernel forced to one CTA per unit.
python tools/decode_variants.py [workload ...]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from easykv_b200 import _lib  # noqa: E402
lib = _lib.load()
dev = torch.device("cuda", 0)
hbm, tf, _ = bench.peaks()
for name in sys.argv[1:] or ["c2", "c2_b8"]:
    for label, variant, cluster in (("auto", 0, 0), ("one group per CTA, 2 CTAs/SM", 1, -1), ("ping-pong groups", 2, -1), ("cluster kernel C=1", 0, 1), ("cluster kernel C=2", 0, 2)):
        lib.ekv_debug_set_dispatch(variant, cluster)
        try:
            r = bench.run_sweep_item(name, dev, torch.cuda.synchronize, hbm, tf, target_s=0.15)
            print(json.dumps(dict(workload=name, dispatch=label, us=r["us_per_layer_forward"], frac=r["frac"])), flush=True)
        except Exception as exc:
            print(json.dumps(dict(workload=name, dispatch=label, error=f"{type(exc).__name__}: {exc}"[:120])), flush=True)
        finally:
            lib.ekv_debug_set_dispatch(0, 0)
