"""Diagnostic (not a test): replays every golden trace through the CUDA path (both kernels) and prints
victim mismatches with the oracle's decision margins, plus a first timing of the decode kernel."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import replay  # noqa: E402
from engines import CudaEngine  # noqa: E402

out = {}
for name in replay.list_golden():
    for kernel in (0, 1):
        try:
            rep = replay.replay(name, lambda *a: CudaEngine(*a, kernel=kernel), resync=True, shadow=replay.OracleEngine)
            out[f"{name}/k{kernel}"] = dict(
                fwd=rep.n_forwards, events=rep.n_events, mismatches=len(rep.victim_mismatch),
                ties=len(rep.tie_ambiguous), max_out_err=rep.max_out_err, final_equal=rep.final_cache_equal,
                min_margin=rep.min_margin,
                mm=[(f, l, m) for f, l, _, _, m in rep.victim_mismatch[:5]])
        except Exception as e:  # noqa: BLE001
            out[f"{name}/k{kernel}"] = dict(error=repr(e))
        print(name, kernel, out[f"{name}/k{kernel}"], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gpu_report.json"), "w"), indent=1, default=str)
