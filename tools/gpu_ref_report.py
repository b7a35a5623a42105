"""Replay statistics of every golden (CPU-recorded tests/golden, B200-recorded tests/golden_gpu or a directory given
on the command line) through the CUDA path with arith = 0 / 1: events, victim mismatches with the shadow oracle's
decision margins, exact ties, output error.  Development aid for setting the parity tests' thresholds.

    python tools/gpu_ref_report.py [golden_dir]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import replay  # noqa: E402
import engines  # noqa: E402
from easykv_b200 import build  # noqa: E402

build.build()
gdir = sys.argv[1] if len(sys.argv) > 1 else replay.GOLDEN_GPU_DIR
names = sorted(f[:-4] for f in os.listdir(gdir) if f.endswith(".npz") and not f.startswith("sampling"))
for name in names:
    for arith in (0, 1):
        shadow = lambda *a: replay.OracleEngine(*a, scale_mul=bool(arith))
        rep = replay.replay(name, lambda *a: engines.CudaEngine(*a, arith=arith), resync=True, shadow=shadow, golden_dir=gdir)
        mm = [(f, l, int((r != g).sum()), [float(x) for x in m]) for f, l, r, g, m in rep.victim_mismatch]
        print(json.dumps(dict(name=name, arith=arith, events=rep.n_events, mismatches=len(rep.victim_mismatch),
                              exact_ties=len(rep.tie_ambiguous), max_out_err=rep.max_out_err, cache_equal=rep.final_cache_equal,
                              mismatch_detail=mm[:6])), flush=True)
