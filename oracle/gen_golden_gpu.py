"""TEST INFRASTRUCTURE — runs ON THE GPU BOX.

Records runs of the UNMODIFIED reference (vendored by tools/vendor_ref.sh into oracle/_ref/, or /root/reference when
mounted) executing in fp16 / bf16 ON CUDA — the arithmetic the product's default `arith=1` reproduces (ATen's CUDA
kernels: logits * (1/sqrt(d)), exp / sum; cuBLAS fp16 GEMMs) — on the 4.36-shaped scaffold (oracle/scaffold.py), in the
same format as tests/golden/ (oracle/gen_golden.py).  Output: gpurun_out/golden_gpu/*.npz, committed as
tests/golden_gpu/*.npz and replayed by tests/test_gpu_reference_goldens.py with arith=1.

    gpurun -- 'python -m oracle.gen_golden_gpu'

Also the SURVEY A.5 probe: which of several EQUAL keys `torch.topk(largest=False)` / `argmin` pick on CUDA
(gpurun_out/golden_gpu/topk_tie_probe.json).
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gen_golden  # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden_gpu")

# small enough to commit (<= ~2 MB each; the reference requires hidden_size == H*d, llama_patch.py:232)
CASES = {
    "gpu_llama_auto_roco_fp16": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=208, dtype="float16",
                                     mode="auto", stride=16, max_new_tokens=40, gen=dict(budget=96, kv_policy="roco")),
    "gpu_mistral_g4_auto_roco_fp16": dict(arch="mistral", L=2, H=8, Hkv=2, d=128, seq=160, dtype="float16",
                                          mode="auto", stride=8, max_new_tokens=32, gen=dict(budget=64, kv_policy="roco")),
    "gpu_mistral_g8_enc_h2o_fp16": dict(arch="mistral", L=1, H=8, Hkv=1, d=128, seq=208, dtype="float16",
                                        mode="encoding", stride=16, max_new_tokens=2, gen=dict(budget=0.5, kv_policy="h2o_head")),
    "gpu_llama_decoding_roco_fp16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=48, dtype="float16",
                                         mode="decoding", stride=1, max_new_tokens=88, gen=dict(budget=40, kv_policy="roco")),
    "gpu_llama_enc_roco_bf16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=160, dtype="bfloat16",
                                    mode="encoding", stride=8, max_new_tokens=2, gen=dict(budget=0.5, kv_policy="roco")),
    "gpu_gqa_llama_auto_tova_fp16": dict(arch="llama", L=1, H=8, Hkv=2, d=128, seq=128, dtype="float16",
                                         mode="auto", stride=8, max_new_tokens=24, gen=dict(budget=48, kv_policy="tova")),
    "gpu_mistral_enc_h2o_keep_fp16": dict(arch="mistral", L=1, H=8, Hkv=2, d=128, seq=164, dtype="float16",
                                          mode="encoding", stride=4, max_new_tokens=2,
                                          gen=dict(budget=0.5, kv_policy="h2o_head", keep_attention=True)),
    "gpu_llama_ppl_roco_fp16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=136, dtype="float16",
                                    mode="ppl", stride=8, max_new_tokens=0, gen=dict(budget=0.4, kv_policy="roco")),
}


def topk_tie_probe():
    """the reference's victims are `topk(largest=False)` / `argmin` indices; among EQUAL keys the pick is
    unspecified.  This package defines it as lowest index first; record what ATen's CUDA kernels do."""
    dev = "cuda"
    out = {}
    g = torch.Generator().manual_seed(0)
    for n, k in ((144, 8), (1089, 762), (1089, 1), (8224, 16), (8320, 64)):
        x = torch.randint(0, 6, (4, n), generator=g).float().to(dev)            # heavy ties
        idx = torch.topk(x, k=k, dim=-1, largest=False)[1]
        # stable expectation: (value asc, index asc)
        exp = torch.sort(x, dim=-1, stable=True)[1][:, :k]
        same_set = all(set(idx[h].tolist()) == set(exp[h].tolist()) for h in range(4))
        same_order = bool(torch.equal(idx, exp))
        am = torch.argmin(x, dim=-1)
        first_min = torch.stack([(x[h] == x[h].min()).nonzero()[0, 0] for h in range(4)])
        out[f"n{n}_k{k}"] = dict(topk_set_is_lowest_index_first=same_set, topk_order_is_stable=same_order,
                                 argmin_is_first_minimum=bool(torch.equal(am, first_min)))
    out["torch"] = torch.__version__
    out["gpu"] = torch.cuda.get_device_name(0)
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    if not only:
        json.dump(topk_tie_probe(), open(os.path.join(OUT, "topk_tie_probe.json"), "w"), indent=1)
        print(open(os.path.join(OUT, "topk_tie_probe.json")).read())
    for name, c in CASES.items():
        if only and name not in only:
            continue
        tr = gen_golden.run_case(name, c, device="cuda", out_dir=OUT)
        sz = os.path.getsize(os.path.join(OUT, name + ".npz")) / 1e6
        last = tr.printed.strip().splitlines()[-1] if tr.printed.strip() else ""
        print(f"{name}: {len(tr.forwards)} forwards, {len(tr.events)} eviction events, {sz:.2f} MB | {last}", flush=True)
