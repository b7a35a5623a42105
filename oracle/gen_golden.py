"""TEST INFRASTRUCTURE — build-container only.

Generates `tests/golden/*.npz` by running the UNMODIFIED reference (`/root/reference/easykv`)
on the 4.36-shaped scaffold via `oracle/ref_harness.py`.  Each file freezes, for one
(model shape, mode, policy) case: the densely prefilled K/V, then for every later forward and
layer the post-RoPE q, the new k/v and the attention output entering o_proj, every eviction
event the reference issued (victim ids exactly as passed to `truncate_kv_cache_*`), the final
cache and the console line with the retained-cache ratio.

    python -m oracle.gen_golden            # regenerate everything (≈1 min, CPU)

Reproducible within this image (torch 2.11.0 CPU); victim ids are what the parity tests pin.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness, scaffold  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
RNG_SEED = 20240229

CASES = {
    # BASELINE.json configs[0]: the reference's own CPU-runnable case
    "c1_llama_enc_roco_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=256, dtype="float32",
                                   mode="encoding", stride=8, max_new_tokens=4,
                                   gen=dict(budget=0.5, kv_policy="roco")),
    "gqa_mistral_auto_roco_fp32": dict(arch="mistral", L=2, H=8, Hkv=2, d=128, seq=160, dtype="float32",
                                       mode="auto", stride=8, max_new_tokens=24,
                                       gen=dict(budget=64, kv_policy="roco")),
    "llama_decoding_roco_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=32, dtype="float32",
                                     mode="decoding", stride=1, max_new_tokens=72,
                                     gen=dict(budget=40, kv_policy="roco")),
    "llama_decoding_h2o_fp32": dict(arch="llama", L=1, H=4, Hkv=2, d=128, seq=24, dtype="float32",
                                    mode="decoding", stride=1, max_new_tokens=40,
                                    gen=dict(budget=16, kv_policy="h2o_head")),
    "gqa_mistral_enc_h2o_fp32": dict(arch="mistral", L=2, H=8, Hkv=2, d=128, seq=132, dtype="float32",
                                     mode="encoding", stride=4, max_new_tokens=3,
                                     gen=dict(budget=0.5, kv_policy="h2o_head")),
    "llama_enc_tova_fp32": dict(arch="llama", L=2, H=4, Hkv=2, d=128, seq=120, dtype="float32",
                                mode="encoding", stride=8, max_new_tokens=3,
                                gen=dict(budget=0.5, kv_policy="tova")),
    "llama_auto_tova_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=96, dtype="float32",
                                 mode="auto", stride=8, max_new_tokens=20,
                                 gen=dict(budget=40, kv_policy="tova")),
    "llama_auto_recency_fp32": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=64, dtype="float32",
                                    mode="auto", stride=4, max_new_tokens=8,
                                    gen=dict(budget=32, kv_policy="recency")),
    "llama_ppl_roco_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=144, dtype="float32",
                                mode="ppl", stride=8, max_new_tokens=0,
                                gen=dict(budget=0.4, kv_policy="roco")),
    # kv_policy='random': victims drawn from torch's CPU generator (seeded below), easykv.py:353-357, :494-499
    "llama_decoding_random_fp32": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=24, dtype="float32",
                                       mode="decoding", stride=1, max_new_tokens=40,
                                       gen=dict(budget=16, kv_policy="random")),
    "llama_enc_random_fp32": dict(arch="llama", L=1, H=4, Hkv=2, d=128, seq=120, dtype="float32",
                                  mode="encoding", stride=8, max_new_tokens=2,
                                  gen=dict(budget=0.5, kv_policy="random")),
    # streaming=True (llama_forward_stream / mistral_forward_stream): un-rotated keys in the cache, RoPE re-applied at
    # cache-relative positions every forward; q / k in these traces are UN-rotated
    "llama_auto_roco_stream_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=96, dtype="float32",
                                        mode="auto", stride=8, max_new_tokens=12,
                                        gen=dict(budget=40, kv_policy="roco", streaming=True)),
    "llama_decoding_roco_stream_fp32": dict(arch="llama", L=1, H=4, Hkv=2, d=128, seq=32, dtype="float32",
                                            mode="decoding", stride=1, max_new_tokens=64,
                                            gen=dict(budget=36, kv_policy="roco", streaming=True)),
    "gqa_mistral_enc_h2o_stream_fp16": dict(arch="mistral", L=1, H=8, Hkv=2, d=128, seq=132, dtype="float16",
                                            mode="encoding", stride=4, max_new_tokens=3,
                                            gen=dict(budget=0.5, kv_policy="h2o_head", streaming=True)),
    # keep_attention=True: state seeded from the dense prefill's attention map (h2o_head_score, easykv.py:173-186)
    "llama_enc_roco_keep_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=128, seq=200, dtype="float32",
                                     mode="encoding", stride=8, max_new_tokens=2,
                                     gen=dict(budget=0.5, kv_policy="roco", keep_attention=True)),
    "gqa_mistral_enc_h2o_keep_fp32": dict(arch="mistral", L=2, H=8, Hkv=2, d=128, seq=164, dtype="float32",
                                          mode="encoding", stride=4, max_new_tokens=2,
                                          gen=dict(budget=0.5, kv_policy="h2o_head", keep_attention=True)),
    "gqa_mistral_enc_h2o_keep_fp16": dict(arch="mistral", L=1, H=8, Hkv=2, d=128, seq=164, dtype="float16",
                                          mode="encoding", stride=4, max_new_tokens=2,
                                          gen=dict(budget=0.5, kv_policy="h2o_head", keep_attention=True)),
    # 16-bit rounding points
    "gqa_mistral_auto_roco_fp16": dict(arch="mistral", L=2, H=8, Hkv=2, d=128, seq=160, dtype="float16",
                                       mode="auto", stride=8, max_new_tokens=24,
                                       gen=dict(budget=64, kv_policy="roco")),
    "llama_decoding_roco_fp16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=32, dtype="float16",
                                     mode="decoding", stride=1, max_new_tokens=64,
                                     gen=dict(budget=36, kv_policy="roco")),
    "gqa_llama_auto_tova_fp16": dict(arch="llama", L=1, H=8, Hkv=2, d=128, seq=96, dtype="float16",
                                     mode="auto", stride=8, max_new_tokens=16,
                                     gen=dict(budget=40, kv_policy="tova")),
    "gqa8_mistral_enc_h2o_fp16": dict(arch="mistral", L=1, H=8, Hkv=1, d=128, seq=144, dtype="float16",
                                      mode="encoding", stride=16, max_new_tokens=2,
                                      gen=dict(budget=0.5, kv_policy="h2o_head")),
    "gqa_llama_enc_tova_fp16": dict(arch="llama", L=1, H=4, Hkv=2, d=128, seq=120, dtype="float16",
                                    mode="encoding", stride=8, max_new_tokens=2,
                                    gen=dict(budget=0.5, kv_policy="tova")),
    "llama_ppl_roco_bf16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=136, dtype="bfloat16",
                                mode="ppl", stride=8, max_new_tokens=0,
                                gen=dict(budget=0.4, kv_policy="roco")),
    "llama_enc_roco_fp16": dict(arch="llama", L=1, H=4, Hkv=4, d=128, seq=128, dtype="float16",
                                mode="encoding", stride=8, max_new_tokens=2,
                                gen=dict(budget=0.5, kv_policy="roco")),
    # head_dim 64 / 96 (the reference is generic in it, llama_patch.py:169-172)
    "llama_auto_roco_d64_fp32": dict(arch="llama", L=2, H=4, Hkv=4, d=64, seq=96, dtype="float32",
                                     mode="auto", stride=8, max_new_tokens=16,
                                     gen=dict(budget=40, kv_policy="roco")),
    "gqa_mistral_enc_h2o_d96_fp16": dict(arch="mistral", L=1, H=8, Hkv=2, d=96, seq=132, dtype="float16",
                                         mode="encoding", stride=4, max_new_tokens=3,
                                         gen=dict(budget=0.5, kv_policy="h2o_head")),
    "gqa_llama_decoding_roco_stream_d64_fp16": dict(arch="llama", L=1, H=8, Hkv=2, d=64, seq=32, dtype="float16",
                                                    mode="decoding", stride=1, max_new_tokens=48,
                                                    gen=dict(budget=36, kv_policy="roco", streaming=True)),
}


def run_trace(c, device="cpu"):
    """One traced run of the unmodified reference on the scaffold (`device`: 'cpu', or 'cuda' on the GPU box)."""
    dtype = getattr(torch, c["dtype"])
    model = scaffold.build(c["arch"], seed=0, dtype=dtype, device=device, L=c["L"], H=c["H"], Hkv=c["Hkv"], d=c["d"],
                           vocab=512, inter=c.get("inter"), max_pos=c.get("max_pos", 4096))
    ids = torch.randint(3, 512, (1, c["seq"]), generator=torch.Generator().manual_seed(1)).to(device)
    gen = dict(temperature=1e-9, top_p=1.0, max_new_tokens=c["max_new_tokens"], **c["gen"])
    ppl = c["mode"] == "ppl"
    torch.manual_seed(RNG_SEED)                      # the global CPU generator the reference's 'random' policy draws from
    return ref_harness.run_reference(model, ids, gen, mode="encoding" if ppl else c["mode"], stride=c["stride"], ppl=ppl)


def trace_arrays(name, c, tr, device="cpu"):
    """(meta, arrays) of one traced run — what a golden file holds."""
    dtype = getattr(torch, c["dtype"])
    arrs = {}
    npdt = np.float32 if dtype in (torch.float32, torch.bfloat16) else np.float16      # bf16 values are exact in fp32
    for l, (k, v) in enumerate(tr.prefill_cache):
        arrs[f"prefill_K_{l}"], arrs[f"prefill_V_{l}"] = k.float().numpy().astype(npdt), v.float().numpy().astype(npdt)
    fmeta = []
    for f, fw in enumerate(tr.forwards):
        fmeta.append(dict(q_len=fw["q_len"], recorded=len(fw["layers"]) > 0 and f > 0,
                          pos0=None if fw["position_ids"] is None else int(fw["position_ids"][0])))
        if f == 0:
            continue
        for l, rec in enumerate(fw["layers"]):
            for key in "qkvo":
                arrs[f"f{f}_l{l}_{key}"] = rec[key].float().numpy().astype(npdt)
    emeta = []
    for e, ev in enumerate(tr.events):
        arrs[f"ev{e}_ids"] = ev["ids"].numpy().astype(np.int64)
        emeta.append(dict(kind=ev["kind"], fwd=ev["fwd"], n_before=ev["n_before"]))
    if tr.seed is not None:
        arrs["seed_S"], arrs["seed_SQ"] = tr.seed[0].numpy().astype(np.float32), tr.seed[1].numpy().astype(np.float32)
    for l, kv in enumerate(tr.final_cache):
        arrs[f"final_K_{l}"], arrs[f"final_V_{l}"] = (kv[0][0].detach().cpu().float().numpy().astype(npdt),
                                                      kv[1][0].detach().cpu().float().numpy().astype(npdt))
    meta = dict(name=name, case=c, rng_seed=RNG_SEED, forwards=fmeta, events=emeta, printed=tr.printed, tokens=tr.tokens,
                result=tr.result if isinstance(tr.result, float) else str(tr.result), device=str(device),
                torch=torch.__version__, reference_commit="a1d71cae3b562d9a709dda3741bd63e46a09ad31")
    if str(device).startswith("cuda"):
        meta["gpu"] = torch.cuda.get_device_name(0)
    return meta, arrs


def run_case(name, c, device="cpu", out_dir=None):
    tr = run_trace(c, device)
    meta, arrs = trace_arrays(name, c, tr, device)
    arrs["meta"] = np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8)
    out_dir = out_dir or OUT
    os.makedirs(out_dir, exist_ok=True)
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrs)
    return tr


SAMPLING_CASES = [  # (vocab, rows, temperature, top_p, logit scale)
    (1000, 3, 0.7, 0.9, 3.0), (4096, 2, 1.0, 0.95, 2.0), (512, 2, 1e-9, 1.0, 3.0), (2000, 2, 0.3, 0.5, 1.0),
    (777, 2, 1.3, 0.0, 2.0), (3000, 2, 1.0, 1.0, 4.0), (32000, 1, 0.8, 0.9, 2.5)]


def sampling_tail():
    """The reference's own `logits_adapter` (easykv/easykv.py:115-134) and the loss it feeds the perplexity with
    (:782, :896-899) on seeded logits -> tests/golden/sampling_tail.npz."""
    ref_main, _, _ = ref_harness.import_reference()
    g = torch.Generator().manual_seed(RNG_SEED)
    arrs, meta = {}, []
    for i, (V, R, T, tp, sc) in enumerate(SAMPLING_CASES):
        x = torch.randn(R, V, generator=g) * sc
        x[:, 5] = x[:, 9]                                  # an exact tie in every row
        x[0, V // 2:V // 2 + 8] = x[0].max()               # a run of equal maxima the nucleus boundary can cut
        final, raw = ref_main.logits_adapter(x.clone(), T, tp)
        t = torch.randint(0, V, (R,), generator=g)
        nll = torch.nn.CrossEntropyLoss(reduction="none")(x, t)
        arrs[f"c{i}_logits"], arrs[f"c{i}_final"], arrs[f"c{i}_raw"] = x.numpy(), final.numpy(), raw.numpy()
        arrs[f"c{i}_targets"], arrs[f"c{i}_nll"] = t.numpy(), nll.numpy()
        meta.append(dict(vocab=V, rows=R, temperature=T, top_p=tp))
    arrs["meta"] = np.frombuffer(json.dumps(dict(cases=meta, torch=torch.__version__,
                                                 reference_commit="a1d71cae3b562d9a709dda3741bd63e46a09ad31")).encode(), dtype=np.uint8)
    np.savez_compressed(os.path.join(OUT, "sampling_tail.npz"), **arrs)
    print(f"sampling_tail: {len(meta)} cases, {os.path.getsize(os.path.join(OUT, 'sampling_tail.npz')) / 1e6:.2f} MB")


if __name__ == "__main__":
    torch.set_num_threads(8)
    only = sys.argv[1:]
    if not only or "sampling_tail" in only:
        sampling_tail()
    for name, c in CASES.items():
        if only and name not in only:
            continue
        tr = run_case(name, c)
        sz = os.path.getsize(os.path.join(OUT, name + ".npz")) / 1e6
        print(f"{name}: {len(tr.forwards)} forwards, {len(tr.events)} eviction events, {sz:.2f} MB | "
              + tr.printed.strip().splitlines()[-1])
