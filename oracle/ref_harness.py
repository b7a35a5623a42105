"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Runs ONLY in the build container.

Executes the reference (`/root/reference/easykv`, imported unmodified) against the
4.36-shaped scaffold (`oracle/scaffold.py`) and records everything the parity tests
need (SURVEY.md §8c "Shims needed"):

  * spies on `truncate_kv_cache_silo / _liso / truncate_kv_cache`
    (reference `easykv/easykv.py:56-82,105-112`) — the only place victim ids surface;
  * a spy on the rotary helper + cache append to capture, per forward and per layer,
    the post-RoPE `q`, the new `k`, `v` (reference `easykv/llama_patch.py:190-196`)
    and the attention output that enters `o_proj` (`:230-242`);
  * `attn_device` forced to the model's device (the reference hard-codes 'cuda',
    `easykv/easykv.py:254`), `Tensor.to('cuda')` mapped to a no-op on CPU for
    `h2o_head_score` (`:182`);
  * `torch.multinomial` replaced by argmax so token choice is deterministic
    (every reference script runs `temperature=1e-9`, e.g. `test_decoding.py:41`).

`/root/reference` does not exist on the GPU box.  `tools/vendor_ref.sh` (also run by `__graft_entry__.build()`) copies the
reference's `easykv/*.py` UNMODIFIED into `oracle/_ref/` (git-ignored, travels with gpurun) so that the same harness can
run the reference on the B200 itself — fp16 / bf16 on CUDA, the arithmetic the product's default (`arith=1`) reproduces.
Importers: `tests/`, `oracle/gen_golden*.py`, `bench.py --impl reference` / its `gpu_reference` leg.  Never the product.
"""
from __future__ import annotations

import contextlib
import io
import sys
import time

import torch

import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_CANDIDATES = ("/root/reference", os.path.join(_HERE, "_ref"))


def reference_root():
    """Where the unmodified reference package lives: the mounted checkout (build container) or the vendored copy."""
    for r in REF_CANDIDATES:
        if os.path.isfile(os.path.join(r, "easykv", "easykv.py")):
            return r
    return None


def import_reference():
    root = reference_root()
    if root is None:
        raise RuntimeError("the reference is neither mounted at /root/reference nor vendored into oracle/_ref "
                           "(run tools/vendor_ref.sh in the build container)")
    if root not in sys.path:
        sys.path.insert(0, root)
    import easykv  # noqa: F401  (the reference package)
    import easykv.easykv as ref_main
    import easykv.llama_patch as ref_llama
    import easykv.mistral_patch as ref_mistral
    return ref_main, ref_llama, ref_mistral


class Trace:
    """forwards: list of dict(q_len, layers=[dict(q,k,v,o)]); events: list of eviction events."""

    def __init__(self):
        self.forwards, self.events, self.final_cache, self.printed, self.result = [], [], None, "", None
        self.tokens = []
        self.prefill_cache = None
        self.seed = None          # keep_attention: (S, SQ) returned by h2o_head_score, [L, Hkv, idx + stride]


@contextlib.contextmanager
def _patched(model, trace: Trace, record_tensors=True):
    ref_main, ref_llama, ref_mistral = import_reference()
    from oracle import scaffold
    saved = {}

    def save(obj, name):
        saved[(obj, name)] = getattr(obj, name)

    # (1) eviction spies -------------------------------------------------------------
    for name in ("truncate_kv_cache_silo", "truncate_kv_cache_liso"):
        save(ref_main, name)
        orig = saved[(ref_main, name)]

        def spy(kv_cache, eviction_ids, _orig=orig, _name=name):
            ids = torch.as_tensor(eviction_ids).detach().cpu().long().clone()
            trace.events.append(dict(kind=_name.rsplit("_", 1)[1], fwd=len(trace.forwards) - 1, ids=ids,
                                     n_before=int(kv_cache[0][0].shape[2])))
            out = _orig(kv_cache, eviction_ids)
            trace.final_cache = out
            return out
        setattr(ref_main, name, spy)
    save(ref_main, "truncate_kv_cache")
    orig_range = saved[(ref_main, "truncate_kv_cache")]

    def spy_range(kv_cache, start, end):
        trace.events.append(dict(kind="range", fwd=len(trace.forwards) - 1, ids=torch.tensor([start, end]),
                                 n_before=int(kv_cache[0][0].shape[2])))
        out = orig_range(kv_cache, start, end)
        trace.final_cache = out
        return out
    ref_main.truncate_kv_cache = spy_range
    save(ref_main, "h2o_head_score")
    orig_seed = saved[(ref_main, "h2o_head_score")]

    def seed_spy(attention_map, device, stride, budget, num_layers, num_heads, empty=False):
        out = orig_seed(attention_map, device, stride, budget, num_layers, num_heads, empty=empty)
        if not empty:
            trace.seed = (out[0].detach().cpu().clone(), out[1].detach().cpu().clone())
        return out
    ref_main.h2o_head_score = seed_spy

    # (2) device shims ---------------------------------------------------------------
    for fname in ("llama_forward", "llama_forward_stream", "mistral_forward", "mistral_forward_stream"):
        save(ref_main, fname)
        orig_f = saved[(ref_main, fname)]

        def fwd(self, *a, _orig=orig_f, **k):
            k["attn_device"] = self.q_proj.weight.device
            return _orig(self, *a, **k)
        setattr(ref_main, fname, fwd)
    save(torch.Tensor, "to")
    orig_to = torch.Tensor.to

    def to_shim(self, *a, **k):
        if a and isinstance(a[0], str) and a[0] == "cuda" and not torch.cuda.is_available():
            return self
        return orig_to(self, *a, **k)
    torch.Tensor.to = to_shim

    # (3) per-layer q/k/v/o capture ----------------------------------------------------
    cur = {}
    for mod in (ref_llama, ref_mistral):
        save(mod, "apply_rotary_pos_emb")
        orig_rope = saved[(mod, "apply_rotary_pos_emb")]

        def rope_spy(q, k, cos, sin, position_ids, *a, _orig=orig_rope, **kw):
            qe, ke = _orig(q, k, cos, sin, position_ids, *a, **kw)
            cur["q"] = qe
            return qe, ke
        mod.apply_rotary_pos_emb = rope_spy
        # streaming variant: q and k are rotated by two separate calls (q first, llama_patch.py:326-327); the replay
        # needs the UN-rotated q (the keys in the cache are un-rotated too)
        save(mod, "apply_rotary_pos_emb_sep")
        orig_sep = saved[(mod, "apply_rotary_pos_emb_sep")]

        def sep_spy(x, cos, sin, position_ids, *a, _orig=orig_sep, **kw):
            if "q" not in cur:
                cur["q"] = x
            return _orig(x, cos, sin, position_ids, *a, **kw)
        mod.apply_rotary_pos_emb_sep = sep_spy
    save(scaffold.DynamicCache, "update")
    orig_update = scaffold.DynamicCache.update

    def update_spy(self, k, v, layer_idx, cache_kwargs=None):
        if cache_kwargs is not None:  # only calls coming from the reference's patched forward
            cur["k"], cur["v"], cur["layer"] = k, v, layer_idx
        return orig_update(self, k, v, layer_idx, cache_kwargs)
    scaffold.DynamicCache.update = update_spy
    hooks = []
    for l, layer in enumerate(model.layers):
        def pre(mod, args, _l=l):
            if cur.get("layer") == _l and "q" in cur and record_tensors:
                trace.forwards[-1]["layers"].append(dict(
                    q=cur["q"][0].detach().cpu().clone(), k=cur["k"][0].detach().cpu().clone(),
                    v=cur["v"][0].detach().cpu().clone(), o=args[0][0].detach().cpu().clone()))
            cur.clear()
        hooks.append(layer.self_attn.o_proj.register_forward_pre_hook(pre))
    save(scaffold.ScaffoldLM, "forward")
    orig_model_fwd = scaffold.ScaffoldLM.forward

    def model_fwd(self, *a, **k):
        ids = k.get("input_ids", a[0] if a else None)
        pos = k.get("position_ids")
        trace.forwards.append(dict(q_len=int(ids.shape[1]), layers=[], t_begin=time.perf_counter(),
                                   input_ids=ids[0].detach().cpu().clone(),
                                   position_ids=None if pos is None else pos[0].detach().cpu().clone()))
        out = orig_model_fwd(self, *a, **k)
        if out.past_key_values is not None:
            trace.final_cache = out.past_key_values
            if len(trace.forwards) == 1:  # the dense prefill: keep its K/V as the replay's starting cache
                trace.prefill_cache = [(k[0].detach().cpu().clone(), v[0].detach().cpu().clone())
                                       for k, v in out.past_key_values]
        return out
    scaffold.ScaffoldLM.forward = model_fwd

    # (4) deterministic token choice ---------------------------------------------------
    save(torch, "multinomial")

    def greedy(prob, num_samples=1, **kw):
        t = prob.argmax(dim=-1, keepdim=True)
        trace.tokens.append(int(t[0, 0]))
        return t
    torch.multinomial = greedy
    try:
        yield ref_main
    finally:
        for (obj, name), val in saved.items():
            setattr(obj, name, val)
        for h in hooks:
            h.remove()
        # the reference leaves its patched forward bound on the instances; unbind
        for layer in model.layers:
            layer.self_attn.__dict__.pop("forward", None)


def run_reference(model, input_ids, generation_config, mode, stride=1, ppl=False, record_tensors=True) -> Trace:
    """`enable_fixed_kv(model, tok, mode, stride)` + `easykv_generate` / `easykv_ppl`, traced."""
    from oracle import scaffold
    trace = Trace()
    with _patched(model, trace, record_tensors) as ref_main:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            ref_main.enable_fixed_kv(model, scaffold.StubTokenizer(), mode=mode, stride=stride)
            fn = model.easykv_ppl if ppl else model.easykv_generate
            trace.result = fn(input_ids=input_ids, generation_config=dict(generation_config))
        trace.printed = buf.getvalue()
    return trace
