"""The operator seam: a per-attention-module `forward` replacement that owns RoPE and the budgeted cache
and calls the CUDA library — the counterpart of the reference's `llama_forward` / `mistral_forward`
(easykv/llama_patch.py:125-248, easykv/mistral_patch.py:90-186) and of its patcher
`modify_method_of_instance` (easykv/utils.py:5-51).

Differences from the reference, by design:
  * modules are matched by capability (`q_proj, k_proj, v_proj, o_proj`), not by the class *name*
    `LlamaAttention` / `MistralAttention` (utils.py:29), so the same seam binds to transformers-4.36-shaped
    modules (`rotary_emb(x, seq_len=)`, 3-tuple return) and to the installed transformers 5.x modules
    (`position_embeddings=(cos, sin)`, 2-tuple return);
  * the probability matrix never leaves the kernel: `attn_weights` is always None.  The statistics the
    reference derives from it (easykv.py:271-300) are accumulated inside `ekv_attend_evict`;
  * the cache is the session's `BudgetedKVCache`; `past_key_value` is ignored and returned untouched.

Projections and o_proj stay in PyTorch; RoPE of q / new k and the head-major re-layout
are one `ekv_rope_qk` launch.
"""
from __future__ import annotations

import types

import torch

_PROJ = ("q_proj", "k_proj", "v_proj", "o_proj")


def find_attention_modules(model):
    """Attention modules in layer order (the reference walks the object graph by class name, utils.py:5-51)."""
    mods = [m for m in model.modules() if all(hasattr(m, p) for p in _PROJ)]
    if not mods:
        raise ValueError("no attention module exposing q_proj/k_proj/v_proj/o_proj found")
    if all(getattr(m, "layer_idx", None) is not None for m in mods):
        mods.sort(key=lambda m: m.layer_idx)
    return mods


def geometry(module, config=None):
    """(H, Hkv, d) of one attention module."""
    d = getattr(module, "head_dim", None) or getattr(config, "head_dim", None)
    if d is None:
        d = config.hidden_size // config.num_attention_heads
    return module.q_proj.out_features // d, module.k_proj.out_features // d, int(d)


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q, k, cos, sin):
    """easykv/llama_patch.py:47-72 — elementwise in the model dtype.  cos/sin `[B, q_len, d]`."""
    cos, sin = cos.unsqueeze(1), sin.unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


def budgeted_attention_forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                               output_attentions=False, use_cache=False, position_embeddings=None, **kwargs):
    sess = self._ekv_session
    l = self._ekv_layer
    H, Hkv, d = self._ekv_geometry
    b, ql, _ = hidden_states.shape
    q_in, k_in, v_in = self.q_proj(hidden_states), self.k_proj(hidden_states), self.v_proj(hidden_states)
    if sess.streaming:
        # streaming variant (llama_patch.py:251-379): un-rotated keys in the cache, every key rotated at its index in
        # the arrival-ordered cache and the queries at n_before .. n_before + q_len - 1, on every forward
        cos, sin = sess.stream_table(self, v_in, sess.cache.n[l] + ql)
        out, victims = sess.cache.step_stream(l, sess.step, q_in, k_in, v_in, cos, sin)
        sess.record(l, victims)
        attn_output = self.o_proj(out.transpose(1, 2).reshape(b, ql, H * d))
        if position_embeddings is not None or sess.two_tuple:
            return attn_output, None
        return attn_output, None, past_key_value
    if position_embeddings is not None:                       # transformers >= 4.48: the model computed them, per token
        cos, sin = position_embeddings
        pos = None
        if cos.shape[0] != b:
            cos, sin = cos.expand(b, -1, -1), sin.expand(b, -1, -1)
    else:
        # the table is sized by the largest *position id*, not by the cache length, so positions stay
        # valid after evictions (llama_patch.py:187-189); the session knows it without a device sync
        cos, sin = self.rotary_emb(v_in, seq_len=sess.max_position + 1)
        pos = position_ids if position_ids is not None else sess.position_ids(hidden_states.device)
        if pos.shape[0] != b:
            pos = pos.expand(b, -1)
    # RoPE at the explicit positions + head-major layout in one launch (ekv_rope_qk)
    q, k, v = sess.cache.rope_qkv(q_in, k_in, v_in, cos, sin, pos)
    out, victims = sess.cache.step(l, sess.step, q, k, v)
    sess.record(l, victims)
    attn_output = self.o_proj(out.transpose(1, 2).reshape(b, ql, H * d))
    if position_embeddings is not None or sess.two_tuple:
        return attn_output, None
    return attn_output, None, past_key_value


class patched_attention:
    """Context manager: bind `budgeted_attention_forward` on every attention instance of `model` for the
    duration of one `generate` call (the reference patches lazily inside generate, easykv.py:253-256, and
    leaves the patch in place; here it is removed again)."""

    def __init__(self, model, session):
        self.mods = find_attention_modules(model)
        self.session = session

    def __enter__(self):
        cfg = getattr(self.session, "config", None)
        for i, m in enumerate(self.mods):
            m._ekv_session, m._ekv_layer, m._ekv_geometry = self.session, i, geometry(m, cfg)
            m.forward = types.MethodType(budgeted_attention_forward, m)
        return self.mods

    def __exit__(self, *exc):
        for m in self.mods:
            m.__dict__.pop("forward", None)
            for a in ("_ekv_session", "_ekv_layer", "_ekv_geometry"):
                m.__dict__.pop(a, None)
        return False
