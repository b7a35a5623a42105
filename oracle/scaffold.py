"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

A minimal "transformers-4.36-shaped" Llama/Mistral decoder so that the reference's
hot path (`/root/reference/easykv/easykv.py`, `llama_patch.py`, `mistral_patch.py`)
can be executed *unmodified* inside this container, where the installed
transformers (5.5.0) no longer has the attention/cache API the reference binds to
(SURVEY.md §8c).  Only `tests/`, `oracle/gen_golden.py`, `__graft_entry__.smoke()`
and `bench.py`'s reference / cpu_baseline legs may import this module.

What the reference needs from the host model (cited per item):
  * attention modules whose class is literally named `LlamaAttention` /
    `MistralAttention` (reference `easykv/utils.py:29`) exposing `num_heads`,
    `head_dim`, `num_key_value_heads`, `num_key_value_groups`, `hidden_size`,
    `q_proj/k_proj/v_proj/o_proj`, `rotary_emb(x, seq_len=)`, `layer_idx`,
    `attention_dropout`, `config.pretraining_tp` (`easykv/llama_patch.py:143-242`);
  * a cache with `get_usable_length` / `update` (`easykv/llama_patch.py:184-196`);
  * a model `__call__(input_ids, past_key_values, attention_mask, position_ids,
    use_cache, output_attentions)` returning `.logits/.past_key_values/.attentions`
    with legacy `[layer][0|1] -> [1,Hkv,n,d]` caches (`easykv/easykv.py:232-277`).
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
from torch import nn


class DynamicCache:
    """Append-by-concatenation cache with the 4.36 method names."""

    def __init__(self):
        self.key_cache, self.value_cache = [], []

    def get_seq_length(self, layer_idx=0):
        return 0 if len(self.key_cache) <= layer_idx else self.key_cache[layer_idx].shape[-2]

    def get_usable_length(self, new_seq_length, layer_idx=0):
        return self.get_seq_length(layer_idx)

    def update(self, k, v, layer_idx, cache_kwargs=None):
        if len(self.key_cache) <= layer_idx:
            self.key_cache.append(k)
            self.value_cache.append(v)
        else:
            self.key_cache[layer_idx] = torch.cat([self.key_cache[layer_idx], k], dim=-2)
            self.value_cache[layer_idx] = torch.cat([self.value_cache[layer_idx], v], dim=-2)
        return self.key_cache[layer_idx], self.value_cache[layer_idx]

    @classmethod
    def from_legacy_cache(cls, past):
        c = cls()
        if past is not None:
            for l, kv in enumerate(past):
                c.update(kv[0], kv[1], l)
        return c

    def to_legacy_cache(self):
        return tuple((k, v) for k, v in zip(self.key_cache, self.value_cache))


class RotaryEmbedding(nn.Module):
    """cos/sin table `[seq, d]` in the model dtype, fp32 `inv_freq`, regrown on demand."""

    def __init__(self, dim, max_position_embeddings=4096, base=10000.0):
        super().__init__()
        self.dim, self.base = dim, base
        self.max_position_embeddings = max_position_embeddings
        inv_freq = 1.0 / (base ** (torch.arange(0, dim, 2, dtype=torch.float32) / dim))
        self.register_buffer("inv_freq", inv_freq, persistent=False)
        self._set_cos_sin_cache(max_position_embeddings, inv_freq.device, torch.get_default_dtype())

    def _set_cos_sin_cache(self, seq_len, device, dtype):
        self.max_seq_len_cached = seq_len
        t = torch.arange(seq_len, device=device, dtype=torch.float32)
        freqs = torch.outer(t, self.inv_freq.to(device=device, dtype=torch.float32))
        emb = torch.cat((freqs, freqs), dim=-1)
        self.register_buffer("cos_cached", emb.cos().to(dtype), persistent=False)
        self.register_buffer("sin_cached", emb.sin().to(dtype), persistent=False)

    def forward(self, x, seq_len=None):
        if seq_len > self.max_seq_len_cached:
            self._set_cos_sin_cache(seq_len, x.device, x.dtype)
        return self.cos_cached[:seq_len].to(dtype=x.dtype), self.sin_cached[:seq_len].to(dtype=x.dtype)


def _rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


class _Attention(nn.Module):
    """Stock eager attention (what runs before the reference patches `forward`)."""

    def __init__(self, config, layer_idx):
        super().__init__()
        self.config, self.layer_idx = config, layer_idx
        self.hidden_size = config.hidden_size
        self.num_heads = config.num_attention_heads
        self.head_dim = config.head_dim
        self.num_key_value_heads = config.num_key_value_heads
        self.num_key_value_groups = self.num_heads // self.num_key_value_heads
        self.attention_dropout = 0.0
        self.q_proj = nn.Linear(self.hidden_size, self.num_heads * self.head_dim, bias=False)
        self.k_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.v_proj = nn.Linear(self.hidden_size, self.num_key_value_heads * self.head_dim, bias=False)
        self.o_proj = nn.Linear(self.num_heads * self.head_dim, self.hidden_size, bias=False)
        self.rotary_emb = RotaryEmbedding(self.head_dim, config.max_position_embeddings, config.rope_theta)

    def forward(self, hidden_states, attention_mask=None, position_ids=None, past_key_value=None,
                output_attentions=False, use_cache=False, **kw):
        b, q, _ = hidden_states.shape
        qs = self.q_proj(hidden_states).view(b, q, self.num_heads, self.head_dim).transpose(1, 2)
        ks = self.k_proj(hidden_states).view(b, q, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        vs = self.v_proj(hidden_states).view(b, q, self.num_key_value_heads, self.head_dim).transpose(1, 2)
        kv_len = q + (past_key_value.get_usable_length(q, self.layer_idx) if past_key_value is not None else 0)
        cos, sin = self.rotary_emb(vs, seq_len=max(kv_len, int(position_ids.max()) + 1))
        cos, sin = cos[position_ids].unsqueeze(1), sin[position_ids].unsqueeze(1)
        qs, ks = qs * cos + _rotate_half(qs) * sin, ks * cos + _rotate_half(ks) * sin
        if past_key_value is not None:
            ks, vs = past_key_value.update(ks, vs, self.layer_idx)
        g = self.num_key_value_groups
        if g > 1:
            ks = ks[:, :, None].expand(b, self.num_key_value_heads, g, kv_len, self.head_dim).reshape(b, -1, kv_len, self.head_dim)
            vs = vs[:, :, None].expand(b, self.num_key_value_heads, g, kv_len, self.head_dim).reshape(b, -1, kv_len, self.head_dim)
        w = torch.matmul(qs, ks.transpose(2, 3)) / math.sqrt(self.head_dim)
        if attention_mask is not None:
            w = w + attention_mask
        w = torch.softmax(w, dim=-1, dtype=torch.float32).to(qs.dtype)
        o = torch.matmul(w, vs).transpose(1, 2).reshape(b, q, -1)
        return self.o_proj(o), (w if output_attentions else None), past_key_value


class LlamaAttention(_Attention):
    pass


class MistralAttention(_Attention):
    pass


class RMSNorm(nn.Module):
    def __init__(self, n, eps):
        super().__init__()
        self.weight, self.eps = nn.Parameter(torch.ones(n)), eps

    def forward(self, x):
        dt = x.dtype
        x = x.float()
        x = x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + self.eps)
        return self.weight * x.to(dt)


class MLP(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.gate_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.up_proj = nn.Linear(config.hidden_size, config.intermediate_size, bias=False)
        self.down_proj = nn.Linear(config.intermediate_size, config.hidden_size, bias=False)

    def forward(self, x):
        return self.down_proj(nn.functional.silu(self.gate_proj(x)) * self.up_proj(x))


class DecoderLayer(nn.Module):
    def __init__(self, config, layer_idx, attn_cls):
        super().__init__()
        self.self_attn = attn_cls(config, layer_idx)
        self.mlp = MLP(config)
        self.input_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.post_attention_layernorm = RMSNorm(config.hidden_size, config.rms_norm_eps)

    def forward(self, h, attention_mask, position_ids, past_key_value, output_attentions, use_cache):
        a, w, _ = self.self_attn(self.input_layernorm(h), attention_mask=attention_mask,
                                 position_ids=position_ids, past_key_value=past_key_value,
                                 output_attentions=output_attentions, use_cache=use_cache)
        h = h + a
        return h + self.mlp(self.post_attention_layernorm(h)), w


def causal_mask_4d(q_len, past_len, dtype, device):
    """Restatement of HF `_prepare_4d_causal_attention_mask` for an all-ones 2-D mask:
    zeros over the `past_len` cached keys, `finfo.min` strictly above the diagonal of the
    trailing `q_len x q_len` block."""
    m = torch.zeros(q_len, past_len + q_len, dtype=dtype, device=device)
    if q_len > 1:
        tri = torch.full((q_len, q_len), torch.finfo(dtype).min, dtype=dtype, device=device).triu(1)
        m[:, past_len:] = tri
    return m[None, None]


def make_config(arch="llama", L=2, H=4, Hkv=4, d=128, hidden=None, inter=None, vocab=512,
                rope_theta=10000.0, max_pos=4096):
    hidden = hidden or H * d
    return SimpleNamespace(
        architectures=["LlamaForCausalLM" if arch == "llama" else "MistralForCausalLM"],
        num_hidden_layers=L, num_attention_heads=H, num_key_value_heads=Hkv, head_dim=d,
        hidden_size=hidden, intermediate_size=inter or 2 * hidden, vocab_size=vocab,
        rope_theta=rope_theta, max_position_embeddings=max_pos, rms_norm_eps=1e-6,
        pretraining_tp=1, initializer_range=0.02)


class ScaffoldLM(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        attn_cls = LlamaAttention if "llama" in config.architectures[0].lower() else MistralAttention
        self.embed_tokens = nn.Embedding(config.vocab_size, config.hidden_size)
        self.layers = nn.ModuleList([DecoderLayer(config, l, attn_cls) for l in range(config.num_hidden_layers)])
        self.norm = RMSNorm(config.hidden_size, config.rms_norm_eps)
        self.lm_head = nn.Linear(config.hidden_size, config.vocab_size, bias=False)

    @property
    def device(self):
        return self.embed_tokens.weight.device

    @property
    def dtype(self):
        return self.embed_tokens.weight.dtype

    def init_weights(self, seed=0):
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for name, p in self.named_parameters():
                if p.ndim >= 2:
                    p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32) * self.config.initializer_range)
        return self

    def forward(self, input_ids=None, past_key_values=None, attention_mask=None, position_ids=None,
                use_cache=True, output_attentions=False, **kw):
        cache = DynamicCache.from_legacy_cache(past_key_values)
        past = cache.get_seq_length()
        q = input_ids.shape[1]
        if position_ids is None:
            position_ids = torch.arange(past, past + q, device=input_ids.device)[None]
        h = self.embed_tokens(input_ids)
        mask = causal_mask_4d(q, past, h.dtype, h.device)
        atts = []
        for layer in self.layers:
            h, w = layer(h, mask, position_ids, cache if use_cache else None, output_attentions, use_cache)
            atts.append(w)
        logits = self.lm_head(self.norm(h)).float()
        return SimpleNamespace(logits=logits,
                               past_key_values=cache.to_legacy_cache() if use_cache else None,
                               attentions=tuple(atts) if output_attentions else None)


class StubTokenizer:
    eos_token_id = -1

    def decode(self, ids, skip_special_tokens=True):
        return " ".join(str(int(i)) for i in ids)

    def convert_ids_to_tokens(self, ids):
        return [str(int(i)) for i in ids]


def build(arch="llama", seed=0, dtype=torch.float32, device="cpu", **cfg):
    model = ScaffoldLM(make_config(arch=arch, **cfg)).init_weights(seed).to(dtype=dtype, device=device).eval()
    for layer in model.layers:  # tables in model dtype, built once in fp32
        r = layer.self_attn.rotary_emb
        r._set_cos_sin_cache(r.max_seq_len_cached, device, dtype)
    return model
