"""Sampling / perplexity tail on the device: `ekv_sample_top_p` and `ekv_token_nll`.

Replaces, per generated token, the reference's `logits_adapter` (easykv/easykv.py:115-134: softmax, two sorts, cumsum, a
boolean-mask assignment that syncs the host, sum, div, gather, a second softmax) and `torch.multinomial` (:258, :509,
:671: two more host syncs inside ATen) with ONE launch; and, per prompt chunk in `ppl` mode, the retention of every
chunk's `[q_len, vocab]` logits for a final CrossEntropyLoss (:826-827, :896-899) with one launch that keeps `q_len`
floats.  CUDA tensors only — like the rest of the package there is no PyTorch fallback.
"""
from __future__ import annotations

import torch

from . import _lib


def _prep(logits):
    if not logits.is_cuda:
        raise RuntimeError("easykv_b200.sampling needs CUDA tensors (no CPU path)")
    x = logits.reshape(-1, logits.shape[-1])
    if x.dtype != torch.float32 or not x.is_contiguous():
        x = x.float().contiguous()                   # the 4.36 model classes return fp32 logits (modeling_llama: logits.float())
    return x


def sample_top_p(logits, temperature, top_p, arith=1, draw=True, want_prob=False, want_raw=False, generator=None):
    """logits `[..., vocab]` -> (token `[rows, 1]` int64 or None, prob or None, raw softmax or None).

    The token is the one `torch.multinomial(prob, 1)` returns from the same generator state: ATen draws one Exp(1)
    variate per element and takes argmax(prob / q); the variates are drawn here with the same call and the division +
    argmax are fused into the kernel."""
    lib = _lib.load()
    x = _prep(logits)
    R, V = x.shape
    need_prob = want_prob or V * 4 > 200 * 1024      # large vocabularies use the prob buffer as the kernel's workspace
    prob = torch.empty_like(x) if need_prob else None
    raw = torch.empty_like(x) if want_raw else None
    q = token = None
    if draw:
        q = torch.empty_like(x).exponential_(1, generator=generator)
        token = torch.empty(R, 1, dtype=torch.int64, device=x.device)
    _lib.check(lib.ekv_sample_top_p(x.data_ptr(), R, V, float(temperature), float(top_p), int(arith),
                                    None if q is None else q.data_ptr(), None if prob is None else prob.data_ptr(),
                                    None if raw is None else raw.data_ptr(), None if token is None else token.data_ptr(),
                                    torch.cuda.current_stream(x.device).cuda_stream))
    shape = logits.shape
    return token, (prob.reshape(shape) if want_prob else None), (raw.reshape(shape) if want_raw else None)


def logits_adapter(logits, temperature, top_p, arith=1):
    """Same contract as easykv/easykv.py:115-134: (top-p renormalised sampling distribution, raw softmax)."""
    _, prob, raw = sample_top_p(logits, temperature, top_p, arith=arith, draw=False, want_prob=True, want_raw=True)
    return prob, raw


def token_nll(logits, targets):
    """Per-row negative log likelihood `[rows]` fp32 of int64 `targets` under fp32 `logits [rows, vocab]`
    (CrossEntropyLoss(reduction='none'), easykv.py:782, :896-899)."""
    lib = _lib.load()
    x = _prep(logits)
    t = targets.reshape(-1).to(device=x.device, dtype=torch.int64).contiguous()
    if t.numel() != x.shape[0]:
        raise ValueError(f"{t.numel()} targets for {x.shape[0]} rows")
    out = torch.empty(x.shape[0], dtype=torch.float32, device=x.device)
    _lib.check(lib.ekv_token_nll(x.data_ptr(), t.data_ptr(), x.shape[0], x.shape[1], out.data_ptr(),
                                 torch.cuda.current_stream(x.device).cuda_stream))
    return out
