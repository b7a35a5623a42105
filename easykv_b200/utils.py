"""Small helpers of the reference's `easykv/utils.py` that belong to the user surface."""
from __future__ import annotations

from .attention import find_attention_modules


def set_dynamicntk_rope_length(model, max_length):
    """easykv/utils.py:53-57: pre-size the cos/sin table of 4.36-style rotary modules (DynamicNTK scaling
    recomputes the base from the table length).  Modules without `_set_cos_sin_cache` (transformers 5.x
    computes RoPE per call from position ids) need nothing."""
    device = next(model.parameters()).device
    for m in find_attention_modules(model):
        rope = getattr(m, "rotary_emb", None)
        if rope is not None and hasattr(rope, "_set_cos_sin_cache"):
            rope._set_cos_sin_cache(max_length, device=device, dtype=rope.inv_freq.dtype)
    print(f"DynamicNTKRoPE max length reset to {max_length}")
