"""easykv_b200 — B200-native KV-budgeted attention + eviction behind the EasyKV user surface.

    from easykv_b200 import enable_fixed_kv
    enable_fixed_kv(model, tokenizer, mode="auto", stride=64)
    text = model.easykv_generate(input_ids=ids, generation_config=dict(budget=1024, kv_policy="roco"))

The compute path is `libeasykv_b200.so` (hand-written sm_100a CUDA behind the C ABI of
`include/easykv_b200.h`); there is no PyTorch or CPU fallback.
"""
from .easykv import enable_fixed_kv, generate, logits_adapter  # noqa: F401
from .utils import set_dynamicntk_rope_length  # noqa: F401
from .cache import BudgetedKVCache, SteadyDecode  # noqa: F401
from .plan import StepParams, resolve_plan, schedule  # noqa: F401
