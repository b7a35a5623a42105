"""Host-side sequencing of the budgeted-cache path: the reference's budget arithmetic and the
per-forward decisions of its four mode loops, expressed as a list of `StepParams` that the CUDA
library executes (one `ekv_step` per forward).  Pure Python integers — no tensors, no device work.

Reference (paths relative to the reference root):
  mode dispatch             easykv/easykv.py:220-227
  budget / idx / r_idx      :385-395 (encoding), :544-556 (encoding_decoding), :773-783 (ppl)
  decoding loop             :257-363      strided loops  :426-500, :587-661, :816-892
  enc-dec decode loop       :670-748      counters       :244-247, :304, :412-418, :460, :469, :708
"""
from __future__ import annotations

from dataclasses import dataclass, asdict

from . import _lib

POLICIES = ("roco", "h2o_head", "tova", "recency", "random", "full")
_ALIASES = {"h2o": "h2o_head"}          # BASELINE.json spells it 'h2o'; the reference's name is 'h2o_head'
_SCORED = ("roco", "h2o_head", "tova")
_POLICY_ENUM = {"roco": _lib.POLICY_ROCO, "h2o_head": _lib.POLICY_H2O, "tova": _lib.POLICY_TOVA,
                "recency": _lib.POLICY_RANGE, "random": _lib.POLICY_RANGE, "full": _lib.POLICY_NONE}


def canonical_policy(name: str) -> str:
    """The reference silently keeps 100 % of the cache for an unknown policy name
    (easykv.py:310-362 has no else branch); here that is an error."""
    name = _ALIASES.get(name, name)
    if name not in POLICIES:
        raise ValueError(f"unknown kv_policy {name!r}; expected one of {POLICIES} (or 'h2o')")
    return name


@dataclass(frozen=True)
class StepParams:
    """Field-for-field the C struct `ekv_step` (include/easykv_b200.h), policy by name.  Immutable (derive variants
    with `dataclasses.replace`): the cache memoises the C struct on the object."""
    policy: str = "full"
    accumulate: bool = False
    evict: int = 0
    score_offset: int = 0
    counter_add: float = 0.0
    c_new0: float = 0.0
    c_new_step: float = 0.0
    k_feasible: int = 0
    protect_last: int = 10
    sink_protect: int = 0
    win_lo: int = 0
    win_recent: int = 0
    tova_head_mean: bool = False
    range_start: int = 0
    raw_colsum: bool = False
    budget_gate: int = 0      # ragged batches: a sequence evicts only when its scored slots after the append exceed this

    def to_c(self, apply=True, arith=0) -> "_lib.Step":
        pol = _POLICY_ENUM[self.policy]
        return _lib.Step(policy=pol, accumulate=int(self.accumulate and self.policy in _SCORED),
                         evict=int(self.evict), apply=int(apply), score_offset=int(self.score_offset),
                         counter_add=float(self.counter_add), c_new0=float(self.c_new0),
                         c_new_step=float(self.c_new_step), k_feasible=int(self.k_feasible),
                         protect_last=int(self.protect_last), sink_protect=int(self.sink_protect),
                         win_lo=int(self.win_lo), win_recent=int(self.win_recent),
                         range_start=int(self.range_start), arith=int(arith),
                         tova_head_mean=int(self.tova_head_mean and self.policy == "tova"),
                         raw_colsum=int(self.raw_colsum), budget_gate=int(self.budget_gate))

    @classmethod
    def from_fields(cls, obj) -> "StepParams":
        """Build from any object with the same attribute names (e.g. the test oracle's Step)."""
        return cls(**{k: getattr(obj, k) for k in asdict(cls()).keys() if hasattr(obj, k)})


@dataclass
class Plan:
    mode: str                 # decoding | encoding | encoding_decoding | ppl | dense
    length: int               # prompt length
    stride: int
    budget: object            # augmented budget (strided modes) / decode budget ('decoding')
    idx: int = 0              # retained cache size after the strided phase
    r_idx: int = 0            # tokens prefilled densely before the strided loop
    recent_window: int = 0
    sink: int = 0

    @property
    def capacity(self) -> int:
        """Physical slots per (sequence, kv head) this plan can ever occupy (excluding tokens
        generated after an 'encoding' prefill, which grow the cache unboundedly in the reference,
        easykv.py:508-526 — the caller adds max_new_tokens there)."""
        if self.mode == "decoding":
            return self.length + int(self.budget) + 1
        if self.mode == "dense":
            return self.length
        return self.idx + self.stride


def required_capacity(n_dense: int, sched) -> int:
    """Physical slots per (sequence, kv head) a run needs: the largest cache length any forward reaches right after
    its append, over the dense prefill (n_dense tokens) and the schedule `sched` (a list of `schedule()` items).  With
    kv_policy='full' (or whenever the schedule never evicts, e.g. the summarisation / ppl baselines the reference
    runs with a budget below the prompt length, easykv.py:459, :850) the cache simply grows to the whole prompt —
    `Plan.capacity` alone (idx + stride) would be too small."""
    n = cap = n_dense
    for _, q_len, st in sched:
        cap = max(cap, n + q_len)
        n += q_len - int(st.evict)
    return cap


def resolve_plan(kv_mode, length, budget, stride, recent_ratio=0.1, temp_length=4) -> Plan:
    if kv_mode == "auto":
        if type(budget) is not int:
            raise AssertionError("mode='auto' needs an integer budget")          # easykv.py:222
        if budget > length:
            kv_mode, budget = "decoding", budget - length
        else:
            kv_mode = "encoding_decoding"
    if kv_mode == "decoding":
        return Plan("decoding", length, stride, budget)
    if kv_mode == "encoding":
        if (type(budget) is float and budget >= 1.0) or (type(budget) is int and budget >= length):
            return Plan("dense", length, stride, budget)                          # easykv.py:372-377
    elif kv_mode == "ppl":
        if budget >= 1.0:
            return Plan("dense", length, stride, budget)                          # easykv.py:759-765
    elif kv_mode == "encoding_decoding":
        if type(budget) is not int or budget > length:
            raise AssertionError("encoding_decoding needs an integer budget <= prompt length")   # :535
        if stride <= 1:
            raise AssertionError("encoding_decoding needs stride > 1 (the reference asserts at easykv.py:666-669)")
    else:
        raise ValueError(f"unknown mode {kv_mode!r}")
    if type(budget) is float:
        aug = int(length * budget) + stride
    else:
        aug = budget + stride
        if kv_mode == "encoding_decoding" and aug >= length:
            aug -= stride
    idx = aug
    while (length - idx) % stride:
        idx -= 1
    if kv_mode == "encoding":
        r_idx = idx - stride
        if r_idx < 1:             # the reference's search range(idx-1, -1, -1) finds nothing: it would prefill 0 tokens
            raise AssertionError("budget too small for this stride (no dense prefix left)")
    else:
        r_idx = idx % stride or stride
        if r_idx >= idx:          # the reference's search range(1, idx) comes up empty
            raise AssertionError("budget too small for this stride")
    return Plan(kv_mode, length, stride, aug, idx, r_idx, int(aug * recent_ratio), temp_length)


def initial_counter(plan: Plan, keep_attention=False):
    """C of the r_idx slots that exist after the dense prefill: a python list, or None
    when the mode keeps no state for the prompt."""
    if plan.mode in ("decoding", "dense"):
        return None
    if keep_attention:
        return [float(plan.idx - i) for i in range(plan.r_idx)]
    return [0.0] * plan.r_idx


def schedule(plan: Plan, policy: str, max_new_tokens: int, keep_attention=False):
    """Yield (kind, q_len, StepParams) for every forward after the dense prefill, in the order the
    reference issues them; kind is 'chunk' or 'decode'."""
    policy = canonical_policy(policy)
    scored = policy in _SCORED
    evicts = policy != "full"
    L, stride = plan.length, plan.stride
    if plan.mode == "encoding_decoding" and policy not in ("random", "recency", "tova", "roco"):
        raise AssertionError(f"kv_policy {policy!r} is not allowed in encoding_decoding / auto (easykv.py:536-537)")
    if plan.mode == "dense":
        for _ in range(max_new_tokens):
            yield "decode", 1, StepParams()
        return
    if plan.mode == "decoding":
        B = int(plan.budget)
        recent = int(B * 0.3)                       # recent_ratio is overridden, easykv.py:308-309
        for t in range(max_new_tokens):
            ev = evicts and (t + 1) > B             # generated tokens in the cache exceed the budget, :303
            yield "decode", 1, StepParams(
                policy=policy, accumulate=scored, evict=int(ev), score_offset=L,
                counter_add=1.0 if ev else 0.0, c_new0=float(B - t) if t <= B else 0.0,
                k_feasible=B - recent, win_recent=recent if policy == "h2o_head" else 0)
        return
    idx, sink, recent = plan.idx, plan.sink, plan.recent_window
    n = plan.r_idx
    for _ in range(plan.r_idx, L, stride):          # one forward per `stride` prompt tokens
        before = n
        over = before + stride > idx                # kv_len > idx, easykv.py:443,459
        ev = stride if (over and evicts) else 0
        yield "chunk", stride, StepParams(
            policy=policy, accumulate=scored and (over or keep_attention),
            evict=ev, counter_add=float(stride),
            c_new0=float(idx - before) if keep_attention else 0.0,
            c_new_step=1.0 if (keep_attention or before >= idx) else 0.0,
            k_feasible=max(plan.budget - recent - sink, stride), sink_protect=sink,
            win_lo=sink, win_recent=recent, tova_head_mean=plan.mode in ("encoding", "ppl"),
            range_start=sink)
        n = before + stride - ev
    if plan.mode == "ppl":
        return
    if plan.mode == "encoding":
        for _ in range(max_new_tokens):             # plain decode, no eviction, easykv.py:508-526
            yield "decode", 1, StepParams()
        return
    B = int(plan.budget)
    recent = int(B * 0.3)                           # easykv.py:709-710
    for _ in range(max_new_tokens):
        yield "decode", 1, StepParams(policy=policy, accumulate=scored, evict=int(evicts), counter_add=1.0,
                                      k_feasible=B - recent, win_recent=recent if policy == "h2o_head" else 0,
                                      range_start=sink)


def random_range_start(plan: Plan, n_state: int) -> int:
    """kv_policy='random': the reference draws `torch.rand(...)` from torch's *CPU* default generator inside its loop
    and evicts at the argmax (easykv.py:353-357 decoding: over the generated slots in the cache; :494-499 / :886-891
    strided: over the state length with the last `stride` entries excluded).  Drawing the same shape at the same point
    reproduces its choice exactly under the same `torch.manual_seed`.  Returns range_start (relative to score_offset)."""
    import torch
    scores = torch.rand(n_state)
    if plan.mode != "decoding":
        scores[-plan.stride:] = -1e9
    return int(torch.topk(scores, k=1, dim=-1)[1][0])
