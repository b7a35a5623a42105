"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

CPU restatement (torch ops on CPU tensors, logical slot order) of the reference's
KV-budgeted attention + eviction hot path.  It is the checker for the CUDA path:
only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
legs may import it.  The product (`easykv_b200/`) never does.

PARITY PIN: the reference ships no tests/golden vectors for this path (SURVEY.md §4,
§8c: "parity unpinned" by the reference's own tests).  This restatement is therefore
pinned against outputs of the reference itself, executed unmodified in the build
container by `oracle/ref_harness.py` and frozen in `tests/golden/*.npz`
(`oracle/gen_golden.py`); `tests/test_oracle_vs_reference.py` replays those traces
through this file and requires identical victim ids.  The only model-independent
known answers the reference publishes (retained-cache arithmetic, reference
`README.md:153,211,314`) are checked in `tests/test_budget.py`.

Each function cites the reference lines it restates (paths relative to /root/reference).

Tie order.  `torch.topk` leaves the order of equal keys unspecified; this
restatement *defines* it as (value ascending, NaN last, slot index ascending) by using a
stable sort, and `decision_margin()` reports when a reference decision was tie-ambiguous.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch

POLICIES = ("roco", "h2o_head", "tova", "recency", "random", "full")


# ----------------------------------------------------------------------------------------
# A.1 budget arithmetic                                         easykv/easykv.py:220-227,
#                                                     :385-395, :544-556, :773-783
# ----------------------------------------------------------------------------------------
@dataclass
class Plan:
    mode: str            # resolved mode: decoding | encoding | encoding_decoding | ppl | dense
    length: int          # prompt length
    stride: int
    budget: object       # augmented budget (int) or the decode budget in 'decoding'
    idx: int = 0         # steady-state retained cache size after the strided phase
    r_idx: int = 0       # tokens prefilled densely before the strided loop
    recent_window: int = 0
    sink: int = 0


def resolve_plan(kv_mode, length, budget, stride, recent_ratio=0.1, temp_length=4) -> Plan:
    if kv_mode == "auto":                                   # easykv.py:220-227
        assert type(budget) == int
        if budget > length:
            kv_mode, budget = "decoding", budget - length
        else:
            kv_mode = "encoding_decoding"
    if kv_mode == "decoding":
        return Plan("decoding", length, stride, budget)
    if kv_mode == "encoding":                               # easykv.py:372
        if (type(budget) == float and budget >= 1.0) or (type(budget) == int and budget >= length):
            return Plan("dense", length, stride, budget)
    if kv_mode == "ppl" and budget >= 1.0:                  # easykv.py:759
        return Plan("dense", length, stride, budget)
    if kv_mode == "encoding_decoding":
        assert type(budget) == int and budget <= length     # easykv.py:535
    if type(budget) == float:                               # easykv.py:385-388
        budget = int(length * budget) + stride
    else:
        budget = budget + stride
        if kv_mode == "encoding_decoding" and budget >= length:   # easykv.py:548
            budget -= stride
    idx = next(i for i in range(budget, -1, -1) if (length - i) % stride == 0)      # :389-390
    if kv_mode == "encoding":
        r_idx = next(r for r in range(idx - 1, -1, -1) if (idx - r) % stride == 0)  # :391-392
    else:
        r_idx = next(r for r in range(1, idx) if (idx - r) % stride == 0)           # :551-552
    return Plan(kv_mode, length, stride, budget, idx, r_idx, int(budget * recent_ratio), temp_length)


# ----------------------------------------------------------------------------------------
# attention core with the reference's rounding points      easykv/llama_patch.py:198-222
# ----------------------------------------------------------------------------------------
def chunk_mask(q_len, n, dtype):
    """HF 4-D mask for an all-ones 2-D mask: zeros over the cache, causal inside the chunk."""
    m = torch.zeros(q_len, n, dtype=dtype)
    if q_len > 1:
        m[:, n - q_len:] = torch.full((q_len, q_len), torch.finfo(dtype).min, dtype=dtype).triu(1)
    return m


def attend(q, K, V, scale_mul=False):
    """q [H,ql,d]; K,V [Hkv,n,d] in logical order, the ql new keys already appended.
    Returns (out [H,ql,d], probs [H,ql,n]) in q.dtype.  `scale_mul` mimics ATen's CUDA
    scalar-division kernel (multiplication by the fp32 reciprocal) instead of CPU division."""
    H, ql, d = q.shape
    Hkv, n, _ = K.shape
    g = H // Hkv
    Kr = K[:, None].expand(Hkv, g, n, d).reshape(1, H, n, d)        # repeat_kv, llama_patch.py:19-29
    Vr = V[:, None].expand(Hkv, g, n, d).reshape(1, H, n, d)
    w = torch.matmul(q[None], Kr.transpose(2, 3))                    # :201
    if scale_mul and w.dtype != torch.float32:
        w = (w.float() * (1.0 / torch.tensor(math.sqrt(d), dtype=torch.float32))).to(w.dtype)
    else:
        w = w / math.sqrt(d)                                         # :202
    w = w + chunk_mask(ql, n, w.dtype).to(w.device)[None, None]      # :215
    p = torch.softmax(w, dim=-1, dtype=torch.float32).to(q.dtype)    # :218-219
    o = torch.matmul(p, Vr)                                          # :222
    return o[0], p[0]


def rope(x, cos, sin, pos):
    """apply_rotary_pos_emb_sep, easykv/llama_patch.py:74-98: x [heads, rows, d] at positions pos [rows], evaluated
    in the model dtype."""
    c, s_ = cos[pos][None], sin[pos][None]
    h = x.shape[-1] // 2
    rot = torch.cat((-x[..., h:], x[..., :h]), dim=-1)
    return (x * c) + (rot * s_)


def fold_gqa(p, Hkv):
    """process_for_mqa_gqa, easykv/easykv.py:188-196: mean over the g query heads of a KV head,
    evaluated in the model dtype."""
    H, ql, n = p.shape
    return p.reshape(1, Hkv, H // Hkv, ql, n).mean(dim=2)[0]


# ----------------------------------------------------------------------------------------
# per-forward step parameters (what the mode loops of easykv.py decide for one forward)
# ----------------------------------------------------------------------------------------
@dataclass
class Step:
    policy: str = "full"
    accumulate: bool = False
    evict: int = 0              # victims per (layer, kv head); 0 = none
    score_offset: int = 0       # 'decoding': the first P slots (the prompt) carry no state
    counter_add: float = 0.0    # added to C of every scored slot before select
    c_new0: float = 0.0         # C of the i-th new slot = c_new0 - i * c_new_step (before counter_add)
    c_new_step: float = 0.0
    k_feasible: int = 0         # roco: size of the low-std candidate set
    protect_last: int = 10      # roco: std[-10:] = 1e9
    sink_protect: int = 0       # roco strided: std[:sink] = 1e9
    win_lo: int = 0             # h2o_head / tova: candidates are state[win_lo : n_s - win_recent]
    win_recent: int = 0
    tova_head_mean: bool = False
    range_start: int = 0        # recency / random: evict logical [range_start, range_start + evict)
    raw_colsum: bool = False    # keep_attention seeding: add fp32 column sums, no per-forward rounding (:173-186)


# ----------------------------------------------------------------------------------------
# accumulate                                   easykv/easykv.py:288-300,443-457,603-618,693-707
# ----------------------------------------------------------------------------------------
def accumulate(st: Step, S, SQ, pf):
    """pf: folded probabilities [Hkv, ql, n] (model dtype).  S,SQ: fp32 [Hkv, n_s], n_s = n - P."""
    P = st.score_offset
    if st.raw_colsum and st.policy in ("h2o_head", "roco"):  # h2o_head_score, :183-184, one chunk of the dense map
        S += pf[:, :, P:].float().sum(dim=1)
        SQ += (pf[:, :, P:] ** 2).float().sum(dim=1)
    elif st.policy == "h2o_head":
        S += pf[:, :, P:].sum(dim=1) if pf.shape[1] > 1 else pf[:, 0, P:]
    elif st.policy == "roco":
        if pf.shape[1] > 1:                                  # strided: row sums in the model dtype
            S += pf[:, :, P:].sum(dim=1)
            SQ += (pf[:, :, P:] ** 2).sum(dim=1)
        else:
            S += pf[:, 0, P:]
            SQ += pf[:, 0, P:] ** 2
    elif st.policy == "tova":
        last = pf[:, -1, P:]
        if st.tova_head_mean:                                # easykv.py:456
            last = last.mean(dim=0).unsqueeze(0).repeat(pf.shape[0], 1)
        S.copy_(last)


# ----------------------------------------------------------------------------------------
# select                       easykv/easykv.py:310-347,462-493,623-654,711-742 (+ A.5 order)
# ----------------------------------------------------------------------------------------
def _smallest(values, k):
    """Indices of the k smallest along the last dim in (value asc, NaN last, index asc) order."""
    return torch.sort(values, dim=-1, stable=True)[1][..., :k]


def select(st: Step, S, SQ, C):
    """Victim ids relative to the scored region, [Hkv, evict] (int64).  Call after counter_add."""
    n_s = S.shape[-1]
    v = st.evict
    if st.policy == "roco":
        std = torch.sqrt(SQ / C - (S / C) ** 2)              # :320 / :471
        std[:, n_s - st.protect_last:] = 1e9                 # :321
        if st.sink_protect:
            std[:, :st.sink_protect] = 1e9                   # :473
        feasible = _smallest(std, st.k_feasible)             # :322  topk(largest=False)
        mean = S.gather(-1, feasible) / C.gather(-1, feasible)
        pick = _smallest(mean, v)                            # :323 argmin  / :475 topk
        return feasible.gather(-1, pick)
    if st.policy in ("h2o_head", "tova"):
        win = S[:, st.win_lo: n_s - st.win_recent]           # :311,335,463,485
        return _smallest(win, v) + st.win_lo
    if st.policy in ("recency", "random"):
        return torch.arange(st.range_start, st.range_start + v).repeat(S.shape[0], 1)
    raise ValueError(st.policy)


def decision_margin(st: Step, S, SQ, C):
    """Smallest gap that decided this step's victims (0.0 => the reference's own choice was
    tie-ambiguous).  Returns (margin_at_feasible_cut, margin_at_victim_cut)."""
    n_s = S.shape[-1]
    if st.policy == "roco":
        std = torch.sqrt(SQ / C - (S / C) ** 2)
        std[:, n_s - st.protect_last:] = 1e9
        if st.sink_protect:
            std[:, :st.sink_protect] = 1e9
        sv, si = torch.sort(std, dim=-1, stable=True)
        k = st.k_feasible
        m1 = (sv[:, k] - sv[:, k - 1]).abs().min().item() if k < n_s else float("inf")
        mean = S.gather(-1, si[:, :k]) / C.gather(-1, si[:, :k])
        mv = torch.sort(mean, dim=-1, stable=True)[0]
        m2 = (mv[:, st.evict] - mv[:, st.evict - 1]).abs().min().item() if st.evict < k else float("inf")
        return m1, m2
    if st.policy in ("h2o_head", "tova"):
        win = S[:, st.win_lo: n_s - st.win_recent]
        mv = torch.sort(win, dim=-1, stable=True)[0]
        return float("inf"), (mv[:, st.evict] - mv[:, st.evict - 1]).abs().min().item()
    return float("inf"), float("inf")


# ----------------------------------------------------------------------------------------
# compaction                                            easykv/easykv.py:56-82,105-112 (K/V)
#                                     :315-318,328-333,465-469,478-483 (state) — order-preserving
# ----------------------------------------------------------------------------------------
def keep_mask(n, ids):
    m = torch.ones(ids.shape[0], n, dtype=torch.bool)
    m.scatter_(1, ids, False)
    return m


def compact_rows(x, mask):
    """x [Hkv, n, ...]; drop the rows where mask is False, per head, preserving order."""
    Hkv = x.shape[0]
    return x[mask].view(Hkv, -1, *x.shape[2:])


# ----------------------------------------------------------------------------------------
# one layer's budgeted cache in the reference's logical order
# ----------------------------------------------------------------------------------------
@dataclass
class LayerOracle:
    Hkv: int
    d: int
    dtype: torch.dtype
    K: torch.Tensor = None       # [Hkv, n, d]
    V: torch.Tensor = None
    S: torch.Tensor = None       # [Hkv, n_s] fp32 — only the slots that exist (n - score_offset)
    SQ: torch.Tensor = None
    C: torch.Tensor = None
    last_margin: tuple = field(default=(float("inf"), float("inf")))

    def __post_init__(self):
        z = lambda *s, dt=torch.float32: torch.zeros(*s, dtype=dt)
        self.K, self.V = z(self.Hkv, 0, self.d, dt=self.dtype), z(self.Hkv, 0, self.d, dt=self.dtype)
        self.S, self.SQ, self.C = z(self.Hkv, 0), z(self.Hkv, 0), z(self.Hkv, 0)

    def load_prefill(self, K, V, n_scored=None, C_init=None):
        """Start from a densely prefilled cache.  State covers the last `n_scored` slots."""
        self.K, self.V = K.clone(), V.clone()
        n_s = K.shape[1] if n_scored is None else n_scored
        self.S, self.SQ = torch.zeros(self.Hkv, n_s), torch.zeros(self.Hkv, n_s)
        self.C = torch.zeros(self.Hkv, n_s) if C_init is None else C_init.clone().float().expand(self.Hkv, n_s).clone()

    def forward(self, st: Step, q, k_new, v_new, scale_mul=False, force=None, stream_table=None):
        """q [H,ql,d]; k_new, v_new [Hkv,ql,d].  Returns (out [H,ql,d], victims [Hkv,evict] or None).
        Victim ids index the cache after the append and before the deletion.
        `force` ([Hkv,evict] cache-relative ids): delete these instead of the selected ones.
        `stream_table` = (cos, sin) [rows, d] in the model dtype: the streaming variant
        (llama_forward_stream, easykv/llama_patch.py:251-379) — q and k_new arrive UN-rotated, the cache holds
        un-rotated keys, and every forward rotates all keys at their cache-relative positions 0..n-1 and the
        queries at n_before..n-1 (:310-327)."""
        ql = q.shape[1]
        self.K = torch.cat([self.K, k_new], dim=1)           # DynamicCache.update, llama_patch.py:195
        self.V = torch.cat([self.V, v_new], dim=1)
        new_c = st.c_new0 - st.c_new_step * torch.arange(ql, dtype=torch.float32)
        self.S = torch.cat([self.S, torch.zeros(self.Hkv, ql)], dim=1)
        self.SQ = torch.cat([self.SQ, torch.zeros(self.Hkv, ql)], dim=1)
        self.C = torch.cat([self.C, new_c.repeat(self.Hkv, 1)], dim=1)
        if stream_table is not None:
            n = self.K.shape[1]
            cos, sin = stream_table
            kpos = torch.arange(n)
            out, p = attend(rope(q, cos, sin, kpos[n - ql:]), rope(self.K, cos, sin, kpos), self.V, scale_mul)
        else:
            out, p = attend(q, self.K, self.V, scale_mul)
        if st.accumulate and st.policy in ("roco", "h2o_head", "tova"):
            accumulate(st, self.S, self.SQ, fold_gqa(p, self.Hkv))
        if not st.evict:
            return out, None
        self.C += st.counter_add
        self.last_margin = decision_margin(st, self.S, self.SQ, self.C)
        ids = select(st, self.S, self.SQ, self.C) + st.score_offset
        dele = ids if force is None else force.long()
        sm = keep_mask(self.S.shape[1], dele - st.score_offset)
        self.S, self.SQ, self.C = (compact_rows(x, sm) for x in (self.S, self.SQ, self.C))
        km = keep_mask(self.K.shape[1], dele)
        self.K, self.V = compact_rows(self.K, km), compact_rows(self.V, km)
        return out, ids


# ----------------------------------------------------------------------------------------
# the mode loops of easykv.py as a schedule of Steps
# ----------------------------------------------------------------------------------------
def schedule(plan: Plan, policy, max_new_tokens, keep_attention=False):
    """Yields (kind, q_len, Step) for every forward after the dense prefill, in the order the
    reference issues them.  kind in {'chunk', 'decode'}.  Restates the control flow of
    easykv.py:257-363 (decoding), :426-526 (encoding), :587-748 (encoding_decoding),
    :816-892 (ppl).  `n` tracks the cache length, `t` the number of generated tokens."""
    L, stride = plan.length, plan.stride
    scored = policy in ("roco", "h2o_head", "tova")
    if plan.mode == "dense":
        for _ in range(max_new_tokens):
            yield "decode", 1, Step()
        return
    if plan.mode == "decoding":
        B = plan.budget
        rw = int(B * 0.3)                                                   # :308-309
        for t in range(max_new_tokens):          # forward t appends generated token t
            evict = (t + 1) > B and policy != "full"                        # :303
            st = Step(policy=policy, accumulate=scored, evict=1 if evict else 0, score_offset=L,
                      counter_add=1.0 if evict else 0.0, c_new0=float(B - t) if t <= B else 0.0,
                      k_feasible=B - rw, win_lo=0, win_recent=rw if policy == "h2o_head" else 0,
                      range_start=0)
            if policy not in POLICIES or policy == "full":
                st.evict = 0
            yield "decode", 1, st
        return
    # strided phase, shared by encoding / encoding_decoding / ppl
    idx, sink, rw = plan.idx, plan.sink, plan.recent_window
    n = plan.r_idx
    for _ in range(plan.r_idx, L, stride):
        n += stride
        over = n > idx
        st = Step(policy=policy, accumulate=scored and (over or keep_attention),      # :443
                  evict=stride if (over and policy in POLICIES and policy != "full") else 0,   # :459
                  counter_add=float(stride), c_new0=float(idx - (n - stride)) if keep_attention else 0.0,
                  c_new_step=1.0 if (keep_attention or n - stride >= idx) else 0.0, k_feasible=max(plan.budget - rw - sink, stride),    # :474
                  sink_protect=sink, win_lo=sink, win_recent=rw,
                  tova_head_mean=plan.mode in ("encoding", "ppl"), range_start=sink)  # :456 vs :617
        yield "chunk", stride, st
        if st.evict:
            n -= stride
    if plan.mode == "ppl":
        return
    if plan.mode == "encoding":                                              # :508-526
        for _ in range(max_new_tokens):
            yield "decode", 1, Step()
        return
    B = plan.budget                                                          # enc-dec decode, :670-748
    rw = int(B * 0.3)                                                        # :709-710
    for _ in range(max_new_tokens):
        yield "decode", 1, Step(policy=policy, accumulate=scored, evict=1, counter_add=1.0,
                                k_feasible=B - rw, win_lo=0, win_recent=rw if policy == "h2o_head" else 0,
                                range_start=sink)


def initial_counter(plan: Plan, keep_attention=False):
    """C for the slots that exist right after the dense prefill (A.2).  The not-yet-existing
    tail of the reference's fixed-length state tensor is represented by Step.c_new0/c_new_step."""
    if plan.mode == "decoding":
        return None
    n0 = plan.r_idx
    if keep_attention:                                                       # :413-414
        return (plan.idx - torch.arange(n0, dtype=torch.float32))
    return torch.zeros(n0)                                                   # :416


# ---------------------------------------------------------------------------------------------------------------
# Sampling / perplexity tail
# ---------------------------------------------------------------------------------------------------------------
def logits_adapter(logits, temperature, top_p, scale_mul=False):
    """easykv/easykv.py:115-134 restated without the two sorts' round trip: softmax(logits / temperature), then in
    descending order (stable, i.e. index ascending among equals) keep every token whose EXCLUSIVE cumulative mass is
    <= top_p (:125-127), zero the rest and renormalise by the kept mass (:128).  Returns (final, raw softmax) like
    the reference.  `scale_mul`: ATen's CUDA kernels divide by a host scalar as a multiply by the fp32 reciprocal."""
    shape = logits.shape
    x = logits.reshape(-1, shape[-1])
    if scale_mul:      # ATen BinaryDivTrueKernel.cu: opmath_t(1.0) / scalar_value<opmath_t>() — both already fp32
        scaled = x * (torch.tensor(1.0, dtype=torch.float32) / torch.tensor(temperature, dtype=torch.float32)).to(x.dtype)
    else:
        scaled = x / temperature
    prob = torch.softmax(scaled, dim=-1)
    order = torch.sort(prob, dim=-1, descending=True, stable=True).indices
    sp = torch.gather(prob, -1, order)
    excl = torch.cumsum(sp, dim=-1) - sp
    keep_sorted = ~(excl > top_p)
    keep = torch.zeros_like(keep_sorted).scatter_(-1, order, keep_sorted)
    kept = torch.where(keep, prob, torch.zeros_like(prob))
    # the reference sums the sorted, masked vector (:128); same values, same order as `sp` with the tail zeroed
    z = torch.where(keep_sorted, sp, torch.zeros_like(sp)).sum(dim=-1, keepdim=True)
    return (kept / z).reshape(shape), torch.softmax(x, dim=-1).reshape(shape)


def top_p_margin(logits, temperature, top_p):
    """Smallest |exclusive cumulative mass - top_p| over a row's tokens: when it is within fp32 summation noise the
    kept set is decided by rounding and two correct implementations may differ by the boundary token."""
    x = logits.reshape(-1, logits.shape[-1])
    sp = torch.sort(torch.softmax(x / temperature, dim=-1), dim=-1, descending=True).values
    return ((torch.cumsum(sp, -1) - sp) - top_p).abs().min(dim=-1).values


def draw_from_exponentials(prob, q_exp):
    """torch.multinomial(prob, 1) (easykv.py:258, :509, :671) as ATen computes it for one draw: argmax(prob / q) with
    q ~ Exp(1) drawn per element (ATen/native/Sampling: the 'gumbel' fast path), first index on ties."""
    return torch.argmax(prob / q_exp, dim=-1, keepdim=True)


def token_nll(logits, targets):
    """CrossEntropyLoss(reduction='none')(all_logits[:-1], all_ids[1:]) (easykv.py:896-899), one row per target."""
    lse = torch.logsumexp(logits.float(), dim=-1)
    return lse - logits.float().gather(-1, targets.view(-1, 1).long())[:, 0]
