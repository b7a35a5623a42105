"""TEST INFRASTRUCTURE — NOT PRODUCT CODE.

Teacher-forced replay of a golden trace (`tests/golden/*.npz`, produced from the unmodified
reference by `oracle/gen_golden.py`) through any *engine* exposing

    engine.load_prefill(layer, K[Hkv,n,d], V[Hkv,n,d], n_scored, C_init or None)
    engine.forward(layer, step: restate.Step, q[H,ql,d], k[Hkv,ql,d], v[Hkv,ql,d], force=None)
        -> (out[H,ql,d], victims[Hkv,evict] int64 or None)
       (`force`: victim ids to APPLY instead of the engine's own choice — the engine still
        reports its own choice; used to re-synchronise after a tie-ambiguous reference step)
    engine.export(layer) -> (K[Hkv,n,d], V[Hkv,n,d]) in the reference's logical order

Every forward gets the reference's own q/k/v, so each eviction decision
is tested on identical inputs; victims are compared as sorted sets per (layer, head)
(SURVEY A.5 "What to compare").
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np
import torch

from . import restate

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def list_golden():
    """Recorded reference runs (sampling_tail.npz holds function-level vectors, not a run: `load_sampling_tail`)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and f != "sampling_tail.npz")


def load_sampling_tail():
    """Cases of tests/golden/sampling_tail.npz (oracle/gen_golden.py: the reference's logits_adapter / loss outputs)."""
    z = np.load(os.path.join(GOLDEN_DIR, "sampling_tail.npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return [dict(m, **{k: torch.from_numpy(z[f"c{i}_{k}"]) for k in ("logits", "final", "raw", "targets", "nll")})
            for i, m in enumerate(meta["cases"])]


GOLDEN_GPU_DIR = os.path.join(os.path.dirname(GOLDEN_DIR), "golden_gpu")


def list_golden_gpu():
    """Runs of the unmodified reference recorded ON A B200 (fp16 / bf16 on CUDA; oracle/gen_golden_gpu.py)."""
    if not os.path.isdir(GOLDEN_GPU_DIR):
        return []
    return sorted(f[:-4] for f in os.listdir(GOLDEN_GPU_DIR) if f.endswith(".npz"))


def load_golden(name, golden_dir=None):
    z = np.load(os.path.join(golden_dir or GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["meta"]).decode())
    return meta, z


class OracleEngine:
    """The CPU restatement behind the engine interface."""

    def __init__(self, L, H, Hkv, d, dtype, scale_mul=False):
        self.layers = [restate.LayerOracle(Hkv, d, dtype) for _ in range(L)]
        self.scale_mul = scale_mul

    def load_prefill(self, l, K, V, n_scored, C_init, S_init=None, SQ_init=None):
        self.layers[l].load_prefill(K, V, n_scored, C_init)
        if S_init is not None:                       # keep_attention seeding (h2o_head_score)
            self.layers[l].S, self.layers[l].SQ = S_init.clone().float(), SQ_init.clone().float()

    def forward(self, l, st, q, k, v, force=None, stream_table=None):
        return self.layers[l].forward(st, q, k, v, self.scale_mul, force=force, stream_table=stream_table)

    def export(self, l):
        return self.layers[l].K, self.layers[l].V

    def margin(self, l):
        return self.layers[l].last_margin


@dataclass
class Report:
    name: str
    n_forwards: int = 0
    n_events: int = 0
    victim_mismatch: list = field(default_factory=list)   # (fwd, layer, ref ids, got ids, margin)
    tie_ambiguous: list = field(default_factory=list)     # same, but the reference's margin was 0
    max_out_err: float = 0.0
    max_out_rel: float = 0.0                              # error relative to max(1, largest |reference output|) of the forward
    final_cache_equal: bool = True
    min_margin: tuple = (float("inf"), float("inf"))
    retained: int = 0

    @property
    def ok(self):
        return not self.victim_mismatch and self.final_cache_equal


def case_plan(meta):
    c = meta["case"]
    gen = c["gen"]
    mode = c["mode"]
    plan = restate.resolve_plan(mode, c["seq"], gen["budget"], c["stride"],
                                gen.get("recent_ratio", 0.1), gen.get("temp_length", 4))
    return plan, gen["kv_policy"], c["max_new_tokens"]


def case_stream_table(meta):
    """(cos, sin) [rows, d] in the case's dtype for generation_config['streaming'] runs, else None — the table the
    scaffold's rotary module hands the reference (fp32 build, one cast)."""
    c = meta["case"]
    if not c["gen"].get("streaming", False):
        return None
    d, dtype = c["d"], getattr(torch, c["dtype"])
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, d, 2, dtype=torch.float32) / d))
    inv_freq = inv_freq.to(dtype).float()        # the scaffold's inv_freq buffer was cast with the model (`.to(dtype)`)
    emb = torch.outer(torch.arange(4096, dtype=torch.float32), inv_freq)
    emb = torch.cat((emb, emb), dim=-1)
    return emb.cos().to(dtype), emb.sin().to(dtype)


def case_keep_attention(meta):
    return bool(meta["case"]["gen"].get("keep_attention", False))


def replay(name, engine_factory, resync=True, shadow=None, tie_eps=0.0, golden_dir=None, trace=None) -> Report:
    """`resync`: after comparing, apply the REFERENCE's victims so every later step is again
    tested on identical state.  `shadow`: an OracleEngine factory run in lock-step (always
    forced to the reference's victims) whose decision margins classify a mismatch as
    tie-ambiguous (margin <= tie_eps) or real.  `trace`: (meta, arrays) of a run recorded in this
    process (oracle/gen_golden.trace_arrays) instead of a file."""
    meta, z = trace if trace is not None else load_golden(name, golden_dir)
    c = meta["case"]
    dtype = getattr(torch, c["dtype"])
    L, H, Hkv, d = c["L"], c["H"], c["Hkv"], c["d"]
    plan, policy, max_new = case_plan(meta)
    eng = engine_factory(L, H, Hkv, d, dtype)
    sh = shadow(L, H, Hkv, d, dtype) if shadow is not None else None
    rep = Report(name)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dtype)
    keep = case_keep_attention(meta)
    table = case_stream_table(meta)
    skw = {} if table is None else {"stream_table": table}
    C0 = restate.initial_counter(plan, keep)
    for l in range(L):
        K, V = T(z[f"prefill_K_{l}"]), T(z[f"prefill_V_{l}"])
        n_scored = 0 if plan.mode == "decoding" else K.shape[1]
        seeds = ()
        if keep:                                     # the reference's own seeds for the prefilled slots
            n0 = K.shape[1]
            seeds = (torch.from_numpy(z["seed_S"][l][:, :n0].copy()), torch.from_numpy(z["seed_SQ"][l][:, :n0].copy()))
        eng.load_prefill(l, K, V, n_scored, C0, *seeds)
        if sh is not None:
            sh.load_prefill(l, K, V, n_scored, C0, *seeds)
    events = {}
    for e, ev in enumerate(meta["events"]):
        events[ev["fwd"]] = (ev, z[f"ev{e}_ids"])
    sched = list(restate.schedule(plan, policy, max_new, keep))
    fwds = meta["forwards"][1:]
    assert len(sched) == len(fwds), (len(sched), len(fwds))
    for f, ((kind, ql, st), fm) in enumerate(zip(sched, fwds), start=1):
        assert ql == fm["q_len"]
        rep.n_forwards += 1
        ev = events.get(f)
        assert (ev is not None) == bool(st.evict), (f, st)
        for l in range(L):
            if not fm["recorded"]:
                raise AssertionError("golden forward without tensors")
            q, k, v, o = (T(z[f"f{f}_l{l}_{key}"]) for key in "qkvo")
            ref_ids = None
            if ev is not None:
                kindname, ids = ev[0]["kind"], torch.from_numpy(ev[1])
                if kindname == "range":
                    ref_ids = torch.arange(int(ids[0]), int(ids[1])).repeat(Hkv, 1)
                else:
                    ref_ids = ids[l].reshape(Hkv, -1)
            if policy == "random" and ev is not None:      # the range is the host's draw: replay the reference's
                st.range_start = int(ev[1][0]) - st.score_offset
            force = ref_ids if resync else None
            out, vic = eng.forward(l, st, q, k, v, force=force, **skw)
            o_ref = o.view(ql, H, d).transpose(0, 1).float()
            err = (out.float().cpu() - o_ref).abs().max().item()
            rep.max_out_err = max(rep.max_out_err, err)
            rep.max_out_rel = max(rep.max_out_rel, err / max(1.0, o_ref.abs().max().item()))
            margin = None
            meng = sh if sh is not None else (eng if hasattr(eng, "margin") else None)
            if sh is not None:
                sh.forward(l, st, q, k, v, force=ref_ids, **skw)
            if meng is not None and st.evict:
                margin = meng.margin(l)
                rep.min_margin = (min(rep.min_margin[0], margin[0]), min(rep.min_margin[1], margin[1]))
            if ev is not None:
                got = torch.sort(vic.cpu().long(), dim=-1)[0]
                ref_sorted = torch.sort(ref_ids, dim=-1)[0]
                if not torch.equal(got, ref_sorted):
                    rec = (f, l, ref_sorted, got, margin)
                    if margin is not None and min(margin) <= tie_eps:
                        rep.tie_ambiguous.append(rec)
                    else:
                        rep.victim_mismatch.append(rec)
        if ev is not None:
            rep.n_events += 1
        if (rep.victim_mismatch or rep.tie_ambiguous) and not resync:
            rep.final_cache_equal = False
            return rep
    for l in range(L):
        K, V = eng.export(l)
        Kr, Vr = T(z[f"final_K_{l}"]), T(z[f"final_V_{l}"])
        rep.retained = K.shape[1]
        if K.shape != Kr.shape or not (torch.equal(K.cpu(), Kr) and torch.equal(V.cpu(), Vr)):
            rep.final_cache_equal = False
    return rep
