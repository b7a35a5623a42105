"""Budget arithmetic: the model-independent known answers the reference publishes in its README
(cache sizes, README.md:153,211,314,274,288), and product host logic == oracle restatement."""
import itertools

import pytest

from easykv_b200 import plan as P
from oracle import restate as R


@pytest.mark.parametrize("length,budget,stride,mode,retained", [
    (5144, 0.5, 24, "encoding", 2576),     # README.md:153
    (9994, 0.5, 96, "encoding", 5002),     # README.md:211
    (10253, 0.5, 96, "ppl", 5165),         # README.md:314
])
def test_readme_cache_sizes(length, budget, stride, mode, retained):
    for mod in (P, R):
        assert mod.resolve_plan(mode, length, budget, stride).idx == retained


def test_readme_decoding_budgets():
    # README.md:274,288: budget 300 -> 300 generated slots kept; 150 -> 150
    for budget, new in ((300, 516 - 0), (150, 530)):
        pl = P.resolve_plan("decoding", 64, budget, 1)
        n = 0
        for kind, q, st in P.schedule(pl, "roco", new):
            n += q - st.evict
        assert n == budget


def test_survey_derived_sizes():
    c1 = P.resolve_plan("encoding", 256, 0.5, 8)
    assert (c1.budget, c1.idx, c1.r_idx, c1.recent_window, c1.sink) == (136, 136, 128, 13, 4)
    c2 = P.resolve_plan("auto", 4096, 1024, 64)
    assert (c2.mode, c2.budget, c2.idx, c2.r_idx, c2.recent_window) == ("encoding_decoding", 1088, 1088, 64, 108)
    c3 = P.resolve_plan("encoding", 16384, 0.5, 16)
    assert (c3.budget, c3.idx, c3.r_idx, c3.recent_window) == (8208, 8208, 8192, 820)


def _same_step(a, b):
    for f in ("policy", "evict", "score_offset", "counter_add", "c_new0", "c_new_step", "k_feasible", "protect_last",
              "sink_protect", "win_lo", "win_recent", "range_start"):
        assert getattr(a, f) == getattr(b, f), f
    assert bool(a.accumulate) == bool(b.accumulate)
    assert bool(a.tova_head_mean) == bool(b.tova_head_mean)


@pytest.mark.parametrize("mode,length,budget,stride", [
    ("encoding", 256, 0.5, 8), ("encoding", 300, 100, 7), ("ppl", 144, 0.4, 8), ("auto", 160, 64, 8),
    ("auto", 96, 200, 4), ("decoding", 32, 40, 1), ("encoding_decoding", 500, 128, 16), ("encoding", 64, 1.0, 8),
])
def test_plan_and_schedule_match_oracle(mode, length, budget, stride):
    a, b = P.resolve_plan(mode, length, budget, stride), R.resolve_plan(mode, length, budget, stride)
    assert (a.mode, a.budget, a.idx, a.r_idx, a.recent_window, a.sink) == (b.mode, b.budget, b.idx, b.r_idx, b.recent_window, b.sink)
    for pol, keep in itertools.product(("roco", "h2o_head", "tova", "recency", "full"), (False, True)):
        if a.mode == "encoding_decoding" and pol in ("h2o_head", "full"):      # whitelist, easykv.py:536-537
            with pytest.raises(AssertionError):
                list(P.schedule(a, pol, 50, keep))
            continue
        sa, sb = list(P.schedule(a, pol, 50, keep)), list(R.schedule(b, pol, 50, keep))
        assert len(sa) == len(sb)
        for (ka, qa, xa), (kb, qb, xb) in zip(sa, sb):
            assert (ka, qa) == (kb, qb)
            _same_step(xa, xb)
    ca, cb = P.initial_counter(a), R.initial_counter(b)
    assert (ca or []) == ([] if cb is None else cb.tolist())


def test_policy_names():
    assert P.canonical_policy("h2o") == "h2o_head"          # BASELINE.json's spelling
    with pytest.raises(ValueError):
        P.canonical_policy("h2o_head_std_avg")               # stale name in the reference's test_passkey.py:56
    with pytest.raises(AssertionError):
        P.resolve_plan("auto", 100, 0.5, 8)                  # easykv.py:222
    with pytest.raises(AssertionError):
        P.resolve_plan("encoding_decoding", 100, 64, 1)      # stride 1 asserts in the reference (:666-669)


def test_plan_and_schedule_match_oracle_property():
    """Randomised: the product's integer-only host logic and the restatement agree on every resolvable
    (mode, length, budget, stride) and raise together on the ones the reference asserts on."""
    import random
    rng = random.Random(7)
    checked = 0
    for _ in range(400):
        mode = rng.choice(["encoding", "ppl", "auto", "decoding", "encoding_decoding"])
        length = rng.randint(16, 3000)
        stride = rng.choice([1, 2, 4, 7, 8, 16, 24, 64, 96])
        budget = rng.choice([rng.randint(8, 2 * length), round(rng.uniform(0.05, 1.2), 3)])
        if mode in ("auto", "decoding", "encoding_decoding"):
            budget = int(budget) if isinstance(budget, int) else rng.randint(8, 2 * length)
        try:
            b = R.resolve_plan(mode, length, budget, stride)
        except (AssertionError, StopIteration):
            with pytest.raises((AssertionError, ValueError)):
                P.resolve_plan(mode, length, budget, stride)
            continue
        try:
            a = P.resolve_plan(mode, length, budget, stride)
        except AssertionError:
            # the reference fails later: stride 1 asserts at :666-669; an empty dense prefix (r_idx == 0) cannot be run
            assert (b.mode == "encoding_decoding" and stride == 1) or (b.mode == "encoding" and b.r_idx == 0)
            continue
        assert (a.mode, a.budget, a.idx, a.r_idx, a.recent_window, a.sink) == (b.mode, b.budget, b.idx, b.r_idx, b.recent_window, b.sink)
        pol = rng.choice(["roco", "tova", "recency"] + ([] if a.mode == "encoding_decoding" else ["h2o_head", "full"]))
        keep = rng.random() < 0.3
        sa, sb = list(P.schedule(a, pol, 12, keep)), list(R.schedule(b, pol, 12, keep))
        assert len(sa) == len(sb)
        for (ka, qa, xa), (kb, qb, xb) in zip(sa, sb):
            assert (ka, qa) == (kb, qb)
            _same_step(xa, xb)
        checked += 1
    assert checked > 150


@pytest.mark.parametrize("mode,length,budget,stride,policy", [
    ("decoding", 70, 16, 1, "roco"), ("decoding", 70, 16, 1, "tova"), ("decoding", 70, 16, 1, "recency"),
    ("decoding", 70, 16, 1, "h2o_head"), ("auto", 70, 40, 8, "roco"), ("auto", 70, 40, 8, "recency"),
    ("auto", 4096, 1024, 64, "roco"), ("auto", 70, 200, 8, "roco")])
def test_decode_schedule_reaches_a_steady_state(mode, length, budget, stride, policy):
    """What `easykv.GraphedDecodeStep` relies on: once eviction has started, every remaining decode step carries the SAME
    parameters (easykv.py:303-362: `C += 1`, one victim per step, constant `k`; :708-748) — so the step's launches are
    identical from token to token and can be replayed from a CUDA graph.  In `decoding` mode that state begins when the
    generated tokens exceed the budget (:303); in `encoding_decoding` it holds from the first decode step."""
    new = 200
    plan = P.resolve_plan(mode, length, budget, stride)
    decodes = [s for s in P.schedule(plan, policy, new) if s[0] == "decode"]
    assert len(decodes) == new
    tail = decodes[-1][2]
    assert tail.evict == 1
    first_steady = next(i for i, s in enumerate(decodes) if s[2] == tail)
    assert all(s[2] == tail for s in decodes[first_steady:])
    assert all(s[2].evict == 0 for s in decodes[:first_steady])
    if plan.mode == "decoding":
        assert first_steady == int(plan.budget)
    else:
        assert first_steady == 0
