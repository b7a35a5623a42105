"""The host driver (`easykv_b200.easykv`: enable_fixed_kv / easykv_generate / easykv_ppl and the attention
seam) against the reference's own runs.

CPU part (`-m "not gpu"`): the driver's sequencing, RoPE, seam binding, sampling and printing are exercised
with the device cache replaced by an oracle-backed stand-in (tests only — the product has no CPU path), so
that the whole call must reproduce what the UNMODIFIED reference returned, printed and evicted in
tests/golden/*.npz.  GPU part: the same calls through the real CUDA path.
"""
import contextlib
import io
import json

import pytest
import torch

from oracle import replay, restate, scaffold

import easykv_b200
from easykv_b200 import easykv as drv

CASES = replay.list_golden()


class OracleCache:
    """Stand-in for BudgetedKVCache with the same interface, computing with the CPU restatement."""

    def __init__(self, L, B, H, Hkv, d, capacity, dtype=torch.float32, device="cpu", arith=0):
        assert B == 1
        self.L, self.Hkv = L, Hkv
        self.layers = [restate.LayerOracle(Hkv, d, dtype) for _ in range(L)]
        self.n = [0] * L
        self.cap = capacity
        self.K_raw = None
        self.scale_mul = bool(arith)

    def sample(self, logits, temperature, top_p):
        prob, _ = restate.logits_adapter(logits.float(), temperature, top_p, self.scale_mul)
        return torch.multinomial(prob, num_samples=1)

    def token_nll(self, logits, targets):
        return restate.token_nll(logits, targets)

    def enable_streaming(self, adopt_rotated=False):
        self.K_raw = True                             # the layers' K simply IS the (un-rotated) cache from now on

    def step_stream(self, l, sp, q_in, k_in, v_in, cos, sin, apply=True, kernel=0):
        lo = self.layers[l]
        b, ql = q_in.shape[:2]
        q = q_in.view(b, ql, -1, lo.d).transpose(1, 2)[0]
        k = k_in.view(b, ql, -1, lo.d).transpose(1, 2)[0]
        v = v_in.view(b, ql, -1, lo.d).transpose(1, 2)[0]
        table = (cos.to(lo.dtype), sin.to(lo.dtype))
        if sp.policy == "full":
            lo.K, lo.V = torch.cat([lo.K, k], 1), torch.cat([lo.V, v], 1)
            n = lo.K.shape[1]
            pos = torch.arange(n)
            out, _ = restate.attend(restate.rope(q, *table, pos[n - ql:]), restate.rope(lo.K, *table, pos), lo.V, self.scale_mul)
            ids = None
        else:
            st = restate.Step(**{f: getattr(sp, f) for f in restate.Step.__dataclass_fields__})
            out, ids = lo.forward(st, q, k, v, self.scale_mul, stream_table=table)
            if ids is not None:
                ids = torch.sort(ids, dim=-1)[0][None].int()
        self.n[l] = lo.K.shape[1]
        return out[None], ids

    def set_counter(self, l, values):
        lo, n = self.layers[l], len(values)
        if lo.S.shape[-1] != n:                       # no state yet (the dense prefill ran without a policy)
            lo.S, lo.SQ = torch.zeros(self.Hkv, n), torch.zeros(self.Hkv, n)
        lo.C = torch.tensor(values, dtype=torch.float32).repeat(self.Hkv, 1)

    def rope_qkv(self, q_in, k_in, v_in, cos, sin, positions=None):
        """The torch ops of easykv/llama_patch.py:47-72, 169-171 (what ekv_rope_qk fuses)."""
        from easykv_b200.attention import apply_rope
        b, ql = q_in.shape[:2]
        d = self.layers[0].d
        q = q_in.view(b, ql, -1, d).transpose(1, 2)
        k = k_in.view(b, ql, -1, d).transpose(1, 2)
        v = v_in.view(b, ql, -1, d).transpose(1, 2)
        if positions is not None:
            cos, sin = cos[positions], sin[positions]
        q, k = apply_rope(q, k, cos.to(q.dtype), sin.to(q.dtype))
        return q, k, v

    def round_state(self, l):
        lo = self.layers[l]
        lo.S, lo.SQ = lo.S.to(lo.dtype).float(), lo.SQ.to(lo.dtype).float()

    def step(self, l, sp, q, k, v, apply=True, kernel=0):
        lo = self.layers[l]
        if sp.policy == "full":                       # dense prefill / plain decode: no policy state
            lo.K, lo.V = torch.cat([lo.K, k[0]], 1), torch.cat([lo.V, v[0]], 1)
            out, _ = restate.attend(q[0], lo.K, lo.V, self.scale_mul)
            ids = None
        else:
            st = restate.Step(**{f: getattr(sp, f) for f in restate.Step.__dataclass_fields__})
            out, ids = lo.forward(st, q[0], k[0], v[0], self.scale_mul)
            if ids is not None:
                ids = torch.sort(ids, dim=-1)[0][None].int()
        self.n[l] = lo.K.shape[1]
        if self.n[l] + (0 if ids is None else ids.shape[-1]) > self.cap:      # what BudgetedKVCache.step raises
            raise ValueError(f"cache capacity {self.cap} exceeded")
        return out[None], ids


def _model_and_ids(meta, device="cpu"):
    c = meta["case"]
    model = scaffold.build(c["arch"], seed=0, dtype=getattr(torch, c["dtype"]), device=device, L=c["L"], H=c["H"],
                           Hkv=c["Hkv"], d=c["d"], vocab=512)
    ids = torch.randint(3, 512, (1, c["seq"]), generator=torch.Generator().manual_seed(1))
    return model, ids


def _run(meta, model, ids, aten_arith):
    c = meta["case"]
    gen = dict(temperature=1e-9, top_p=1.0, max_new_tokens=c["max_new_tokens"], aten_arith=aten_arith, **c["gen"])
    buf = io.StringIO()
    torch.manual_seed(meta.get("rng_seed", 0))        # kv_policy='random' draws from torch's CPU generator
    with contextlib.redirect_stdout(buf):
        ppl = c["mode"] == "ppl"
        easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="encoding" if ppl else c["mode"], stride=c["stride"])
        fn = model.easykv_ppl if ppl else model.easykv_generate
        result = fn(input_ids=ids, generation_config=gen)
    return result, buf.getvalue(), model.easykv_last


def _golden_events(meta, z, Hkv):
    out = []
    for e, ev in enumerate(meta["events"]):
        ids = torch.from_numpy(z[f"ev{e}_ids"])
        if ev["kind"] == "range":
            L = meta["case"]["L"]
            ids = torch.arange(int(ids[0]), int(ids[1])).repeat(L, Hkv, 1)
        else:
            ids = ids.reshape(ids.shape[0], Hkv, -1)
        out.append(torch.sort(ids, dim=-1)[0])
    return out


def _compare(meta, z, result, printed, sess, exact_events=True):
    c = meta["case"]
    ref_events = _golden_events(meta, z, c["Hkv"])
    got = [torch.sort(ids[:, 0].cpu().long(), dim=-1)[0] for _, ids in sess.events]
    assert len(got) == len(ref_events)
    diverged = next((i for i, (a, b) in enumerate(zip(got, ref_events)) if not torch.equal(a, b)), None)
    if exact_events:
        assert diverged is None, f"eviction event {diverged} differs"
        if c["mode"] == "ppl":
            assert result == pytest.approx(float(meta["result"]), rel=1e-6)
        else:
            assert result == meta["result"]
        ratio = [ln for ln in meta["printed"].splitlines() if "atio" in ln and "%" in ln]
        for ln in ratio:
            assert ln in printed
    return diverged


@pytest.mark.parametrize("name", [n for n in CASES if "fp32" in n])
def test_driver_reproduces_reference_runs_cpu(name, monkeypatch):
    """Same text / perplexity, same printed retained-cache line, same eviction ids as the reference, for
    every mode (decoding, encoding, auto -> encoding_decoding, ppl) and policy in the fp32 golden set."""
    meta, z = replay.load_golden(name)
    monkeypatch.setattr(drv, "BudgetedKVCache", OracleCache)
    monkeypatch.setattr(drv, "DENSE_CHUNK", 1 << 30)       # one dense forward, as the reference issues it
    # the golden runs chose tokens greedily without touching the CPU generator (oracle/ref_harness.py)
    monkeypatch.setattr(torch, "multinomial", lambda p, num_samples=1, **kw: p.argmax(dim=-1, keepdim=True))
    model, ids = _model_and_ids(meta)
    result, printed, sess = _run(meta, model, ids, "cpu")
    _compare(meta, z, result, printed, sess)
    assert "forward" not in model.layers[0].self_attn.__dict__      # the seam is unbound again


def test_driver_chunked_dense_prefill_cpu(monkeypatch):
    """The dense prefill issued as 64-token causal chunks gives the same run (C1)."""
    meta, z = replay.load_golden("c1_llama_enc_roco_fp32")
    monkeypatch.setattr(drv, "BudgetedKVCache", OracleCache)
    model, ids = _model_and_ids(meta)
    result, printed, sess = _run(meta, model, ids, "cpu")
    assert "53.12%(136/256)" in printed
    assert len(sess.events) == 15


@pytest.mark.parametrize("mode,budget", [("encoding", 0.5), ("ppl", 0.5), ("encoding", 1.0), ("ppl", 1.0)])
def test_driver_full_policy_baseline_matches_reference_cpu(mode, budget, monkeypatch):
    """kv_policy='full' with a budget below the prompt length is the full-cache baseline the reference's ppl and
    summarisation scripts run (easykv.py:459, :850 `mode != 'full'`): nothing is ever evicted, the cache grows to the
    whole prompt (the capacity must follow the schedule, not idx + stride), and 'encoding' prints its ratio line even
    when the prefill is dense (:501-503).  Compared with a fresh run of the unmodified reference."""
    from oracle import ref_harness
    if ref_harness.reference_root() is None:
        pytest.skip("reference not available")
    c = dict(arch="llama", L=1, H=4, Hkv=2, d=128, seq=100, dtype="float32", mode=mode, stride=8,
             max_new_tokens=0 if mode == "ppl" else 3, gen=dict(budget=budget, kv_policy="full"))
    model = scaffold.build(c["arch"], seed=0, dtype=torch.float32, L=c["L"], H=c["H"], Hkv=c["Hkv"], d=c["d"], vocab=512)
    ids = torch.randint(3, 512, (1, c["seq"]), generator=torch.Generator().manual_seed(1))
    gen = dict(temperature=1e-9, top_p=1.0, max_new_tokens=c["max_new_tokens"], **c["gen"])
    ref = ref_harness.run_reference(model, ids, gen, mode="encoding", stride=c["stride"], ppl=mode == "ppl", record_tensors=False)
    monkeypatch.setattr(drv, "BudgetedKVCache", OracleCache)
    monkeypatch.setattr(torch, "multinomial", lambda p, num_samples=1, **kw: p.argmax(dim=-1, keepdim=True))
    meta = dict(case=c)
    result, printed, sess = _run(meta, model, ids, "cpu")
    if mode == "ppl":
        assert result == pytest.approx(float(ref.result), rel=1e-5)
    else:
        assert result == str(ref.result)
    for ln in [ln for ln in ref.printed.splitlines() if "atio" in ln and "%" in ln]:
        assert ln in printed, (ln, printed)
    assert not sess.events


def test_driver_rejects_what_the_reference_silently_ignores():
    meta, _ = replay.load_golden("c1_llama_enc_roco_fp32")
    model, ids = _model_and_ids(meta)
    easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="encoding", stride=8)
    with pytest.raises(ValueError):
        model.easykv_generate(input_ids=ids, generation_config=dict(kv_policy="h2o_head_std_avg"))
    with pytest.raises(ValueError):
        easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="bogus")


def test_sampling_tail_has_no_cpu_path():
    """`logits_adapter` keeps the reference's name and contract but is a CUDA launch: CPU tensors are refused loudly."""
    with pytest.raises(RuntimeError, match="CUDA"):
        drv.logits_adapter(torch.randn(2, 50), 0.7, 0.9)


@pytest.mark.parametrize("arch", ["llama", "mistral", "mistral_sliding"])
def test_seam_binds_to_installed_transformers_cpu(monkeypatch, arch):
    """The attention seam on the installed transformers (5.x) Llama / Mistral classes — `position_embeddings`, 2-tuple
    return, keyword-only decoder-layer call, the already-4D mask that skips their causal-mask construction — with the
    stand-in cache: without a policy, greedy generation through `easykv_generate` must equal the model's own eager
    generation."""
    transformers = pytest.importorskip("transformers")
    monkeypatch.setattr(drv, "BudgetedKVCache", OracleCache)
    monkeypatch.setattr(torch, "multinomial", lambda p, num_samples=1, **kw: p.argmax(dim=-1, keepdim=True))
    kw = dict(hidden_size=256, intermediate_size=512, num_hidden_layers=2, num_attention_heads=4, num_key_value_heads=2,
              head_dim=64, vocab_size=300, max_position_embeddings=512, attn_implementation="eager")
    if arch == "llama":
        cfg, cls = transformers.LlamaConfig(**kw), transformers.LlamaForCausalLM
    else:       # a window longer than the sequence: same attention, but the sliding-window mask constructor is the one bypassed
        cfg = transformers.MistralConfig(sliding_window=None if arch == "mistral" else 256, **kw)
        cls = transformers.MistralForCausalLM
    torch.manual_seed(0)
    model = cls(cfg).eval()
    ids = torch.randint(3, 300, (1, 40), generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = model.generate(ids, max_new_tokens=8, do_sample=False, pad_token_id=0)[0, 40:].tolist()
    with contextlib.redirect_stdout(io.StringIO()):
        easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="decoding")
        text = model.easykv_generate(input_ids=ids, generation_config=dict(
            temperature=1e-9, max_new_tokens=8, budget=512, kv_policy="full", aten_arith="cpu"))
    assert [int(t) for t in text.split()] == ref
    assert "forward" not in model.model.layers[0].self_attn.__dict__
    assert model.easykv_last.model_kwargs["attention_mask"].dim() == 4


# ---------------------------------------------------------------------------------------------------------
# GPU: the same user-level calls through the CUDA library
# ---------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", [n for n in CASES if "fp32" in n])
def test_driver_end_to_end_gpu(name, ekv_lib):
    """Free-running on the GPU (projections by cuBLAS, attention + eviction by the CUDA path) against the
    reference's CPU run.  Inputs are no longer bit-identical (GEMM summation order), so eviction ids may
    fork at a near-tie; until the first fork every event must be identical, and the retained-cache line is
    model-independent."""
    meta, z = replay.load_golden(name)
    model, ids = _model_and_ids(meta, device="cuda")
    result, printed, sess = _run(meta, model, ids, "cpu")
    diverged = _compare(meta, z, result, printed, sess, exact_events=False)
    ratio = [ln for ln in meta["printed"].splitlines() if "atio" in ln and "%" in ln]
    for ln in ratio:
        assert ln in printed
    n_ev = len(meta["events"])
    print(json.dumps(dict(case=name, events=n_ev, first_divergence=diverged)))
    assert diverged is None or diverged >= n_ev // 2, f"{name}: diverged at event {diverged} of {n_ev}"
    if diverged is None and meta["case"]["mode"] != "ppl":
        assert result == meta["result"]


@pytest.mark.gpu
def test_driver_batch_of_identical_prompts(ekv_lib):
    """The reference is hard-wired to batch 1 (easykv.py:66-67, 290, 430); here sequences are independent units of
    the same launches: a batch of identical prompts must evict and generate exactly what one prompt does."""
    meta, z = replay.load_golden("gqa_mistral_auto_roco_fp32")
    model, ids = _model_and_ids(meta, device="cuda")
    res1, _, sess1 = _run(meta, model, ids, "cpu")
    ev1 = [e.clone() for _, e in sess1.events]
    res3, _, sess3 = _run(meta, model, ids.repeat(3, 1), "cpu")
    assert res3 == [res1] * 3
    assert len(sess3.events) == len(ev1)
    for (_, e3), e1 in zip(sess3.events, ev1):
        for b in range(3):
            assert torch.equal(e3[:, b], e1[:, 0])


@pytest.mark.gpu
def test_driver_binds_to_installed_transformers(ekv_lib):
    """The seam binds to transformers-5.x attention modules (position_embeddings, 2-tuple return): with
    kv_policy='full' nothing is evicted, so greedy generation must equal HF's own eager generation."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=128, vocab_size=512, max_position_embeddings=1024,
                                   attn_implementation="eager")
    torch.manual_seed(0)
    model = transformers.LlamaForCausalLM(cfg).cuda().eval()
    ids = torch.randint(3, 512, (1, 70), generator=torch.Generator().manual_seed(1)).cuda()
    with torch.no_grad():
        ref = model.generate(ids, max_new_tokens=12, do_sample=False, pad_token_id=0)[0, 70:].tolist()
    easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="decoding")
    text = model.easykv_generate(input_ids=ids, generation_config=dict(
        temperature=1e-9, max_new_tokens=12, budget=1024, kv_policy="full"))
    assert [int(t) for t in text.split()] == ref
    # and with a real budget the cache is bounded
    text = model.easykv_generate(input_ids=ids, generation_config=dict(
        temperature=1e-9, max_new_tokens=40, budget=16, kv_policy="roco"))
    assert model.easykv_last.cache.n[0] == 70 + 16
    assert len(model.easykv_last.events) == 40 - 16


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,mode,policy,bsz", [("float32", "decoding", "roco", 1), ("float16", "auto", "roco", 2),
                                                   ("float16", "decoding", "tova", 1), ("bfloat16", "auto", "recency", 1)])
def test_graph_captured_decode_step_equals_eager(ekv_lib, dtype, mode, policy, bsz):
    """Steady-state decode steps replayed from one CUDA graph of the whole model step (GraphedDecodeStep) must generate
    the same tokens and evict the same slots as the eager loop, on the installed transformers Llama classes."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=128, vocab_size=512, max_position_embeddings=1024,
                                   attn_implementation="eager")
    torch.manual_seed(0)
    model = transformers.LlamaForCausalLM(cfg).to(getattr(torch, dtype)).cuda().eval()
    ids = torch.randint(3, 512, (bsz, 70), generator=torch.Generator().manual_seed(1)).cuda()
    easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode=mode, stride=8)
    budget = 16 if mode == "decoding" else 40
    runs = []
    for graph in (False, True):
        torch.manual_seed(7)
        text = model.easykv_generate(input_ids=ids, generation_config=dict(
            temperature=0.8, top_p=0.9, max_new_tokens=90, budget=budget, kv_policy=policy, cuda_graph=graph, cuda_graph_min_steps=8))
        sess = model.easykv_last
        runs.append((text, [(f, e.clone()) for f, e in sess.events], sess.cache.n[0], sess.graphed_steps, sess.graph_error,
                     [sess.cache.export(l)[0].clone() for l in range(2)]))
    (t0, ev0, n0, g0, _, k0), (t1, ev1, n1, g1, err1, k1) = runs
    assert g0 == 0 and err1 is None and g1 >= 40, (g1, err1)
    assert t1 == t0
    assert n1 == n0 and len(ev1) == len(ev0)
    for (f0, e0), (f1, e1) in zip(ev0, ev1):
        assert f0 == f1 and torch.equal(e0, e1)
    for a, b in zip(k0, k1):
        assert torch.equal(a, b)                        # the retained keys, in logical order


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,mode,policy,bsz,stride", [("float16", "auto", "roco", 1, 8), ("float16", "encoding", "h2o_head", 2, 16),
                                                          ("bfloat16", "auto", "recency", 1, 8), ("float16", "ppl", "roco", 1, 8)])
def test_graph_captured_strided_chunks_equal_eager(ekv_lib, dtype, mode, policy, bsz, stride):
    """Steady strided-prefill chunks replayed from one CUDA graph of the whole model forward (easykv.py:426-500,
    :587-661 as one launch per chunk) must evict the same slots, leave the same cache and produce the same tokens /
    perplexity as the eager chunk loop."""
    transformers = pytest.importorskip("transformers")
    cfg = transformers.LlamaConfig(hidden_size=512, intermediate_size=1024, num_hidden_layers=2, num_attention_heads=4,
                                   num_key_value_heads=2, head_dim=128, vocab_size=512, max_position_embeddings=2048,
                                   attn_implementation="eager")
    torch.manual_seed(0)
    model = transformers.LlamaForCausalLM(cfg).to(getattr(torch, dtype)).cuda().eval()
    length = 40 + stride * 30
    ids = torch.randint(3, 512, (bsz, length), generator=torch.Generator().manual_seed(1)).cuda()
    ppl = mode == "ppl"
    easykv_b200.enable_fixed_kv(model, scaffold.StubTokenizer(), mode="encoding" if ppl else mode, stride=stride)
    budget = 64 if mode == "auto" else 0.4
    runs = []
    for graph in (False, True):
        torch.manual_seed(7)
        gen = dict(temperature=1e-9, max_new_tokens=0 if ppl else 6, budget=budget, kv_policy=policy, cuda_graph=graph,
                   cuda_graph_min_chunks=3, cuda_graph_min_steps=1000)
        out = (model.easykv_ppl if ppl else model.easykv_generate)(input_ids=ids, generation_config=gen)
        sess = model.easykv_last
        runs.append((out, [(f, e.clone()) for f, e in sess.events], sess.cache.n[0], sess.graphed_chunks, sess.graph_error,
                     [sess.cache.export(l)[0].clone() for l in range(2)]))
    (o0, ev0, n0, g0, _, k0), (o1, ev1, n1, g1, err1, k1) = runs
    assert g0 == 0 and g1 >= 10, (g1, err1)
    assert n0 == n1 and len(ev0) == len(ev1)
    for (f0, e0), (f1, e1) in zip(ev0, ev1):
        assert f0 == f1 and torch.equal(torch.sort(e0.long(), -1)[0], torch.sort(e1.long(), -1)[0])
    for a_, b_ in zip(k0, k1):
        assert torch.equal(a_, b_)
    if ppl:
        assert o0 == pytest.approx(o1, rel=1e-6)
    else:
        assert o0 == o1
