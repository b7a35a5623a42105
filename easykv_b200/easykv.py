"""User surface of the budgeted-KV path — same names, signatures, config keys and return types as the
reference's `easykv/easykv.py`:

    enable_fixed_kv(model, tokenizer, mode, stride=1, verbose=False)        easykv/easykv.py:903-908
    model.easykv_generate(input_ids=LongTensor[1, len], generation_config=dict) -> str    :199-753
    model.easykv_ppl(input_ids=..., generation_config=dict) -> float                      :754-901

What is different underneath: the reference's loops call the HF model with `output_attentions=True`,
pull `[1, H, q, n]` probability tensors back into Python, fold / accumulate / select with ~135 host
syncs per token and rebuild the cache by boolean indexing (easykv.py:271-362).  Here every forward is ONE
`ekv_attend_evict` launch per layer (issued from the attention seam, `attention.py`) driven by a
`StepParams` computed on the host from integers only (`plan.py`); nothing is read back except the
sampled token.

Mode loops mirrored (control flow only — sequencing and sampling are host glue):
  decoding           :228-366     encoding            :367-529
  encoding_decoding  :530-753     ppl                 :754-901     auto dispatch :220-227
Sampling (`logits_adapter` + torch.multinomial; :115-134, :258) and the perplexity loss (:896-899) are one launch each
(easykv_b200/sampling.py).
"""
from __future__ import annotations

import dataclasses
import functools
import math
import statistics
import time
import traceback

import torch

from . import plan as P
from .attention import find_attention_modules, geometry, patched_attention
from .cache import BudgetedKVCache


def logits_adapter(logits: torch.Tensor, temperature: float, top_p: float):
    """Temperature scaling and top-p renormalisation; returns (sampling distribution, raw softmax) — the contract of
    easykv/easykv.py:115-134, computed by one `ekv_sample_top_p` launch (easykv_b200/sampling.py; CUDA tensors only)."""
    from . import sampling
    return sampling.logits_adapter(logits, temperature, top_p)


class Session:
    """State shared between the driver loop and the attention seam during one `generate` call."""

    def __init__(self, model, cache: BudgetedKVCache, record=True):
        self.config = getattr(model, "config", None)
        self.cache = cache
        self.step = P.StepParams()
        self.pos0 = 0
        self.q_len = 0
        self.max_position = 0
        self.two_tuple = False
        self.streaming = False
        self.rotary = getattr(getattr(model, "model", model), "rotary_emb", None)   # transformers 5.x: model-level RoPE module
        self._table = None
        self.fwd = 0
        self.events = [] if record else None     # (forward index, int32 [L, B, Hkv, evict] victim ids)
        self.model_kwargs = {}
        self.graph_error = None                  # why the decode step could not be captured into a CUDA graph, if so
        self.graphed_steps = 0
        self.graphed_chunks = 0
        self.graph_capture_s = 0.0
        self.token_times = []                    # host time at which each generated token reached the host
        self._cur = None

    def begin(self, step: P.StepParams, pos0: int, q_len: int):
        self.step, self.pos0, self.q_len = step, pos0, q_len
        self.max_position = max(self.max_position, pos0 + q_len - 1)
        self.fwd += 1
        self._cur = [] if (self.events is not None and step.evict) else None

    def stream_table(self, module, x, rows):
        """cos / sin `[rows, d]` for cache-relative positions 0..rows-1 (streaming variant)."""
        if hasattr(module, "rotary_emb") and self.rotary is None:          # 4.36-style per-module table
            return module.rotary_emb(x, seq_len=rows)
        if self._table is None or self._table[0].shape[0] < rows:
            want = max(rows, self.cache.cap + 1)
            pos = torch.arange(want, device=x.device)[None]
            cos, sin = self.rotary(x, pos)
            self._table = (cos[0], sin[0])
        return self._table

    def position_ids(self, device):
        return torch.arange(self.pos0, self.pos0 + self.q_len, device=device)[None]

    def record(self, layer, victims):
        if self._cur is not None and victims is not None:
            self._cur.append(victims)

    def end(self):
        if self._cur:
            self.events.append((self.fwd, torch.stack(self._cur)))
        self._cur = None


class GraphedDecodeStep:
    """The whole model's steady step (`q_len` = 1: a decode step; `q_len` = stride: a strided-prefill chunk) — every projection, MLP, norm and the per-layer `ekv_rope_qk` + `ekv_attend_evict`
    launches — captured once into a CUDA graph and replayed per generated token.  Valid in the steady state of decoding
    (append one, evict one, identical step parameters — `BudgetedKVCache.enable_steady`), where no shape, pointer or host
    scalar changes from step to step: the token and its position are refreshed in two static device buffers.  The
    reference pays ~135 host syncs per token in this loop (easykv.py:271-362); eagerly this package pays the HF modules'
    launch overhead (≈10 ms per token at batch 1 on a 7B model); replayed, the step is one launch."""

    MIN_STEPS = 128       # capture costs 0.06-0.5 s (measured, 7B shape) and saves ~4 ms per token at batch 1
    MIN_CHUNKS = 16       # strided-prefill chunks left for a chunk capture to pay off

    def __init__(self, model, sess, cache, step, bsz, device, q_len=1):
        self.model, self.sess, self.cache, self.step, self.q_len = model, sess, cache, step, q_len
        self.ids = torch.zeros(bsz, q_len, dtype=torch.int64, device=device)
        self.pos = torch.zeros(bsz, q_len, dtype=torch.int64, device=device)
        self.offs = torch.arange(q_len, dtype=torch.int64, device=device)[None]
        self.graph = self.logits = None

    def capture(self):
        cache, sess = self.cache, self.sess
        self.victims = cache.enable_steady(self.q_len)
        sess.begin(self.step, sess.pos0, self.q_len)
        sess.fwd -= 1                                   # capturing executes nothing
        sess._cur = None
        # the raw capture API: `torch.cuda.graph` would first gc.collect() and empty the caching allocator — seconds
        # after a long prefill (measured 0.06-3.4 s), more than the capture itself
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=self.ids.device)
        side.wait_stream(torch.cuda.current_stream(self.ids.device))
        with torch.cuda.stream(side):
            graph.capture_begin()
            try:
                logits = self.model(input_ids=self.ids, position_ids=self.pos, use_cache=False, **sess.model_kwargs).logits
            except BaseException:
                try:
                    graph.capture_end()                # leave the stream out of capture mode before reporting
                except Exception:
                    pass
                raise
            graph.capture_end()
        torch.cuda.current_stream(self.ids.device).wait_stream(side)
        self.graph, self.logits = graph, logits
        return self

    def run(self, ids, pos0, all_rows=False):
        sess = self.sess
        self.ids.copy_(ids)
        if self.q_len == 1:
            self.pos.fill_(pos0)
        else:
            torch.add(self.offs, pos0, out=self.pos[:1])
            if self.pos.shape[0] > 1:
                self.pos[1:] = self.pos[:1]
        sess.begin(self.step, pos0, self.q_len)
        sess._cur = None
        self.graph.replay()
        if sess.events is not None:
            sess.events.append((sess.fwd, self.victims.clone()))
        return self.logits if all_rows else self.logits[:, -1, :]


DENSE_CHUNK = 256     # tokens per forward of the dense (no-eviction) prefill (generation_config['dense_chunk']): each forward
                      # streams the whole model's weights once, so fewer, larger forwards (64 -> 256: 7B 4096-token prefill 0.60 s)


@torch.inference_mode()
def generate(self, input_ids, generation_config, kv_mode="encoding", stride=1, report_decoding_latency=False):
    cfg = generation_config
    temperature = cfg.get("temperature", 1.0)                 # easykv/easykv.py:201-210
    top_p = cfg.get("top_p", 1.0)
    max_new_tokens = cfg.get("max_new_tokens", 1024)
    budget = cfg.get("budget", 0.5)
    policy = cfg.get("kv_policy", "recency")
    temp_length = cfg.get("temp_length", 4)
    recent_ratio = cfg.get("recent_ratio", 0.1)
    keep_attention = cfg.get("keep_attention", False)
    eos_token_ids = cfg.get("eos_token_ids", [self.tokenizer.eos_token_id])
    streaming = bool(cfg.get("streaming", False))             # llama_forward_stream / mistral_forward_stream
    policy = P.canonical_policy(policy)
    if input_ids.dim() == 1:
        input_ids = input_ids[None]
    bsz, length = input_ids.shape
    ppl_mode = kv_mode == "ppl"
    plan = P.resolve_plan(kv_mode, length, budget, stride, recent_ratio, temp_length)
    # keep_attention=True: the reference seeds S / SQ with the column sums of the dense prefill's [r_idx, r_idx]
    # attention map (h2o_head_score, easykv.py:173-186; 137 GB of maps for Mistral at 16K).  Here the dense
    # prefill's chunks accumulate those sums in-kernel (raw_colsum) and nothing is materialised.
    seed = bool(keep_attention) and plan.mode in ("encoding", "encoding_decoding", "ppl") and policy in ("roco", "h2o_head")

    mods = find_attention_modules(self)
    H, Hkv, d = geometry(mods[0], getattr(self, "config", None))
    param = next(self.parameters())
    device, dtype = param.device, param.dtype
    input_ids = input_ids.to(device)
    sched = list(P.schedule(plan, policy, max_new_tokens, keep_attention))
    n_dense = length if plan.mode in ("decoding", "dense") else plan.r_idx
    # the largest cache length any forward reaches (a schedule that never evicts — kv_policy='full' — grows to the prompt)
    capacity = max(plan.capacity + (max_new_tokens if plan.mode in ("dense", "encoding") else 0),
                   P.required_capacity(n_dense, sched)) + 1
    # "aten_arith": which ATen flavour's two non-associative spots to reproduce (include/easykv_b200.h,
    # ekv_step.arith): the CUDA kernels' (default — what the reference does on a GPU) or the CPU kernels'
    arith = {"cuda": 1, "cpu": 0}[cfg.get("aten_arith", "cuda")]
    cache = BudgetedKVCache(len(mods), bsz, H, Hkv, d, capacity, dtype=dtype, device=device, arith=arith)
    # an evicting forward selects among all of a head's entries inside one CTA (strided chunks) or one cluster (decode
    # steps): a schedule that would exceed that fails here, before the first forward, not in the middle of the prompt
    if hasattr(cache, "check_schedule"):
        cache.check_schedule(n_dense, sched)
    sess = Session(self, cache, record=cfg.get("record_evictions", True))
    if sess.rotary is not None:
        # transformers >= 4.48 builds a causal mask per forward — element-wise launches plus a host sync (its packed-
        # sequence check, masking_utils.find_packed_sequence_indices) — that the seam never reads; an already-4D mask
        # makes it return at once, which also keeps the forward capturable into a CUDA graph
        sess.model_kwargs = {"attention_mask": torch.zeros(bsz, 1, 1, 1, dtype=torch.bool, device=device)}
    self.easykv_last = sess                                   # eviction trace / cache of the last call
    if streaming and plan.mode != "decoding":
        # un-rotated keys in the cache, RoPE re-applied at cache-relative positions on every forward
        cache.enable_streaming()
        sess.streaming = True

    def forward(ids, pos0, step):
        if step.policy == "random" and step.evict:
            # the victim range is the host's draw, exactly where the reference draws it (plan.random_range_start)
            n_after = cache.n[0] + ids.shape[1]
            n_state = n_after - step.score_offset if plan.mode == "decoding" else plan.idx + plan.stride
            step = dataclasses.replace(step, range_start=P.random_range_start(plan, n_state))
        sess.begin(step, pos0, ids.shape[1])
        pos = torch.arange(pos0, pos0 + ids.shape[1], device=device)[None].expand(bsz, -1)
        out = self(input_ids=ids, position_ids=pos, use_cache=False, **sess.model_kwargs)
        sess.end()
        return out.logits

    def dense_prefill(upto):
        """Causal attention over tokens [0, upto) with no policy: the reference's unpatched / patched dense
        forward (easykv.py:232, :396, :557).  Chunked so that every forward is one fused launch per layer."""
        logits = None
        sp = P.StepParams(policy=policy, accumulate=True, raw_colsum=True) if seed else P.StepParams()
        dc = int(cfg.get("dense_chunk", DENSE_CHUNK))
        for t0 in range(0, upto, dc):
            logits = forward(input_ids[:, t0:min(t0 + dc, upto)], t0, sp)
        if seed:
            for l in range(cache.L):
                cache.round_state(l)
        return logits

    def sample(last_logits):
        # logits_adapter + torch.multinomial (easykv.py:115-134, :258): one launch, same generator draw, no host sync
        return cache.sample(last_logits, temperature, top_p)

    with patched_attention(self, sess):
        chunks = [s for s in sched if s[0] == "chunk"]
        decodes = [s for s in sched if s[0] == "decode"]
        all_nll = []

        def next_ids(t0, q_len):
            """Targets of rows t0 .. t0+q_len-1: the following token; the prompt's last row has none (dropped, :899)."""
            t = input_ids[0, t0 + 1:t0 + q_len + 1]
            return t if t.numel() == q_len else torch.cat([t, t.new_zeros(q_len - t.numel())])

        if ppl_mode and plan.mode == "dense":                 # easykv.py:759-765
            lp = [cache.token_nll(forward(input_ids[:, t0:t0 + DENSE_CHUNK], t0, P.StepParams())[0].float(),
                                  next_ids(t0, min(DENSE_CHUNK, length - t0))) for t0 in range(0, length, DENSE_CHUNK)]
            return math.exp(statistics.mean(torch.cat(lp)[:-1].cpu().numpy().tolist()))
        # ---- prompt --------------------------------------------------------------------------------------
        logits = dense_prefill(n_dense)
        if streaming and plan.mode == "decoding":
            # the reference prefills the prompt with the stock forward BEFORE patching (easykv.py:232 vs :253): the
            # cache holds keys rotated at their true positions, which the streaming forward then treats as
            # un-rotated and rotates again on every step — reproduced as is
            cache.enable_streaming(adopt_rotated=True)
            sess.streaming = True
        C0 = P.initial_counter(plan, keep_attention)
        if C0 is not None:                                     # strided modes keep state for the prompt
            for l in range(cache.L):
                cache.set_counter(l, C0)
        cur = n_dense
        # steady strided prefill: once the cache has its size every chunk appends `stride` rows and evicts `stride`
        # (identical step parameters, the victims' slots are the next chunk's new slots): the whole model forward of a
        # chunk is captured into one CUDA graph and replayed (the reference pays ~135 host syncs per chunk here)
        steady_chunk_from = len(chunks)
        while steady_chunk_from > 0 and chunks[steady_chunk_from - 1][2] == chunks[-1][2] and chunks[-1][2].evict == chunks[-1][1]:
            steady_chunk_from -= 1
        chunk_graph_ok = (bool(cfg.get("cuda_graph", True)) and device.type == "cuda" and hasattr(cache, "enable_steady")
                          and sess.rotary is not None and not sess.streaming and bool(chunks) and policy != "random"
                          and not (chunks[-1][2].policy == "tova" and chunks[-1][2].tova_head_mean))
        graphed_chunk = None
        for ci, (_, q_len, st) in enumerate(chunks):           # easykv.py:426-500 / :587-661 / :816-892
            if (chunk_graph_ok and graphed_chunk is None and ci > steady_chunk_from
                    and len(chunks) - ci >= int(cfg.get("cuda_graph_min_chunks", GraphedDecodeStep.MIN_CHUNKS))
                    and all(cache.free_count(l) == q_len for l in range(cache.L))):
                try:
                    t_cap = time.perf_counter()
                    graphed_chunk = GraphedDecodeStep(self, sess, cache, st, bsz, device, q_len=q_len).capture()
                    sess.graph_capture_s += time.perf_counter() - t_cap
                except Exception as exc:
                    chunk_graph_ok, graphed_chunk = False, None
                    sess.graph_error = f"{type(exc).__name__}: {exc}\n" + "".join(traceback.format_tb(exc.__traceback__)[-6:])
                    torch.cuda.synchronize(device)
                    cache.disable_steady()
            if graphed_chunk is not None:
                logits = graphed_chunk.run(input_ids[:, cur:cur + q_len], cur, all_rows=True)
                sess.graphed_chunks += 1
            else:
                logits = forward(input_ids[:, cur:cur + q_len], cur, st)
            if ppl_mode:                                       # easykv.py:826-827: this chunk's rows of the final loss
                all_nll.append(cache.token_nll(logits[0].float(), next_ids(cur, q_len)))
            cur += q_len
        if graphed_chunk is not None:
            cache.disable_steady()                             # the decode phase pins its own buffers
            logits = logits.clone()
        retained = cache.n[0]
        if plan.mode in ("encoding", "ppl") or (plan.mode == "dense" and kv_mode == "encoding"):
            # easykv.py:501-503 sits outside the budget if/else: 'encoding' prints the line for a dense prefill too
            print(f"KV cache budget ratio: {retained / length * 100:.2f}%({retained}/{length})")
        if ppl_mode:                                           # easykv.py:896-901
            return math.exp(statistics.mean(torch.cat(all_nll)[:-1].cpu().numpy().tolist()))
        # ---- generation ------------------------------------------------------------------------------------
        if cfg.get("record_timing", False):                    # for benchmarks: when the prompt phase was complete
            torch.cuda.synchronize(device)
            sess.t_prompt_done = time.perf_counter()
        last = logits[:, -1, :]
        output_ids, times = [], []
        cur_pos = length
        if plan.mode == "encoding_decoding" and policy == "random" and decodes:
            # the reference itself fails here (UnboundLocalError: positions_tensor, easykv.py:744)
            raise NotImplementedError("kv_policy='random' has no decode phase in encoding_decoding / auto mode")
        # steady state: from `steady_from` on every step has the same parameters and evicts one slot per head
        steady_from = len(decodes)
        while steady_from > 0 and decodes[steady_from - 1][2] == decodes[-1][2]:
            steady_from -= 1
        graph_ok = (bool(cfg.get("cuda_graph", True)) and device.type == "cuda" and hasattr(cache, "enable_steady")
                    and sess.rotary is not None and not sess.streaming and bool(decodes)
                    and decodes[-1][2].evict == 1 and decodes[-1][2].policy != "random")
        graphed = None
        for i, (_, _, st) in enumerate(decodes):               # easykv.py:257-363 / :508-526 / :670-748
            nxt = sample(last)
            output_ids.append(nxt[:, 0].tolist())
            sess.token_times.append(time.perf_counter())      # the read-back above is the step's only host sync
            if bsz == 1 and output_ids[-1][0] in eos_token_ids:
                break
            t0 = time.time()
            if (graph_ok and graphed is None and i > steady_from and len(decodes) - i >= int(cfg.get("cuda_graph_min_steps", GraphedDecodeStep.MIN_STEPS))
                    and all(cache.free_count(l) == 1 for l in range(cache.L))):
                try:
                    t_cap = time.perf_counter()
                    graphed = GraphedDecodeStep(self, sess, cache, st, bsz, device).capture()
                    sess.graph_capture_s = time.perf_counter() - t_cap
                except Exception as exc:               # e.g. a host sync inside this model class's forward: stay eager
                    graph_ok, graphed = False, None
                    sess.graph_error = f"{type(exc).__name__}: {exc}\n" + "".join(traceback.format_tb(exc.__traceback__)[-6:])
                    torch.cuda.synchronize(device)
            if graphed is not None:
                last = graphed.run(nxt, cur_pos)
                sess.graphed_steps += 1
            else:
                last = forward(nxt, cur_pos, st)[:, -1, :]
            if report_decoding_latency:
                torch.cuda.synchronize(device)
                times.append(time.time() - t0)
            cur_pos += 1
        n_out = len(output_ids)
        size = cache.n[0]
        if plan.mode == "decoding":
            kept = size - length
            print(f"KV cache budget ratio: {kept / max(n_out, 1) * 100:.2f}%({kept}/{n_out})")
        elif plan.mode == "encoding_decoding":
            print(f"KV Cache Budget ratio {size / (length + n_out) * 100:.2f}%[{size}/({length}+{n_out})]")
        if report_decoding_latency and len(times) > 1:
            print(f"Per-step decoding latency: {statistics.mean(times[1:]):.3f}")
    texts = [self.tokenizer.decode([step[b] for step in output_ids], skip_special_tokens=True).strip()
             for b in range(bsz)]
    return texts[0] if bsz == 1 else texts


def enable_fixed_kv(model, tokenizer, mode, stride=1, verbose=False):
    """easykv/easykv.py:903-908."""
    if mode not in ("decoding", "encoding", "auto", "ppl", "encoding_decoding"):
        raise ValueError(f"unknown mode {mode!r}")
    model.tokenizer = tokenizer
    model.easykv_generate = functools.partial(generate, self=model, kv_mode=mode, stride=stride,
                                              report_decoding_latency=verbose)
    model.easykv_ppl = functools.partial(generate, self=model, kv_mode="ppl", stride=stride)
    print(f"Fixed KV Cache for {mode} enabled")
