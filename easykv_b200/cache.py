"""Device-resident budgeted KV cache of one model (all layers) and the per-forward call into the
CUDA library.

HBM layout per layer (include/easykv_b200.h): K, V `[B, Hkv, cap, d]` in *physical* slot order —
rows never move; eviction frees a slot that the next appended token overwrites — plus fp32 policy
state S, SQ, C `[B, Hkv, cap]` and the int32 map `lidx` physical slot -> logical (arrival-order)
index that carries the reference's order-dependent semantics (protected last-10 / sink / recent
windows, tie order, reported eviction ids).  torch is used for allocation and streams only.

Replaces, together with the kernels: HF `DynamicCache.update` as called from
easykv/llama_patch.py:193-196, `truncate_kv_cache_silo/_liso/truncate_kv_cache`
(easykv/easykv.py:56-82,105-112) and the state tensors of easykv.py:242-247, :412-418.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .plan import StepParams

_DTYPES = {torch.float16: _lib.F16, torch.bfloat16: _lib.BF16, torch.float32: _lib.F32}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class BudgetedKVCache:
    def __init__(self, num_layers, batch, num_heads, num_kv_heads, head_dim, capacity, dtype=torch.float16,
                 device="cuda", arith=0):
        if dtype not in _DTYPES:
            raise ValueError(f"unsupported dtype {dtype}")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("easykv_b200 runs on CUDA devices only (no CPU path)")
        self.lib = _lib.load()
        capacity = (int(capacity) + 7) // 8 * 8      # 16-byte aligned per-head rows of the int32/fp32 arrays
        self.L, self.B, self.H, self.Hkv, self.d, self.cap = num_layers, batch, num_heads, num_kv_heads, head_dim, capacity
        self.dtype, self.device, self.arith = dtype, dev, arith
        self._steady = None                              # (free-slot views, victim-id views) once enable_steady() ran
        kv = dict(dtype=dtype, device=dev)
        st = dict(dtype=torch.float32, device=dev)
        self.K = [torch.zeros(batch, num_kv_heads, capacity, head_dim, **kv) for _ in range(num_layers)]
        self.V = [torch.zeros(batch, num_kv_heads, capacity, head_dim, **kv) for _ in range(num_layers)]
        self.S = [torch.zeros(batch, num_kv_heads, capacity, **st) for _ in range(num_layers)]
        self.SQ = [torch.zeros(batch, num_kv_heads, capacity, **st) for _ in range(num_layers)]
        self.Cn = [torch.zeros(batch, num_kv_heads, capacity, **st) for _ in range(num_layers)]
        self.lidx = [torch.full((batch, num_kv_heads, capacity), -1, dtype=torch.int32, device=dev)
                     for _ in range(num_layers)]
        self.n = [0] * num_layers          # valid slots
        self.n_phys = [0] * num_layers     # streamed physical extent
        self.free = [None] * num_layers    # int32 [B, Hkv, f]: free physical slots inside [0, n_phys)
        self.scratch = None
        self.K_raw = None                  # streaming variant: the un-rotated keys, same physical layout as K
        # streaming variant, decode steps: rotate the cached rows inside the attention kernel (True), always in a separate
        # pass (False), or "auto": fused while the batch is small enough for the cluster-split kernel (latency-bound, where
        # one launch less and one cache pass less win); the persistent kernel's K phase is instruction-bound with the
        # rotation in it (measured 598 vs 434 us per layer-step at 64 sequences), so large batches stay two-pass
        self.fused_streaming = "auto"
        self._shape_cache = [self._shape(l, 1) for l in range(num_layers)]
        self._io_cache = [self._io(l) for l in range(num_layers)]
        self._rope_shape = self._shape(0, 1)

    # ------------------------------------------------------------------------------------------
    def _shape(self, l, q_len):
        return _lib.Shape(dtype=_DTYPES[self.dtype], B=self.B, H=self.H, Hkv=self.Hkv, d=self.d, q_len=q_len,
                          cap=self.cap, n_before=self.n[l], n_phys=self.n_phys[l])

    def entry_limit(self, q_len, evict, kernel=0):
        """Entries per (sequence, kv head) an evicting forward of `q_len` rows can hold (`ekv_chunk_entry_limit`): the
        strided chunk's per-unit tail and the exact kernel select in one CTA's shared memory."""
        return int(self.lib.ekv_chunk_entry_limit(C.byref(self._shape(0, q_len)), int(evict), int(kernel)))

    def check_schedule(self, n_dense, sched, kernel=0):
        """Fail BEFORE the first forward when some evicting forward of the schedule (`plan.schedule` items) would hold
        more entries than the kernels can select among — instead of `EKV_ERR_UNSUPPORTED` in the middle of a prompt."""
        n, limits = n_dense, {}
        for _, q_len, st in sched:
            ev = int(st.evict)
            if ev > 0 and getattr(st, "policy", "roco") not in ("none", "full"):
                key = (q_len, ev)
                if key not in limits:
                    limits[key] = self.entry_limit(q_len, ev, kernel)
                if n + q_len > limits[key]:
                    raise NotImplementedError(
                        f"a forward of {q_len} tokens that evicts {ev} per head over {n + q_len} cache entries exceeds what the "
                        f"kernels select among ({limits[key]} entries for this dtype / head layout; include/easykv_b200.h, "
                        f"ekv_chunk_entry_limit): use a smaller budget or stride, or kv_policy='full'")
            n += q_len - ev
        return True

    def _io(self, l, **kw):
        io = _lib.LayerIO(K=_ptr(self.K[l]), V=_ptr(self.V[l]), S=_ptr(self.S[l]), SQ=_ptr(self.SQ[l]),
                          C=_ptr(self.Cn[l]), lidx=_ptr(self.lidx[l]))
        for k, v in kw.items():
            setattr(io, k, _ptr(v))
        return io

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def free_count(self, l):
        return 0 if self.free[l] is None else self.free[l].shape[-1]

    # ------------------------------------------------------------------------------------------
    def load_prefill(self, l, K, V, n_scored=None, C_init=None):
        """Start layer `l` from a densely prefilled cache K, V `[B, Hkv, n, d]` (or `[Hkv, n, d]`).
        Policy state is zero; C of the last `n_scored` slots is `C_init` (sequence of length
        n_scored) or zero."""
        if K.dim() == 3:
            K, V = K[None], V[None]
        n = K.shape[2]
        if n > self.cap:
            raise ValueError(f"prefill of {n} slots exceeds capacity {self.cap}")
        self.K[l][:, :, :n].copy_(K)
        if self.K_raw is not None:
            self.K_raw[l][:, :, :n].copy_(K)
        self.V[l][:, :, :n].copy_(V)
        self.S[l].zero_(); self.SQ[l].zero_(); self.Cn[l].zero_()
        if C_init is not None and len(C_init):
            c = torch.as_tensor(C_init, dtype=torch.float32, device=self.device)
            self.Cn[l][:, :, n - c.numel():n] = c
        self.lidx[l].fill_(-1)
        self.lidx[l][:, :, :n] = torch.arange(n, dtype=torch.int32, device=self.device)
        self.n[l] = self.n_phys[l] = n
        self.free[l] = None

    def set_counter(self, l, values):
        """C of the first len(values) logical slots (easykv.py:412-418).  Only valid while the layout is
        still dense (right after a prefill: physical slot == logical index)."""
        if self.free_count(l) or self.n[l] != self.n_phys[l]:
            raise RuntimeError("set_counter needs a dense layout")
        c = torch.as_tensor(values, dtype=torch.float32, device=self.device)
        self.Cn[l][:, :, :c.numel()] = c

    def rope_qkv(self, q_in, k_in, v_in, cos, sin, positions=None):
        """RoPE of q and k at explicit positions fused with the re-layout of the projections' outputs
        (`ekv_rope_qk`): q_in `[B, q_len, H*d]`, k_in / v_in `[B, q_len, Hkv*d]` -> q `[B, H, q_len, d]`, k / v
        `[B, Hkv, q_len, d]`.  cos / sin: `[rows, d]` tables indexed by `positions` (int `[B, q_len]`), or already
        gathered per token `[B, q_len, d]` when positions is None.  Replaces apply_rotary_pos_emb
        (easykv/llama_patch.py:47-72) and the transposes around it."""
        B, ql = q_in.shape[0], q_in.shape[1]
        q_in, k_in, v_in = q_in.contiguous(), k_in.contiguous(), v_in.contiguous()
        if cos.dtype != self.dtype:
            cos, sin = cos.to(self.dtype), sin.to(self.dtype)
        cos, sin = cos.contiguous(), sin.contiguous()
        nq, nk = B * self.H * ql * self.d, B * self.Hkv * ql * self.d
        buf = torch.empty(nq + 2 * nk, dtype=self.dtype, device=self.device)           # one allocation for q, k, v
        q = buf[:nq].view(B, self.H, ql, self.d)
        k = buf[nq:nq + nk].view(B, self.Hkv, ql, self.d)
        v = buf[nq + nk:].view(B, self.Hkv, ql, self.d)
        pos = None
        if positions is not None:
            pos = positions if positions.dtype == torch.int32 else positions.to(torch.int32)
            pos = pos.contiguous()
        shape = self._rope_shape
        shape.B, shape.q_len = B, ql
        _lib.check(self.lib.ekv_rope_qk(C.byref(shape), q_in.data_ptr(), k_in.data_ptr(), v_in.data_ptr(), cos.data_ptr(),
                                        sin.data_ptr(), None if pos is None else pos.data_ptr(), q.data_ptr(), k.data_ptr(),
                                        v.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return q, k, v

    # ---- streaming variant (generation_config['streaming'], llama_forward_stream) -----------------------------------
    def enable_streaming(self, adopt_rotated=False):
        """Keep un-rotated keys in a second buffer `K_raw`; `step_stream` re-rotates the whole cache at
        cache-relative positions before every forward (easykv/llama_patch.py:310-327).  `adopt_rotated`: take the
        rows currently in K as the un-rotated ones — what the reference does in 'decoding' mode, where the prompt is
        prefilled by the stock forward (rotated at true positions) before the streaming forward is patched in
        (easykv.py:232 vs :253) and then rotated again on every step."""
        if self.K_raw is None:
            self.K_raw = [torch.zeros_like(k) for k in self.K]
        if adopt_rotated:
            for kr, k in zip(self.K_raw, self.K):
                kr.copy_(k)

    def step_stream(self, l, sp: StepParams, q_in, k_in, v_in, cos, sin, apply=True, kernel=0):
        """One streaming forward of layer `l`.  q_in `[B, q_len, H*d]`, k_in / v_in `[B, q_len, Hkv*d]` are the
        projections' UN-rotated outputs; cos / sin `[rows, d]` the model's table."""
        if self.K_raw is None:
            raise RuntimeError("enable_streaming() first")
        B, ql = q_in.shape[0], q_in.shape[1]
        if self.free_count(l) > ql or 0 < self.free_count(l) < ql:
            self.defragment(l)
        if cos.dtype != self.dtype:
            cos, sin = cos.to(self.dtype), sin.to(self.dtype)
        cos, sin = cos.contiguous(), sin.contiguous()
        n = self.n[l]
        pos = torch.arange(n, n + ql, dtype=torch.int32, device=self.device)[None].expand(B, -1)
        q, k_rot, v = self.rope_qkv(q_in, k_in, v_in, cos, sin, pos)
        k_raw = k_in.view(B, ql, self.Hkv, self.d).transpose(1, 2)
        fused = self.fused_streaming is True or (self.fused_streaming == "auto" and 2 * B * self.Hkv <= 148)
        if fused and ql == 1 and kernel == 0 and self.dtype != torch.float32 and self.d == 128 and n:
            # decode steps: the kernels read K_raw and rotate every cached row at its cache-relative position on the fly
            # (ekv_layer_io.rope_cos / rope_sin / k_new_raw) — no rotated copy of the cache is written or re-read
            try:
                return self.step(l, sp, q, k_rot, v, apply=apply, kernel=kernel, rope=(cos, sin, k_raw.contiguous()))
            except NotImplementedError:
                pass                                      # a shape the fused kernels decline: the two-pass path below
        if n:                                             # K = rope(K_raw, position = logical index)
            shape = self._shape(l, 0)
            io = self._io(l)
            _lib.check(self.lib.ekv_rope_cache(C.byref(shape), C.byref(io), self.K_raw[l].data_ptr(), cos.data_ptr(),
                                               sin.data_ptr(), torch.cuda.current_stream().cuda_stream))
        if self.free_count(l) == ql:                      # the slots the fused kernel will append to
            slots = self.free[l].long()
        else:
            slots = torch.arange(self.n_phys[l], self.n_phys[l] + ql, device=self.device).expand(B, self.Hkv, ql)
        out, vl = self.step(l, sp, q, k_rot, v, apply=apply, kernel=kernel)
        self.K_raw[l].scatter_(2, slots[..., None].expand(B, self.Hkv, ql, self.d), k_raw)
        return out, vl

    # ---- sampling / perplexity tail (easykv_b200/sampling.py) --------------------------------------
    def sample(self, logits, temperature, top_p):
        """Next token `[rows, 1]` int64 from `logits [rows, vocab]`: logits_adapter + torch.multinomial
        (easykv/easykv.py:115-134, :258) as one launch, drawing from torch's CUDA generator exactly as multinomial does."""
        from . import sampling
        return sampling.sample_top_p(logits, temperature, top_p, arith=self.arith)[0]

    def token_nll(self, logits, targets):
        """Per-row cross entropy of one chunk's logits (easykv.py:896-899) — nothing `[prompt, vocab]`-sized is kept."""
        from . import sampling
        return sampling.token_nll(logits, targets)

    def round_state(self, l):
        """Round S / SQ to the model dtype once (what `torch.sum(attention_map, dim=1)` does for the whole dense
        map in h2o_head_score, easykv.py:183-184) after a dense prefill issued with `raw_colsum` chunks."""
        if self.dtype != torch.float32:
            self.S[l].copy_(self.S[l].to(self.dtype).float())
            self.SQ[l].copy_(self.SQ[l].to(self.dtype).float())

    def step(self, l, sp: StepParams, q, k_new, v_new, apply=True, kernel=0, rope=None):
        """One forward of layer `l`: q `[B, H, q_len, d]`, k_new / v_new `[B, Hkv, q_len, d]`
        (post-RoPE).  Returns (out `[B, H, q_len, d]`, victim_lidx `[B, Hkv, evict]` int32 or None).
        With `apply=False` the victims are only reported (the caller may `evict()` others)."""
        q_len = q.shape[2]
        q, k_new, v_new = q.contiguous(), k_new.contiguous(), v_new.contiguous()
        if self.free_count(l) > q_len or 0 < self.free_count(l) < q_len:
            self.defragment(l)
        new_slots = self.free[l] if self.free_count(l) == q_len else None
        if new_slots is None and self.n_phys[l] + q_len > self.cap:
            raise ValueError(f"cache capacity {self.cap} exceeded")
        out = torch.empty_like(q)
        evict = int(sp.evict)
        vs = vl = None
        if evict and self._steady is not None and new_slots is self._steady[0][l] and evict == q_len and apply:
            # steady state (enable_steady): this forward's victim slots overwrite the slot ids it appended at, in place,
            # so every pointer of the launch is the same from forward to forward and the whole model step can be replayed
            # from a CUDA graph (decode: q_len = 1; strided prefill: q_len = stride)
            vs, vl = new_slots, self._steady[1][l]
        elif evict:
            vv = torch.empty(2, self.B, self.Hkv, evict, dtype=torch.int32, device=self.device)
            vs, vl = vv[0], vv[1]
        # the C structs are cached (per layer / per StepParams object) and only the fields that change are
        # rewritten: at batch 1 this call's host time is comparable to the kernels it launches
        shape = self._shape_cache[l]
        shape.q_len, shape.n_before, shape.n_phys = q_len, self.n[l], self.n_phys[l]
        key = (apply, self.arith)
        cached = sp.__dict__.get("_c")
        if cached is None or cached[0] != key:
            cached = sp.__dict__["_c"] = (key, sp.to_c(apply=apply, arith=self.arith))
        cstep = cached[1]
        need = 0
        if q_len > 1 or cstep.tova_head_mean:           # only the chunk kernels and tova's head mean use scratch
            need = self.lib.ekv_scratch_bytes(C.byref(shape), C.byref(cstep))
            if need and (self.scratch is None or self.scratch.numel() < need):
                self.scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        io = self._io_cache[l]
        io.q, io.k_new, io.v_new, io.out = q.data_ptr(), k_new.data_ptr(), v_new.data_ptr(), out.data_ptr()
        io.new_slots = None if new_slots is None else new_slots.data_ptr()
        io.victim_slots = None if vs is None else vs.data_ptr()
        io.victim_lidx = None if vl is None else vl.data_ptr()
        io.scratch = self.scratch.data_ptr() if need else None
        if rope is not None:                              # fused streaming variant: K_raw is the cache, rotated while read
            io.K, io.rope_cos, io.rope_sin, io.k_new_raw = self.K_raw[l].data_ptr(), rope[0].data_ptr(), rope[1].data_ptr(), rope[2].data_ptr()
        try:
            _lib.check(self.lib.ekv_attend_evict(C.byref(shape), C.byref(io), C.byref(cstep), kernel, self._stream()))
        finally:
            if rope is not None:
                io.K, io.rope_cos, io.rope_sin, io.k_new_raw = self.K[l].data_ptr(), None, None, None
        self.n[l] += q_len
        if new_slots is None:
            self.n_phys[l] += q_len
        self.free[l] = None
        if evict and apply:
            self.n[l] -= evict
            self.free[l] = vs
        return out, vl

    def enable_steady(self, q_len=1):
        """Pin the per-layer free-slot / victim buffers for a steady state (append q_len, evict q_len, per forward —
        decoding once the budget is reached, easykv.py:257-363, :670-748; the strided prefill once the cache has its
        size, :426-500, :587-661): needs exactly q_len free slots per (sequence, kv head) in every layer.  From here on
        `step()` launches with identical pointers and shapes every forward, which is what lets `easykv.generate` capture
        the whole model step into one CUDA graph.  Returns the `[L, B, Hkv, q_len]` buffer that holds each forward's
        victim ids.  `disable_steady()` returns to per-forward buffers (mode transitions)."""
        if any(self.free_count(l) != q_len for l in range(self.L)):
            raise RuntimeError(f"enable_steady needs exactly {q_len} free slot(s) per (sequence, kv head) in every layer")
        slots = torch.empty(self.L, self.B, self.Hkv, q_len, dtype=torch.int32, device=self.device)
        victims = torch.empty_like(slots)
        views = ([slots[l] for l in range(self.L)], [victims[l] for l in range(self.L)])
        for l in range(self.L):
            views[0][l].copy_(self.free[l])
            self.free[l] = views[0][l]
        self._steady = views
        self.steady_victims = victims
        return victims

    def disable_steady(self):
        """Leave the steady state: the current free-slot lists move to buffers of their own (the pinned ones stay with
        whatever CUDA graph captured them)."""
        if self._steady is not None:
            for l in range(self.L):
                if self.free[l] is not None:
                    self.free[l] = self.free[l].clone()
            self._steady = None

    def evict(self, l, victims):
        """Delete logical ids `victims` `[B, Hkv, e]` (what truncate_kv_cache_* did)."""
        victims = victims.to(device=self.device, dtype=torch.int32).contiguous()
        if victims.dim() == 2:
            victims = victims[None]
        e = victims.shape[-1]
        vs = torch.empty(self.B, self.Hkv, e, dtype=torch.int32, device=self.device)
        shape = self._shape(l, 0)
        io = self._io(l, victim_slots=vs)
        _lib.check(self.lib.ekv_evict_explicit(C.byref(shape), C.byref(io), _ptr(victims), e, self._stream()))
        self.n[l] -= e
        self.free[l] = vs if self.free[l] is None else torch.cat([self.free[l], vs], dim=-1)

    def select(self, l, sp: StepParams, apply=False):
        """Victims for the current state without a forward (ekv_select)."""
        evict = int(sp.evict)
        vs = torch.empty(self.B, self.Hkv, evict, dtype=torch.int32, device=self.device)
        vl = torch.empty(self.B, self.Hkv, evict, dtype=torch.int32, device=self.device)
        shape = self._shape(l, 0)
        cstep = sp.to_c(apply=apply, arith=self.arith)
        io = self._io(l, victim_slots=vs, victim_lidx=vl)
        _lib.check(self.lib.ekv_select(C.byref(shape), C.byref(io), C.byref(cstep), self._stream()))
        if apply:
            self.n[l] -= evict
            self.free[l] = vs if self.free[l] is None else torch.cat([self.free[l], vs], dim=-1)
        return vl

    def export(self, l, with_state=False):
        """K, V `[B, Hkv, n, d]` in the reference's logical (arrival) order."""
        n = self.n[l]
        Ko = torch.empty(self.B, self.Hkv, n, self.d, dtype=self.dtype, device=self.device)
        Vo = torch.empty_like(Ko)
        st = [torch.empty(self.B, self.Hkv, n, dtype=torch.float32, device=self.device) for _ in range(3)] if with_state else [None] * 3
        shape = self._shape(l, 0)
        io = self._io(l)
        if self.K_raw is not None:                        # streaming: the cache's keys are the un-rotated ones
            io.K = self.K_raw[l].data_ptr()
        _lib.check(self.lib.ekv_export_logical(C.byref(shape), C.byref(io), _ptr(Ko), _ptr(Vo), _ptr(st[0]),
                                               _ptr(st[1]), _ptr(st[2]), self._stream()))
        return (Ko, Vo, *st) if with_state else (Ko, Vo)

    def defragment(self, l):
        """Make the physical layout dense again (valid slots at [0, n) in logical order).  Used at
        mode transitions (e.g. after the strided phase of encoding_decoding leaves stride-1 free
        slots that a one-token-per-step decode loop would otherwise stream forever)."""
        n = self.n[l]
        Ko, Vo, S, SQ, Cn = self.export(l, with_state=True)
        (self.K if self.K_raw is None else self.K_raw)[l][:, :, :n].copy_(Ko); self.V[l][:, :, :n].copy_(Vo)
        self.S[l][:, :, :n].copy_(S); self.SQ[l][:, :, :n].copy_(SQ); self.Cn[l][:, :, :n].copy_(Cn)
        self.lidx[l].fill_(-1)
        self.lidx[l][:, :, :n] = torch.arange(n, dtype=torch.int32, device=self.device)
        self.n_phys[l] = n
        self.free[l] = None


class SteadyStep:
    """The steady state of a budgeted loop: every forward appends `q_len` tokens per sequence and evicts `q_len` per
    (sequence, layer, kv head), so shapes never change — the decode step of `encoding_decoding` / `decoding`
    (q_len = 1; easykv.py:670-748, :257-363 once the budget is reached) and the strided-prefill chunk once the cache
    has reached its size (q_len = stride; easykv.py:426-500, :587-661).  All C-ABI arguments are built once; the
    per-layer victim buffer of step t is the new-slot buffer of step t+1 (no host round trip), and the whole L-layer
    step can be captured into one CUDA graph."""

    def __init__(self, cache: BudgetedKVCache, sp: StepParams, q, k_new, v_new, out=None):
        """q `[L, B, H, q_len, d]`, k_new / v_new `[L, B, Hkv, q_len, d]`: device buffers the caller refills
        before each `run()` (e.g. the projections' outputs)."""
        c = self.cache = cache
        ql = self.q_len = q.shape[3]
        assert sp.evict == ql, "steady state: as many victims as appended tokens"
        self.q, self.k_new, self.v_new = q, k_new, v_new
        self.out = torch.empty_like(q) if out is None else out
        self.victim_lidx = torch.empty(c.L, c.B, c.Hkv, ql, dtype=torch.int32, device=c.device)
        self.slots = torch.empty(c.L, c.B, c.Hkv, ql, dtype=torch.int32, device=c.device)
        self.cstep = sp.to_c(apply=True, arith=c.arith)
        self.calls = []
        for l in range(c.L):
            if c.free_count(l) != ql:
                # bring the layer into the steady state: one append-mode evicting step
                o, _ = c.step(l, sp, q[l], k_new[l], v_new[l])
                self.out[l].copy_(o)
            self.slots[l].copy_(c.free[l])
            c.free[l] = self.slots[l]
            shape = c._shape(l, ql)
            need = c.lib.ekv_scratch_bytes(C.byref(shape), C.byref(self.cstep)) if (ql > 1 or self.cstep.tova_head_mean) else 0
            if need and (c.scratch is None or c.scratch.numel() < need):
                c.scratch = torch.empty(need, dtype=torch.uint8, device=c.device)
            io = c._io(l, q=q[l], k_new=k_new[l], v_new=v_new[l], out=self.out[l], new_slots=self.slots[l],
                       victim_slots=self.slots[l], victim_lidx=self.victim_lidx[l])
            self.calls.append((shape, io))
        for shape, io in self.calls:                  # the scratch may have been re-allocated while the list was built
            io.scratch = c.scratch.data_ptr() if c.scratch is not None else None
        self.graph = None

    def run(self, stream=None):
        lib, cs = self.cache.lib, self.cstep
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        for shape, io in self.calls:
            rc = lib.ekv_attend_evict(C.byref(shape), C.byref(io), C.byref(cs), 0, s)
            if rc:
                _lib.check(rc)

    def run_layer(self, l, stream=None):
        """One layer's launch (for callers that pipeline per-layer host<->device copies around the step)."""
        shape, io = self.calls[l]
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream if stream is None else stream)
        rc = self.cache.lib.ekv_attend_evict(C.byref(shape), C.byref(io), C.byref(self.cstep), 0, s)
        if rc:
            _lib.check(rc)

    def capture(self, pre=None):
        """Capture one whole step (L forwards) into a CUDA graph; `replay()` then costs one launch.  `pre`: device-side
        work captured ahead of the step (e.g. refreshing q / k_new / v_new from a pool, as bench.py does)."""
        if pre is not None:
            pre()
        self.run()                                   # warm: function attributes are set outside capture
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            if pre is not None:
                pre()
            self.run()
        return self

    def replay(self):
        self.graph.replay()


SteadyDecode = SteadyStep


class RaggedDecode:
    """Decode steps over a batch whose sequences hold DIFFERENT numbers of slots (the reference is hard-wired to one
    sequence; SURVEY §8b proposed per-sequence lengths).  ABI: `ekv_layer_io.seq_n_before` + `ekv_step.budget_gate` — a
    sequence below the budget only accumulates and appends (at its own physical extent), one above it evicts and the
    next token reuses the victim's slot, exactly as `if cur_kv_size - len(prefix) > budget` (easykv.py:303) decides for
    a single sequence.  No host synchronisation: the per-sequence counts follow from the lengths and the gate."""

    def __init__(self, cache: BudgetedKVCache, lengths, budget_gate: int):
        c = self.cache = cache
        assert len(lengths) == c.B
        self.gate = int(budget_gate)
        self.n = [[int(x) for x in lengths] for _ in range(c.L)]             # valid slots per layer and sequence
        self.ext = [[int(x) for x in lengths] for _ in range(c.L)]           # physical extent per layer and sequence
        self.victims = [None] * c.L                                          # last step's victim slots [B, Hkv, 1] (-1: none)
        self.had = [[False] * c.B for _ in range(c.L)]                       # ... and whether sequence b evicted in it

    def load_prefill(self, l, Ks, Vs, C_inits=None):
        """Sequence b starts from Ks[b], Vs[b] `[Hkv, n_b, d]`; slots beyond n_b are free."""
        c = self.cache
        c.S[l].zero_(); c.SQ[l].zero_(); c.Cn[l].zero_(); c.lidx[l].fill_(-1)
        for b, (K, V) in enumerate(zip(Ks, Vs)):
            n = K.shape[1]
            assert n == self.n[l][b]
            c.K[l][b, :, :n].copy_(K); c.V[l][b, :, :n].copy_(V)
            c.lidx[l][b, :, :n] = torch.arange(n, dtype=torch.int32, device=c.device)
            if C_inits is not None:
                ci = torch.as_tensor(C_inits[b], dtype=torch.float32, device=c.device)
                c.Cn[l][b, :, n - ci.numel():n] = ci
        c.n[l] = c.n_phys[l] = max(self.n[l])
        c.free[l] = None
        self.victims[l] = None
        self.had[l] = [False] * c.B

    def step(self, l, sp: StepParams, q, k_new, v_new, kernel=0):
        """q `[B, H, 1, d]`, k_new / v_new `[B, Hkv, 1, d]`.  Returns (out, victim_lidx `[B, Hkv, 1]`, -1 where the
        sequence did not evict)."""
        from dataclasses import replace
        c = self.cache
        sp = replace(sp, budget_gate=self.gate)
        n, ext = self.n[l], self.ext[l]
        dev = c.device
        seq_n = torch.tensor(n, dtype=torch.int32, device=dev)
        own = torch.tensor(ext, dtype=torch.int32, device=dev)[:, None, None].expand(c.B, c.Hkv, 1)
        slots = own.contiguous() if self.victims[l] is None else torch.where(self.victims[l] >= 0, self.victims[l], own).contiguous()
        out = torch.empty_like(q)
        vv = torch.empty(2, c.B, c.Hkv, 1, dtype=torch.int32, device=dev)
        shape = _lib.Shape(dtype=_DTYPES[c.dtype], B=c.B, H=c.H, Hkv=c.Hkv, d=c.d, q_len=1, cap=c.cap, n_before=max(n),
                           n_phys=max(ext))
        io = c._io(l, q=q.contiguous(), k_new=k_new.contiguous(), v_new=v_new.contiguous(), out=out, new_slots=slots,
                   victim_slots=vv[0], victim_lidx=vv[1], seq_n_before=seq_n)
        cstep = sp.to_c(apply=True, arith=c.arith)
        _lib.check(c.lib.ekv_attend_evict(C.byref(shape), C.byref(io), C.byref(cstep), kernel, c._stream()))
        had = self.had[l]
        for b in range(c.B):
            if not had[b]:                                                   # appended at its own extent (no victim slot to reuse)
                ext[b] += 1
            if self.gate > 0 and n[b] + 1 - sp.score_offset <= self.gate:    # below the budget: grew by one, nothing evicted
                n[b] += 1
                had[b] = False
            else:
                had[b] = True
        self.victims[l] = vv[0]
        c.n[l], c.n_phys[l] = max(n), max(ext)
        return out, vv[1]

