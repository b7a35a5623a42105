"""Adapters that put the CUDA path behind the engine interface of `oracle/replay.py`."""
import torch

from easykv_b200.cache import BudgetedKVCache
from easykv_b200.plan import StepParams


class CudaEngine:
    """Every call goes through the C ABI (easykv_b200._lib -> libeasykv_b200.so)."""

    def __init__(self, L, H, Hkv, d, dtype, kernel=0, capacity=1024, arith=0):
        self.cache = BudgetedKVCache(L, 1, H, Hkv, d, capacity, dtype=dtype, device="cuda", arith=arith)
        self.kernel = kernel

    def load_prefill(self, l, K, V, n_scored, C_init, S_init=None, SQ_init=None):
        c = None
        if C_init is not None and n_scored:
            c = C_init[-n_scored:] if n_scored < len(C_init) else C_init
        self.cache.load_prefill(l, K.cuda(), V.cuda(), n_scored, c)
        if S_init is not None:                       # keep_attention seeding
            n = S_init.shape[-1]
            self.cache.S[l][0, :, :n] = S_init.cuda().float()
            self.cache.SQ[l][0, :, :n] = SQ_init.cuda().float()

    def forward(self, l, st, q, k, v, force=None, stream_table=None):
        sp = StepParams.from_fields(st)
        if stream_table is not None:                 # streaming golden: q / k are un-rotated (oracle/replay.py)
            if self.cache.K_raw is None:
                self.cache.enable_streaming(adopt_rotated=True)      # what load_prefill put into K is the raw cache
            flat = lambda x: x.transpose(0, 1).reshape(1, x.shape[1], -1).cuda()       # [heads, ql, d] -> [1, ql, heads*d]
            out, vl = self.cache.step_stream(l, sp, flat(q), flat(k), flat(v), stream_table[0].cuda(), stream_table[1].cuda(),
                                             apply=(force is None), kernel=self.kernel)
        else:
            out, vl = self.cache.step(l, sp, q[None].cuda(), k[None].cuda(), v[None].cuda(),
                                      apply=(force is None), kernel=self.kernel)
        if force is not None and st.evict:
            self.cache.evict(l, force[None])
        return out[0].cpu(), (None if vl is None else vl[0].cpu().long())

    def export(self, l):
        K, V = self.cache.export(l)
        return K[0].cpu(), V[0].cpu()
