// Per-(sequence, kv-head) policy-state update, budgeted victim select and in-place eviction,
// executed by one thread group out of shared memory.  Shared by the fused attention kernels
// and the standalone select / explicit-evict kernels.
//
// Reference semantics (paths relative to the reference root):
//   accumulate   easykv/easykv.py:288-300, 443-457, 603-618, 693-707
//   counter      :304, :460, :708
//   roco select  :320-324 (decode), :471-476 (strided)      std -> k smallest -> smallest mean(s)
//   h2o / tova   :311, :335, :463, :485                      smallest S inside a logical window
//   recency      :343-347, :491-493, :741-742                contiguous logical range
//   compaction   :56-82 (K/V), :315-333, :465-483 (state)    here: renumber lidx, free the slots
// Ordering: every "k smallest" is taken in (value ascending, NaN last, logical index
// ascending) order; the second roco stage breaks equal means by (std, logical index), which is
// what argmin/topk over a std-sorted candidate list yields.
//
// No sort: each "k smallest" is an MSB-first 8-bit radix select over order-preserving uint32
// keys held in shared memory (4 passes), followed by tie levels only when the cut falls inside a
// run of equal keys.
#pragma once
#include "ekv_common.cuh"
#include "ekv_bucket.cuh"

namespace ekv {

struct UnitState {      // this (sequence, kv head)'s slices
  float* S;
  float* SQ;
  float* C;
  const float* S_in = nullptr;   // optional staged copies (shared memory) of S/SQ/C[0, n_phys) to read from
  const float* SQ_in = nullptr;
  const float* C_in = nullptr;
  int32_t* lidx;
  const int32_t* new_slots;   // [q_len] or nullptr
  int32_t* victim_slots;      // [evict] or nullptr
  int32_t* victim_lidx;       // [evict] or nullptr
};

struct SelScratch {     // shared memory, carved by the caller; NE = n_phys + q_len entries
  int32_t* lj;          // [NE]  absolute logical index of entry e, -1 = free slot
  uint32_t* keyA;       // [NE]
  uint32_t* keyB;       // [NE]
  uint8_t* flag;        // [NE]  bit0 candidate, bit1 feasible, bit2 chosen
  uint32_t* hist;       // [256]
  int32_t* vl;          // [2 * max(evict,1)]  victim logical ids: unsorted | sorted
  int32_t* vs;          // [max(evict,1)]      victim physical slots (unsorted)
  int32_t* misc;        // [8]
  unsigned long long* red;  // [2 * 32]
  unsigned long long* dbg = nullptr;   // profiling hook: 8 clock64 stamps of the tail's stages
  static __host__ __device__ size_t bytes(int NE, int evict) {
    int ev = evict > 0 ? evict : 1;
    size_t b = (size_t)NE * 4 * 2;                 // keyA, keyB   (lj is carved separately by callers that preload it)
    b += ((size_t)NE + 15) / 16 * 16;              // flag
    b += 256 * 4 + (size_t)ev * 3 * 4 + 8 * 4 + 64 * 8 + 64;
    return (b + 15) / 16 * 16;
  }
  // carve everything except lj from `base` (16-byte aligned)
  __device__ void carve(void* base, int NE, int evict) {
    int ev = evict > 0 ? evict : 1;
    char* p = reinterpret_cast<char*>(base);
    red = reinterpret_cast<unsigned long long*>(p); p += 64 * 8;
    keyA = reinterpret_cast<uint32_t*>(p); p += (size_t)NE * 4;
    keyB = reinterpret_cast<uint32_t*>(p); p += (size_t)NE * 4;
    hist = reinterpret_cast<uint32_t*>(p); p += 256 * 4;
    vl = reinterpret_cast<int32_t*>(p); p += (size_t)ev * 2 * 4;
    vs = reinterpret_cast<int32_t*>(p); p += (size_t)ev * 4;
    misc = reinterpret_cast<int32_t*>(p); p += 8 * 4;
    flag = reinterpret_cast<uint8_t*>(p);
  }
};

enum { F_CAND = 1, F_FEAS = 2, F_CHOSEN = 4, F_REJ = 8 };

// m-th smallest (1-based) of key(e) over {e : pred(e)}: returns the threshold T, how many of the
// entries equal to T belong to the m smallest (`need`), and how many entries equal T (`tcount`).
template <class Key, class Pred>
__device__ __forceinline__ void radix_select(int NE, int m, Key key, Pred pred, const SelScratch& c, const Grp& g,
                                             uint32_t& T, int& need, int& tcount) {
  uint32_t prefix = 0, mask = 0;
  int rem = m;
  tcount = 0;
#pragma unroll 1
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = g.tid; i < 256; i += g.n) c.hist[i] = 0;
    g.sync();
    for (int e = g.tid; e < NE; e += g.n) {
      if (pred(e)) {
        uint32_t k = key(e);
        if ((k & mask) == prefix) atomicAdd(&c.hist[(k >> shift) & 255u], 1u);
      }
    }
    g.sync();
    if (g.tid < 32) {
      const int lane = g.tid;
      uint32_t loc[8], s = 0;
#pragma unroll
      for (int j = 0; j < 8; ++j) { loc[j] = c.hist[lane * 8 + j]; s += loc[j]; }
      uint32_t inc = s;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      uint32_t exc = inc - s;
      uint32_t total = __shfl_sync(0xffffffffu, inc, 31);
      uint32_t want = (uint32_t)rem;
      if (want > total) want = total;          // fewer candidates than requested: take them all
      if (want == 0) want = 1;
      if (exc < want && want <= inc) {
        uint32_t cum = exc;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (want <= cum + loc[j]) {
            c.misc[0] = lane * 8 + j; c.misc[1] = (int)cum; c.misc[2] = (int)loc[j];
            break;
          }
          cum += loc[j];
        }
      }
      if (lane == 0 && total == 0) { c.misc[0] = 255; c.misc[1] = 0; c.misc[2] = 0; }
    }
    g.sync();
    prefix |= (uint32_t)c.misc[0] << shift;
    mask |= 255u << shift;
    rem -= c.misc[1];
    tcount = c.misc[2];
  }
  T = prefix;
  need = rem < tcount ? rem : tcount;
  if (need < 0) need = 0;
}

// One scored entry's share of a forward: policy-state update (accumulate, counter) and, when evicting,
// its selection keys and candidate flags.  j = logical index relative to score_offset, n_s = scored
// slots after the append.  Shared by every kernel so that the arithmetic is the same everywhere.
__device__ __forceinline__ void entry_update(const ekv_step& st, int j, int n_s, bool is_new, float ds, float dsq,
                                             float& s, float& sq, float& cc, uint32_t& ka, uint32_t& kb, uint8_t& f,
                                             bool& dirty) {
  const int policy = st.policy;
  const bool evicting = st.evict > 0;
  ka = 0; kb = 0; f = 0;
  dirty = is_new;
  if (st.accumulate) {
    if (policy == EKV_POLICY_ROCO) { s = __fadd_rn(s, ds); sq = __fadd_rn(sq, dsq); dirty = true; }
    else if (policy == EKV_POLICY_H2O) { s = __fadd_rn(s, ds); dirty = true; }
    else if (policy == EKV_POLICY_TOVA) { s = ds; dirty = true; }
  }
  if (evicting && st.counter_add != 0.f) { cc = __fadd_rn(cc, st.counter_add); dirty = true; }
  if (evicting) {
    if (policy == EKV_POLICY_ROCO) {
      const float mean = __fdiv_rn(s, cc);
      const float var = __fsub_rn(__fdiv_rn(sq, cc), __fmul_rn(mean, mean));
      // sqrt of a negative variance (p**2 underflowed in the model dtype) is NaN: said directly, so that the IEEE
      // square root's out-of-line slow path is not entered by the many such slots of a long fp16 cache
      float sd = var < 0.f ? __int_as_float(0x7fc00000) : __fsqrt_rn(var);
      if (j >= n_s - st.protect_last || j < st.sink_protect) sd = 1e9f;
      ka = order_key(sd);
      kb = order_key(mean);
      f = F_CAND;
    } else if (policy == EKV_POLICY_H2O || policy == EKV_POLICY_TOVA) {
      kb = order_key(s);
      if (j >= st.win_lo && j < n_s - st.win_recent) f = F_CAND | F_FEAS;
    } else if (policy == EKV_POLICY_RANGE) {
      if (j >= st.range_start && j < st.range_start + st.evict) f = F_CAND | F_FEAS | F_CHOSEN;
    }
  }
}

// Acc: void operator()(int e, float& ds, float& dsq) — this forward's (folded, rounded)
// contribution of entry e to S and SQ.
template <class Acc>
__device__ void state_select_apply(const ekv_step& st, const UnitState& u, int n_before, int n_phys, int q_len,
                                   bool lj_preloaded, Acc acc, SelScratch& c, const Grp& g, bool keys_ready = false) {
  const int NE = n_phys + q_len;
  const int n_after = n_before + q_len;
  const int P = st.score_offset;
  const int n_s = n_after - P;
  const bool evicting = st.evict > 0;
  const int policy = st.policy;
  auto stamp = [&](int i) { if (c.dbg && g.tid == 0) c.dbg[i] = clock64(); };
  stamp(0);

  // ---- pass 1: state update + keys --------------------------------------------------------
  // FCH entries per thread and trip, in three stages — all loads, then all arithmetic, then all stores —
  // so that one DRAM round trip covers the whole trip and the IEEE div/sqrt chains of the entries
  // overlap.  When a single trip covers the unit (NE <= FCH * threads) the keys stay in registers for
  // the single-victim fast path below.
  constexpr int FCH = 5;
  // keys_ready: pass 1 already ran elsewhere (the chunk path does it in a chip-wide kernel): c.lj / keyA / keyB /
  // flag are filled and the state is written back; only the select and the renumbering are left
  const bool one_trip = !keys_ready && NE <= FCH * g.n;
  int rl[FCH];
  uint32_t rka[FCH], rkb[FCH];
  uint8_t rf[FCH];
  for (int base = 0; base < (keys_ready ? 0 : NE); base += FCH * g.n) {
    float s[FCH], sq[FCH], cc[FCH], ds[FCH], dsq[FCH];
    int ph[FCH];
#pragma unroll
    for (int k = 0; k < FCH; ++k) {
      const int e = base + k * g.n + g.tid;
      rl[k] = -1; s[k] = 0.f; sq[k] = 0.f; cc[k] = 1.f; ds[k] = 0.f; dsq[k] = 0.f; ph[k] = e;
      if (e < NE) {
        if (e >= n_phys) {
          rl[k] = n_before + (e - n_phys);
          cc[k] = __fsub_rn(st.c_new0, __fmul_rn((float)(e - n_phys), st.c_new_step));
          ph[k] = u.new_slots ? u.new_slots[e - n_phys] : e;
        } else {
          rl[k] = lj_preloaded ? c.lj[e] : u.lidx[e];
          if (!lj_preloaded || rl[k] >= P) {
            if (u.S_in) { s[k] = u.S_in[e]; sq[k] = u.SQ_in[e]; cc[k] = u.C_in[e]; }
            else { s[k] = u.S[e]; sq[k] = u.SQ[e]; cc[k] = u.C[e]; }
          }
        }
        if (st.accumulate) acc(e, ds[k], dsq[k]);
      }
    }
    bool dirty[FCH];
#pragma unroll
    for (int k = 0; k < FCH; ++k) {
      const int e = base + k * g.n + g.tid;
      rka[k] = 0; rkb[k] = 0; rf[k] = 0; dirty[k] = false;
      if (e < NE && rl[k] >= 0 && rl[k] >= P)
        entry_update(st, rl[k] - P, n_s, e >= n_phys, ds[k], dsq[k], s[k], sq[k], cc[k], rka[k], rkb[k], rf[k], dirty[k]);
    }
#pragma unroll
    for (int k = 0; k < FCH; ++k) {
      const int e = base + k * g.n + g.tid;
      if (e < NE) {
        if (dirty[k]) { u.S[ph[k]] = s[k]; u.SQ[ph[k]] = sq[k]; u.C[ph[k]] = cc[k]; }
        c.lj[e] = rl[k]; c.keyA[e] = rka[k]; c.keyB[e] = rkb[k]; c.flag[e] = rf[k];
      }
    }
  }
  stamp(1);
  if (!evicting || policy == EKV_POLICY_NONE) {
    for (int e = n_phys + g.tid; e < NE; e += g.n) {
      const int i_new = e - n_phys;
      const int phys = u.new_slots ? u.new_slots[i_new] : n_phys + i_new;
      u.lidx[phys] = n_before + i_new;
    }
    return;
  }

  // ---- single victim (decode), keys in registers -------------------------------------------------------
  // roco: walk the candidates in (mean, std, index) order and take the first whose std rank is below
  // k_feasible — the same slot as argmin over the k smallest std (easykv.py:322-324), found with one
  // block argmin + one counting pass per attempt (the lowest-mean slot is usually among the 70 % lowest
  // std) instead of a radix select.  h2o / tova: one block argmin over the window.  Victim bookkeeping
  // and the renumbering also run from registers.  After MAX_TRY rejected candidates the general path
  // below takes over.
  if (one_trip && st.evict == 1 && policy != EKV_POLICY_RANGE) {
    constexpr int MAX_TRY = 4;
    const int w = g.tid >> 5, nw = (g.n + 31) >> 5;
    const uint8_t need_flag = policy == EKV_POLICY_ROCO ? F_CAND : F_FEAS;
    bool found = false;
    int e_c = -1;
    uint32_t l_c = 0;
    for (int attempt = 0; attempt < MAX_TRY; ++attempt) {
      Tuple128 best; best.hi = ~0ull; best.lo = ~0ull;
#pragma unroll
      for (int k = 0; k < FCH; ++k) {
        if ((rf[k] & (need_flag | F_REJ)) == need_flag) {
          Tuple128 t;
          t.hi = ((unsigned long long)rkb[k] << 32) | rka[k];
          t.lo = ((unsigned long long)(uint32_t)rl[k] << 32) | (uint32_t)(k * g.n + g.tid);
          if (tuple_less(t, best)) best = t;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Tuple128 t;
        t.hi = __shfl_xor_sync(0xffffffffu, best.hi, o);
        t.lo = __shfl_xor_sync(0xffffffffu, best.lo, o);
        if (tuple_less(t, best)) best = t;
      }
      if ((g.tid & 31) == 0) { c.red[2 * w] = best.hi; c.red[2 * w + 1] = best.lo; }
      g.sync();
      for (int k = 0; k < nw; ++k) {
        Tuple128 t; t.hi = c.red[2 * k]; t.lo = c.red[2 * k + 1];
        if (tuple_less(t, best)) best = t;
      }
      if (best.lo == ~0ull) break;                                // no candidate left (uniform)
      const uint32_t ka_c = (uint32_t)(best.hi & 0xffffffffu);
      l_c = (uint32_t)(best.lo >> 32);
      e_c = (int)(best.lo & 0xffffffffu);
      if (policy != EKV_POLICY_ROCO) { found = true; break; }
      int cnt = 0;                                                // std rank of the candidate
#pragma unroll
      for (int k = 0; k < FCH; ++k)
        cnt += ((rf[k] & F_CAND) && (rka[k] < ka_c || (rka[k] == ka_c && (uint32_t)rl[k] < l_c))) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if ((g.tid & 31) == 0) c.hist[(attempt & 1) * 32 + w] = (uint32_t)cnt;
      g.sync();
      int rank = 0;
      for (int k = 0; k < nw; ++k) rank += (int)c.hist[(attempt & 1) * 32 + k];
      if (rank < st.k_feasible) { found = true; break; }
#pragma unroll
      for (int k = 0; k < FCH; ++k)
        if (k * g.n + g.tid == e_c) rf[k] |= F_REJ;
    }
    stamp(2);
    if (found) {
#pragma unroll
      for (int k = 0; k < FCH; ++k) {
        const int e = k * g.n + g.tid;
        if (e >= NE || rl[k] < 0) continue;
        const int phys = e >= n_phys ? (u.new_slots ? u.new_slots[e - n_phys] : e) : e;
        if (e == e_c) {
          if (u.victim_lidx) u.victim_lidx[0] = rl[k];
          if (u.victim_slots) u.victim_slots[0] = phys;
          if (st.apply) u.lidx[phys] = -1;
          else if (e >= n_phys) u.lidx[phys] = rl[k];
        } else if (st.apply && (uint32_t)rl[k] > l_c) {
          u.lidx[phys] = rl[k] - 1;
        } else if (e >= n_phys) {
          u.lidx[phys] = rl[k];
        }
      }
      stamp(5);
      return;
    }
  }
  g.sync();
  bool done = false;

  stamp(2);
  // ---- stage 1 (roco): the k_feasible smallest std ------------------------------------------
  if (policy == EKV_POLICY_ROCO && !done) {
    uint32_t T1; int need1, tc1;
    radix_select(NE, st.k_feasible,
                 [&](int e) { return c.keyA[e]; }, [&](int e) { return (c.flag[e] & F_CAND) != 0; }, c, g, T1, need1, tc1);
    uint32_t jT = 0xffffffffu;
    if (need1 < tc1) {
      int nd, tc;
      radix_select(NE, need1, [&](int e) { return (uint32_t)c.lj[e]; },
                   [&](int e) { return (c.flag[e] & F_CAND) && c.keyA[e] == T1; }, c, g, jT, nd, tc);
    }
    for (int e = g.tid; e < NE; e += g.n) {
      uint8_t f = c.flag[e];
      if (f & F_CAND) {
        const uint32_t ka = c.keyA[e];
        if (ka < T1 || (ka == T1 && (uint32_t)c.lj[e] <= jT)) c.flag[e] = f | F_FEAS;
      }
    }
    g.sync();
  }

  // ---- stage 2: the `evict` smallest (keyB, keyA, logical index) among the feasible ------------
  if (policy != EKV_POLICY_RANGE && !done) {
    if (st.evict == 1) {
      Tuple128 best; best.hi = ~0ull; best.lo = ~0ull;
      for (int e = g.tid; e < NE; e += g.n) {
        if (c.flag[e] & F_FEAS) {
          Tuple128 t;
          t.hi = ((unsigned long long)c.keyB[e] << 32) | c.keyA[e];
          t.lo = ((unsigned long long)(uint32_t)c.lj[e] << 32) | (uint32_t)e;
          if (tuple_less(t, best)) best = t;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        Tuple128 t;
        t.hi = __shfl_xor_sync(0xffffffffu, best.hi, o);
        t.lo = __shfl_xor_sync(0xffffffffu, best.lo, o);
        if (tuple_less(t, best)) best = t;
      }
      const int w = g.tid >> 5, nw = (g.n + 31) >> 5;
      if ((g.tid & 31) == 0) { c.red[2 * w] = best.hi; c.red[2 * w + 1] = best.lo; }
      g.sync();
      if (g.tid == 0) {
        for (int k = 1; k < nw; ++k) {
          Tuple128 t; t.hi = c.red[2 * k]; t.lo = c.red[2 * k + 1];
          if (tuple_less(t, best)) best = t;
        }
        if (best.lo != ~0ull) c.flag[(uint32_t)(best.lo & 0xffffffffu)] |= F_CHOSEN;
      }
      g.sync();
    } else {
      uint32_t T2, T3 = 0xffffffffu, jT = 0xffffffffu; int need2, tc2;
      radix_select(NE, st.evict, [&](int e) { return c.keyB[e]; }, [&](int e) { return (c.flag[e] & F_FEAS) != 0; },
                   c, g, T2, need2, tc2);
      if (need2 < tc2) {
        int need3, tc3;
        radix_select(NE, need2, [&](int e) { return c.keyA[e]; },
                     [&](int e) { return (c.flag[e] & F_FEAS) && c.keyB[e] == T2; }, c, g, T3, need3, tc3);
        if (need3 < tc3) {
          int nd, tc;
          radix_select(NE, need3, [&](int e) { return (uint32_t)c.lj[e]; },
                       [&](int e) { return (c.flag[e] & F_FEAS) && c.keyB[e] == T2 && c.keyA[e] == T3; }, c, g, jT, nd, tc);
        }
      }
      for (int e = g.tid; e < NE; e += g.n) {
        uint8_t f = c.flag[e];
        if (f & F_FEAS) {
          const uint32_t kb = c.keyB[e], ka = c.keyA[e];
          if (kb < T2 || (kb == T2 && (ka < T3 || (ka == T3 && (uint32_t)c.lj[e] <= jT)))) c.flag[e] = f | F_CHOSEN;
        }
      }
      g.sync();
    }
  }

  stamp(3);
  // ---- gather victims, order them by logical index, renumber --------------------------------------
  if (g.tid == 0) c.misc[4] = 0;
  g.sync();
  for (int e = g.tid; e < NE; e += g.n) {
    if (c.flag[e] & F_CHOSEN) {
      const int pos = atomicAdd(&c.misc[4], 1);
      if (pos < st.evict) {
        c.vl[pos] = c.lj[e];
        c.vs[pos] = e >= n_phys ? (u.new_slots ? u.new_slots[e - n_phys] : e) : e;
      }
    }
  }
  g.sync();
  const int nv = c.misc[4] < st.evict ? c.misc[4] : st.evict;
  int32_t* vsorted = c.vl + (st.evict > 0 ? st.evict : 1);
  for (int t = g.tid; t < st.evict; t += g.n) {
    if (t < nv) {
      const int l = c.vl[t];
      int rank = 0;
      for (int k = 0; k < nv; ++k) rank += c.vl[k] < l;
      vsorted[rank] = l;
      if (u.victim_lidx) u.victim_lidx[rank] = l;
      if (u.victim_slots) u.victim_slots[rank] = c.vs[t];
    } else {   // fewer candidates than requested (degenerate shapes): pad
      if (u.victim_lidx) u.victim_lidx[t] = -1;
      if (u.victim_slots) u.victim_slots[t] = -1;
    }
  }
  g.sync();
  stamp(4);
  if (st.apply) {
    for (int e = g.tid; e < NE; e += g.n) {
      const int l = c.lj[e];
      if (l < 0) continue;
      const int phys = e >= n_phys ? (u.new_slots ? u.new_slots[e - n_phys] : e) : e;
      if (c.flag[e] & F_CHOSEN) { u.lidx[phys] = -1; continue; }
      int lo = 0, hi = nv;                       // number of victims with logical id < l
      while (lo < hi) { int mid = (lo + hi) >> 1; if (vsorted[mid] < l) lo = mid + 1; else hi = mid; }
      if (lo > 0 || e >= n_phys) u.lidx[phys] = l - lo;
    }
  } else {
    for (int e = n_phys + g.tid; e < NE; e += g.n) {
      const int phys = u.new_slots ? u.new_slots[e - n_phys] : e;
      u.lidx[phys] = c.lj[e];
    }
  }
  stamp(5);
}

}  // namespace ekv
