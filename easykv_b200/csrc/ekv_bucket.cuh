// RoCo victim select in bounded time, shared by the decode kernels (single CTA or a cluster per unit).
//
// Reference semantics (easykv/easykv.py:320-324, :471-476, :722-724): std -> topk(k_feasible smallest) -> argmin(mean)
// over those.  Order as everywhere here: (std key asc, NaN last, logical index asc) for the k smallest,
// (mean key, std key, logical index) for the argmin.
//
// The k-th smallest std is found from histograms that the keys pass fills on the fly (no extra pass over the
// entries): 2048 buckets over the order-preserving std key — 64 per octave from 2^-31 up to ~1.9 (a probability's
// std is <= 0.5), one for everything above (the 1e9 sentinels of the protected slots), one for (0, 2^-31), and two
// PURE buckets whose keys are all equal: std == 0 (slots whose probability underflowed to 0 at every step) and NaN
// (all NaN keys are equal, order_key; on long fp16 caches most slots have a NaN std because p**2 underflows,
// easykv.py:296).  16-bit counters, two per word.  A cut inside a pure bucket is resolved by logical index from a second
// on-the-fly histogram (one per pure bucket).  After the scan every entry is classified in one pass:
//     bucket below the cut            feasible         -> running argmin of (mean, std, index)
//     the cut's bucket                boundary         -> listed; at most CAP of them, ranked exactly afterwards
//     above                           infeasible
// A crowded regular cut bucket (> CAP entries) is first refined by one more histogram pass over its entries (the next
// 11 key bits).  A cluster shares: the histograms (peers read them over distributed shared memory), one exchange of the
// per-warp argmins and the boundary list.  What is still crowded after that (hundreds of slots with bit-identical
// non-zero std) reports FALLBACK and the caller runs its MSB-first radix select.
#pragma once
#include "ekv_common.cuh"

namespace ekv {

struct Tuple128 { unsigned long long hi, lo; };
__device__ __forceinline__ bool tuple_less(const Tuple128& a, const Tuple128& b) {
  return a.hi < b.hi || (a.hi == b.hi && a.lo < b.lo);
}

// smallest 128-bit tuple of a (converged) warp: four 32-bit min-reductions (redux.sync), most significant word first
__device__ __forceinline__ Tuple128 warp_argmin(const Tuple128& t) {
  const uint32_t w3 = (uint32_t)(t.hi >> 32), w2 = (uint32_t)t.hi, w1 = (uint32_t)(t.lo >> 32), w0 = (uint32_t)t.lo;
  const uint32_t m3 = __reduce_min_sync(0xffffffffu, w3);
  bool on = w3 == m3;
  const uint32_t m2 = __reduce_min_sync(0xffffffffu, on ? w2 : 0xffffffffu);
  on = on && w2 == m2;
  const uint32_t m1 = __reduce_min_sync(0xffffffffu, on ? w1 : 0xffffffffu);
  on = on && w1 == m1;
  const uint32_t m0 = __reduce_min_sync(0xffffffffu, on ? w0 : 0xffffffffu);
  Tuple128 r;
  r.hi = ((unsigned long long)m3 << 32) | m2; r.lo = ((unsigned long long)m1 << 32) | m0;
  return r;
}

// explicit shared-memory accesses: the scratch pointers are carved at run-time offsets, for which the compiler falls
// back to GENERIC loads (LD.E, twice the latency of LDS) — measured 5 us for a 17-iteration classify pass
__device__ __forceinline__ void lds_2u64(uint32_t addr, unsigned long long& a, unsigned long long& b) {
  asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "r"(addr));
}
__device__ __forceinline__ unsigned long long lds_u64(uint32_t addr) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

namespace bk {
constexpr int NBIN = 2048, NWORD = NBIN / 2, NAN_BIN = NBIN - 1, TOP_BIN = NBIN - 2, ZERO_BIN = 0;
constexpr uint32_t KBASE = 0xB0000000u;  // order_key(2^-31)
constexpr int BIN_SHIFT = 17;            // 2^23 keys per octave / 64 buckets
constexpr int SUB_SHIFT = BIN_SHIFT - 11;   // refinement: the next 11 bits
constexpr int CAP = 256;                 // boundary entries ranked by brute force
constexpr int NTHR = 256, NW = NTHR / 32;   // threads taking part in the scan: 8 buckets each
enum { OK = 0, NONE = 1, FALLBACK = 2 };

__device__ __forceinline__ int std_bin(uint32_t ka) {
  if (ka == 0xffffffffu) return NAN_BIN;
  if (ka <= 0x80000000u) return ZERO_BIN;                      // zero (a std is never negative: sqrt, or NaN)
  if (ka < KBASE) return 1;                                    // (0, 2^-31)
  const uint32_t b = 2u + ((ka - KBASE) >> BIN_SHIFT);
  return b > (uint32_t)TOP_BIN ? TOP_BIN : (int)b;
}
// first / last key of a bucket (std_bin is monotone: a bucket is a key interval)
__device__ __forceinline__ uint32_t bin_lo(int b) {
  return b <= 0 ? 0u : b == 1 ? 0x80000001u : b >= NAN_BIN ? 0xffffffffu : KBASE + ((uint32_t)(b - 2) << BIN_SHIFT);
}
__device__ __forceinline__ uint32_t bin_last(int b) { return b >= NAN_BIN ? 0xffffffffu : bin_lo(b + 1) - 1u; }
__device__ __forceinline__ int sub_bin(uint32_t ka) { return (int)(((ka - KBASE) >> SUB_SHIFT) & (NBIN - 1)); }
// logical indices (< n_after) -> NBIN buckets
__host__ __device__ inline int lidx_shift(int n_after) {
  int s = 0;
  while (((n_after - 1) >> s) > NBIN - 1) ++s;
  return s;
}
__device__ __forceinline__ void add16(uint32_t* hist, int b) { atomicAdd(&hist[b >> 1], 1u << ((b & 1) * 16)); }
__device__ __forceinline__ uint32_t get16(const uint4& w, int j) {       // bucket j (0..7) of four packed words
  const uint32_t x = j < 2 ? w.x : j < 4 ? w.y : j < 6 ? w.z : w.w;
  return (x >> ((j & 1) * 16)) & 0xffffu;
}
}  // namespace bk

struct BucketScratch {
  uint32_t* histA;             // [NWORD] std-key buckets
  uint32_t* histN;             // [NWORD] logical-index buckets of the NaN-keyed candidates (reused by the refinement pass)
  uint32_t* histZ;             // [NWORD] logical-index buckets of the candidates whose std is exactly 0
  unsigned long long* blist;   // [2 * CAP] boundary entries (cluster-wide, same content in every CTA)
  unsigned long long* xg;      // [2 * MAXC * NW] gathered per-warp argmins
  unsigned long long* wred;    // [2 * NW]
  int* misc;                   // [16]: warp sums [8] | bucket, count below, count inside, list offset | list cursor
  static __host__ __device__ constexpr int bytes(int maxc) { return 3 * bk::NWORD * 4 + bk::CAP * 16 + xgn(maxc) * 16 + bk::NW * 16 + 64; }
  static __host__ __device__ constexpr int xgn(int maxc) { return maxc * bk::NW < 32 ? 32 : maxc * bk::NW; }   // gathered tuples (a single CTA: one per warp)
  __device__ void carve(void* base, int maxc) {      // 16-byte aligned
    char* p = reinterpret_cast<char*>(base);
    histA = reinterpret_cast<uint32_t*>(p); p += bk::NWORD * 4;
    histN = reinterpret_cast<uint32_t*>(p); p += bk::NWORD * 4;
    histZ = reinterpret_cast<uint32_t*>(p); p += bk::NWORD * 4;
    blist = reinterpret_cast<unsigned long long*>(p); p += bk::CAP * 16;
    xg = reinterpret_cast<unsigned long long*>(p); p += xgn(maxc) * 16;
    wred = reinterpret_cast<unsigned long long*>(p); p += bk::NW * 16;
    misc = reinterpret_cast<int*>(p);
  }
  // every thread of the group, before the keys pass (a group barrier must follow)
  __device__ void clear(int tid, int nthr) const {
    for (int i = tid; i < 3 * bk::NWORD / 4; i += nthr) reinterpret_cast<uint4*>(histA)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) misc[12] = 0;
  }
  // one candidate's keys -> histograms (any thread, any order)
  __device__ __forceinline__ void add(uint32_t ka, uint32_t l, int lsh) const {
    const int b = bk::std_bin(ka);
    bk::add16(histA, b);
    if (b == bk::NAN_BIN) bk::add16(histN, (int)(l >> lsh));
    else if (b == bk::ZERO_BIN) bk::add16(histZ, (int)(l >> lsh));
  }
  // the same for a whole warp (converged): a pure bucket is one address for every lane — one add per warp
  __device__ __forceinline__ void add_warp(bool cand, uint32_t ka, uint32_t l, int lsh, int lane) const {
    const int b = cand ? bk::std_bin(ka) : -1;
    const uint32_t mn = __ballot_sync(0xffffffffu, b == bk::NAN_BIN), mz = __ballot_sync(0xffffffffu, b == bk::ZERO_BIN);
    if (mn && lane == (__ffs(mn) - 1)) atomicAdd(&histA[bk::NAN_BIN >> 1], (uint32_t)__popc(mn) << ((bk::NAN_BIN & 1) * 16));
    if (mz && lane == (__ffs(mz) - 1)) atomicAdd(&histA[bk::ZERO_BIN >> 1], (uint32_t)__popc(mz) << ((bk::ZERO_BIN & 1) * 16));
    if (b == bk::NAN_BIN) bk::add16(histN, (int)(l >> lsh));
    else if (b == bk::ZERO_BIN) bk::add16(histZ, (int)(l >> lsh));
    if (b > bk::ZERO_BIN && b < bk::NAN_BIN) bk::add16(histA, b);
  }
};

struct BucketSel {
  int status;        // bk::OK | NONE | FALLBACK
  int bsel, rb;      // the cut's std bucket and how many of its entries are feasible (in (std key, index) order)
  int kind;          // how the cut's bucket is split further: 0 not at all, 1 pure bucket, by logical index, 2 refined by key bits
  int csel;          // kinds 1, 2: the cut's sub-bucket
  int mtot, my_off;  // boundary entries in the cluster; where this CTA's share of the list starts
};

// The cut.  Called by every thread of the group (nthr >= 256, `sync` = the group's barrier) once all CTAs' histograms
// are complete and visible; `ld(ptr, peer)` loads 16 bytes at the same offset of peer's shared memory.
template <int MAXC, class Sync, class Ld>
__device__ __forceinline__ void bucket_scan_one(const BucketScratch& s, const uint32_t* hist, int want_in, int C, int rank, int tid,
                                                Sync sync, Ld ld, int& sel, int& below, int& inside, int& off, int& total) {
  const int lane = tid & 31, w = tid >> 5;
  uint4 wv[MAXC];
  uint32_t loc[8], sum = 0, inc = 0;
  if (tid < bk::NTHR) {
#pragma unroll
    for (int p = 0; p < MAXC; ++p) wv[p] = p < C ? ld(hist + 4 * tid, p) : make_uint4(0, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t t = 0;
#pragma unroll
      for (int p = 0; p < MAXC; ++p) t += bk::get16(wv[p], j);
      loc[j] = t; sum += t;
    }
    inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s.misc[w] = (int)inc;
  }
  sync();
  uint32_t pre = 0, tot = 0;
#pragma unroll
  for (int q = 0; q < bk::NW; ++q) { const uint32_t v = (uint32_t)s.misc[q]; pre += q < w ? v : 0u; tot += v; }
  if (tid < bk::NTHR) {
    inc += pre;
    const uint32_t exc = inc - sum;
    uint32_t want = want_in < 1 ? 1u : (uint32_t)want_in;      // k <= 0 still yields one (radix_select's rule)
    if (want > tot) want = tot;                                 // fewer candidates than requested: all of them
    if (tot > 0 && exc < want && want <= inc) {
      uint32_t cum = exc;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (want <= cum + loc[j]) {
          int o = 0;
#pragma unroll
          for (int p = 0; p < MAXC; ++p) o += p < rank ? (int)bk::get16(wv[p], j) : 0;
          s.misc[8] = 8 * tid + j; s.misc[9] = (int)cum; s.misc[10] = (int)loc[j]; s.misc[11] = o;
          break;
        }
        cum += loc[j];
      }
    }
  }
  sync();
  sel = s.misc[8]; below = s.misc[9]; inside = s.misc[10]; off = s.misc[11]; total = (int)tot;
}

// get(e, ka, kb, l) -> entry e is a candidate (with its keys); hist_sync() returns once every CTA's refinement histogram
// is complete and visible (no-op for a single CTA).
template <int MAXC, class Sync, class Ld, class Get, class HistSync>
__device__ __forceinline__ BucketSel bucket_scan(const BucketScratch& s, int k_feasible, int C, int rank, int tid, int nthr, int NEl,
                                                 Sync sync, Ld ld, Get get, HistSync hist_sync) {
  BucketSel b;
  b.status = bk::OK; b.csel = 0; b.kind = 0; b.mtot = 0; b.my_off = 0; b.bsel = 0; b.rb = 0;
  int below, inside, off, total;
  bucket_scan_one<MAXC>(s, s.histA, k_feasible, C, rank, tid, sync, ld, b.bsel, below, inside, off, total);
  if (total == 0) { b.status = bk::NONE; b.bsel = -1; return b; }          // no candidate: nothing is feasible
  int want = k_feasible < 1 ? 1 : k_feasible;
  if (want > total) want = total;
  b.rb = want - below;                                                     // 1 .. inside
  b.mtot = inside; b.my_off = off;
  int below2, inside2, off2, total2;
  if (b.bsel == bk::NAN_BIN || b.bsel == bk::ZERO_BIN) {                   // a run of equal keys: lowest logical index first
    b.kind = 1;
    bucket_scan_one<MAXC>(s, b.bsel == bk::NAN_BIN ? s.histN : s.histZ, b.rb, C, rank, tid, sync, ld, b.csel, below2, inside2, off2, total2);
    b.rb -= below2;
    b.mtot = inside2; b.my_off = off2;
  } else if (b.mtot > bk::CAP && b.bsel >= 2 && b.bsel < bk::TOP_BIN) {     // crowded: split the bucket by its next 11 key bits
    b.kind = 2;
    for (int i = tid; i < bk::NWORD / 4; i += nthr) reinterpret_cast<uint4*>(s.histN)[i] = make_uint4(0, 0, 0, 0);
    sync();
    for (int e = tid; e < NEl; e += nthr) {
      uint32_t ka, kb, l;
      if (get(e, ka, kb, l) && bk::std_bin(ka) == b.bsel) bk::add16(s.histN, bk::sub_bin(ka));
    }
    sync();
    hist_sync();
    bucket_scan_one<MAXC>(s, s.histN, b.rb, C, rank, tid, sync, ld, b.csel, below2, inside2, off2, total2);
    b.rb -= below2;
    b.mtot = inside2; b.my_off = off2;
  }
  if (b.mtot > bk::CAP) b.status = bk::FALLBACK;
  return b;
}

// How an entry is judged: mode 0 — every keyed entry is feasible (h2o_head / tova window); 1 — buckets; 2 — explicit
// thresholds from a radix select (feasible: std key < T1, or == T1 with logical index <= jT).
struct Feasibility {
  int mode;
  BucketSel b;
  int lsh;
  uint32_t T1, jT;
};
// ... as key intervals, so that the pass judges an entry with a handful of compares and no divergence: feasible iff
// ka < a_lo, or ka <= a_last and l < l_lo; listed (boundary) iff a_lo <= ka <= a_last and l_lo <= l < l_hi.
struct FeasCut {
  // 32-bit forms (logical indices stay far below 2^32 - 1, so the 64-bit bounds saturate): in = ka - a_lo <= a_span;
  // feasible = ka < a_lo || (in && l < l_lo); listed = in && l - l_lo < l_span (the unsigned wrap rejects l < l_lo)
  uint32_t a_lo, a_span, l_lo, l_span;
  __device__ __forceinline__ explicit FeasCut(const Feasibility& f) {
    uint32_t alo = 0u, alast = 0u;
    unsigned long long llo = 0ull, lhi = 0ull;                                 // nothing feasible, nothing listed
    if (f.mode == 0) { alast = 0xffffffffu; llo = 1ull << 32; lhi = llo; }      // every keyed entry is feasible
    else if (f.mode == 2) { alo = f.T1; alast = f.T1; llo = (unsigned long long)f.jT + 1ull; lhi = llo; }
    else if (f.mode == 1 && f.b.bsel >= 0) {
      alo = bk::bin_lo(f.b.bsel); alast = bk::bin_last(f.b.bsel); lhi = 1ull << 32;
      if (f.b.kind == 1) { llo = (unsigned long long)f.b.csel << f.lsh; lhi = (unsigned long long)(f.b.csel + 1) << f.lsh; }
      else if (f.b.kind == 2) { alo += (uint32_t)f.b.csel << bk::SUB_SHIFT; alast = alo + (1u << bk::SUB_SHIFT) - 1u; }
    }
    const uint32_t lo32 = llo > 0xffffffffull ? 0xffffffffu : (uint32_t)llo, hi32 = lhi > 0xffffffffull ? 0xffffffffu : (uint32_t)lhi;
    a_lo = alo; a_span = alast - alo; l_lo = lo32; l_span = hi32 - lo32;
  }
  __device__ __forceinline__ void judge(uint32_t ka, uint32_t l, bool& feas, bool& bound) const {
    const bool in = ka - a_lo <= a_span;
    feas = (ka < a_lo) | (in & (l < l_lo));
    bound = in & (l - l_lo < l_span);
  }
};

// One pass over this CTA's entries: argmin over the feasible ones, boundary entries listed.  get(e, ka, kb, l) -> is a
// candidate; push_peers(slot, hi, lo) stores 16 bytes at `slot` (an address in this CTA's shared memory) in every OTHER
// CTA of the cluster (a no-op for a single CTA).  Boundary entries are first listed in this CTA's own copy and sent to the
// peers afterwards, one entry per thread: remote stores issued from inside the (divergent) entry loop serialise — measured
// 5 us for 48 entries.  Every warp's lane 0 publishes the warp's best at xg[(rank * NW + warp)] (nthr == 256).
template <class Get, class Push, class Sync>
__device__ __forceinline__ void bucket_pass(const BucketScratch& s, const Feasibility& f, int NEl, int rank, int tid, int nthr,
                                            Get get, Push push_peers, Sync sync, unsigned long long* dbg = nullptr) {
  if (nthr > bk::NTHR) nthr = bk::NTHR;
  const bool active = tid < nthr;                                // a wider group idles (but joins the barrier)
  const FeasCut cut(f);
  // two entries in flight per thread.  The pass runs once per launch from a cold instruction cache (a third of the tail's
  // stall samples were instruction fetches), so a compact body beats a deeper unroll: the loads are shared-memory hits
  constexpr int U = 2;
  Tuple128 best; best.hi = ~0ull; best.lo = ~0ull;
  const uint32_t cursor = smem_u32(&s.misc[12]), bl = smem_u32(s.blist);
#pragma unroll 1
  for (int e0 = tid; active && e0 < NEl; e0 += U * nthr) {
    uint32_t ka[U], kb[U], l[U];
    bool c[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int e = e0 + u * nthr;
      ka[u] = 0; kb[u] = 0; l[u] = 0;
      c[u] = e < NEl && get(e, ka[u], kb[u], l[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      // branch-free except for the (rare) listed entries: the pass is a chain of dependent compares on two warps per
      // scheduler, so instructions per entry are what it costs
      bool feas, bound;
      cut.judge(ka[u], l[u], feas, bound);
      const unsigned long long hi = ((unsigned long long)kb[u] << 32) | ka[u];
      const unsigned long long lo = ((unsigned long long)l[u] << 32) | ((uint32_t)rank << 24) | (uint32_t)(e0 + u * nthr);
      const bool less = c[u] & feas & ((hi < best.hi) | ((hi == best.hi) & (lo < best.lo)));
      best.hi = less ? hi : best.hi;
      best.lo = less ? lo : best.lo;
      if (c[u] & bound) {
        uint32_t slot;
        asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(slot) : "r"(cursor) : "memory");
        const int pos = f.b.my_off + (int)slot;
        if (pos < bk::CAP) asm volatile("st.shared.v2.u64 [%0], {%1, %2};" ::"r"(bl + 16u * (uint32_t)pos), "l"(hi), "l"(lo) : "memory");
      }
    }
  }
  if (dbg && tid == 0) { unsigned long long tns; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns)); dbg[0] = tns; }   // profiling hook: loop done
  if (active) {
    best = warp_argmin(best);
    if ((tid & 31) == 0) {
      unsigned long long* slot = &s.xg[2 * (rank * bk::NW + (tid >> 5))];
      slot[0] = best.hi; slot[1] = best.lo;
      push_peers(slot, best.hi, best.lo);
    }
  }
  if (f.mode == 1) {                                             // (uniform) this CTA's listed entries -> the peers, one per thread
    sync();
    const int mine = s.misc[12];
    for (int i = tid; active && i < mine && f.b.my_off + i < bk::CAP; i += nthr) {
      unsigned long long hi, lo;
      lds_2u64(bl + 16u * (uint32_t)(f.b.my_off + i), hi, lo);
      push_peers(&s.blist[2 * (f.b.my_off + i)], hi, lo);
    }
  }
}

// After the exchange: rank the boundary entries, merge with the gathered argmins.  Every CTA derives the same winner.
template <class Sync>
__device__ __forceinline__ bool bucket_final(const BucketScratch& s, const Feasibility& f, int C, int tid, int nthr, Sync sync, Tuple128& win) {
  Tuple128 best; best.hi = ~0ull; best.lo = ~0ull;
  const int m = f.mode == 1 ? (f.b.mtot < bk::CAP ? f.b.mtot : bk::CAP) : 0;
  const int nt = nthr > bk::NTHR ? bk::NTHR : nthr;
  const uint32_t bl = smem_u32(s.blist), xg = smem_u32(s.xg);
  // ranks by brute force, eight lanes per listed entry (each counts every eighth competitor)
  if (tid < nt) {                                                  // (whole warps: nt is a multiple of 32)
    const int sub = tid & 7;
    for (int i0 = 0; i0 < m; i0 += nt / 8) {
      const int i = i0 + (tid >> 3);
      const bool valid = i < m;
      Tuple128 t; t.hi = ~0ull; t.lo = ~0ull;
      if (valid) lds_2u64(bl + 16 * i, t.hi, t.lo);
      const unsigned long long key = (t.hi << 32) | (t.lo >> 32);          // (std key, logical index)
      int rk = 0;
      if (valid) {
#pragma unroll 4
        for (int j = sub; j < m; j += 8) {
          unsigned long long hj, lj2;
          lds_2u64(bl + 16 * j, hj, lj2);
          rk += ((hj << 32) | (lj2 >> 32)) < key ? 1 : 0;
        }
      }
      rk += __shfl_xor_sync(0xffffffffu, rk, 1);
      rk += __shfl_xor_sync(0xffffffffu, rk, 2);
      rk += __shfl_xor_sync(0xffffffffu, rk, 4);
      if (valid && sub == 0 && rk < f.b.rb && tuple_less(t, best)) best = t;
    }
  }
  for (int i = tid; i < C * bk::NW && tid < nt; i += nt) {
    Tuple128 t;
    lds_2u64(xg + 16 * i, t.hi, t.lo);
    if (t.lo != ~0ull && tuple_less(t, best)) best = t;
  }
  best = warp_argmin(best);
  if ((tid & 31) == 0 && (tid >> 5) < bk::NW) { s.wred[2 * (tid >> 5)] = best.hi; s.wred[2 * (tid >> 5) + 1] = best.lo; }
  sync();
  const int nw = nt / 32;
  win.hi = s.wred[0]; win.lo = s.wred[1];
  for (int q = 1; q < nw; ++q) {
    Tuple128 t; t.hi = s.wred[2 * q]; t.lo = s.wred[2 * q + 1];
    if (tuple_less(t, win)) win = t;
  }
  return win.lo != ~0ull;
}

}  // namespace ekv
