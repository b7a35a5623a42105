// Fused decode step (q_len == 1) for one layer.  One persistent, warp-specialised CTA per SM that
// loops over (sequence, kv head) units:
//
//   TMA producer warp : streams each unit's K rows, then its V rows, HBM -> shared memory, as 16 KB
//                       cp.async.bulk tiles through one deep mbarrier ring (all shared memory that the
//                       consumers do not need: ~11 tiles = 176 KB in flight per SM).  The stream is
//                       continuous across unit boundaries.  Each K/V byte is read from HBM exactly
//                       once; the g query heads of a GQA group share the stream.
//   2 consumer groups : of 8 warps each, alternating units (ping-pong): while one group runs the
//                       select tail of unit k, the other already consumes the tiles of unit k+1, so the
//                       tail never idles HBM.  Per unit a group does
//                       header   — q, the new K/V row, the slot map -> shared memory (while idle);
//                       K phase  — 16 lanes per row, 128-bit shared-memory reads, fp32 FMA dot
//                                  products, transposing warp-shuffle reduction, logits rounded and
//                                  masked at the reference's rounding points;
//                       softmax  — fp32, warp-shuffle + named-barrier reductions, probabilities
//                                  rounded to the model dtype;
//                       V phase  — fp32 FMA accumulate of p·V, cross-warp reduction, out;
//                       tail     — GQA fold, policy accumulate, victim select, in-place eviction
//                                  (ekv_select.cuh) and the append of the new K/V row.
//
// Replaces (reference paths): llama_patch.py:193-230 / mistral_patch.py:137-170 (cache append,
// repeat_kv, QK^T, mask, softmax, PV) and easykv.py:271-362 / :683-748 (fold, accumulate, select,
// truncate_kv_cache_silo, state compaction) for one layer of one decode forward.
#include "ekv_decode_common.cuh"
#include "ekv_mma.cuh"

namespace ekv {

template <typename T> struct DecodeSmem {
  // byte offsets inside dynamic shared memory; g_* are relative to a consumer group's block
  int off_bar, off_grp, grp_bytes, g_red, g_q, g_k, g_v, g_ns, g_lj, g_plog, g_scr, off_ring, fixed, nep;
  __host__ __device__ DecodeSmem(int G, int n_phys, int evict, int ngroups) {
    using Cfg = DecodeCfg<T>;
    const int NE = n_phys + 1;
    nep = align_up(NE, 8);
    int o = 0;
    off_bar = o; o += (Cfg::MAX_GROUPS + 1) * Cfg::MAX_STAGES * 8;   // full[group][stage], empty[stage]
    o = align_up(o, 128);
    off_grp = o;
    int h = 0;
    g_red = h; h += 2 * 8 * Cfg::NWARP * 4;
    g_q = h; h += G * Cfg::D * (int)sizeof(T);
    g_k = h; h += Cfg::ROW_BYTES;
    g_v = h; h += Cfg::ROW_BYTES;
    g_ns = h; h += 16;
    g_lj = h; h += align_up(NE * 4, 16);
    g_plog = h; h += align_up(G * nep * (int)sizeof(T), 16);
    h = align_up(h, 128);
    g_scr = h;
    size_t scr = SelScratch::bytes(NE, evict);
    size_t outp = (size_t)Cfg::NWARP * G * Cfg::D * 4;
    h += (int)(scr > outp ? scr : outp);
    grp_bytes = align_up(h, 128);
    o += ngroups * grp_bytes;
    off_ring = o;
    fixed = o;
  }
};

// NG consumer groups per CTA: 2 = one CTA per SM whose groups ping-pong over one shared ring;
// 1 = a lighter CTA (one group, its own ring) of which two are resident per SM and stream independently.
// EXT: the instantiation that also serves the fused streaming variant (io.rope_cos: cached rows rotated while read) and
// ragged batches (io.seq_n_before / step.budget_gate); the plain one carries none of that code (96 registers, no spills).
template <typename T, int G, int NG, bool EXT>
__global__ void __launch_bounds__(NG * DecodeCfg<T>::NCONS + 32, NG == 1 ? 2 : 1)
decode_kernel(const KernelArgs a, const int stages) {
  constexpr int ngroups = NG;
  using Cfg = DecodeCfg<T>;
  constexpr int D = Cfg::D, NWARP = Cfg::NWARP, NCONS = Cfg::NCONS, RPT = Cfg::RPT;
  constexpr int TILE_ROWS = Cfg::TILE_ROWS;
  pdl_trigger();            // programmatic dependent launch: the next kernel may be placed now; this one touches global memory
  pdl_wait();               // only once the previous kernel of the stream has completed (ekv_common.cuh)
  extern __shared__ __align__(128) unsigned char smem[];
  const DecodeSmem<T> L(G, a.n_phys, a.st.evict, NG);
  // full[g][s]: tile in ring slot s has landed, signalled to the consumer group g that owns the tile —
  // one barrier per (group, slot) so that each waiter tracks the phase of a barrier only it consumes
  // (a parity wait must never run ahead of the barrier's previous phase); empty[s]: slot s released.
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* empty = full + Cfg::MAX_GROUPS * Cfg::MAX_STAGES;
  unsigned char* ring = smem + L.off_ring;

  const int U = a.B * a.Hkv;
  const int n_phys = a.n_phys, NE = n_phys + 1, nep = L.nep;
  const bool stream_rope = EXT && sizeof(T) == 2 && a.rope_cos != nullptr;
  const int nt = (n_phys + TILE_ROWS - 1) / TILE_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      for (int g = 0; g < NG; ++g) mbar_init(&full[g * Cfg::MAX_STAGES + s], 1);
      mbar_init(&empty[s], NWARP);
    }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NG * NWARP) {
    // ===== TMA producer: one continuous tile stream over all of this CTA's units ======================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      int s = 0, use = 0;                          // ring slot and how often it has been used
      int k_unit = 0;
      long long pwait = 0;
      const long long pstart = clock64();
      for (int unit = blockIdx.x; unit < U; unit += gridDim.x, ++k_unit) {
        uint64_t* gfull = full + (k_unit % ngroups) * Cfg::MAX_STAGES;
        const T* Kg = reinterpret_cast<const T*>(a.K) + (size_t)unit * a.cap * D;
        const T* Vg = reinterpret_cast<const T*>(a.V) + (size_t)unit * a.cap * D;
        {
          // this unit's header (slot map, q, new K / V rows) -> L2 now: its consumer group reaches the header a tail and
          // most of a ring later, and would otherwise wait for DRAM behind the saturated tile stream (~6 K cycles)
          const uint32_t lb = (uint32_t)(((n_phys * 4 + 15) / 16) * 16);
          if (((size_t)unit * a.cap * 4) % 16 == 0) bulk_prefetch_l2(a.lidx + (size_t)unit * a.cap, lb);
          bulk_prefetch_l2(reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D, (uint32_t)(G * Cfg::ROW_BYTES));
          bulk_prefetch_l2(reinterpret_cast<const T*>(a.k_new) + (size_t)unit * D, (uint32_t)Cfg::ROW_BYTES);
          bulk_prefetch_l2(reinterpret_cast<const T*>(a.v_new) + (size_t)unit * D, (uint32_t)Cfg::ROW_BYTES);
        }
        for (int i = 0; i < 2 * nt; ++i) {
          if (use > 0) {
            if (a.timeline) {
              const long long c0 = clock64();
              mbar_wait(&empty[s], (use - 1) & 1);
              pwait += clock64() - c0;
            } else {
              mbar_wait(&empty[s], (use - 1) & 1);
            }
          }
          const int tt = i < nt ? i : i - nt;
          const int rows = min(TILE_ROWS, n_phys - tt * TILE_ROWS);
          const uint32_t bytes = (uint32_t)rows * Cfg::ROW_BYTES;
          const T* src = (i < nt ? Kg : Vg) + (size_t)tt * TILE_ROWS * D;
          mbar_arrive_expect_tx(&gfull[s], bytes);
          tma_bulk_g2s(ring + (size_t)s * Cfg::TILE_BYTES, src, bytes, &gfull[s], pol);
          if (++s == stages) { s = 0; ++use; }
        }
      }
      if (a.timeline) {                            // profiling: cycles the ring was full (nothing to request)
        unsigned long long* tlp = a.timeline + (size_t)blockIdx.x * 16 * 8 + 15 * 8;
        tlp[0] = (unsigned long long)pwait;
        tlp[1] = (unsigned long long)(clock64() - pstart);
      }
    }
    return;
  }

  // ===== consumers ===============================================================================
  const int gid = warp / NWARP;                 // consumer group
  const int tid = threadIdx.x - gid * NCONS, gw = warp - gid * NWARP;
  const Grp grp{tid, NCONS, 1 + gid};
  const int hw = tid >> 4, l16 = tid & 15;      // 16 half-warps per group
  unsigned char* gb = smem + L.off_grp + gid * L.grp_bytes;
  float* red = reinterpret_cast<float*>(gb + L.g_red);
  T* qh = reinterpret_cast<T*>(gb + L.g_q);
  T* kh = reinterpret_cast<T*>(gb + L.g_k);
  T* vh = reinterpret_cast<T*>(gb + L.g_v);
  int32_t* ns = reinterpret_cast<int32_t*>(gb + L.g_ns);
  int32_t* lj = reinterpret_cast<int32_t*>(gb + L.g_lj);
  T* plog = reinterpret_cast<T*>(gb + L.g_plog);
  unsigned char* scr = gb + L.g_scr;

  auto finish_logit = [&](float dot, bool valid) -> T {
    float x = Tr<T>::round_f(dot);                                           // llama_patch.py:201
    x = a.st.arith ? __fmul_rn(x, a.scale_mul) : __fdiv_rn(x, a.scale_div);  // :202
    return valid ? Tr<T>::from_f(x) : neg_inf<T>();                          // free slots are masked
  };

  // the k-th unit of this CTA is consumed by group k % ngroups; its tiles are [k*2nt, (k+1)*2nt)
  uint64_t* gfull = full + gid * Cfg::MAX_STAGES;
  uint32_t par = 0;                             // bit s: parity of this group's next wait on gfull[s]
  int k_unit = gid;
  unsigned long long* tl = a.timeline ? a.timeline + (size_t)blockIdx.x * 16 * 8 : nullptr;
  auto stamp = [&](int k, int slot) {
    if (tl && tid == 0 && k < 16) tl[k * 8 + slot] = clock64();
  };
  if (tl && threadIdx.x == 0) {
    unsigned long long gt;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    tl[15 * 8 + 7] = gt;
  }
  for (int unit = blockIdx.x + gid * gridDim.x; unit < U; unit += ngroups * gridDim.x, k_unit += ngroups) {
    int s = (int)(((long long)k_unit * 2 * nt) % stages);   // ring slot of this unit's first tile
    stamp(k_unit, 0);
    // ragged batches: this sequence's own count of valid slots; it evicts only past the budget gate (easykv.py:303)
    const int nb = EXT && a.seq_n_before ? a.seq_n_before[unit / a.Hkv] : a.n_before;
    const bool gated_off = EXT && a.st.budget_gate > 0 && nb + 1 - a.st.score_offset <= a.st.budget_gate;
    // ---- header: q, new K/V row, slot map (this group is idle until its first tile lands) --------------
    {
      const uint4* qg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D);
      const uint4* kg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.k_new) + (size_t)unit * D);
      const uint4* vg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.v_new) + (size_t)unit * D);
      constexpr int QCH = G * Cfg::ROW_BYTES / 16, RCH = Cfg::ROW_BYTES / 16;
      for (int i = tid; i < QCH + 2 * RCH; i += NCONS) {
        if (i < QCH) reinterpret_cast<uint4*>(qh)[i] = qg[i];
        else if (i < QCH + RCH) reinterpret_cast<uint4*>(kh)[i - QCH] = kg[i - QCH];
        else reinterpret_cast<uint4*>(vh)[i - QCH - RCH] = vg[i - QCH - RCH];
      }
      const int32_t* lg = a.lidx + (size_t)unit * a.cap;
      for (int e = tid; e < n_phys; e += NCONS) lj[e] = lg[e];
      if (tid == 0) {
        ns[0] = a.new_slots ? a.new_slots[unit] : n_phys;
        lj[n_phys] = nb;
      }
      // the tail reads this unit's S / SQ / C once: pull the lines into L2 now
      if (a.st.policy != EKV_POLICY_NONE && a.st.policy != EKV_POLICY_RANGE) {
        const int lines = (n_phys * 4 + 127) / 128;
        for (int i = tid; i < 3 * lines; i += NCONS) {
          const float* base = (i < lines ? a.S : (i < 2 * lines ? a.SQ : a.C)) + (size_t)unit * a.cap;
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + (size_t)(i % lines) * 128));
        }
      }
    }
    grp.sync();
    stamp(k_unit, 1);
    Row8<T> qr[G], knew;
#pragma unroll
    for (int g = 0; g < G; ++g) qr[g].load(qh + g * D, l16);
    knew.load(kh, l16);

    // ---- K phase ----------------------------------------------------------------------------------
    // Per tile a half-warp owns RPT rows -> NVT = RPT*G per-lane partial dots.  They are reduced across
    // the 16 lanes 16 values at a time (over TB tiles when a tile yields fewer), so that afterwards
    // every lane finishes exactly one logit.
    constexpr int NVT = RPT * G;
    constexpr int TB = NVT >= 16 ? 1 : 16 / NVT;       // tiles per batch
    constexpr int NB = NVT >= 16 ? NVT / 16 : 1;       // 16-value reductions per batch
    static_assert(NVT * TB == 16 * NB, "batch must be a whole number of 16-value reductions");
    const int vi0 = bitrev_idx<16>(l16);               // value index this lane finishes in each reduction
    float mloc = -INFINITY;                            // running max of the logits this lane stores
    for (int i0 = 0; i0 < nt; i0 += TB) {
      float part[NVT * TB];
#pragma unroll
      for (int tb = 0; tb < TB; ++tb) {
        if (i0 + tb < nt) {
          mbar_wait(&gfull[s], (par >> s) & 1u);
          par ^= 1u << s;
          if (i0 + tb == 0) stamp(k_unit, 2);
          const T* tile = reinterpret_cast<const T*>(ring + (size_t)s * Cfg::TILE_BYTES);
          Row8<T> x[RPT];
#pragma unroll
          for (int k = 0; k < RPT; ++k) x[k].load(tile + (hw * RPT + k) * D, l16);
          if (stream_rope) {                                   // fused streaming variant: rotate at the cache-relative position
#pragma unroll
            for (int k = 0; k < RPT; ++k) {
              const int e = (i0 + tb) * TILE_ROWS + hw * RPT + k;
              const int pos = e < n_phys ? max(lj[e], 0) : 0;
              rope_row8<T>(x[k], l16, a.rope_cos, a.rope_sin, pos);
            }
          }
#pragma unroll
          for (int k = 0; k < RPT; ++k)
#pragma unroll
            for (int g = 0; g < G; ++g) part[tb * NVT + k * G + g] = dot8(x[k], qr[g], 0.f);
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);
          s = s + 1 == stages ? 0 : s + 1;
        } else {
#pragma unroll
          for (int j = 0; j < NVT; ++j) part[tb * NVT + j] = 0.f;
        }
      }
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = part[b * 16 + j];
        const float r = transpose_reduce16<16>(v, l16);
        const int vi = b * 16 + vi0;                   // index into this batch's NVT*TB values
        const int tb = vi / NVT, k = (vi % NVT) / G, g = vi % G;
        const int e = (i0 + tb) * TILE_ROWS + hw * RPT + k;
        if (e < n_phys) {
          const T x = finish_logit(r, lj[e] >= 0);
          plog[g * nep + e] = x;
          mloc = fmaxf(mloc, Tr<T>::to_f(x));
        }
      }
    }
    // the appended token's own key (the reference attends it: llama_patch.py:193-196).  Every
    // half-warp computes it (the shuffles need all 32 lanes); half-warp 0 stores it.
    float xnew[G];
    {
      float v[G];
#pragma unroll
      for (int g = 0; g < G; ++g) v[g] = dot8(knew, qr[g], 0.f);
      const float r = transpose_reduce16<G>(v, l16);
      const T x = finish_logit(r, true);          // lanes l16 < G: head bitrev_idx<G>(l16)
      if (hw == 0 && l16 < G) plog[bitrev_idx<G>(l16) * nep + n_phys] = x;
#pragma unroll
      for (int g = 0; g < G; ++g)
        xnew[g] = __shfl_sync(0xffffffffu, Tr<T>::to_f(x), (lane & 16) | bitrev_idx<G>(g));
    }

    stamp(k_unit, 3);
    // ---- softmax (fp32 over the model-dtype logits; llama_patch.py:210-219) ---------------------------
    {
      float mx[G], inv[G], rcp[G];
      const int my_g = vi0 % G;                                  // head of the logits this lane stored
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float m = fmaxf(warp_max(my_g == g ? mloc : -INFINITY), xnew[g]);
        if (lane == 0) red[g * NWARP + gw] = m;
      }
      grp.sync();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float v = red[g * NWARP];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) v = fmaxf(v, red[g * NWARP + w]);
        mx[g] = v;
      }
      float* red2 = red + 8 * NWARP;
      float sacc[G];
#pragma unroll
      for (int g = 0; g < G; ++g) sacc[g] = 0.f;
#pragma unroll 5
      for (int e = tid; e < NE; e += NCONS)
#pragma unroll
        for (int g = 0; g < G; ++g) sacc[g] += expf(Tr<T>::to_f(plog[g * nep + e]) - mx[g]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        sacc[g] = warp_sum(sacc[g]);
        if (lane == 0) red2[g * NWARP + gw] = sacc[g];
      }
      grp.sync();
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float v = red2[g * NWARP];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) v += red2[g * NWARP + w];
        inv[g] = a.st.arith ? v : __fdiv_rn(1.0f, v);
        rcp[g] = __frcp_rn(v);
      }
#pragma unroll 5
      for (int e = tid; e < NE; e += NCONS)
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float ex = expf(Tr<T>::to_f(plog[g * nep + e]) - mx[g]);
          plog[g * nep + e] = Tr<T>::from_f(a.st.arith ? div_rn_by(ex, inv[g], rcp[g]) : __fmul_rn(ex, inv[g]));
        }
    }
    grp.sync();
    stamp(k_unit, 4);

    // ---- V phase ----------------------------------------------------------------------------------
    float oacc[G][8];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) oacc[g][j] = 0.f;
    // free slots contribute p == 0 exactly; their rows are stale but finite (the cache buffers are
    // zero-initialised and only ever hold rows that were valid), so no branch on p is needed
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&gfull[s], (par >> s) & 1u);
      par ^= 1u << s;
      const T* tile = reinterpret_cast<const T*>(ring + (size_t)s * Cfg::TILE_BYTES);
      const int e0 = i * TILE_ROWS + hw * RPT;
      Row8<T> x[RPT];
      T pv[RPT][G];
#pragma unroll
      for (int k = 0; k < RPT; ++k) {
        x[k].load(tile + (hw * RPT + k) * D, l16);
#pragma unroll
        for (int g = 0; g < G; ++g) pv[k][g] = plog[g * nep + min(e0 + k, n_phys)];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      s = s + 1 == stages ? 0 : s + 1;
#pragma unroll
      for (int k = 0; k < RPT; ++k)
        if (e0 + k < n_phys) {
#pragma unroll
          for (int g = 0; g < G; ++g) axpy8(pv[k][g], x[k], oacc[g]);
        }
    }
    stamp(k_unit, 5);
    if (hw == 0) {
      Row8<T> vnew;
      vnew.load(vh, l16);
#pragma unroll
      for (int g = 0; g < G; ++g) axpy8(plog[g * nep + n_phys], vnew, oacc[g]);
    }
    {
      float* part = reinterpret_cast<float*>(scr);          // [NWARP][G][D]
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float v = oacc[g][j] + __shfl_xor_sync(0xffffffffu, oacc[g][j], 16);
          if (lane < 16) part[(gw * G + g) * D + dim_of<T>(l16, j)] = v;
        }
      grp.sync();
      T* og = reinterpret_cast<T*>(a.out) + (size_t)unit * G * D;
      for (int i = tid; i < G * D; i += NCONS) {
        float v = part[i];
#pragma unroll
        for (int w = 1; w < NWARP; ++w) v += part[w * G * D + i];
        og[i] = Tr<T>::from_f(v);                                              // llama_patch.py:222
      }
    }
    grp.sync();
    stamp(k_unit, 6);

    // ---- tail: fold, accumulate, select, evict, append ------------------------------------------------
    SelScratch sc;
    sc.lj = lj;
    sc.carve(scr, NE, a.st.evict);
    if (tl && k_unit < 7) sc.dbg = tl + (8 + k_unit) * 8;
    UnitState u;
    u.S = a.S + (size_t)unit * a.cap; u.SQ = a.SQ + (size_t)unit * a.cap; u.C = a.C + (size_t)unit * a.cap;
    u.lidx = a.lidx + (size_t)unit * a.cap;
    u.new_slots = ns;
    u.victim_slots = a.victim_slots ? a.victim_slots + (size_t)unit * a.st.evict : nullptr;
    u.victim_lidx = a.victim_lidx ? a.victim_lidx + (size_t)unit * a.st.evict : nullptr;
    const float inv_g = 1.0f / (float)G;
    auto acc = [&](int e, float& ds, float& dsq) {
      float pf;
      if (G == 1) pf = Tr<T>::to_f(plog[e]);
      else {                                                  // process_for_mqa_gqa, easykv.py:188-196
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) sum += Tr<T>::to_f(plog[g * nep + e]);
        pf = Tr<T>::round_f(__fmul_rn(sum, inv_g));
      }
      ds = pf;
      dsq = Tr<T>::round_f(__fmul_rn(pf, pf));               // p**2 in the model dtype, easykv.py:296
    };
    // append the new row (this unit's K/V stream has been fully consumed)
    if (hw == 0) {
      const int slot = ns[0];
      float x[8];
      if (stream_rope) load_row8<T>(reinterpret_cast<const T*>(a.k_new_raw) + (size_t)unit * D, l16, x);   // the cache keeps un-rotated keys
      else load_row8<T>(kh, l16, x);
      store_row8<T>(reinterpret_cast<T*>(a.K) + ((size_t)unit * a.cap + slot) * D, l16, x);   // bit-exact round trip
      load_row8<T>(vh, l16, x);
      store_row8<T>(reinterpret_cast<T*>(a.V) + ((size_t)unit * a.cap + slot) * D, l16, x);
    }
    if (gated_off) {                                          // below the budget: accumulate and append only
      ekv_step stu = a.st;
      stu.evict = 0;
      state_select_apply(stu, u, nb, n_phys, 1, /*lj_preloaded=*/true, acc, sc, grp);
      if (tid == 0) {
        if (u.victim_lidx) u.victim_lidx[0] = -1;
        if (u.victim_slots) u.victim_slots[0] = -1;
      }
    } else {
      state_select_apply(a.st, u, nb, n_phys, 1, /*lj_preloaded=*/true, acc, sc, grp);
    }
    grp.sync();                                   // the header / scratch are rewritten for the next unit
    stamp(k_unit, 7);
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T, int G, int NG, bool EXT>
static int launch_decode_cfg_x(const KernelArgs& a, int grid, int stages, int smem_bytes, int dev, cudaStream_t stream) {
  static thread_local int configured[16] = {0};
  cudaError_t err;
  if (!configured[dev]) {
    err = cudaFuncSetAttribute(decode_kernel<T, G, NG, EXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(decode)", err);
    configured[dev] = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(NG * DecodeCfg<T>::NCONS + 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the kernel waits (pdl_wait) before its first global access
  attr[0].val.programmaticStreamSerializationAllowed = pdl_allowed();
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  err = cudaLaunchKernelEx(&cfg, decode_kernel<T, G, NG, EXT>, a, stages);
  if (err != cudaSuccess) return set_cuda_error("decode_kernel launch", err);
  count_launch();
  return EKV_OK;
}
template <typename T, int G, int NG>
static int launch_decode_cfg(const KernelArgs& a, int grid, int stages, int smem_bytes, int dev, cudaStream_t stream) {
  if (a.rope_cos || a.seq_n_before) return launch_decode_cfg_x<T, G, NG, true>(a, grid, stages, smem_bytes, dev, stream);
  return launch_decode_cfg_x<T, G, NG, false>(a, grid, stages, smem_bytes, dev, stream);
}

int decode_variant();   // ekv_api.cu (env EKV_DECODE_VARIANT): 0 = automatic, 1 = one group per CTA (2 CTAs/SM), 2 = ping-pong groups

template <typename T, int G> static int launch_decode_tg(const KernelArgs& a, cudaStream_t stream) {
  using Cfg = DecodeCfg<T>;
  static thread_local int sm_count[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (!sm_count[dev]) {
    cudaError_t err = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return set_cuda_error("cudaDeviceGetAttribute", err);
  }
  const int U = a.B * a.Hkv, sms = sm_count[dev];
  const int sm_total = 227 * 1024;
  const int variant = decode_variant() >= 5 ? 0 : decode_variant();     // 5 / 6 only steer the tcgen05 kernel (launch_decode)
  // automatic: more than one unit per SM -> one CTA per SM with two ping-pong consumer groups over a
  // shared ring (measured best: 0.87-0.93 of the HBM roofline vs 0.79-0.89 for two independent CTAs per
  // SM); otherwise one light CTA per unit.
  bool pingpong = variant == 2 || (variant == 0 && U > sms);
  if (pingpong) {
    const DecodeSmem<T> L2(G, a.n_phys, a.st.evict, 2);
    if (sm_total < L2.fixed || (sm_total - L2.fixed) / Cfg::TILE_BYTES < 4) pingpong = false;
  }
  if (!pingpong) {
    // one consumer group per CTA; with variant 1 two CTAs per SM (each with its own ring) once there is
    // more than one unit per SM and at least 3 ring stages fit in half an SM's shared memory
    const DecodeSmem<T> L(G, a.n_phys, a.st.evict, 1);
    int per_sm = (variant == 1 && U > sms) ? 2 : 1;
    int budget = per_sm == 2 ? sm_total / 2 - 1024 : sm_total;
    int stages = (budget - L.fixed) / Cfg::TILE_BYTES;
    if (per_sm == 2 && stages < 3) { per_sm = 1; budget = sm_total; stages = (budget - L.fixed) / Cfg::TILE_BYTES; }
    if (budget < L.fixed || stages < 2) return EKV_ERR_UNSUPPORTED;
    if (stages > Cfg::MAX_STAGES) stages = Cfg::MAX_STAGES;
    const int grid = U < sms * per_sm ? U : sms * per_sm;
    return launch_decode_cfg<T, G, 1>(a, grid, stages, L.fixed + stages * Cfg::TILE_BYTES, dev, stream);
  }
  const DecodeSmem<T> L(G, a.n_phys, a.st.evict, 2);
  int stages = (sm_total - L.fixed) / Cfg::TILE_BYTES;
  if (sm_total < L.fixed || stages < 4) return EKV_ERR_UNSUPPORTED;
  if (stages > Cfg::MAX_STAGES) stages = Cfg::MAX_STAGES;
  const int grid = U < sms ? U : sms;
  return launch_decode_cfg<T, G, 2>(a, grid, stages, L.fixed + stages * Cfg::TILE_BYTES, dev, stream);
}

template <typename T> static int launch_decode_t(const KernelArgs& a, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_decode_tg<T, 1>(a, stream);
    case 2: return launch_decode_tg<T, 2>(a, stream);
    case 4: return launch_decode_tg<T, 4>(a, stream);
    case 8: return launch_decode_tg<T, 8>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

int decode_cluster_size();   // ekv_api.cu (env EKV_DECODE_CLUSTER)

static int launch_decode_single(const KernelArgs& a, cudaStream_t stream) {
  switch (a.dtype) {
    case EKV_F16: return launch_decode_t<__half>(a, stream);
    case EKV_BF16: return launch_decode_t<__nv_bfloat16>(a, stream);
    case EKV_F32: return launch_decode_t<float>(a, stream);
    default: return EKV_ERR_INVALID;
  }
}

// Dispatch between the two decode kernels: fewer units than half the SMs -> split each unit over a
// cluster (ekv_decode_cluster.cu) so the whole chip streams; otherwise the persistent kernel above; and
// the cluster kernel again for units too large for one CTA's shared memory.
static int device_sm_count() {
  static thread_local int sm_count[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (!sm_count[dev] && cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count[dev] = 148;
  return sm_count[dev];
}

int launch_decode(const KernelArgs& a, cudaStream_t stream) {
  if (a.q_len != 1 || a.d != 128 || a.st.tova_head_mean) return EKV_ERR_UNSUPPORTED;
  if (a.rope_cos && a.dtype == EKV_F32) return EKV_ERR_UNSUPPORTED;       // fused streaming: 16-bit dtypes
  if (a.st.evict <= 1) {
    const int G = a.H / a.Hkv;
    // grouped-query layouts in a 16-bit dtype: the tcgen05 kernel (decode_variant 5 forces it for any g, 6 forbids it)
    const int dv = decode_variant();
    // automatic: long caches (>= 2048 slots) with enough units to fill the chip; short caches stay with the cluster kernel,
    // whose whole unit fits one light CTA
    if (!a.rope_cos && (dv == 5 || (dv == 0 && decode_cluster_size() == 0 && G >= 2 && a.dtype != EKV_F32 && a.B * a.Hkv >= 32 && a.n_phys >= 2048))) {
      const int rc = launch_decode_umma(a, stream);
      if (rc != EKV_ERR_UNSUPPORTED || dv == 5) return rc;
    }
    if (decode_cluster_size() > 0) return launch_decode_cluster(a, false, stream);
    // fewer units than half the SMs: split each unit over a cluster.  g >= 2: the persistent kernel's 544-thread
    // CTAs leave 96 registers per thread and spill (measured 0.40-0.45 of the roofline on the Mistral layout at any
    // cache size, 0.66 at g = 2); the cluster kernel's 288-thread CTAs do not (0.56-0.67 / 0.74, C = 1 included)
    if (decode_cluster_size() == 0 && (a.B * a.Hkv * 2 <= device_sm_count() || G >= 2)) {
      const int rc = launch_decode_cluster(a, G < 2, stream);
      if (rc != EKV_ERR_UNSUPPORTED) return rc;
    }
  }
  const int rc = launch_decode_single(a, stream);
  if (rc != EKV_ERR_UNSUPPORTED || a.st.evict > 1) return rc;
  return launch_decode_cluster(a, false, stream);
}

}  // namespace ekv
