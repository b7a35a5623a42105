// Fused decode step (q_len == 1) for one layer: one CTA per (sequence, kv head).
//
//   TMA producer warp : streams the head's K rows, then its V rows, HBM -> shared memory, as 8 KB
//                       cp.async.bulk tiles through a STAGES-deep mbarrier ring (each K/V byte is
//                       read from HBM exactly once; the g query heads of a GQA group share it).
//   4 consumer warps  : K phase  — 16 lanes per row, 128-bit shared-memory reads, fp32 FMA dot
//                                  products, transposing warp-shuffle reduction, logits rounded
//                                  at the reference's rounding points;
//                       softmax  — fp32, warp-shuffle + named-barrier reductions, probabilities
//                                  rounded to the model dtype;
//                       V phase  — fp32 FMA accumulate of p·V, cross-warp reduction, out;
//                       tail     — GQA fold, policy accumulate, victim select, in-place
//                                  eviction (ekv_select.cuh) and the append of the new K/V row.
//
// Replaces (reference paths): llama_patch.py:193-230 / mistral_patch.py:137-170 (cache append,
// repeat_kv, QK^T, mask, softmax, PV) and easykv.py:271-362 / :683-748 (fold, accumulate, select,
// truncate_kv_cache_silo, state compaction) for one layer of one decode forward.
#include "ekv_select.cuh"
#include "ekv_kernels.h"

namespace ekv {

template <typename T> struct DecodeCfg {
  static constexpr int D = 128;
  static constexpr int NWARP = 4;                       // consumer warps
  static constexpr int NCONS = NWARP * 32;
  static constexpr int NTHREADS = NCONS + 32;           // + the TMA producer warp
  static constexpr int ROW_BYTES = D * (int)sizeof(T);
  static constexpr int TILE_BYTES = 8192;
  static constexpr int TILE_ROWS = TILE_BYTES / ROW_BYTES;   // 32 (16-bit) / 16 (fp32)
  static constexpr int RPT = TILE_ROWS / (NWARP * 2);        // rows per 16-lane group per tile
  static constexpr int STAGES = 4;
};

// sum over the 16 lanes of a half-warp of NV per-lane values; afterwards lane l (< NV) of the
// group holds the total of value index bitrev_{log2 NV}(l).  NV-1 + log2(16/NV) shuffles instead
// of 4*NV.
template <int NV> __device__ __forceinline__ float transpose_reduce16(float (&v)[NV], int l16) {
  int bit = 1;
#pragma unroll
  for (int w = NV / 2; w >= 1; w >>= 1) {
    const bool up = (l16 & bit) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const float send = up ? v[i] : v[i + w];
      const float keep = up ? v[i + w] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
    bit <<= 1;
  }
  float r = v[0];
#pragma unroll
  for (; bit < 16; bit <<= 1) r += __shfl_xor_sync(0xffffffffu, r, bit);
  return r;
}
template <int NV> __device__ __forceinline__ int bitrev_idx(int l) {
  int r = 0;
#pragma unroll
  for (int w = NV / 2, b = 1; w >= 1; w >>= 1, b <<= 1) r += (l & b) ? w : 0;
  return r;
}

template <typename T> struct DecodeSmem {
  // byte offsets inside dynamic shared memory
  int off_bar, off_q, off_red, off_ns, off_lj, off_plog, off_pool, total;
  int nep;   // padded entries per head in plog
  __host__ __device__ DecodeSmem(int G, int n_phys, int evict) {
    using Cfg = DecodeCfg<T>;
    const int NE = n_phys + 1;
    nep = (NE + 7) / 8 * 8;
    int o = 0;
    off_bar = o; o += 2 * Cfg::STAGES * 8;
    off_q = o; o += G * Cfg::D * 4;
    off_red = o; o += 8 * Cfg::NWARP * 4 * 2;
    off_ns = o; o += 16;
    off_lj = o; o += (NE * 4 + 15) / 16 * 16;
    off_plog = o; o += (G * nep * (int)sizeof(T) + 15) / 16 * 16;
    o = (o + 127) / 128 * 128;
    off_pool = o;
    size_t pool = (size_t)Cfg::STAGES * Cfg::TILE_BYTES;
    size_t sel = SelScratch::bytes(NE, evict);
    size_t outp = (size_t)Cfg::NWARP * 2 * G * Cfg::D * 4;
    if (sel > pool) pool = sel;
    if (outp > pool) pool = outp;
    total = o + (int)pool;
  }
};

template <typename T, int G>
__global__ void __launch_bounds__(DecodeCfg<T>::NTHREADS)
decode_kernel(const KernelArgs a) {
  using Cfg = DecodeCfg<T>;
  constexpr int D = Cfg::D, NWARP = Cfg::NWARP, NCONS = Cfg::NCONS, RPT = Cfg::RPT, STAGES = Cfg::STAGES;
  constexpr int TILE_ROWS = Cfg::TILE_ROWS;
  extern __shared__ __align__(128) unsigned char smem[];
  const DecodeSmem<T> L(G, a.n_phys, a.st.evict);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* empty = full + STAGES;
  float* qs = reinterpret_cast<float*>(smem + L.off_q);
  float* red = reinterpret_cast<float*>(smem + L.off_red);
  int32_t* ns = reinterpret_cast<int32_t*>(smem + L.off_ns);
  int32_t* lj = reinterpret_cast<int32_t*>(smem + L.off_lj);
  T* plog = reinterpret_cast<T*>(smem + L.off_plog);
  unsigned char* pool = smem + L.off_pool;

  const int unit = blockIdx.x;                 // b * Hkv + h
  const int n_phys = a.n_phys, NE = n_phys + 1, nep = L.nep;
  const int nt = (n_phys + TILE_ROWS - 1) / TILE_ROWS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const T* Kg = reinterpret_cast<const T*>(a.K) + (size_t)unit * a.cap * D;
  const T* Vg = reinterpret_cast<const T*>(a.V) + (size_t)unit * a.cap * D;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NWARP); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == NWARP) {
    // ===== TMA producer ======================================================================
    if (lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      for (int t = 0; t < 2 * nt; ++t) {
        const int s = t % STAGES, use = t / STAGES;
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        const int tt = t < nt ? t : t - nt;
        const int rows = min(TILE_ROWS, n_phys - tt * TILE_ROWS);
        const uint32_t bytes = (uint32_t)rows * Cfg::ROW_BYTES;
        const T* src = (t < nt ? Kg : Vg) + (size_t)tt * TILE_ROWS * D;
        mbar_arrive_expect_tx(&full[s], bytes);
        tma_bulk_g2s(pool + (size_t)s * Cfg::TILE_BYTES, src, bytes, &full[s], pol);
      }
    }
    return;
  }

  // ===== consumers ===============================================================================
  const Grp grp{tid, NCONS, 1};
  const int hw = tid >> 4, l16 = tid & 15;      // 8 half-warps
  const size_t unit_q = (size_t)unit * G * D;   // q/out: [B, H, 1, D] with H = Hkv * G
  const size_t unit_kv = (size_t)unit * D;      // k_new/v_new: [B, Hkv, 1, D]

  // stage q (fp32), lidx, the new slot; preload the new token's K/V chunk
  {
    const T* qg = reinterpret_cast<const T*>(a.q) + unit_q;
    for (int i = tid; i < G * D; i += NCONS) qs[i] = Tr<T>::to_f(qg[i]);
    const int32_t* lg = a.lidx + (size_t)unit * a.cap;
    for (int e = tid; e < n_phys; e += NCONS) lj[e] = lg[e];
    if (tid == 0) {
      ns[0] = a.new_slots ? a.new_slots[unit] : n_phys;
      lj[n_phys] = a.n_before;
    }
  }
  float knew[8], vnew[8];
  load_row8<T>(reinterpret_cast<const T*>(a.k_new) + unit_kv, l16, knew);
  load_row8<T>(reinterpret_cast<const T*>(a.v_new) + unit_kv, l16, vnew);
  grp.sync();
  float qr[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) qr[g][i] = qs[g * D + dim_of<T>(l16, i)];

  auto finish_logit = [&](float dot, bool valid) -> T {
    float x = Tr<T>::round_f(dot);                                           // llama_patch.py:201
    x = a.st.arith ? __fmul_rn(x, a.scale_mul) : __fdiv_rn(x, a.scale_div);  // :202
    return valid ? Tr<T>::from_f(x) : neg_inf<T>();
  };

  // ---- K phase ------------------------------------------------------------------------------------
  constexpr int NVT = RPT * G;                       // values per half-warp per tile
  constexpr int NV = NVT < 16 ? NVT : 16;            // values per transposing reduction
  constexpr int NB = NVT / NV;
  for (int t = 0; t < nt; ++t) {
    const int s = t % STAGES;
    mbar_wait(&full[s], (t / STAGES) & 1);
    const T* tile = reinterpret_cast<const T*>(pool + (size_t)s * Cfg::TILE_BYTES);
    float part[NVT];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      float x[8];
      load_row8<T>(tile + (hw * RPT + k) * D, l16, x);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(x[i], qr[g][i], acc);
        part[k * G + g] = acc;
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = part[b * NV + i];
      const float r = transpose_reduce16<NV>(v, l16);
      if (l16 < NV) {
        const int vi = b * NV + bitrev_idx<NV>(l16);
        const int k = vi / G, g = vi % G;
        const int e = t * TILE_ROWS + hw * RPT + k;
        if (e < n_phys) plog[g * nep + e] = finish_logit(r, lj[e] >= 0);
      }
    }
  }
  // the appended token's own key (the reference attends it: llama_patch.py:193-196).  Every
  // half-warp computes it (the shuffles need all 32 lanes); half-warp 0 stores it.
  {
    float v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) acc = fmaf(knew[i], qr[g][i], acc);
      v[g] = acc;
    }
    const float r = transpose_reduce16<G>(v, l16);
    if (hw == 0 && l16 < G) plog[bitrev_idx<G>(l16) * nep + n_phys] = finish_logit(r, true);
  }
  grp.sync();

  // ---- softmax (fp32 over the model-dtype logits; llama_patch.py:218-219) -------------------------
  float mx[G], inv[G];
  {
    float m[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
      m[g] = -INFINITY;
      for (int e = tid; e < NE; e += NCONS) m[g] = fmaxf(m[g], Tr<T>::to_f(plog[g * nep + e]));
      m[g] = warp_max(m[g]);
      if (lane == 0) red[g * NWARP + warp] = m[g];
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
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float sacc = 0.f;
      for (int e = tid; e < NE; e += NCONS) sacc += expf(Tr<T>::to_f(plog[g * nep + e]) - mx[g]);
      sacc = warp_sum(sacc);
      if (lane == 0) red2[g * NWARP + warp] = sacc;
    }
    grp.sync();
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float v = red2[g * NWARP];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) v += red2[g * NWARP + w];
      inv[g] = a.st.arith ? v : __fdiv_rn(1.0f, v);
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
      for (int e = tid; e < NE; e += NCONS) {
        const float ex = expf(Tr<T>::to_f(plog[g * nep + e]) - mx[g]);
        plog[g * nep + e] = Tr<T>::from_f(a.st.arith ? __fdiv_rn(ex, inv[g]) : __fmul_rn(ex, inv[g]));
      }
  }
  grp.sync();

  // ---- V phase ------------------------------------------------------------------------------------
  float oacc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int i = 0; i < 8; ++i) oacc[g][i] = 0.f;
  for (int t = nt; t < 2 * nt; ++t) {
    const int s = t % STAGES;
    mbar_wait(&full[s], (t / STAGES) & 1);
    const T* tile = reinterpret_cast<const T*>(pool + (size_t)s * Cfg::TILE_BYTES);
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      const int e = (t - nt) * TILE_ROWS + hw * RPT + k;
      if (e < n_phys) {
        float pv[G];
        bool any = false;
#pragma unroll
        for (int g = 0; g < G; ++g) { pv[g] = Tr<T>::to_f(plog[g * nep + e]); any |= pv[g] != 0.f; }
        if (any) {
          float x[8];
          load_row8<T>(tile + (hw * RPT + k) * D, l16, x);
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int i = 0; i < 8; ++i) oacc[g][i] = fmaf(pv[g], x[i], oacc[g][i]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
  }
  if (hw == 0) {
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float p = Tr<T>::to_f(plog[g * nep + n_phys]);
#pragma unroll
      for (int i = 0; i < 8; ++i) oacc[g][i] = fmaf(p, vnew[i], oacc[g][i]);
    }
  }
  grp.sync();                                   // every stage buffer is consumed: the pool is free
  {
    float* part = reinterpret_cast<float*>(pool);          // [8 half-warps][G][D]
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int i = 0; i < 8; ++i) part[(hw * G + g) * D + dim_of<T>(l16, i)] = oacc[g][i];
    grp.sync();
    T* og = reinterpret_cast<T*>(a.out) + unit_q;
    for (int i = tid; i < G * D; i += NCONS) {
      float v = part[i];
#pragma unroll
      for (int h = 1; h < NWARP * 2; ++h) v += part[h * G * D + i];
      og[i] = Tr<T>::from_f(v);                                              // llama_patch.py:222
    }
    grp.sync();
  }

  // ---- tail: fold, accumulate, select, evict, append ------------------------------------------------
  SelScratch sc;
  sc.lj = lj;
  sc.carve(pool, NE, a.st.evict);
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
  // append the new row (nobody reads K/V any more in this launch)
  if (hw == 0) {
    const int slot = ns[0];
    store_row8<T>(reinterpret_cast<T*>(a.K) + ((size_t)unit * a.cap + slot) * D, l16, knew);
    store_row8<T>(reinterpret_cast<T*>(a.V) + ((size_t)unit * a.cap + slot) * D, l16, vnew);
  }
  state_select_apply(a.st, u, a.n_before, n_phys, 1, /*lj_preloaded=*/true, acc, sc, grp);
}

// ---------------------------------------------------------------------------------------------------
template <typename T, int G> static int launch_decode_tg(const KernelArgs& a, cudaStream_t stream) {
  const DecodeSmem<T> L(G, a.n_phys, a.st.evict);
  if (L.total > 227 * 1024) return EKV_ERR_UNSUPPORTED;
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t err;
  if (dev < 16 && configured[dev] < L.total) {
    err = cudaFuncSetAttribute(decode_kernel<T, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(decode)", err);
    configured[dev] = 227 * 1024;
  }
  decode_kernel<T, G><<<a.B * a.Hkv, DecodeCfg<T>::NTHREADS, L.total, stream>>>(a);
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("decode_kernel launch", err);
  count_launch();
  return EKV_OK;
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

int launch_decode(const KernelArgs& a, cudaStream_t stream) {
  if (a.q_len != 1 || a.d != 128 || a.st.tova_head_mean) return EKV_ERR_UNSUPPORTED;
  switch (a.dtype) {
    case EKV_F16: return launch_decode_t<__half>(a, stream);
    case EKV_BF16: return launch_decode_t<__nv_bfloat16>(a, stream);
    case EKV_F32: return launch_decode_t<float>(a, stream);
    default: return EKV_ERR_INVALID;
  }
}

}  // namespace ekv
