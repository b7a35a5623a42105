// General fused forward for one layer: any q_len >= 1 (strided-prefill chunk with the causal mask
// inside the chunk, or a decode step), any supported dtype including fp32 — the exact-arithmetic
// CUDA-core path.  One CTA per (sequence, kv head); the g query heads of a GQA group and all q_len
// query rows share one pass over the head's K/V.
//
// Three sweeps over the key tiles per block of 64 (query, head) rows — row max, row sum with the
// final max (so the sum is formed exactly as softmax forms it), then probabilities -> P·V and the
// per-key column statistics — so the [H, q_len, n] probability tensor the reference materialises
// (easykv/llama_patch.py:244-246, kept alive for all layers) never exists.
//
// Replaces: llama_patch.py:193-230 / mistral_patch.py:137-170 and easykv.py:439-499 / :599-661 /
// :830-892 (fold, strided accumulate with model-dtype row sums, select, truncate_kv_cache_liso,
// state compaction) for one layer of one forward.
#include "ekv_select.cuh"
#include "ekv_kernels.h"

namespace ekv {

namespace gen {
constexpr int NT = 256;          // threads
constexpr int NW = NT / 32;      // warps
constexpr int RB = 64;           // (query, head) rows per block = NW * 8
constexpr int TK = 32;           // keys per tile (one per lane)
constexpr int PS = TK + 1;
}  // namespace gen
// head_dim D is a template parameter of this kernel (64, 96, 128: the reference is generic in it, llama_patch.py:169-172);
// the tensor-core and decode kernels are built for 128 and decline anything else, which then lands here.

struct GenSmem {
  int off_ns, off_lj, off_colS, off_colSQ, off_pool, total;
  // inside the pool (attention phase)
  int off_q, off_k, off_v, off_p, off_cpart;
  __host__ __device__ GenSmem(int n_phys, int q_len, int evict, int D) {
    using namespace gen;
    const int KS = D + 4;   // padded fp32 row stride of the K tile (conflict-free 128-bit reads: KS % 32 == 4 for D = 64, 96, 128)
    const int NE = n_phys + q_len;
    int o = 0;
    off_ns = o; o += (q_len * 4 + 15) / 16 * 16;
    off_lj = o; o += (NE * 4 + 15) / 16 * 16;
    off_colS = o; o += (NE * 4 + 15) / 16 * 16;
    off_colSQ = o; o += (NE * 4 + 15) / 16 * 16;
    o = (o + 127) / 128 * 128;
    off_pool = o;
    int p = 0;
    off_q = p; p += RB * D * 4;
    off_k = p; p += TK * KS * 4;
    off_v = p; p += TK * D * 4;
    off_p = p; p += RB * PS * 4;
    off_cpart = p; p += NW * TK * 2 * 4;
    size_t pool = (size_t)p;
    size_t sel = SelScratch::bytes(NE, evict);
    if (sel > pool) pool = sel;
    total = o + (int)pool;
  }
};

template <typename T> __device__ __forceinline__ void load4(const T* p, float (&x)[4]);
template <> __device__ __forceinline__ void load4<__half>(const __half* p, float (&x)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = Tr<__half>::to_f2(u.x), b = Tr<__half>::to_f2(u.y);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <> __device__ __forceinline__ void load4<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[4]) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = Tr<__nv_bfloat16>::to_f2(u.x), b = Tr<__nv_bfloat16>::to_f2(u.y);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <> __device__ __forceinline__ void load4<float>(const float* p, float (&x)[4]) {
  float4 u = *reinterpret_cast<const float4*>(p);
  x[0] = u.x; x[1] = u.y; x[2] = u.z; x[3] = u.w;
}

template <typename T, int G, int D>
__global__ void __launch_bounds__(gen::NT) general_kernel(const KernelArgs a) {
  using namespace gen;
  constexpr int KS = D + 4;
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_phys = a.n_phys, QL = a.q_len, NE = n_phys + QL, R = QL * G;
  const GenSmem L(n_phys, QL, a.st.evict, D);
  int32_t* ns = reinterpret_cast<int32_t*>(smem + L.off_ns);
  int32_t* lj = reinterpret_cast<int32_t*>(smem + L.off_lj);
  float* colS = reinterpret_cast<float*>(smem + L.off_colS);
  float* colSQ = reinterpret_cast<float*>(smem + L.off_colSQ);
  unsigned char* pool = smem + L.off_pool;
  float* qblk = reinterpret_cast<float*>(pool + L.off_q);
  float* ktile = reinterpret_cast<float*>(pool + L.off_k);
  float* vtile = reinterpret_cast<float*>(pool + L.off_v);
  float* ptile = reinterpret_cast<float*>(pool + L.off_p);
  float* cpart = reinterpret_cast<float*>(pool + L.off_cpart);

  const int unit = blockIdx.x, b = unit / a.Hkv, h = unit % a.Hkv;
  const int nb = a.seq_n_before ? a.seq_n_before[b] : a.n_before;            // ragged batches: this sequence's valid slots
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const Grp grp{tid, NT, 0};
  const T* Kg = reinterpret_cast<const T*>(a.K) + (size_t)unit * a.cap * D;
  const T* Vg = reinterpret_cast<const T*>(a.V) + (size_t)unit * a.cap * D;
  const T* kn = reinterpret_cast<const T*>(a.k_new) + (size_t)unit * QL * D;
  const T* vn = reinterpret_cast<const T*>(a.v_new) + (size_t)unit * QL * D;
  const T* qg = reinterpret_cast<const T*>(a.q) + (size_t)b * a.H * QL * D;      // [H, QL, D]
  T* og = reinterpret_cast<T*>(a.out) + (size_t)b * a.H * QL * D;

  {
    const int32_t* lg = a.lidx + (size_t)unit * a.cap;
    for (int e = tid; e < n_phys; e += NT) lj[e] = lg[e];
    for (int i = tid; i < QL; i += NT) {
      ns[i] = a.new_slots ? a.new_slots[(size_t)unit * QL + i] : n_phys + i;
      lj[n_phys + i] = nb + i;
    }
    for (int e = tid; e < NE; e += NT) { colS[e] = 0.f; colSQ[e] = 0.f; }
  }
  __syncthreads();

  const int ntile = (NE + TK - 1) / TK;
  const bool tova = a.st.policy == EKV_POLICY_TOVA;
  const float inv_g = 1.0f / (float)G;

  auto load_tile = [&](int tile, const T* base_old, const T* base_new, float* dst, int stride, bool zero_free) {
    for (int idx = tid; idx < TK * (D / 4); idx += NT) {
      const int kk = idx / (D / 4), c4 = idx % (D / 4);
      const int e = tile * TK + kk;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (e < NE && !(zero_free && e < n_phys && lj[e] < 0)) {
        const T* src = e < n_phys ? base_old + (size_t)e * D : base_new + (size_t)(e - n_phys) * D;
        load4<T>(src + c4 * 4, x);
      }
      *reinterpret_cast<float4*>(dst + kk * stride + c4 * 4) = make_float4(x[0], x[1], x[2], x[3]);
    }
  };

  for (int rb = 0; rb * RB < R; ++rb) {
    const int row0 = rb * RB + warp * 8;         // this warp's first (query-major) row
    // ---- q block -> fp32 shared ----------------------------------------------------------------
    __syncthreads();
    for (int idx = tid; idx < RB * (D / 4); idx += NT) {
      const int rr = idx / (D / 4), c4 = idx % (D / 4);
      const int r = rb * RB + rr;
      float x[4] = {0.f, 0.f, 0.f, 0.f};
      if (r < R) {
        const int i = r / G, g = r % G;
        load4<T>(qg + ((size_t)(h * G + g) * QL + i) * D + c4 * 4, x);
      }
      *reinterpret_cast<float4*>(qblk + rr * D + c4 * 4) = make_float4(x[0], x[1], x[2], x[3]);
    }
    float m[8], lsum[8], acc[8][4];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      m[r] = -INFINITY; lsum[r] = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    }

    // logits of this lane's key against the warp's 8 rows, at the reference's rounding points
    auto logits8 = [&](int tile, float (&x)[8]) {
      const int e = tile * TK + lane;
      float s[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) s[r] = 0.f;
      const float* kr = ktile + lane * KS;
      const float* qr = qblk + warp * 8 * D;
#pragma unroll 4
      for (int c = 0; c < D; c += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kr + c);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float4 qv = *reinterpret_cast<const float4*>(qr + r * D + c);
          s[r] = fmaf(kv.x, qv.x, s[r]); s[r] = fmaf(kv.y, qv.y, s[r]);
          s[r] = fmaf(kv.z, qv.z, s[r]); s[r] = fmaf(kv.w, qv.w, s[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int row = row0 + r;
        bool vis = row < R && e < NE;
        if (vis) vis = e < n_phys ? lj[e] >= 0 : (e - n_phys) <= row / G;   // causal inside the chunk
        float v = Tr<T>::round_f(s[r]);                                             // llama_patch.py:201
        v = a.st.arith ? __fmul_rn(v, a.scale_mul) : __fdiv_rn(v, a.scale_div);     // :202
        x[r] = vis ? Tr<T>::round_f(v) : -INFINITY;                                 // :210-215
      }
    };

    // ---- sweep A: row max --------------------------------------------------------------------------
    for (int tile = 0; tile < ntile; ++tile) {
      __syncthreads();
      load_tile(tile, Kg, kn, ktile, KS, false);
      __syncthreads();
      float x[8];
      logits8(tile, x);
#pragma unroll
      for (int r = 0; r < 8; ++r) m[r] = fmaxf(m[r], x[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) m[r] = warp_max(m[r]);
    // ---- sweep B: row sum of exp(x - max) (llama_patch.py:218) --------------------------------------
    for (int tile = 0; tile < ntile; ++tile) {
      __syncthreads();
      load_tile(tile, Kg, kn, ktile, KS, false);
      __syncthreads();
      float x[8];
      logits8(tile, x);
#pragma unroll
      for (int r = 0; r < 8; ++r) lsum[r] += (x[r] == -INFINITY) ? 0.f : expf(x[r] - m[r]);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      lsum[r] = warp_sum(lsum[r]);
      if (!a.st.arith) lsum[r] = __fdiv_rn(1.0f, lsum[r]);
    }
    // ---- sweep C: probabilities -> column statistics and P·V --------------------------------------
    for (int tile = 0; tile < ntile; ++tile) {
      __syncthreads();
      load_tile(tile, Kg, kn, ktile, KS, false);
      load_tile(tile, Vg, vn, vtile, D, true);
      __syncthreads();
      float x[8], p[8];
      logits8(tile, x);
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float ex = (x[r] == -INFINITY) ? 0.f : expf(x[r] - m[r]);
        ex = a.st.arith ? __fdiv_rn(ex, lsum[r]) : __fmul_rn(ex, lsum[r]);
        p[r] = (row0 + r < R) ? Tr<T>::round_f(ex) : 0.f;                          // :219
        ptile[(warp * 8 + r) * PS + lane] = p[r];
      }
      // GQA fold (easykv.py:188-196) and this warp's share of the row sums (:450-451)
      float cs = 0.f, csq = 0.f;
#pragma unroll
      for (int qi = 0; qi < 8 / G; ++qi) {
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < G; ++g) sum += p[qi * G + g];
        const float pf = (G == 1) ? sum : Tr<T>::round_f(__fmul_rn(sum, inv_g));
        const int i = (row0 + qi * G) / G;
        if (row0 + qi * G < R && (!tova || i == QL - 1)) {
          cs += pf;
          csq += Tr<T>::round_f(__fmul_rn(pf, pf));
        }
      }
      cpart[(warp * TK + lane) * 2] = cs;
      cpart[(warp * TK + lane) * 2 + 1] = csq;
      __syncthreads();
      if (warp == 0) {
        const int e = tile * TK + lane;
        if (e < NE) {
          float s1 = colS[e], s2 = colSQ[e];
#pragma unroll
          for (int w = 0; w < NW; ++w) { s1 += cpart[(w * TK + lane) * 2]; s2 += cpart[(w * TK + lane) * 2 + 1]; }
          colS[e] = s1; colSQ[e] = s2;
        }
      }
      // P·V: lane owns output dims [4*lane, 4*lane+4) of the warp's 8 rows (lanes past D / 4 idle for D < 128)
      if (lane * 4 < D) {
#pragma unroll 4
      for (int kk = 0; kk < TK; ++kk) {
        const float4 vv = *reinterpret_cast<const float4*>(vtile + kk * D + lane * 4);
#pragma unroll
        for (int r = 0; r < 8; ++r) {
          const float pr = ptile[(warp * 8 + r) * PS + kk];
          acc[r][0] = fmaf(pr, vv.x, acc[r][0]); acc[r][1] = fmaf(pr, vv.y, acc[r][1]);
          acc[r][2] = fmaf(pr, vv.z, acc[r][2]); acc[r][3] = fmaf(pr, vv.w, acc[r][3]);
        }
      }
      }
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int row = row0 + r;
      if (row < R && lane * 4 < D) {
        const int i = row / G, g = row % G;
        T* o = og + ((size_t)(h * G + g) * QL + i) * D + lane * 4;
#pragma unroll
        for (int c = 0; c < 4; ++c) o[c] = Tr<T>::from_f(acc[r][c]);               // :222
      }
    }
  }
  __syncthreads();

  // ---- append the chunk's K/V rows ---------------------------------------------------------------------
  {
    T* Kw = reinterpret_cast<T*>(a.K) + (size_t)unit * a.cap * D;
    T* Vw = reinterpret_cast<T*>(a.V) + (size_t)unit * a.cap * D;
    for (int idx = tid; idx < QL * D; idx += NT) {
      const int i = idx / D, c = idx % D;
      Kw[(size_t)ns[i] * D + c] = kn[(size_t)i * D + c];
      Vw[(size_t)ns[i] * D + c] = vn[(size_t)i * D + c];
    }
  }

  // ---- tail ---------------------------------------------------------------------------------------------
  SelScratch sc;
  sc.lj = lj;
  sc.carve(pool, NE, a.st.evict);
  UnitState u;
  u.S = a.S + (size_t)unit * a.cap; u.SQ = a.SQ + (size_t)unit * a.cap; u.C = a.C + (size_t)unit * a.cap;
  u.lidx = a.lidx + (size_t)unit * a.cap;
  u.new_slots = ns;
  u.victim_slots = a.victim_slots ? a.victim_slots + (size_t)unit * a.st.evict : nullptr;
  u.victim_lidx = a.victim_lidx ? a.victim_lidx + (size_t)unit * a.st.evict : nullptr;
  auto accf = [&](int e, float& ds, float& dsq) {
    ds = a.st.raw_colsum ? colS[e] : Tr<T>::round_f(colS[e]);       // p.sum(dim=1) is a model-dtype result (easykv.py:450)
    dsq = a.st.raw_colsum ? colSQ[e] : Tr<T>::round_f(colSQ[e]);    // (p**2).sum(dim=1) likewise (:451)
  };
  if (a.st.budget_gate > 0 && nb + QL - a.st.score_offset <= a.st.budget_gate) {      // ragged batches: this sequence is below the budget
    ekv_step stu = a.st;
    stu.evict = 0;
    state_select_apply(stu, u, nb, n_phys, QL, /*lj_preloaded=*/true, accf, sc, grp);
    for (int t = tid; t < a.st.evict; t += gen::NT) {
      if (u.victim_lidx) u.victim_lidx[t] = -1;
      if (u.victim_slots) u.victim_slots[t] = -1;
    }
  } else {
    state_select_apply(a.st, u, nb, n_phys, QL, /*lj_preloaded=*/true, accf, sc, grp);
  }
}

// largest n_phys + q_len the general kernel can hold when it selects `evict` victims (ekv_chunk_entry_limit)
int general_entry_limit(int q_len, int evict, int d) {
  int lo = 0, hi = 1 << 22;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (GenSmem(mid, q_len, evict, d).total <= 227 * 1024) lo = mid; else hi = mid - 1;
  }
  return lo + q_len;
}

template <typename T, int G, int D> static int launch_general_tgd(const KernelArgs& a, cudaStream_t stream) {
  const GenSmem L(a.n_phys, a.q_len, a.st.evict, D);
  if (L.total > 227 * 1024) return set_error(EKV_ERR_UNSUPPORTED, "general kernel: %d bytes of shared memory needed", L.total);
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t err;
  if (dev < 16 && !configured[dev]) {
    err = cudaFuncSetAttribute(general_kernel<T, G, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(general)", err);
    configured[dev] = 1;
  }
  general_kernel<T, G, D><<<a.B * a.Hkv, gen::NT, L.total, stream>>>(a);
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("general_kernel launch", err);
  count_launch();
  return EKV_OK;
}

template <typename T, int G> static int launch_general_tg(const KernelArgs& a, cudaStream_t stream) {
  switch (a.d) {
    case 64: return launch_general_tgd<T, G, 64>(a, stream);
    case 96: return launch_general_tgd<T, G, 96>(a, stream);
    case 128: return launch_general_tgd<T, G, 128>(a, stream);
    default: return set_error(EKV_ERR_UNSUPPORTED, "head_dim %d (built: 64, 96, 128)", a.d);
  }
}

template <typename T> static int launch_general_t(const KernelArgs& a, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_general_tg<T, 1>(a, stream);
    case 2: return launch_general_tg<T, 2>(a, stream);
    case 4: return launch_general_tg<T, 4>(a, stream);
    case 8: return launch_general_tg<T, 8>(a, stream);
    default: return set_error(EKV_ERR_UNSUPPORTED, "GQA group size %d not in {1,2,4,8}", a.H / a.Hkv);
  }
}

int launch_general(const KernelArgs& a, cudaStream_t stream) {
  switch (a.dtype) {
    case EKV_F16: return launch_general_t<__half>(a, stream);
    case EKV_BF16: return launch_general_t<__nv_bfloat16>(a, stream);
    case EKV_F32: return launch_general_t<float>(a, stream);
    default: return set_error(EKV_ERR_INVALID, "dtype %d", a.dtype);
  }
}

}  // namespace ekv
