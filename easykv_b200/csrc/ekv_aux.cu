// Stand-alone pieces of the path: select over existing state, explicit eviction, tova's
// cross-head mean, and the export of the reference's arrival-ordered view.
#include "ekv_select.cuh"
#include "ekv_kernels.h"

namespace ekv {

constexpr int AUX_NT = 256;

struct AuxSmem {
  int off_lj, off_pool, total;
  __host__ __device__ AuxSmem(int NE, int evict) {
    off_lj = 0;
    off_pool = ((NE * 4 + 15) / 16 * 16 + 127) / 128 * 128;
    total = off_pool + (int)SelScratch::bytes(NE, evict);
  }
};

// ---- ekv_select: easykv.py:310-347 / :462-493 in isolation ---------------------------------------
__global__ void __launch_bounds__(AUX_NT) select_kernel(const KernelArgs a) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int unit = blockIdx.x;
  const AuxSmem L(a.n_phys, a.st.evict);
  SelScratch sc;
  sc.lj = reinterpret_cast<int32_t*>(smem + L.off_lj);
  sc.carve(smem + L.off_pool, a.n_phys, a.st.evict);
  UnitState u;
  u.S = a.S + (size_t)unit * a.cap; u.SQ = a.SQ + (size_t)unit * a.cap; u.C = a.C + (size_t)unit * a.cap;
  u.lidx = a.lidx + (size_t)unit * a.cap;
  u.new_slots = nullptr;
  u.victim_slots = a.victim_slots ? a.victim_slots + (size_t)unit * a.st.evict : nullptr;
  u.victim_lidx = a.victim_lidx ? a.victim_lidx + (size_t)unit * a.st.evict : nullptr;
  const Grp grp{(int)threadIdx.x, AUX_NT, 0};
  ekv_step st = a.st;
  st.accumulate = 0;
  auto none = [](int, float& ds, float& dsq) { ds = 0.f; dsq = 0.f; };
  state_select_apply(st, u, a.n_before, a.n_phys, 0, false, none, sc, grp);
}

int launch_select(const KernelArgs& a, cudaStream_t stream) {
  const AuxSmem L(a.n_phys, a.st.evict);
  if (L.total > 227 * 1024) return set_error(EKV_ERR_UNSUPPORTED, "select: %d bytes of shared memory needed", L.total);
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  cudaError_t err;
  if (dev < 16 && !configured[dev]) {
    err = cudaFuncSetAttribute(select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(select)", err);
    configured[dev] = 1;
  }
  select_kernel<<<a.B * a.Hkv, AUX_NT, L.total, stream>>>(a);
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("select_kernel launch", err);
  count_launch();
  return EKV_OK;
}

// ---- explicit eviction: truncate_kv_cache_silo/_liso/truncate_kv_cache (easykv.py:56-82,105-112) ---
__global__ void __launch_bounds__(AUX_NT) evict_explicit_kernel(const KernelArgs a, const int32_t* victims, int evict) {
  extern __shared__ __align__(128) unsigned char smem[];
  int32_t* vsorted = reinterpret_cast<int32_t*>(smem);          // [evict]
  const int unit = blockIdx.x, tid = threadIdx.x;
  const int32_t* vin = victims + (size_t)unit * evict;
  int32_t* lidx = a.lidx + (size_t)unit * a.cap;
  for (int t = tid; t < evict; t += AUX_NT) {
    const int l = vin[t];
    int rank = 0;
    for (int k = 0; k < evict; ++k) rank += vin[k] < l;
    vsorted[rank] = l;
  }
  __syncthreads();
  for (int e = tid; e < a.n_phys; e += AUX_NT) {
    const int l = lidx[e];
    if (l < 0) continue;
    int lo = 0, hi = evict;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (vsorted[mid] < l) lo = mid + 1; else hi = mid; }
    if (lo < evict && vsorted[lo] == l) {
      lidx[e] = -1;
      if (a.victim_slots) a.victim_slots[(size_t)unit * evict + lo] = e;
      if (a.victim_lidx) a.victim_lidx[(size_t)unit * evict + lo] = l;
    } else if (lo > 0) {
      lidx[e] = l - lo;
    }
  }
}

int launch_evict_explicit(const KernelArgs& a, const int32_t* victims, int evict, cudaStream_t stream) {
  if (evict <= 0) return EKV_OK;
  if (evict * 4 > 48 * 1024) return set_error(EKV_ERR_UNSUPPORTED, "evict_explicit: too many victims (%d)", evict);
  evict_explicit_kernel<<<a.B * a.Hkv, AUX_NT, evict * 4, stream>>>(a, victims, evict);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("evict_explicit_kernel launch", err);
  count_launch();
  return EKV_OK;
}

// ---- tova in 'encoding'/'ppl': S = mean over KV heads of the last row's folded probabilities ----------
// (easykv.py:454-457, :845-848).  Logical index l lives in a different physical slot in every head, so
// the per-head values are first laid out by logical index in `scratch` [B, Hkv, n] fp32.
__global__ void tova_gather_kernel(const KernelArgs a, int n) {
  const int unit = blockIdx.y;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_phys) return;
  const int l = a.lidx[(size_t)unit * a.cap + e];
  if (l >= 0 && l < n) reinterpret_cast<float*>(a.scratch)[(size_t)unit * n + l] = a.S[(size_t)unit * a.cap + e];
}
template <typename T> __global__ void tova_mean_scatter_kernel(const KernelArgs a, int n) {
  const int unit = blockIdx.y, b = unit / a.Hkv;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_phys) return;
  const int l = a.lidx[(size_t)unit * a.cap + e];
  if (l < 0 || l >= n) return;
  const float* tmp = reinterpret_cast<const float*>(a.scratch) + (size_t)b * a.Hkv * n + l;
  float s = 0.f;
  for (int hh = 0; hh < a.Hkv; ++hh) s += tmp[(size_t)hh * n];
  a.S[(size_t)unit * a.cap + e] = Tr<T>::round_f(__fmul_rn(s, 1.0f / (float)a.Hkv));   // fp16 mean over heads
}

int launch_tova_head_mean(const KernelArgs& a, cudaStream_t stream) {
  const int n = a.n_before;     // valid slots (the caller passes the post-append count)
  dim3 grid((a.n_phys + 255) / 256, a.B * a.Hkv);
  tova_gather_kernel<<<grid, 256, 0, stream>>>(a, n);
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("tova_gather_kernel launch", err);
  count_launch();
  switch (a.dtype) {
    case EKV_F16: tova_mean_scatter_kernel<__half><<<grid, 256, 0, stream>>>(a, n); break;
    case EKV_BF16: tova_mean_scatter_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, n); break;
    default: tova_mean_scatter_kernel<float><<<grid, 256, 0, stream>>>(a, n); break;
  }
  err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("tova_mean_scatter_kernel launch", err);
  count_launch();
  return EKV_OK;
}

// ---- export of the arrival-ordered view -----------------------------------------------------------------
template <typename T>
__global__ void export_kernel(const KernelArgs a, T* K_out, T* V_out, float* S_out, float* SQ_out, float* C_out) {
  const int unit = blockIdx.y, n = a.n_before, D = a.d;
  const int rows_per_block = blockDim.x / 16;                 // 16 threads per row
  const int e = blockIdx.x * rows_per_block + threadIdx.x / 16;
  if (e >= a.n_phys) return;
  const int l = a.lidx[(size_t)unit * a.cap + e];
  if (l < 0 || l >= n) return;
  const int t16 = threadIdx.x % 16;
  const T* ks = reinterpret_cast<const T*>(a.K) + ((size_t)unit * a.cap + e) * D;
  const T* vs = reinterpret_cast<const T*>(a.V) + ((size_t)unit * a.cap + e) * D;
  T* kd = K_out + ((size_t)unit * n + l) * D;
  T* vd = V_out + ((size_t)unit * n + l) * D;
  for (int c = t16; c < D; c += 16) { kd[c] = ks[c]; vd[c] = vs[c]; }
  if (t16 == 0) {
    if (S_out) S_out[(size_t)unit * n + l] = a.S[(size_t)unit * a.cap + e];
    if (SQ_out) SQ_out[(size_t)unit * n + l] = a.SQ[(size_t)unit * a.cap + e];
    if (C_out) C_out[(size_t)unit * n + l] = a.C[(size_t)unit * a.cap + e];
  }
}

int launch_export(const KernelArgs& a, void* K_out, void* V_out, float* S_out, float* SQ_out, float* C_out,
                  cudaStream_t stream) {
  if (a.n_phys <= 0) return EKV_OK;
  dim3 grid((a.n_phys + 15) / 16, a.B * a.Hkv);
  switch (a.dtype) {
    case EKV_F16:
      export_kernel<__half><<<grid, 256, 0, stream>>>(a, (__half*)K_out, (__half*)V_out, S_out, SQ_out, C_out); break;
    case EKV_BF16:
      export_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(a, (__nv_bfloat16*)K_out, (__nv_bfloat16*)V_out, S_out, SQ_out, C_out); break;
    case EKV_F32:
      export_kernel<float><<<grid, 256, 0, stream>>>(a, (float*)K_out, (float*)V_out, S_out, SQ_out, C_out); break;
    default: return set_error(EKV_ERR_INVALID, "dtype %d", a.dtype);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("export_kernel launch", err);
  count_launch();
  return EKV_OK;
}

}  // namespace ekv

// ---- RoPE of q and the new k at explicit positions + head-major re-layout ----------------------------
// Replaces apply_rotary_pos_emb (easykv/llama_patch.py:47-72, mistral_patch.py:62-87) and the transpose /
// contiguous copies around it (llama_patch.py:169-171): the projections' [B, q_len, heads, d] outputs are rotated
// and written as the [B, heads, q_len, d] tensors ekv_attend_evict takes; v is only re-laid out.  Arithmetic as the
// reference evaluates it in the model dtype: x*cos and rotate_half(x)*sin are each rounded, then their sum.
namespace ekv {

template <typename T>
__global__ void rope_qk_kernel(const T* __restrict__ q_in, const T* __restrict__ k_in, const T* __restrict__ v_in,
                               const T* __restrict__ cos_t, const T* __restrict__ sin_t, const int32_t* __restrict__ positions,
                               T* __restrict__ q_out, T* __restrict__ k_out, T* __restrict__ v_out,
                               int B, int H, int Hkv, int QL, int d) {
  const int heads = H + 2 * Hkv;                       // q heads, then k heads, then v heads
  const int half = d >> 1;
  const long long total = (long long)B * QL * heads * half;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % half);
    const int hh = (int)((idx / half) % heads);
    const int i = (int)((idx / ((long long)half * heads)) % QL);
    const int b = (int)(idx / ((long long)half * heads * QL));
    const int row = positions ? positions[b * QL + i] : b * QL + i;
    const T* src; T* dst; int nh, h; bool rot = true;
    if (hh < H) { src = q_in; dst = q_out; nh = H; h = hh; }
    else if (hh < H + Hkv) { src = k_in; dst = k_out; nh = Hkv; h = hh - H; }
    else { src = v_in; dst = v_out; nh = Hkv; h = hh - H - Hkv; rot = false; }
    if (!src || !dst) continue;
    const T* x = src + (((size_t)b * QL + i) * nh + h) * d;
    T* y = dst + (((size_t)b * nh + h) * QL + i) * d;
    const float x1 = Tr<T>::to_f(x[j]), x2 = Tr<T>::to_f(x[j + half]);
    if (!rot) { y[j] = x[j]; y[j + half] = x[j + half]; continue; }
    const float c1 = Tr<T>::to_f(cos_t[(size_t)row * d + j]), c2 = Tr<T>::to_f(cos_t[(size_t)row * d + j + half]);
    const float s1 = Tr<T>::to_f(sin_t[(size_t)row * d + j]), s2 = Tr<T>::to_f(sin_t[(size_t)row * d + j + half]);
    // rotate_half(x) = cat(-x2, x1)   (llama_patch.py:13-17)
    const float a1 = Tr<T>::round_f(__fmul_rn(x1, c1)), b1 = Tr<T>::round_f(__fmul_rn(-x2, s1));
    const float a2 = Tr<T>::round_f(__fmul_rn(x2, c2)), b2 = Tr<T>::round_f(__fmul_rn(x1, s2));
    y[j] = Tr<T>::from_f(__fadd_rn(a1, b1));
    y[j + half] = Tr<T>::from_f(__fadd_rn(a2, b2));
  }
}

int launch_rope_qk(int dtype, const void* q_in, const void* k_in, const void* v_in, const void* cos_t, const void* sin_t,
                   const int32_t* positions, void* q_out, void* k_out, void* v_out, int B, int H, int Hkv, int QL, int d,
                   cudaStream_t stream) {
  const long long total = (long long)B * QL * (H + 2 * Hkv) * (d / 2);
  if (total <= 0) return EKV_OK;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  if (blocks > 148 * 16) blocks = 148 * 16;
  switch (dtype) {
    case EKV_F16:
      rope_qk_kernel<__half><<<(int)blocks, threads, 0, stream>>>((const __half*)q_in, (const __half*)k_in, (const __half*)v_in,
          (const __half*)cos_t, (const __half*)sin_t, positions, (__half*)q_out, (__half*)k_out, (__half*)v_out, B, H, Hkv, QL, d);
      break;
    case EKV_BF16:
      rope_qk_kernel<__nv_bfloat16><<<(int)blocks, threads, 0, stream>>>((const __nv_bfloat16*)q_in, (const __nv_bfloat16*)k_in,
          (const __nv_bfloat16*)v_in, (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, positions, (__nv_bfloat16*)q_out,
          (__nv_bfloat16*)k_out, (__nv_bfloat16*)v_out, B, H, Hkv, QL, d);
      break;
    case EKV_F32:
      rope_qk_kernel<float><<<(int)blocks, threads, 0, stream>>>((const float*)q_in, (const float*)k_in, (const float*)v_in,
          (const float*)cos_t, (const float*)sin_t, positions, (float*)q_out, (float*)k_out, (float*)v_out, B, H, Hkv, QL, d);
      break;
    default: return set_error(EKV_ERR_INVALID, "dtype %d", dtype);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("rope_qk_kernel launch", err);
  count_launch();
  return EKV_OK;
}

}  // namespace ekv

// ---- streaming variant: re-rotate the whole cache at cache-relative positions --------------------
// llama_forward_stream / mistral_forward_stream (easykv/llama_patch.py:251-379, mistral_patch.py:189-286) keep
// UN-rotated keys in the cache and apply RoPE to all of them at positions 0..kv_len-1 — their index in the
// arrival-ordered cache, which shifts with every eviction — on every forward (:310-327).  Here the un-rotated rows
// live in a second buffer K_raw (same physical layout); this kernel writes K[slot] = rope(K_raw[slot], lidx[slot])
// for every valid slot, after which the attention kernels run unchanged.  Same arithmetic as rope_qk_kernel.
namespace ekv {

template <typename T>
__global__ void rope_cache_kernel(const T* __restrict__ K_raw, T* __restrict__ K, const int32_t* __restrict__ lidx,
                                  const T* __restrict__ cos_t, const T* __restrict__ sin_t, int units, int cap, int n_phys, int d) {
  const int half = d >> 1;
  const long long total = (long long)units * n_phys * half;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(idx % half);
    const int slot = (int)((idx / half) % n_phys);
    const int unit = (int)(idx / ((long long)half * n_phys));
    const int row = lidx[(size_t)unit * cap + slot];
    if (row < 0) continue;
    const T* x = K_raw + ((size_t)unit * cap + slot) * d;
    T* y = K + ((size_t)unit * cap + slot) * d;
    const float x1 = Tr<T>::to_f(x[j]), x2 = Tr<T>::to_f(x[j + half]);
    const float c1 = Tr<T>::to_f(cos_t[(size_t)row * d + j]), c2 = Tr<T>::to_f(cos_t[(size_t)row * d + j + half]);
    const float s1 = Tr<T>::to_f(sin_t[(size_t)row * d + j]), s2 = Tr<T>::to_f(sin_t[(size_t)row * d + j + half]);
    const float a1 = Tr<T>::round_f(__fmul_rn(x1, c1)), b1 = Tr<T>::round_f(__fmul_rn(-x2, s1));
    const float a2 = Tr<T>::round_f(__fmul_rn(x2, c2)), b2 = Tr<T>::round_f(__fmul_rn(x1, s2));
    y[j] = Tr<T>::from_f(__fadd_rn(a1, b1));
    y[j + half] = Tr<T>::from_f(__fadd_rn(a2, b2));
  }
}

// 16-bit dtypes: one thread per (slot, 8 dims of the lower half + the matching 8 of the upper half), 128-bit loads
// and stores — the pass is pure streaming (read K_raw, write K) and the table rows come from L2.
template <typename T>
__global__ void rope_cache_vec_kernel(const T* __restrict__ K_raw, T* __restrict__ K, const int32_t* __restrict__ lidx,
                                      const T* __restrict__ cos_t, const T* __restrict__ sin_t, int units, int cap, int n_phys) {
  constexpr int D = 128, CH = D / 2 / 8;             // 8 chunks of 8 dims per half row
  const long long total = (long long)units * n_phys * CH;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(idx % CH);
    const int slot = (int)((idx / CH) % n_phys);
    const int unit = (int)(idx / ((long long)CH * n_phys));
    const int row = lidx[(size_t)unit * cap + slot];
    if (row < 0) continue;
    const T* x = K_raw + ((size_t)unit * cap + slot) * D + c * 8;
    T* y = K + ((size_t)unit * cap + slot) * D + c * 8;
    const T* cr = cos_t + (size_t)row * D + c * 8;
    const T* sr = sin_t + (size_t)row * D + c * 8;
    const uint4 xl = *reinterpret_cast<const uint4*>(x), xh = *reinterpret_cast<const uint4*>(x + D / 2);
    const uint4 cl = *reinterpret_cast<const uint4*>(cr), ch = *reinterpret_cast<const uint4*>(cr + D / 2);
    const uint4 sl = *reinterpret_cast<const uint4*>(sr), sh = *reinterpret_cast<const uint4*>(sr + D / 2);
    const T* xlp = reinterpret_cast<const T*>(&xl); const T* xhp = reinterpret_cast<const T*>(&xh);
    const T* clp = reinterpret_cast<const T*>(&cl); const T* chp = reinterpret_cast<const T*>(&ch);
    const T* slp = reinterpret_cast<const T*>(&sl); const T* shp = reinterpret_cast<const T*>(&sh);
    uint4 ol, oh;
    T* olp = reinterpret_cast<T*>(&ol); T* ohp = reinterpret_cast<T*>(&oh);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float x1 = Tr<T>::to_f(xlp[i]), x2 = Tr<T>::to_f(xhp[i]);
      const float a1 = Tr<T>::round_f(__fmul_rn(x1, Tr<T>::to_f(clp[i]))), b1 = Tr<T>::round_f(__fmul_rn(-x2, Tr<T>::to_f(slp[i])));
      const float a2 = Tr<T>::round_f(__fmul_rn(x2, Tr<T>::to_f(chp[i]))), b2 = Tr<T>::round_f(__fmul_rn(x1, Tr<T>::to_f(shp[i])));
      olp[i] = Tr<T>::from_f(__fadd_rn(a1, b1));
      ohp[i] = Tr<T>::from_f(__fadd_rn(a2, b2));
    }
    *reinterpret_cast<uint4*>(y) = ol;
    *reinterpret_cast<uint4*>(y + D / 2) = oh;
  }
}

int launch_rope_cache(int dtype, const void* K_raw, void* K, const int32_t* lidx, const void* cos_t, const void* sin_t,
                      int units, int cap, int n_phys, int d, cudaStream_t stream) {
  const long long total = (long long)units * n_phys * (d / 2);
  if (total <= 0) return EKV_OK;
  long long blocks = (total + 255) / 256;
  if (blocks > 148 * 32) blocks = 148 * 32;
  const bool vec = d == 128 && dtype != EKV_F32;
  long long vblocks = ((long long)units * n_phys * 8 + 255) / 256;
  if (vblocks > 148 * 32) vblocks = 148 * 32;
  switch (dtype) {
    case EKV_F16:
      if (vec) rope_cache_vec_kernel<__half><<<(int)vblocks, 256, 0, stream>>>((const __half*)K_raw, (__half*)K, lidx, (const __half*)cos_t,
                                                                              (const __half*)sin_t, units, cap, n_phys);
      else rope_cache_kernel<__half><<<(int)blocks, 256, 0, stream>>>((const __half*)K_raw, (__half*)K, lidx, (const __half*)cos_t,
                                                                     (const __half*)sin_t, units, cap, n_phys, d);
      break;
    case EKV_BF16:
      if (vec) rope_cache_vec_kernel<__nv_bfloat16><<<(int)vblocks, 256, 0, stream>>>((const __nv_bfloat16*)K_raw, (__nv_bfloat16*)K, lidx,
          (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, units, cap, n_phys);
      else rope_cache_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, stream>>>((const __nv_bfloat16*)K_raw, (__nv_bfloat16*)K, lidx,
          (const __nv_bfloat16*)cos_t, (const __nv_bfloat16*)sin_t, units, cap, n_phys, d);
      break;
    case EKV_F32:
      rope_cache_kernel<float><<<(int)blocks, 256, 0, stream>>>((const float*)K_raw, (float*)K, lidx, (const float*)cos_t,
                                                               (const float*)sin_t, units, cap, n_phys, d); break;
    default: return set_error(EKV_ERR_INVALID, "dtype %d", dtype);
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) return set_cuda_error("rope_cache_kernel launch", err);
  count_launch();
  return EKV_OK;
}

}  // namespace ekv
