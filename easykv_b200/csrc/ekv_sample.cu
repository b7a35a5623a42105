// Sampling / perplexity tail: one launch per decode step instead of the reference's
// softmax -> sort -> cumsum -> boolean-mask assignment (a host sync) -> sum -> div -> sort -> gather -> softmax chain
// (easykv/easykv.py:115-134) followed by torch.multinomial (two more host syncs, :258), and one launch per prompt chunk
// instead of keeping every chunk's [q_len, vocab] logits until the end for the cross entropy (:826-827, :896-901).
//
// Both kernels are one 1024-thread CTA per row of fp32 logits (the 4.36 model classes return fp32 logits).  A row is
// 128-600 KB: it is read from L2 after the first pass, the probabilities live in shared memory when they fit.
//
// top-p without a sort: the reference keeps, in descending order, every token whose EXCLUSIVE cumulative mass is
// <= top_p.  With M(t) = mass of the tokens whose probability bits are > t (non-increasing in t), the kept set is
// {p > t0} plus the first j ties at t0 (index order = stable sort order), where t0 is the smallest bit pattern with
// M(t0) <= top_p: found by bisection over the 30 significant bits, each probe one deterministic block reduction.
#include <math_constants.h>

#include "ekv_kernels.h"

namespace ekv {

constexpr int SA_NT = 1024;

__device__ __forceinline__ float sa_block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();                                    // red free again
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;                                           // bitwise identical in every thread
}

__device__ __forceinline__ float sa_block_max(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o; o >>= 1) r = fmaxf(r, __shfl_xor_sync(0xffffffffu, r, o));
  return r;
}

// softmax of x (optionally scaled) into dst; returns nothing, every thread ends with the row fully written + synced
template <int ARITH>
__device__ __forceinline__ void sa_softmax(const float* __restrict__ x, int V, float temperature, bool scaled, float* dst,
                                           float* red) {
  const float inv = 1.0f / temperature;               // ATen CUDA divides by a host scalar as a multiply by 1/b
  auto scale = [&](float v) { return !scaled ? v : (ARITH ? v * inv : __fdiv_rn(v, temperature)); };
  float m = -CUDART_INF_F;
  for (int i = threadIdx.x; i < V; i += SA_NT) m = fmaxf(m, scale(x[i]));
  m = sa_block_max(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += SA_NT) {
    const float e = expf(scale(x[i]) - m);
    dst[i] = e;
    s += e;
  }
  s = sa_block_sum(s, red);
  for (int i = threadIdx.x; i < V; i += SA_NT) dst[i] = __fdiv_rn(dst[i], s);   // own elements only: no sync needed
}

template <int ARITH>
__global__ void __launch_bounds__(SA_NT, 1)
logits_adapter_kernel(const float* __restrict__ logits, int V, float temperature, float top_p, const float* __restrict__ q_exp,
                      float* __restrict__ prob, float* __restrict__ raw, long long* __restrict__ token, int use_smem) {
  extern __shared__ float sa_row[];
  __shared__ float red[32];
  __shared__ int wcount[32];
  __shared__ int arg_i[32];
  const int row = blockIdx.x;
  const float* x = logits + (size_t)row * V;
  float* buf = use_smem ? sa_row : prob + (size_t)row * V;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (raw) sa_softmax<ARITH>(x, V, 1.f, false, raw + (size_t)row * V, red);     // the un-tempered softmax (:134)
  sa_softmax<ARITH>(x, V, temperature, true, buf, red);

  // ---- t0 = smallest bit pattern with M(t0) <= top_p --------------------------------------------------------------
  float pmax = 0.f;
  for (int i = threadIdx.x; i < V; i += SA_NT) pmax = fmaxf(pmax, buf[i]);
  pmax = sa_block_max(pmax, red);
  // (an 8-ary variant — seven probes per round sharing one reduction — measured slower: 120 vs 98 us per call; the
  //  rounds are bound by the per-element select-adds, not by the two barriers)
  unsigned lo = 0u, hi = __float_as_uint(pmax);
  float m_hi = 0.f;                                    // M(hi)
  while (lo < hi) {
    const unsigned mid = lo + ((hi - lo) >> 1);
    float s = 0.f;
    for (int i = threadIdx.x; i < V; i += SA_NT) {
      const float p = buf[i];
      s += __float_as_uint(p) > mid ? p : 0.f;
    }
    s = sa_block_sum(s, red);
    if (s <= top_p) { hi = mid; m_hi = s; } else lo = mid + 1;
  }
  const unsigned t0 = hi;
  const float v0 = __uint_as_float(t0);
  // ties at t0: the first `jkeep` of them (in index order) have exclusive mass M + j * v0 <= top_p
  float cnt = 0.f;
  for (int i = threadIdx.x; i < V; i += SA_NT) cnt += __float_as_uint(buf[i]) == t0 ? 1.f : 0.f;
  const int c = (int)sa_block_sum(cnt, red);
  int jkeep = c;
  if (v0 > 0.f && c > 1) {
    const float room = __fdiv_rn(top_p - m_hi, v0);
    jkeep = room >= (float)c ? c : max(1, (int)room + 1);
    jkeep = min(jkeep, c);
  }
  if (jkeep < c) {                                     // rare: the boundary cuts a run of equal probabilities
    int base = 0;
    for (int i0 = 0; i0 < V; i0 += SA_NT) {
      const int i = i0 + threadIdx.x;
      const bool tie = i < V && __float_as_uint(buf[i]) == t0;
      const unsigned bal = __ballot_sync(0xffffffffu, tie);
      __syncthreads();
      if (lane == 0) wcount[warp] = __popc(bal);
      __syncthreads();
      int before = base, total = 0;
      for (int w = 0; w < 32; ++w) {
        const int n = wcount[w];
        before += w < warp ? n : 0;
        total += n;
      }
      if (tie && before + __popc(bal & ((1u << lane) - 1u)) >= jkeep) buf[i] = 0.f;
      base += total;
    }
  }
  // ---- renormalise (:124) and, with q_exp, draw: argmax(final / q) is torch.multinomial's n_sample == 1 path ------
  float z = 0.f;
  for (int i = threadIdx.x; i < V; i += SA_NT) {
    const float p = buf[i];
    z += __float_as_uint(p) >= t0 ? p : 0.f;
  }
  z = sa_block_sum(z, red);
  float best = -1.f;
  int best_i = 0x7fffffff;
  float* out = prob ? prob + (size_t)row * V : nullptr;
  for (int i = threadIdx.x; i < V; i += SA_NT) {
    const float p = buf[i];
    const float f = __float_as_uint(p) >= t0 ? __fdiv_rn(p, z) : 0.f;
    if (out) out[i] = f;
    if (q_exp) {
      const float r = __fdiv_rn(f, q_exp[(size_t)row * V + i]);
      if (r > best) { best = r; best_i = i; }         // i ascending per thread: first maximum kept
    }
  }
  if (token) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
      if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = best; arg_i[warp] = best_i; }
    __syncthreads();
    if (warp == 0) {
      best = red[lane];
      best_i = arg_i[lane];
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, best_i, o);
        if (ob > best || (ob == best && oi < best_i)) { best = ob; best_i = oi; }
      }
      if (lane == 0) token[row] = best_i == 0x7fffffff ? 0 : best_i;
    }
  }
}

// nll[row] = -(x[t] - max - log(sum exp(x - max))): CrossEntropyLoss(reduction='none') on fp32 logits (:896-899)
__global__ void __launch_bounds__(SA_NT, 1)
token_nll_kernel(const float* __restrict__ logits, const long long* __restrict__ targets, int V, float* __restrict__ nll) {
  __shared__ float red[32];
  const int row = blockIdx.x;
  const float* x = logits + (size_t)row * V;
  float m = -CUDART_INF_F;
  for (int i = threadIdx.x; i < V; i += SA_NT) m = fmaxf(m, x[i]);
  m = sa_block_max(m, red);
  float s = 0.f;
  for (int i = threadIdx.x; i < V; i += SA_NT) s += expf(x[i] - m);
  s = sa_block_sum(s, red);
  if (threadIdx.x == 0) {
    const long long t = targets[row];
    nll[row] = (t < 0 || t >= V) ? CUDART_NAN_F : -(x[t] - m - logf(s));
  }
}

int launch_logits_adapter(const float* logits, int rows, int V, float temperature, float top_p, int arith, const float* q_exp,
                          float* prob, float* raw, long long* token, cudaStream_t stream) {
  const size_t row_bytes = (size_t)V * sizeof(float);
  int use_smem = row_bytes <= 200 * 1024;
  if (!use_smem && !prob) return set_error(EKV_ERR_UNSUPPORTED, "vocab %d needs the prob output as workspace", V);
  auto kern = arith ? logits_adapter_kernel<1> : logits_adapter_kernel<0>;
  if (use_smem) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)row_bytes);
    if (e != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(logits_adapter)", e);
  }
  kern<<<rows, SA_NT, use_smem ? row_bytes : 0, stream>>>(logits, V, temperature, top_p, q_exp, prob, raw, token, use_smem);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EKV_OK : set_cuda_error("logits_adapter_kernel", e);
}

int launch_token_nll(const float* logits, const long long* targets, int rows, int V, float* nll, cudaStream_t stream) {
  token_nll_kernel<<<rows, SA_NT, 0, stream>>>(logits, targets, V, nll);
  count_launch();
  cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? EKV_OK : set_cuda_error("token_nll_kernel", e);
}

}  // namespace ekv
