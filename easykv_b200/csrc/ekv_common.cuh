// Device-side building blocks shared by the easykv_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <math.h>

#include "../../include/easykv_b200.h"

namespace ekv {

// ---------------------------------------------------------------------------------------
// element traits: every "rounding point" of the reference goes through these
// ---------------------------------------------------------------------------------------
template <typename T> struct Tr;
template <> struct Tr<__half> {
  static constexpr int kDtype = EKV_F16;
  __device__ __forceinline__ static float to_f(__half x) { return __half2float(x); }
  __device__ __forceinline__ static __half from_f(float x) { return __float2half_rn(x); }
  __device__ __forceinline__ static float round_f(float x) { return __half2float(__float2half_rn(x)); }
  __device__ __forceinline__ static float2 to_f2(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};
template <> struct Tr<__nv_bfloat16> {
  static constexpr int kDtype = EKV_BF16;
  __device__ __forceinline__ static float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
  __device__ __forceinline__ static __nv_bfloat16 from_f(float x) { return __float2bfloat16_rn(x); }
  __device__ __forceinline__ static float round_f(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
  __device__ __forceinline__ static float2 to_f2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
  }
};
template <> struct Tr<float> {
  static constexpr int kDtype = EKV_F32;
  __device__ __forceinline__ static float to_f(float x) { return x; }
  __device__ __forceinline__ static float from_f(float x) { return x; }
  __device__ __forceinline__ static float round_f(float x) { return x; }
};

// ---------------------------------------------------------------------------------------
// mbarrier + TMA bulk copy (cp.async.bulk -> SASS UBLKCP) — hand-written PTX
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy, completion signalled on `bar` (bytes must be a multiple of 16,
// both addresses 16-byte aligned).  L2 evict-first: the retained cache is streamed once per step.
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                             uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
      ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization): the next kernel of
// the stream may be scheduled while this one still runs (pdl_trigger), and a kernel launched that way must not touch
// global memory before the previous kernel has completed and flushed (pdl_wait).  Every decode kernel waits before its
// first global access, so the stream's semantics are exactly the serial ones; what overlaps is launch latency, CTA
// placement and the shared-memory / tensor-memory set-up — several microseconds per layer of a back-to-back decode step.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// pull `bytes` (a multiple of 16, from a 16-byte aligned address) into L2 without waiting for them
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Ampere-style per-thread async copies (SASS LDGSTS) for the small per-unit headers
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// ---------------------------------------------------------------------------------------
// thread-block cluster primitives (DSMEM)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* smem_ptr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(smem_ptr)), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void st_cluster_u8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared::cluster.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_cluster_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// ---------------------------------------------------------------------------------------
// ordered keys: ascending uint32 order == (value ascending, -0 == +0, NaN last)   SURVEY A.5
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t order_key(float x) {
  if (x != x) return 0xffffffffu;
  if (x == 0.f) return 0x80000000u;
  uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// a subset of the CTA's warps synchronising on a named barrier (the TMA producer warp of the
// decode kernel never joins the consumers' barriers)
struct Grp {
  int tid, n, bar;
  __device__ __forceinline__ void sync() const {
    asm volatile("bar.sync %0, %1;" ::"r"(bar), "r"(n) : "memory");
  }
};

// 8 consecutive-dimension elements owned by lane `l16` of a 16-lane group reading one row of
// D=128 elements: 16-bit types own dims [8*l16, 8*l16+8) (one 128-bit access); fp32 owns
// [4*l16, 4*l16+4) and [64+4*l16, 64+4*l16+4) (two 128-bit accesses, each conflict-free).
template <typename T> __device__ __forceinline__ int dim_of(int l16, int i) {
  if (sizeof(T) == 2) return 8 * l16 + i;
  return (i < 4) ? 4 * l16 + i : 64 + 4 * l16 + (i - 4);
}
template <typename T> __device__ __forceinline__ void load_row8(const T* row, int l16, float (&x)[8]);
template <> __device__ __forceinline__ void load_row8<__half>(const __half* row, int l16, float (&x)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(row + 8 * l16);
  float2 a = Tr<__half>::to_f2(u.x), b = Tr<__half>::to_f2(u.y), c = Tr<__half>::to_f2(u.z), d = Tr<__half>::to_f2(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
template <> __device__ __forceinline__ void load_row8<__nv_bfloat16>(const __nv_bfloat16* row, int l16, float (&x)[8]) {
  uint4 u = *reinterpret_cast<const uint4*>(row + 8 * l16);
  float2 a = Tr<__nv_bfloat16>::to_f2(u.x), b = Tr<__nv_bfloat16>::to_f2(u.y), c = Tr<__nv_bfloat16>::to_f2(u.z),
         d = Tr<__nv_bfloat16>::to_f2(u.w);
  x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y; x[4] = c.x; x[5] = c.y; x[6] = d.x; x[7] = d.y;
}
template <> __device__ __forceinline__ void load_row8<float>(const float* row, int l16, float (&x)[8]) {
  float4 a = *reinterpret_cast<const float4*>(row + 4 * l16);
  float4 b = *reinterpret_cast<const float4*>(row + 64 + 4 * l16);
  x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <typename T> __device__ __forceinline__ void store_row8(T* row, int l16, const float (&x)[8]);
template <> __device__ __forceinline__ void store_row8<__half>(__half* row, int l16, const float (&x)[8]) {
  __half2 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
  *reinterpret_cast<uint4*>(row + 8 * l16) = *reinterpret_cast<uint4*>(h);
}
template <> __device__ __forceinline__ void store_row8<__nv_bfloat16>(__nv_bfloat16* row, int l16, const float (&x)[8]) {
  __nv_bfloat162 h[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(x[2 * i], x[2 * i + 1]);
  *reinterpret_cast<uint4*>(row + 8 * l16) = *reinterpret_cast<uint4*>(h);
}
template <> __device__ __forceinline__ void store_row8<float>(float* row, int l16, const float (&x)[8]) {
  *reinterpret_cast<float4*>(row + 4 * l16) = make_float4(x[0], x[1], x[2], x[3]);
  *reinterpret_cast<float4*>(row + 64 + 4 * l16) = make_float4(x[4], x[5], x[6], x[7]);
}

// ---------------------------------------------------------------------------------------
// Row8<T>: the 8 elements of a D=128 row owned by one lane of a 16-lane group, kept in the
// storage format.  16-bit types are never converted: the products go through the sm_100
// mixed-precision FMA (PTX fma.rn.f32.f16 / .bf16 -> SASS FHFMA with .H0/.H1 operand selectors),
// which multiplies two 16-bit values exactly and adds into fp32 with one rounding — bit-identical
// to fmaf(float(a), float(b), c) at a third of the instructions.
// ---------------------------------------------------------------------------------------
template <typename T> struct Row8;
template <> struct Row8<float> {
  float x[8];
  __device__ __forceinline__ void load(const float* row, int l16) { load_row8<float>(row, l16, x); }
};
template <> struct Row8<__half> {
  uint4 u;
  __device__ __forceinline__ void load(const __half* row, int l16) { u = *reinterpret_cast<const uint4*>(row + 8 * l16); }
};
template <> struct Row8<__nv_bfloat16> {
  uint4 u;
  __device__ __forceinline__ void load(const __nv_bfloat16* row, int l16) { u = *reinterpret_cast<const uint4*>(row + 8 * l16); }
};
__device__ __forceinline__ float fma2_f16(uint32_t a, uint32_t b, float acc) {      // acc += a.lo*b.lo; acc += a.hi*b.hi
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.f16 %0, al, bl, %0;\n\tfma.rn.f32.f16 %0, ah, bh, %0;\n\t}" : "+f"(acc) : "r"(a), "r"(b));
  return acc;
}
__device__ __forceinline__ float fma2_bf16(uint32_t a, uint32_t b, float acc) {
  asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
      "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %0, ah, bh, %0;\n\t}" : "+f"(acc) : "r"(a), "r"(b));
  return acc;
}
__device__ __forceinline__ void axpy2_f16(uint16_t p, uint32_t v, float& o0, float& o1) {   // o0 += p*v.lo; o1 += p*v.hi
  asm("{\n\t.reg .b16 vl, vh;\n\tmov.b32 {vl, vh}, %3;\n\t"
      "fma.rn.f32.f16 %0, %2, vl, %0;\n\tfma.rn.f32.f16 %1, %2, vh, %1;\n\t}" : "+f"(o0), "+f"(o1) : "h"(p), "r"(v));
}
__device__ __forceinline__ void axpy2_bf16(uint16_t p, uint32_t v, float& o0, float& o1) {
  asm("{\n\t.reg .b16 vl, vh;\n\tmov.b32 {vl, vh}, %3;\n\t"
      "fma.rn.f32.bf16 %0, %2, vl, %0;\n\tfma.rn.f32.bf16 %1, %2, vh, %1;\n\t}" : "+f"(o0), "+f"(o1) : "h"(p), "r"(v));
}
// dot of two lanes' 8 elements, accumulated sequentially j = 0..7 into acc
__device__ __forceinline__ float dot8(const Row8<float>& a, const Row8<float>& b, float acc) {
#pragma unroll
  for (int j = 0; j < 8; ++j) acc = fmaf(a.x[j], b.x[j], acc);
  return acc;
}
__device__ __forceinline__ float dot8(const Row8<__half>& a, const Row8<__half>& b, float acc) {
  acc = fma2_f16(a.u.x, b.u.x, acc); acc = fma2_f16(a.u.y, b.u.y, acc);
  acc = fma2_f16(a.u.z, b.u.z, acc); acc = fma2_f16(a.u.w, b.u.w, acc);
  return acc;
}
__device__ __forceinline__ float dot8(const Row8<__nv_bfloat16>& a, const Row8<__nv_bfloat16>& b, float acc) {
  acc = fma2_bf16(a.u.x, b.u.x, acc); acc = fma2_bf16(a.u.y, b.u.y, acc);
  acc = fma2_bf16(a.u.z, b.u.z, acc); acc = fma2_bf16(a.u.w, b.u.w, acc);
  return acc;
}
// o[j] += p * v[j] with p a model-dtype probability
__device__ __forceinline__ void axpy8(float p, const Row8<float>& v, float (&o)[8]) {
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = fmaf(p, v.x[j], o[j]);
}
__device__ __forceinline__ void axpy8(__half p, const Row8<__half>& v, float (&o)[8]) {
  const uint16_t pb = __half_as_ushort(p);
  axpy2_f16(pb, v.u.x, o[0], o[1]); axpy2_f16(pb, v.u.y, o[2], o[3]);
  axpy2_f16(pb, v.u.z, o[4], o[5]); axpy2_f16(pb, v.u.w, o[6], o[7]);
}
__device__ __forceinline__ void axpy8(__nv_bfloat16 p, const Row8<__nv_bfloat16>& v, float (&o)[8]) {
  const uint16_t pb = __bfloat16_as_ushort(p);
  axpy2_bf16(pb, v.u.x, o[0], o[1]); axpy2_bf16(pb, v.u.y, o[2], o[3]);
  axpy2_bf16(pb, v.u.z, o[4], o[5]); axpy2_bf16(pb, v.u.w, o[6], o[7]);
}

// ---------------------------------------------------------------------------------------
// RoPE of one cached row while it is read (fused streaming variant, llama_patch.py:310-327): the row is spread over
// the 16 lanes of a half-warp, 8 consecutive dims per lane (Row8), so lane l16 ^ 8 holds the partner half
// (rotate_half, :47-55).  Packed model-dtype arithmetic = apply_rotary_pos_emb's roundings: rn(x*cos), rn(rot*sin),
// rn(sum) (:70) — bit-identical to rope_cache_kernel, whose fp32 product / sum of two 11-bit (8-bit) operands rounds once.
// NOT inlined on purpose: inlined into the cluster decode kernel's unrolled K loop, the g = 2 and g = 8 instantiations
// (nvcc 12.9, sm_100a, -O3) fetched table rows through garbage addresses (compute-sanitizer: invalid 16-byte global
// reads in this function; g = 1 / 4 and the persistent kernel were fine, and so was any build with one more live
// value in the loop) although the SASS address arithmetic reads correctly; as a call every instantiation is
// bit-identical to the two-pass path (tests/test_gpu_stream_ragged.py covers g = 1, 2, 4, 8 on both kernels).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mul2_rn(uint32_t a, uint32_t b, __half) { uint32_t r; asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t add2_rn(uint32_t a, uint32_t b, __half) { uint32_t r; asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t mul2_rn(uint32_t a, uint32_t b, __nv_bfloat16) { uint32_t r; asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ uint32_t add2_rn(uint32_t a, uint32_t b, __nv_bfloat16) { uint32_t r; asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
template <typename T> __device__ __noinline__ void rope_row8(Row8<T>& x, int l16, const void* cos_t, const void* sin_t, int pos) {
  // table row `pos` = 16 uint4 (d == 128, 16-bit dtype)
  const uint4 c = *(reinterpret_cast<const uint4*>(cos_t) + (size_t)pos * 16 + l16);
  const uint4 sn = *(reinterpret_cast<const uint4*>(sin_t) + (size_t)pos * 16 + l16);
  const uint32_t sg = l16 < 8 ? 0x80008000u : 0u;              // rotate_half: -x2 under the lower half, +x1 under the upper
  uint4 o;
  o.x = __shfl_xor_sync(0xffffffffu, x.u.x, 8) ^ sg; o.y = __shfl_xor_sync(0xffffffffu, x.u.y, 8) ^ sg;
  o.z = __shfl_xor_sync(0xffffffffu, x.u.z, 8) ^ sg; o.w = __shfl_xor_sync(0xffffffffu, x.u.w, 8) ^ sg;
  const T t{};
  x.u.x = add2_rn(mul2_rn(x.u.x, c.x, t), mul2_rn(o.x, sn.x, t), t); x.u.y = add2_rn(mul2_rn(x.u.y, c.y, t), mul2_rn(o.y, sn.y, t), t);
  x.u.z = add2_rn(mul2_rn(x.u.z, c.z, t), mul2_rn(o.z, sn.z, t), t); x.u.w = add2_rn(mul2_rn(x.u.w, c.w, t), mul2_rn(o.w, sn.w, t), t);
}
template <> inline __device__ void rope_row8<float>(Row8<float>&, int, const void*, const void*, int) {}   // (fp32 is not fused)

template <typename T> __device__ __forceinline__ T neg_inf();
template <> __device__ __forceinline__ __half neg_inf<__half>() { return __ushort_as_half(0xfc00); }
template <> __device__ __forceinline__ __nv_bfloat16 neg_inf<__nv_bfloat16>() { return __ushort_as_bfloat16(0xff80); }
template <> __device__ __forceinline__ float neg_inf<float>() { return __uint_as_float(0xff800000u); }

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ekv
