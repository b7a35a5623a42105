// Warp-level tensor-core primitives shared by the chunk kernels (ekv_chunk_tc.cu) and the tensor-core variant
// of the cluster decode kernel (ekv_decode_cluster.cu): ldmatrix fragment loads, mma.sync m16n8k16 with fp32
// accumulation, packed model-dtype rounding and an exact fp32 division by a value whose reciprocal is known.
#pragma once
#include "ekv_common.cuh"

namespace ekv {

// ---- tensor-core primitives ---------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
template <typename T> __device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <> __device__ __forceinline__ void mma16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <> __device__ __forceinline__ void mma16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <typename T> __device__ __forceinline__ uint32_t pack2(float lo, float hi);      // both already model-dtype values
template <> __device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

// round a pair to the model dtype and back with the packed conversions (one F2FP per pair instead of
// two scalar F2F per element)
template <typename T> __device__ __forceinline__ void round2(float& x, float& y);
template <> __device__ __forceinline__ void round2<__half>(float& x, float& y) {
  const float2 f = __half22float2(__floats2half2_rn(x, y));
  x = f.x; y = f.y;
}
template <> __device__ __forceinline__ void round2<__nv_bfloat16>(float& x, float& y) {
  const uint32_t u = pack2<__nv_bfloat16>(x, y);
  x = __uint_as_float(u << 16); y = __uint_as_float(u & 0xffff0000u);
}
// a / b correctly rounded for normal operands, given rb = RN(1/b) (Markstein: q = a*rb; r = a - b*q exactly
// by FMA; q + r*rb).  Here 0 <= a <= 1 <= b, so neither overflow nor a subnormal quotient that would
// survive the rounding to the model dtype can occur.
__device__ __forceinline__ float div_rn_by(float a, float b, float rb) {
  const float q = __fmul_rn(a, rb);
  const float r = __fmaf_rn(-b, q, a);
  return __fmaf_rn(r, rb, q);
}

}  // namespace ekv
