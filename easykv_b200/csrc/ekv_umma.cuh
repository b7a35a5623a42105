// Blackwell (sm_100a) tensor-core building blocks, hand-written PTX: tcgen05.mma with shared-memory operand
// descriptors, tensor memory (TMEM) allocation / loads / stores, tcgen05.commit onto mbarriers, and tensor-map TMA
// (cp.async.bulk.tensor) tile loads with the 128-byte swizzle both engines agree on.  Used by the strided-prefill
// chunk kernel (ekv_chunk_umma.cu) and pinned by a primitive-level probe (ekv_umma_probe.cu, tests/test_gpu_umma.py).
//
// Operand layouts (all 16-bit element types, 128-byte swizzle, buffers 1024-byte aligned):
//   K-major  [rows][64 elements]: row r is 128 contiguous bytes at r*128; inside each 8-row / 1024-byte atom the
//            16-byte chunk c of row r sits at chunk position c ^ (r & 7).  One MMA k-step (16 elements) = 32 bytes
//            along the row: advance the descriptor start address by 32 bytes; SBO = 1024 bytes between 8-row groups.
//   MN-major [k][64 elements along M or N]: the same memory picture, but the 128-byte row is indexed by the
//            contraction index k and holds 64 consecutive M (or N) elements.  SBO = 1024 bytes between groups of 8 k,
//            LBO = distance between successive 64-element groups along M / N.  One MMA k-step (16 k) = 2048 bytes.
// A TMA box {64 elements, rows} with CU_TENSOR_MAP_SWIZZLE_128B produces exactly this picture.
#pragma once
#include <cuda.h>

#include "ekv_common.cuh"

namespace ekv {
namespace umma {

// ---- shared-memory matrix descriptor (PTX ISA "tcgen05 shared memory descriptor") ------------------------------------
// bits [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
// [46,48) version = 1 | [61,64) layout type: 2 = 128-byte swizzle
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// the same without a swizzle (layout type 0).  MN-major operand of 8-element (16-byte) groups: a core matrix is 8 k-rows
// x 16 bytes, 128 contiguous bytes; SBO = distance between successive 8-element groups along M / N, LBO = distance
// between successive groups of 8 k.  (Used for the 16-row P^T operand of the GQA decode kernel.)
__device__ __forceinline__ uint64_t smem_desc_noswizzle(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fffu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// ---- instruction descriptor, kind::f16, fp32 accumulate ---------------------------------------------------------------
// [4,6) D format 1 = f32 | [7,10) A format, [10,13) B format: 0 = f16, 1 = bf16 | bit 15 / 16: A / B is MN-major |
// [17,23) N >> 3 | [24,29) M >> 4
template <typename T> __host__ __device__ constexpr uint32_t fmt_of();
template <> __host__ __device__ constexpr uint32_t fmt_of<__half>() { return 0u; }
template <> __host__ __device__ constexpr uint32_t fmt_of<__nv_bfloat16>() { return 1u; }
template <typename T> __host__ __device__ constexpr uint32_t instr_desc(int M, int N, bool a_mn, bool b_mn) {
  return (1u << 4) | (fmt_of<T>() << 7) | (fmt_of<T>() << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread on behalf of the CTA
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// all MMAs issued so far by this thread arrive on `bar` when they have completed (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (tensor core / TMA reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ---------------------------------------------------------------------------------------------------------
// one warp allocates `cols` (power of two >= 32) columns and writes the base address to *dst (shared memory)
__device__ __forceinline__ void tmem_alloc(uint32_t* dst, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// TMEM address: lane in bits [16,32), column in bits [0,16).  A warp may only touch lanes 32*(warp%4) .. +31;
// with the 32x32b shape thread t of the warp owns lane 32*(warp%4)+t and receives N consecutive columns.
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,"
      "%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}

// ---- tensor-map TMA ----------------------------------------------------------------------------------------------------------
// 2-D tile load global -> shared (this CTA), completion on `bar` (complete_tx::bytes); c0 = element column, c1 = row
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// byte offset of 16-byte chunk `c` (0..7) of 128-byte row `r` in a 128-byte-swizzled block of rows
__device__ __forceinline__ uint32_t swz128(int r, int c) { return (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4); }

// mbarrier helpers with cluster scope (remote arrives / acquire waits for DSMEM exchanges)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
}

}  // namespace umma

// ---- helpers shared by the tcgen05 kernels (chunk: ekv_chunk_umma.cu, GQA decode: ekv_decode_umma.cu) --------------------
template <typename T> __device__ __forceinline__ uint32_t neg_inf2();
template <> __device__ __forceinline__ uint32_t neg_inf2<__half>() { return 0xfc00fc00u; }
template <> __device__ __forceinline__ uint32_t neg_inf2<__nv_bfloat16>() { return 0xff80ff80u; }
template <typename T> __device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b);
template <> __device__ __forceinline__ uint32_t max2<__half>(uint32_t a, uint32_t b) {
  __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
template <> __device__ __forceinline__ uint32_t max2<__nv_bfloat16>(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}

template <int N> struct TmemIO;
template <> struct TmemIO<32> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[32]) { umma::tmem_ld32(a, r); }
};
template <> struct TmemIO<16> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[16]) { umma::tmem_ld16(a, r); }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[16]) { umma::tmem_st16(a, r); }
};
template <> struct TmemIO<8> {
  static __device__ __forceinline__ void ld(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
  }
  static __device__ __forceinline__ void st(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
  }
};

// a latency-tolerant waiter (the TMA / MMA threads): sleep between polls instead of burning the issue slots of the
// scheduler partition it shares with two softmax warps
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity, unsigned ns) {
  while (!mbar_try_wait(bar, parity)) __nanosleep(ns);
}
// 2^x, hardware approximation (MUFU.EX2, ~2^-22 relative): used only for the softmax DENOMINATORS, which are sums of
// hundreds to thousands of terms whose summation order alone moves them by as much; the numerators use expf
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}


// Host side: a 2-D tensor map over a row-major [rows][128] 16-bit matrix (row pitch 256 bytes), box = {64 columns,
// box_rows} with the 128-byte swizzle.  The driver entry point is resolved at run time (no link-time libcuda
// dependency: the library must load on machines without a driver).  Returns 0 on success.
int make_tensor_map_rows128(CUtensorMap* map, const void* base, unsigned long long rows, int box_rows, int dtype);

}  // namespace ekv
