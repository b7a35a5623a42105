// Primitive-level probe of the Blackwell tensor-core path (ekv_umma.cuh), exported as ekv_debug_umma_probe and
// pinned by tests/test_gpu_umma.py against torch: one CTA loads a 128-key K tile and V tile by tensor-map TMA
// (128-byte swizzle), stages Q (K-major) and P^T (MN-major) by hand in the same swizzle, and runs the two
// contractions of the strided-prefill chunk kernel exactly as that kernel issues them:
//     S^T[128 keys][64 rows] = K[128][128] . Q^T          (A K-major, B K-major,  M=128 N=64 K=16 x 8)
//     O^T[128 dims][64 rows] = V^T[128 dims][128 keys] . P^T   (A MN-major, B MN-major, M=128 N=64 K=16 x 8)
// with fp32 accumulators in tensor memory read back through tcgen05.ld.  Also holds the host-side tensor-map
// constructor shared with the chunk kernel.
#include <cudaTypedefs.h>

#include "ekv_kernels.h"
#include "ekv_umma.cuh"

namespace ekv {

int make_tensor_map_rows128(CUtensorMap* map, const void* base, unsigned long long rows, int box_rows, int dtype) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t err = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (err != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn)
      return set_error(EKV_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t gdim[2] = {128, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {256};                               // bytes between rows (dimension 1)
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = dtype == EKV_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  const CUresult r = encode(map, dt, 2, const_cast<void*>(base), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(EKV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return EKV_OK;
}

template <typename T>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, const T* __restrict__ Q,
                  const T* __restrict__ Pt, float* __restrict__ St, float* __restrict__ Ot) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* Ks = smem;                 // 2 x [128 keys][64 dims]   32 KB
  unsigned char* Vs = Ks + 32768;           // 2 x [128 keys][64 dims]   32 KB
  unsigned char* Qs = Vs + 32768;           // 2 x [64 rows][64 dims]    16 KB   (K-major B operand)
  unsigned char* Ps = Qs + 16384;           // [128 keys][64 rows]       16 KB   (MN-major B operand)
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 16384);          // full, done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;

  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 128);
  // Q: row r, 16 chunks of 16 bytes; chunk c -> block c / 8 (dims 0-63 | 64-127), position (c % 8) ^ (r % 8)
  for (int i = tid; i < 64 * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    *reinterpret_cast<uint4*>(Qs + (c >> 3) * 8192 + umma::swz128(r, c & 7)) = reinterpret_cast<const uint4*>(Q + (size_t)r * 128)[c];
  }
  // P^T: key k is one 128-byte row of 64 row-values; chunk c = rows 8c .. 8c+7
  for (int i = tid; i < 128 * 8; i += 128) {
    const int k = i >> 3, c = i & 7;
    *reinterpret_cast<uint4*>(Ps + umma::swz128(k, c)) = reinterpret_cast<const uint4*>(Pt + (size_t)k * 64)[c];
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint64_t pol = l2_policy_evict_first();
    mbar_arrive_expect_tx(&bars[0], 65536);
    umma::tma_load_2d(Ks, &mapK, 0, 0, &bars[0], pol);
    umma::tma_load_2d(Ks + 16384, &mapK, 64, 0, &bars[0], pol);
    umma::tma_load_2d(Vs, &mapV, 0, 0, &bars[0], pol);
    umma::tma_load_2d(Vs + 16384, &mapV, 64, 0, &bars[0], pol);
    mbar_wait(&bars[0], 0);
    umma::fence_after_sync();
    const uint32_t id_qk = umma::instr_desc<T>(128, 64, false, false);
    const uint32_t id_pv = umma::instr_desc<T>(128, 64, true, true);
#pragma unroll
    for (int j = 0; j < 8; ++j) {           // k-step j: dims [16j, 16j+16)
      const uint64_t a = umma::smem_desc(smem_u32(Ks) + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
      const uint64_t b = umma::smem_desc(smem_u32(Qs) + (j >> 2) * 8192 + (j & 3) * 32, 16, 1024);
      umma::mma_ss(tmem, a, b, id_qk, j > 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {           // k-step j: keys [16j, 16j+16)
      const uint64_t a = umma::smem_desc(smem_u32(Vs) + j * 2048, 16384, 1024);
      const uint64_t b = umma::smem_desc(smem_u32(Ps) + j * 2048, 1024, 1024);
      umma::mma_ss(tmem + 64, a, b, id_pv, j > 0);
    }
    umma::commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  umma::fence_after_sync();
  uint32_t r[32];
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + (tid & 31);
#pragma unroll
  for (int half = 0; half < 4; ++half) {    // columns [32 half, +32): S^T in 0..63, O^T in 64..127
    umma::tmem_ld32(tmem + lane_base + half * 32, r);
    umma::tmem_wait_ld();
    float* dst = (half < 2 ? St : Ot) + (size_t)row * 64 + (half & 1) * 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) dst[j] = __uint_as_float(r[j]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

// The GQA decode kernel's shapes: N = 16 rows.  Q [16][128] as the K-major B operand (two 16-row blocks of 128-byte
// rows, 128-byte swizzle); P^T [128 keys][16 rows] as an MN-major B operand WITHOUT swizzle: key k, row group n (8 rows
// = 16 bytes) at (k / 8) * 256 + n * 128 + (k % 8) * 16.
template <typename T>
__global__ void __launch_bounds__(128, 1)
umma_probe16_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapV, const T* __restrict__ Q,
                    const T* __restrict__ Pt, float* __restrict__ St, float* __restrict__ Ot) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* Ks = smem;                 // 32 KB
  unsigned char* Vs = Ks + 32768;           // 32 KB
  unsigned char* Qs = Vs + 32768;           // 2 x [16 rows][128 B] = 4 KB
  unsigned char* Ps = Qs + 4096;            // [128 keys][32 B]   = 4 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(Ps + 4096);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (warp == 0) umma::tmem_alloc(tmem_slot, 32);
  for (int i = tid; i < 16 * 16; i += 128) {
    const int r = i >> 4, c = i & 15;
    *reinterpret_cast<uint4*>(Qs + (c >> 3) * 2048 + umma::swz128(r, c & 7)) = reinterpret_cast<const uint4*>(Q + (size_t)r * 128)[c];
  }
  for (int i = tid; i < 128 * 2; i += 128) {
    const int k = i >> 1, n = i & 1;
    *reinterpret_cast<uint4*>(Ps + (k >> 3) * 256 + n * 128 + (k & 7) * 16) = reinterpret_cast<const uint4*>(Pt + (size_t)k * 16)[n];
  }
  umma::fence_proxy_async_smem();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  if (tid == 0) {
    const uint64_t pol = l2_policy_evict_first();
    mbar_arrive_expect_tx(&bars[0], 65536);
    umma::tma_load_2d(Ks, &mapK, 0, 0, &bars[0], pol);
    umma::tma_load_2d(Ks + 16384, &mapK, 64, 0, &bars[0], pol);
    umma::tma_load_2d(Vs, &mapV, 0, 0, &bars[0], pol);
    umma::tma_load_2d(Vs + 16384, &mapV, 64, 0, &bars[0], pol);
    mbar_wait(&bars[0], 0);
    umma::fence_after_sync();
    const uint32_t id_qk = umma::instr_desc<T>(128, 16, false, false);
    const uint32_t id_pv = umma::instr_desc<T>(128, 16, true, true);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint64_t a = umma::smem_desc(smem_u32(Ks) + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
      const uint64_t b = umma::smem_desc(smem_u32(Qs) + (j >> 2) * 2048 + (j & 3) * 32, 16, 1024);
      umma::mma_ss(tmem, a, b, id_qk, j > 0);
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {           // keys [16j, 16j+16) = two groups of 8 keys = 512 bytes of P^T
      const uint64_t a = umma::smem_desc(smem_u32(Vs) + j * 2048, 16384, 1024);
      const uint64_t b = umma::smem_desc_noswizzle(smem_u32(Ps) + j * 512, 256, 128);
      umma::mma_ss(tmem + 16, a, b, id_pv, j > 0);
    }
    umma::commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  umma::fence_after_sync();
  uint32_t r[16];
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const int row = warp * 32 + (tid & 31);
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    umma::tmem_ld16(tmem + lane_base + half * 16, r);
    umma::tmem_wait_ld();
    float* dst = (half == 0 ? St : Ot) + (size_t)row * 16;
#pragma unroll
    for (int j = 0; j < 16; ++j) dst[j] = __uint_as_float(r[j]);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_dealloc(tmem, 32);
}

template <typename T>
static int probe16_t(const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot, int dtype, cudaStream_t s) {
  CUtensorMap mk, mv;
  int rc = make_tensor_map_rows128(&mk, K, 128, 128, dtype);
  if (rc) return rc;
  rc = make_tensor_map_rows128(&mv, V, 128, 128, dtype);
  if (rc) return rc;
  const int smem = 32768 * 2 + 4096 * 2 + 64 + 1024;
  cudaError_t err = cudaFuncSetAttribute(umma_probe16_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(umma_probe16)", err);
  umma_probe16_kernel<T><<<1, 128, smem, s>>>(mk, mv, reinterpret_cast<const T*>(Q), reinterpret_cast<const T*>(Pt), St, Ot);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("umma_probe16 launch", err);
  count_launch();
  return EKV_OK;
}

template <typename T>
static int probe_t(const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot, int dtype, cudaStream_t s) {
  CUtensorMap mk, mv;
  int rc = make_tensor_map_rows128(&mk, K, 128, 128, dtype);
  if (rc) return rc;
  rc = make_tensor_map_rows128(&mv, V, 128, 128, dtype);
  if (rc) return rc;
  const int smem = 32768 * 2 + 16384 * 2 + 64 + 1024;
  cudaError_t err = cudaFuncSetAttribute(umma_probe_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(umma_probe)", err);
  umma_probe_kernel<T><<<1, 128, smem, s>>>(mk, mv, reinterpret_cast<const T*>(Q), reinterpret_cast<const T*>(Pt), St, Ot);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("umma_probe launch", err);
  count_launch();
  return EKV_OK;
}

int launch_umma_probe(int dtype, const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot, cudaStream_t s) {
  if (dtype & 0x100) {                                            // rows = 16 variant (the GQA decode kernel's shapes)
    dtype &= 0xff;
    if (dtype == EKV_F16) return probe16_t<__half>(K, V, Q, Pt, St, Ot, dtype, s);
    if (dtype == EKV_BF16) return probe16_t<__nv_bfloat16>(K, V, Q, Pt, St, Ot, dtype, s);
    return set_error(EKV_ERR_UNSUPPORTED, "umma probe: 16-bit dtypes only");
  }
  if (dtype == EKV_F16) return probe_t<__half>(K, V, Q, Pt, St, Ot, dtype, s);
  if (dtype == EKV_BF16) return probe_t<__nv_bfloat16>(K, V, Q, Pt, St, Ot, dtype, s);
  return set_error(EKV_ERR_UNSUPPORTED, "umma probe: 16-bit dtypes only");
}

}  // namespace ekv
