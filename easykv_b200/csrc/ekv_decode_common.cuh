// Pieces shared by the two decode kernels (persistent ping-pong: ekv_decode.cu; cluster-split:
// ekv_decode_cluster.cu): tile geometry and the transposing half-warp reduction.
#pragma once
#include "ekv_select.cuh"
#include "ekv_kernels.h"

namespace ekv {

template <typename T> struct DecodeCfg {
  static constexpr int D = 128;
  static constexpr int NWARP = 8;                       // consumer warps per group
  static constexpr int NCONS = NWARP * 32;
  static constexpr int MAX_GROUPS = 2;
  static constexpr int ROW_BYTES = D * (int)sizeof(T);
  static constexpr int TILE_BYTES = 16384;
  static constexpr int TILE_ROWS = TILE_BYTES / ROW_BYTES;   // 64 (16-bit) / 32 (fp32)
  static constexpr int RPT = TILE_ROWS / (NWARP * 2);        // rows per 16-lane group per tile
  static constexpr int MAX_STAGES = 12;
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


static inline __host__ __device__ int align_up(int x, int a) { return (x + a - 1) / a * a; }

}  // namespace ekv
