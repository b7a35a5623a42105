// Scratch layout and grid plan shared by the strided-prefill chunk paths: the tcgen05 cluster kernel
// (ekv_chunk_umma.cu) and the mma.sync two-pass kernels (ekv_chunk_tc.cu), plus the two finishing kernels both use
// (chunk_out_kernel: sum of partial outputs + chip-wide policy-state update; chunk_tail_kernel: append + select).
#pragma once
#include "ekv_kernels.h"

namespace ekv {

struct ChunkPlan {
  int R, RB, Rpad, NE, NEpad, ntiles, splits, tps;      // tps = tiles per split (pass 2)
  int splits1, tps1;                                    // pass 1 (lighter CTAs, three per SM): a finer split
  long long off_stats, off_opart, off_cpart, off_klj, off_ka, off_kb, off_kf, bytes;
  // tcgen05 path: 128-key tiles — nct over the cached slots, nnt over the chunk's own keys; a cluster of `splits`
  // CTAs per (unit, 64-row block), `tps` tiles each
  int nct, nnt;
  int cparts;                                           // column-statistics partials per 64-row block (query groups)
};

ChunkPlan make_chunk_plan(int B, int Hkv, int G, int q_len, int n_phys);            // ekv_chunk_tc.cu
// ekv_chunk_umma.cu: plan for the tcgen05 path; returns false when the shape is outside it (then `p` is untouched)
bool make_umma_plan(int B, int Hkv, int G, int q_len, int n_phys, int sms, ChunkPlan& p);
// the launches after the attention part: out + state update, then the per-unit tail (ekv_chunk_tc.cu)
int launch_chunk_finish(const KernelArgs& a, const ChunkPlan& pl, cudaStream_t stream);
int launch_chunk_umma(const KernelArgs& a, cudaStream_t stream);                   // ekv_chunk_umma.cu
int umma_sm_count();
int chunk_variant();                                                               // ekv_api.cu

}  // namespace ekv
