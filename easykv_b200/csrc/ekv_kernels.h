// Internal: flattened kernel arguments + launcher prototypes shared by the .cu files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/easykv_b200.h"

namespace ekv {

struct KernelArgs {
  // ekv_layer_io
  const void* q; const void* k_new; const void* v_new; void* out;
  void* K; void* V; float* S; float* SQ; float* C; int32_t* lidx;
  const int32_t* new_slots; int32_t* victim_slots; int32_t* victim_lidx; void* scratch;
  const void* rope_cos; const void* rope_sin; const void* k_new_raw;   // fused streaming variant (all null: K is post-RoPE)
  const int32_t* seq_n_before;                                          // ragged batches (null: n_before for all)
  // ekv_shape
  int32_t dtype, B, H, Hkv, d, q_len, cap, n_before, n_phys;
  // derived
  float scale_div;   // (float)sqrt(d)
  float scale_mul;   // 1.0f / scale_div
  ekv_step st;
  unsigned long long* timeline;   // profiling hook (ekv_debug_set_timeline), normally null
};

unsigned long long* debug_timeline();

int set_cuda_error(const char* what, cudaError_t err);
int set_error(int code, const char* fmt, ...);
void count_launch();
int pdl_allowed();    // 1 unless EKV_NO_PDL was set at load (ekv_api.cu)

int launch_decode(const KernelArgs& a, cudaStream_t stream);      // ekv_decode.cu
int launch_decode_cluster(const KernelArgs& a, bool only_if_better, cudaStream_t stream);   // ekv_decode_cluster.cu
int launch_decode_umma(const KernelArgs& a, cudaStream_t stream);                          // ekv_decode_umma.cu (tcgen05, GQA)
int launch_general(const KernelArgs& a, cudaStream_t stream);     // ekv_chunk.cu
int launch_chunk_tc(const KernelArgs& a, cudaStream_t stream);    // ekv_chunk_tc.cu (16-bit dtypes, needs scratch)
long long chunk_tc_scratch_bytes(int B, int Hkv, int G, int q_len, int n_phys);
int chunk_tc_entry_limit(int q_len, int evict);                  // ekv_chunk_tc.cu: entries the evicting chunk tail can hold
int decode_cluster_entry_limit(int dtype, int G);                // ekv_decode_cluster.cu: cached slots a decode step can hold
int general_entry_limit(int q_len, int evict, int d);            // ekv_chunk.cu: ... the general kernel
int launch_select(const KernelArgs& a, cudaStream_t stream);      // ekv_aux.cu
int launch_tova_head_mean(const KernelArgs& a, cudaStream_t stream);
int launch_evict_explicit(const KernelArgs& a, const int32_t* victims, int evict, cudaStream_t stream);
int launch_rope_qk(int dtype, const void* q_in, const void* k_in, const void* v_in, const void* cos_t, const void* sin_t,
                   const int32_t* positions, void* q_out, void* k_out, void* v_out, int B, int H, int Hkv, int QL, int d,
                   cudaStream_t stream);
int launch_rope_cache(int dtype, const void* K_raw, void* K, const int32_t* lidx, const void* cos_t, const void* sin_t,
                      int units, int cap, int n_phys, int d, cudaStream_t stream);
int launch_logits_adapter(const float* logits, int rows, int V, float temperature, float top_p, int arith, const float* q_exp,
                          float* prob, float* raw, long long* token, cudaStream_t stream);   // ekv_sample.cu
int launch_token_nll(const float* logits, const long long* targets, int rows, int V, float* nll, cudaStream_t stream);
int launch_umma_probe(int dtype, const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot,
                      cudaStream_t stream);   // ekv_umma_probe.cu
int launch_export(const KernelArgs& a, void* K_out, void* V_out, float* S_out, float* SQ_out, float* C_out,
                  cudaStream_t stream);

}  // namespace ekv
