// extern "C" boundary of libeasykv_b200.so (see include/easykv_b200.h).  Validation, dispatch and
// error reporting only; no allocation, no synchronisation, no global mutable state besides the
// launch counter and the thread-local error string.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>

#include "ekv_kernels.h"
#include "ekv_chunk_plan.h"

namespace ekv {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
int set_cuda_error(const char* what, cudaError_t err) {
  snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(err));
  return EKV_ERR_CUDA;
}
static std::atomic<unsigned long long*> g_timeline{nullptr};
unsigned long long* debug_timeline() { return g_timeline.load(std::memory_order_relaxed); }
// kernel-selection overrides (development / test hook): environment at load, ekv_debug_set_dispatch later
static std::atomic<int> g_variant{[] { const char* e = getenv("EKV_DECODE_VARIANT"); return e ? atoi(e) : 0; }()};
static std::atomic<int> g_cluster{[] { const char* e = getenv("EKV_DECODE_CLUSTER"); return e ? atoi(e) : 0; }()};
static std::atomic<int> g_chunk{[] { const char* e = getenv("EKV_CHUNK_VARIANT"); return e ? atoi(e) : 0; }()};
int chunk_variant() { return g_chunk.load(std::memory_order_relaxed) & 0xff; }
int umma_force_cluster() { return (g_chunk.load(std::memory_order_relaxed) >> 8) & 0xff; }   // bits 8-15: forced cluster size   // 0 = automatic, 1 = tcgen05 path, 2 = mma.sync two-pass path
int decode_variant() { return g_variant.load(std::memory_order_relaxed); }
int decode_cluster_size() { return g_cluster.load(std::memory_order_relaxed); }
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
// programmatic dependent launch on every attention kernel (EKV_NO_PDL=1 in the environment at load turns it off: A/B hook)
static const int g_pdl = [] { const char* e = getenv("EKV_NO_PDL"); return (e && atoi(e)) ? 0 : 1; }();
int pdl_allowed() { return g_pdl; }

static int build_args(const ekv_shape* sh, const ekv_layer_io* io, const ekv_step* st, KernelArgs& a) {
  if (!sh || !io) return set_error(EKV_ERR_INVALID, "null shape/io");
  if (sh->B <= 0 || sh->H <= 0 || sh->Hkv <= 0 || sh->d <= 0 || sh->cap <= 0)
    return set_error(EKV_ERR_INVALID, "non-positive dimension (B=%d H=%d Hkv=%d d=%d cap=%d)", sh->B, sh->H, sh->Hkv, sh->d, sh->cap);
  if (sh->H % sh->Hkv) return set_error(EKV_ERR_INVALID, "H=%d is not a multiple of Hkv=%d", sh->H, sh->Hkv);
  if (sh->dtype != EKV_F16 && sh->dtype != EKV_BF16 && sh->dtype != EKV_F32) return set_error(EKV_ERR_INVALID, "dtype %d", sh->dtype);
  if (sh->n_before < 0 || sh->n_phys < sh->n_before || sh->n_phys > sh->cap)
    return set_error(EKV_ERR_INVALID, "need 0 <= n_before (%d) <= n_phys (%d) <= cap (%d)", sh->n_before, sh->n_phys, sh->cap);
  a.q = io->q; a.k_new = io->k_new; a.v_new = io->v_new; a.out = io->out;
  a.K = io->K; a.V = io->V; a.S = io->S; a.SQ = io->SQ; a.C = io->C; a.lidx = io->lidx;
  a.new_slots = io->new_slots; a.victim_slots = io->victim_slots; a.victim_lidx = io->victim_lidx; a.scratch = io->scratch;
  a.rope_cos = io->rope_cos; a.rope_sin = io->rope_sin; a.k_new_raw = io->k_new_raw; a.seq_n_before = io->seq_n_before;
  if ((a.rope_cos != nullptr) != (a.rope_sin != nullptr) || (a.rope_cos != nullptr) != (a.k_new_raw != nullptr))
    return set_error(EKV_ERR_INVALID, "rope_cos, rope_sin and k_new_raw go together");
  a.dtype = sh->dtype; a.B = sh->B; a.H = sh->H; a.Hkv = sh->Hkv; a.d = sh->d; a.q_len = sh->q_len; a.cap = sh->cap;
  a.n_before = sh->n_before; a.n_phys = sh->n_phys;
  a.scale_div = (float)std::sqrt((double)sh->d);
  a.scale_mul = 1.0f / a.scale_div;
  a.timeline = debug_timeline();
  if (st) a.st = *st;
  else {
    a.st = ekv_step{};
    a.st.policy = EKV_POLICY_NONE;
  }
  return EKV_OK;
}

// the reference would raise / mis-index on these (SURVEY A.5 "Edge"); here they are errors
static int check_step(const KernelArgs& a, int n_after) {
  const ekv_step& s = a.st;
  if (s.policy < EKV_POLICY_NONE || s.policy > EKV_POLICY_RANGE) return set_error(EKV_ERR_INVALID, "policy %d", s.policy);
  if (s.evict < 0) return set_error(EKV_ERR_INVALID, "evict %d", s.evict);
  if (s.evict == 0 || s.policy == EKV_POLICY_NONE) return EKV_OK;
  const int n_s = n_after - s.score_offset;
  if (n_s <= 0) return set_error(EKV_ERR_INVALID, "no scored slots (n=%d, score_offset=%d)", n_after, s.score_offset);
  if (s.policy == EKV_POLICY_ROCO) {
    if (s.k_feasible < s.evict || s.k_feasible > n_s)
      return set_error(EKV_ERR_INVALID, "roco: need evict (%d) <= k_feasible (%d) <= scored slots (%d)", s.evict, s.k_feasible, n_s);
  } else if (s.policy == EKV_POLICY_H2O || s.policy == EKV_POLICY_TOVA) {
    if (n_s - s.win_recent - s.win_lo < s.evict)
      return set_error(EKV_ERR_INVALID, "window [%d, %d) holds fewer than %d candidates", s.win_lo, n_s - s.win_recent, s.evict);
  } else if (s.policy == EKV_POLICY_RANGE) {
    if (s.range_start < 0 || s.range_start + s.evict > n_s)
      return set_error(EKV_ERR_INVALID, "range [%d, %d) outside the %d scored slots", s.range_start, s.range_start + s.evict, n_s);
  }
  if (s.evict > 0 && (!a.victim_lidx || !a.victim_slots)) return set_error(EKV_ERR_INVALID, "victim_slots / victim_lidx are required when evict > 0");
  return EKV_OK;
}

}  // namespace ekv

using namespace ekv;

extern "C" {

int ekv_abi_version(void) { return EKV_ABI_VERSION; }
const char* ekv_last_error(void) { return g_err; }
int64_t ekv_launch_count(void) { return (int64_t)g_launches.load(std::memory_order_relaxed); }
void ekv_debug_set_dispatch(int32_t decode_variant, int32_t cluster_size) {
  g_variant.store(decode_variant, std::memory_order_relaxed);
  g_cluster.store(cluster_size, std::memory_order_relaxed);
}
void ekv_debug_set_chunk_variant(int32_t chunk_variant) { g_chunk.store(chunk_variant, std::memory_order_relaxed); }
void ekv_debug_set_timeline(void* device_buffer) { g_timeline.store((unsigned long long*)device_buffer, std::memory_order_relaxed); }

static bool chunk_tc_shape(const ekv_shape* sh) {
  const int G = sh->Hkv > 0 ? sh->H / sh->Hkv : 0;
  return sh->q_len > 1 && sh->d == 128 && (sh->dtype == EKV_F16 || sh->dtype == EKV_BF16) &&
         (G == 1 || G == 2 || G == 4 || G == 8);
}
// scratch of the 16-bit chunk paths: the larger of the tcgen05 cluster path's and the two-pass mma.sync path's plans
static int64_t tc_bytes(const ekv_shape* sh) {
  if (!chunk_tc_shape(sh)) return 0;
  const int G = sh->H / sh->Hkv;
  int64_t b = (int64_t)chunk_tc_scratch_bytes(sh->B, sh->Hkv, G, sh->q_len, sh->n_phys);
  ChunkPlan up;
  if (make_umma_plan(sh->B, sh->Hkv, G, sh->q_len, sh->n_phys, umma_sm_count(), up) && up.bytes > b) b = up.bytes;
  return b;
}

int64_t ekv_scratch_bytes(const ekv_shape* sh, const ekv_step* st) {
  if (!sh) return 0;
  int64_t bytes = tc_bytes(sh);                       // tensor-core chunk path: row statistics, partial outputs, column sums
  if (st && st->tova_head_mean && st->policy == EKV_POLICY_TOVA)
    bytes += (int64_t)sh->B * sh->Hkv * (sh->n_before + sh->q_len) * 4;
  return bytes;
}

// q_len > 1 in a 16-bit dtype -> the tensor-core chunk kernels (when the caller supplied scratch); anything
// else, or kernel == 1 -> the exact-arithmetic general kernel
static int launch_chunk_auto(const KernelArgs& a, const ekv_shape* sh, int32_t kernel, cudaStream_t s) {
  if (kernel == 0 && chunk_tc_shape(sh) && a.scratch) {
    // 16-bit chunks: the tcgen05 / TMEM / tensor-map-TMA cluster kernel; the two-pass mma.sync kernels serve what
    // it declines (more than 8 x 12 key tiles per unit) and chunk_variant 2 (development / A-B comparisons)
    int rc = chunk_variant() == 2 ? EKV_ERR_UNSUPPORTED : launch_chunk_umma(a, s);
    if (rc != EKV_ERR_UNSUPPORTED) return rc;
    rc = launch_chunk_tc(a, s);
    if (rc != EKV_ERR_UNSUPPORTED) return rc;
  }
  return launch_general(a, s);
}

int ekv_attend_evict(const ekv_shape* sh, const ekv_layer_io* io, const ekv_step* st, int32_t kernel, void* stream) {
  KernelArgs a;
  int rc = build_args(sh, io, st, a);
  if (rc) return rc;
  if (sh->q_len < 1) return set_error(EKV_ERR_INVALID, "q_len %d", sh->q_len);
  if (!io->q || !io->k_new || !io->v_new || !io->out || !io->K || !io->V || !io->S || !io->SQ || !io->C || !io->lidx)
    return set_error(EKV_ERR_INVALID, "null tensor pointer");
  if (!io->new_slots && sh->n_phys + sh->q_len > sh->cap)
    return set_error(EKV_ERR_INVALID, "appending %d rows at %d overflows cap %d", sh->q_len, sh->n_phys, sh->cap);
  rc = check_step(a, sh->n_before + sh->q_len);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const bool defer = a.st.policy == EKV_POLICY_TOVA && a.st.tova_head_mean && a.st.accumulate;
  if (defer && (a.rope_cos || a.seq_n_before)) return set_error(EKV_ERR_UNSUPPORTED, "tova_head_mean with fused streaming / ragged batches");
  if (defer) {
    // per-head accumulate inside the fused kernel, then the cross-head mean, then the select
    if (!io->scratch) return set_error(EKV_ERR_INVALID, "tova_head_mean needs scratch (ekv_scratch_bytes)");
    const ekv_step full = a.st;
    a.st.evict = 0;
    rc = launch_chunk_auto(a, sh, kernel, s);
    if (rc) return rc;
    KernelArgs b = a;
    b.scratch = reinterpret_cast<unsigned char*>(io->scratch) + tc_bytes(sh);     // the head-mean staging follows the chunk scratch
    b.n_before = sh->n_before + sh->q_len;
    b.n_phys = io->new_slots ? sh->n_phys : sh->n_phys + sh->q_len;
    b.q_len = 0;
    b.new_slots = nullptr;
    b.st = full;
    rc = launch_tova_head_mean(b, s);
    if (rc || full.evict == 0) return rc;
    return launch_select(b, s);
  }
  if (kernel == 0 && sh->q_len == 1) {
    rc = launch_decode(a, s);
    if (rc != EKV_ERR_UNSUPPORTED) return rc;
  }
  if (a.rope_cos)      // only the decode kernels rotate on the fly: the caller falls back to ekv_rope_cache + a second buffer
    return set_error(EKV_ERR_UNSUPPORTED, "fused streaming (rope_cos) needs a decode step (q_len == 1) in a 16-bit dtype with d == 128");
  if (a.seq_n_before && sh->q_len != 1)
    return set_error(EKV_ERR_UNSUPPORTED, "ragged batches (seq_n_before) are served for decode steps (q_len == 1) only");
  return launch_chunk_auto(a, sh, kernel, s);
}

int32_t ekv_chunk_entry_limit(const ekv_shape* sh, int32_t evict, int32_t kernel) {
  if (!sh || sh->q_len < 1 || sh->Hkv < 1) return 0;
  if (evict <= 0) return INT32_MAX;                                       // nothing is selected: no per-entry scratch
  const int G = sh->H / sh->Hkv;
  const bool g_ok = G == 1 || G == 2 || G == 4 || G == 8;
  if (sh->q_len == 1 && kernel == 0 && sh->d == 128 && g_ok && evict == 1)        // the decode kernels split a unit over a cluster of <= 8 CTAs
    return decode_cluster_entry_limit(sh->dtype, G) + 1;
  if (kernel == 0 && chunk_tc_shape(sh)) return chunk_tc_entry_limit(sh->q_len, evict);
  return general_entry_limit(sh->q_len, evict, sh->d);
}

int ekv_select(const ekv_shape* sh, const ekv_layer_io* io, const ekv_step* st, void* stream) {
  KernelArgs a;
  int rc = build_args(sh, io, st, a);
  if (rc) return rc;
  if (!st) return set_error(EKV_ERR_INVALID, "null step");
  if (!io->S || !io->SQ || !io->C || !io->lidx) return set_error(EKV_ERR_INVALID, "null tensor pointer");
  a.q_len = 0;
  a.new_slots = nullptr;
  rc = check_step(a, sh->n_before);
  if (rc) return rc;
  return launch_select(a, (cudaStream_t)stream);
}

int ekv_evict_explicit(const ekv_shape* sh, const ekv_layer_io* io, const int32_t* victims, int32_t evict, void* stream) {
  KernelArgs a;
  int rc = build_args(sh, io, nullptr, a);
  if (rc) return rc;
  if (!io->lidx || (evict > 0 && !victims)) return set_error(EKV_ERR_INVALID, "null tensor pointer");
  if (evict < 0 || evict > sh->n_before) return set_error(EKV_ERR_INVALID, "evict %d of %d", evict, sh->n_before);
  return launch_evict_explicit(a, victims, evict, (cudaStream_t)stream);
}

int ekv_rope_qk(const ekv_shape* sh, const void* q_in, const void* k_in, const void* v_in, const void* cos_t,
                const void* sin_t, const int32_t* positions, void* q_out, void* k_out, void* v_out, void* stream) {
  if (!sh) return set_error(EKV_ERR_INVALID, "null shape");
  if (sh->B <= 0 || sh->H <= 0 || sh->Hkv <= 0 || sh->q_len <= 0 || sh->d <= 0 || (sh->d & 1))
    return set_error(EKV_ERR_INVALID, "bad dimension (B=%d H=%d Hkv=%d q_len=%d d=%d)", sh->B, sh->H, sh->Hkv, sh->q_len, sh->d);
  if (((q_in && q_out) || (k_in && k_out)) && (!cos_t || !sin_t)) return set_error(EKV_ERR_INVALID, "null cos / sin table");
  if ((q_in != nullptr) != (q_out != nullptr) || (k_in != nullptr) != (k_out != nullptr) || (v_in != nullptr) != (v_out != nullptr))
    return set_error(EKV_ERR_INVALID, "each of q, k, v needs both its input and its output (or neither)");
  return launch_rope_qk(sh->dtype, q_in, k_in, v_in, cos_t, sin_t, positions, q_out, k_out, v_out, sh->B, sh->H, sh->Hkv,
                        sh->q_len, sh->d, (cudaStream_t)stream);
}

int ekv_rope_cache(const ekv_shape* sh, const ekv_layer_io* io, const void* K_raw, const void* cos_t, const void* sin_t,
                   void* stream) {
  KernelArgs a;
  int rc = build_args(sh, io, nullptr, a);
  if (rc) return rc;
  if (!io->K || !io->lidx || !K_raw || !cos_t || !sin_t) return set_error(EKV_ERR_INVALID, "null tensor pointer");
  if (sh->d & 1) return set_error(EKV_ERR_INVALID, "odd head dim %d", sh->d);
  return launch_rope_cache(sh->dtype, K_raw, io->K, io->lidx, cos_t, sin_t, sh->B * sh->Hkv, sh->cap, sh->n_phys, sh->d,
                           (cudaStream_t)stream);
}

int ekv_export_logical(const ekv_shape* sh, const ekv_layer_io* io, void* K_out, void* V_out, float* S_out,
                       float* SQ_out, float* C_out, void* stream) {
  KernelArgs a;
  int rc = build_args(sh, io, nullptr, a);
  if (rc) return rc;
  if (!io->K || !io->V || !io->lidx || !K_out || !V_out) return set_error(EKV_ERR_INVALID, "null tensor pointer");
  if ((S_out && !io->S) || (SQ_out && !io->SQ) || (C_out && !io->C)) return set_error(EKV_ERR_INVALID, "state export without state");
  return launch_export(a, K_out, V_out, S_out, SQ_out, C_out, (cudaStream_t)stream);
}

int ekv_sample_top_p(const float* logits, int32_t rows, int32_t vocab, float temperature, float top_p, int32_t arith,
                     const float* q_exp, float* prob, float* raw_prob, int64_t* token, void* stream) {
  if (!logits || rows < 1 || vocab < 1) return set_error(EKV_ERR_INVALID, "logits / rows (%d) / vocab (%d)", rows, vocab);
  if (!(temperature > 0.f) || !(top_p >= 0.f)) return set_error(EKV_ERR_INVALID, "need temperature > 0 and top_p >= 0 (%g, %g)", temperature, top_p);
  if (!prob && !token) return set_error(EKV_ERR_INVALID, "no output requested");
  if (token && !q_exp) return set_error(EKV_ERR_INVALID, "token draw needs q_exp (one Exp(1) variate per logit)");
  return launch_logits_adapter(logits, rows, vocab, temperature, top_p, arith, q_exp, prob, raw_prob, (long long*)token,
                               (cudaStream_t)stream);
}

int ekv_debug_umma_probe(int32_t dtype, const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot,
                         void* stream) {
  if (!K || !V || !Q || !Pt || !St || !Ot) return set_error(EKV_ERR_INVALID, "null tensor pointer");
  return launch_umma_probe(dtype, K, V, Q, Pt, St, Ot, (cudaStream_t)stream);
}

int ekv_token_nll(const float* logits, const int64_t* targets, int32_t rows, int32_t vocab, float* nll, void* stream) {
  if (!logits || !targets || !nll || rows < 1 || vocab < 1) return set_error(EKV_ERR_INVALID, "null pointer or empty shape");
  return launch_token_nll(logits, (const long long*)targets, rows, vocab, nll, (cudaStream_t)stream);
}

}  // extern "C"
