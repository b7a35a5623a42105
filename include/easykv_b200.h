/*
 * easykv_b200 — C ABI of the B200-native KV-budgeted attention + eviction path.
 *
 * Drop-in boundary.  The reference (DRSY/EasyKV) has no FFI: its hot path is
 * PyTorch called from Python.  Each entry point below replaces a *group of reference call
 * sites*; the Python host code in `easykv_b200/` binds them with ctypes (see INTEGRATION.md
 * for the stub a maintainer of the reference would add).  All paths are relative to the
 * reference repository root.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer into caller-owned
 *     (e.g. torch-allocated) memory unless stated otherwise.  The library owns nothing.
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates
 *     nothing, never synchronises, and is CUDA-graph capturable.
 *   - return value: 0 = ok, negative = error (EKV_ERR_*); `ekv_last_error()` gives a
 *     thread-local human-readable message.  Nothing throws.
 *   - dtype: the element type of q/k/v/K/V/out (EKV_F16, EKV_BF16, EKV_F32).  Policy state is
 *     always fp32, slot maps int32.
 *
 * Supported shapes (this build): head dim d = 128 on every kernel, d = 64 and 96 on the exact CUDA-core kernel (the
 * tensor-core and decode kernels decline them and the call lands there: same results, lower throughput); GQA group size
 * H / Hkv in {1, 2, 4, 8} (every Llama-2 / Code Llama / Mistral layout); sequences of a call share n_before / n_phys
 * unless io.seq_n_before is given (decode steps).  Anything else returns EKV_ERR_UNSUPPORTED.
 *
 * HBM layout (per layer; B sequences, Hkv KV heads, `cap` physical slots, head dim d)
 *   K, V        [B, Hkv, cap, d]   dtype      physical slot order — rows NEVER move
 *   S, SQ, C    [B, Hkv, cap]      fp32       per-slot policy state (sum p, sum p^2, counter)
 *   lidx        [B, Hkv, cap]      int32      physical slot -> logical index in the reference's
 *                                             arrival-ordered cache, -1 = free slot
 *   The reference deletes victims order-preservingly (two full cache copies per step,
 *   easykv/easykv.py:56-82); here eviction renumbers `lidx` and the next token overwrites the
 *   victim's physical slot.  K is cached post-RoPE (easykv/llama_patch.py:190-196) so attention
 *   is invariant to physical order.  K/V buffers must be initialised (e.g. zeros): free slots
 *   inside [0, n_phys) are streamed and masked, never skipped.
 */
#ifndef EASYKV_B200_H
#define EASYKV_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EKV_ABI_VERSION 8

#if defined(__GNUC__)
#define EKV_API __attribute__((visibility("default")))
#else
#define EKV_API
#endif

enum { EKV_F16 = 0, EKV_BF16 = 1, EKV_F32 = 2 };

/* kv_policy of easykv/easykv.py:205.  'recency' and 'random' evict a host-chosen contiguous
 * logical range (easykv.py:343-362,491-499,741-747) => EKV_POLICY_RANGE. */
enum {
  EKV_POLICY_NONE = 0,  /* 'full' / plain attention over the retained cache            */
  EKV_POLICY_ROCO = 1,  /* easykv.py:292-296 accumulate, :319-324 / :470-476 select      */
  EKV_POLICY_H2O  = 2,  /* 'h2o_head': :288-291, :310-311 / :462-463                     */
  EKV_POLICY_TOVA = 3,  /* :297-300, :334-335 / :484-485                                 */
  EKV_POLICY_RANGE = 4  /* recency / random: evict logical [range_start, +evict)         */
};

enum {
  EKV_OK = 0,
  EKV_ERR_INVALID = -1,      /* bad argument / unsupported shape  (reference: ValueError, llama_patch.py:204-228) */
  EKV_ERR_UNSUPPORTED = -2,  /* valid request this build cannot serve                                           */
  EKV_ERR_CUDA = -3          /* CUDA runtime error on launch                                                    */
};

/* What the reference's mode loops (easykv.py:257-363, 426-500, 587-748, 816-892) decide for ONE
 * forward. */
typedef struct ekv_step {
  int32_t policy;        /* EKV_POLICY_*                                                          */
  int32_t accumulate;    /* update S/SQ from this forward's probabilities (easykv.py:443)          */
  int32_t evict;         /* victims per (sequence, kv head): 0, 1 (decode) or stride (chunk)       */
  int32_t apply;         /* 1: perform the eviction in place; 0: only report the victims           */
  int32_t score_offset;  /* 'decoding' mode: the first P logical slots carry no state (:294,:324)  */
  float   counter_add;   /* C += counter_add on scored slots before select (:304,:460,:708);
                            applied only when evict > 0                                           */
  float   c_new0;        /* C of the i-th appended slot = c_new0 - i*c_new_step (:244-245,:416,:469)*/
  float   c_new_step;
  int32_t k_feasible;    /* roco: size of the low-std candidate set (:322,:474)                    */
  int32_t protect_last;  /* roco: std[-10:] = 1e9 (:321)                                           */
  int32_t sink_protect;  /* roco strided: std[:sink] = 1e9 (:473)                                  */
  int32_t win_lo;        /* h2o/tova: candidates are state[win_lo : n_s - win_recent] (:311,:463)  */
  int32_t win_recent;
  int32_t range_start;   /* RANGE: first logical index (relative to score_offset) to evict         */
  int32_t arith;         /* which ATen kernels' arithmetic to reproduce at the two places they
                            differ: 0 = CPU (logits / sqrt(d); exp * (1/sum)),
                                    1 = CUDA (logits * (1/sqrt(d)); exp / sum)                    */
  int32_t tova_head_mean;/* tova in 'encoding'/'ppl': state = mean over KV heads of the last
                            query row, broadcast to all heads (:454-457,:845-848)                 */
  int32_t raw_colsum;    /* keep_attention seeding (h2o_head_score, :173-186): the dense prefill is issued
                            as causal chunks whose column sums must be added to S / SQ WITHOUT the
                            per-forward model-dtype rounding of :450-451; the caller rounds S / SQ once
                            when the prefill is complete (torch.sum over the whole map rounds once) */
  int32_t budget_gate;   /* ragged batches (io->seq_n_before): a sequence evicts (and counts, counter_add) only when
                            its scored slots after the append exceed this — the reference's
                            `if cur_kv_size - len(prefix) > budget` (:303, :459, :709) evaluated per sequence.
                            0: every sequence evicts `evict` slots                                   */
} ekv_step;

/* One layer's tensors for one forward.  Replaces the body of llama_forward / mistral_forward
 * between the projections and o_proj (easykv/llama_patch.py:193-230, mistral_patch.py:137-170)
 * AND the per-layer slice of the score/evict block of generate() (easykv/easykv.py:271-362). */
typedef struct ekv_layer_io {
  const void* q;        /* [B, H,   q_len, d] post-RoPE queries                                   */
  const void* k_new;    /* [B, Hkv, q_len, d] post-RoPE keys of the q_len appended tokens          */
  const void* v_new;    /* [B, Hkv, q_len, d]                                                      */
  void*       out;      /* [B, H,   q_len, d] softmax(QK^T/sqrt(d)) V  (input of o_proj)           */
  void*       K;        /* [B, Hkv, cap, d]   in/out: the appended rows are written here           */
  void*       V;
  float*      S;        /* [B, Hkv, cap]      in/out                                               */
  float*      SQ;
  float*      C;
  int32_t*    lidx;     /* [B, Hkv, cap]      in/out                                               */
  const int32_t* new_slots;  /* [B, Hkv, q_len] physical slots the appended tokens go to; they must
                                be free (lidx == -1 or >= n_phys).  NULL: slot n_phys + i.          */
  int32_t*    victim_slots;  /* [B, Hkv, evict] out: physical slots freed (may alias new_slots),
                                ordered like victim_lidx                                          */
  int32_t*    victim_lidx;   /* [B, Hkv, evict] out: the reference's eviction ids, ascending       */
  void*       scratch;       /* >= ekv_scratch_bytes() bytes of device memory, or NULL (see below)  */
  /* --- streaming variant fused into the step (llama_forward_stream, llama_patch.py:310-327): K holds the UN-rotated
   * keys; each cached row is rotated at its cache-relative position lidx[slot] while it is read (same roundings as
   * apply_rotary_pos_emb, :47-72), so no rotated copy of the cache is ever written.  All three NULL: K is post-RoPE.
   * Supported by the decode kernels' FMA paths (q_len == 1, 16-bit dtypes, d == 128, group size <= 4); otherwise the
   * call returns EKV_ERR_UNSUPPORTED and the caller runs ekv_rope_cache into a second buffer first. */
  const void* rope_cos;      /* [rows, d] model-dtype tables, rows > every logical index                */
  const void* rope_sin;
  const void* k_new_raw;     /* [B, Hkv, q_len, d] un-rotated keys of the appended tokens: THESE are written to K;
                                k_new (rotated at the tokens' own positions) only enters this forward's logits */
  /* --- ragged batches: sequences of one call may hold different numbers of valid slots ------------------------- */
  const int32_t* seq_n_before;  /* [B] valid slots of each sequence before the append, each <= shape.n_before;
                                   shape.n_phys bounds every sequence's physical extent (slots beyond a sequence's own
                                   are free: lidx == -1).  NULL: shape.n_before for all.  Decode steps (q_len == 1)
                                   and the general kernel; EKV_ERR_UNSUPPORTED elsewhere.               */
} ekv_layer_io;

typedef struct ekv_shape {
  int32_t dtype;     /* EKV_F16 | EKV_BF16 | EKV_F32                                              */
  int32_t B, H, Hkv, d;
  int32_t q_len;     /* 1 = decode step, >1 = strided prefill chunk (causal inside the chunk)     */
  int32_t cap;       /* physical slots per (sequence, kv head)                                    */
  int32_t n_before;  /* valid slots before this forward's append (the largest, with io->seq_n_before) */
  int32_t n_phys;    /* physical slots [0, n_phys) hold every valid slot and are streamed         */
} ekv_shape;

EKV_API int         ekv_abi_version(void);
EKV_API const char* ekv_last_error(void);

/* Bytes of `scratch` a call with this shape/step uses.  Non-zero for (a) strided chunks (q_len > 1) in a 16-bit
 * dtype — the tensor-core chunk kernels keep their per-split row statistics, partial outputs, column sums and
 * selection keys there; with scratch == NULL such a call falls back to the exact CUDA-core general kernel — and
 * (b) step->tova_head_mean (required).  Contents need not survive the call. */
EKV_API int64_t ekv_scratch_bytes(const ekv_shape* shape, const ekv_step* step);

/* Largest number of entries per (sequence, kv head) — cached slots + the q_len appended rows — that an EVICTING forward
 * of this shape (dtype, d, H / Hkv, q_len; n_before / n_phys are ignored) can hold when `evict` victims are selected:
 * the per-unit tail of a strided chunk (q_len > 1) and the exact general kernel keep one logical index and two
 * selection keys per entry in one CTA's shared memory (about 17.6 K entries in a 16-bit dtype on the tensor-core
 * path, about 11 K on the general kernel, which also serves kernel == 1, fp32 and d != 128).  Decode steps on the
 * decode kernels (q_len == 1, evict == 1, kernel == 0, d == 128) split a unit over a cluster of up to 8 CTAs: tens of
 * thousands of slots, depending on dtype and group size.  Non-evicting forwards have no ceiling: INT32_MAX.
 * A caller checks its schedule against this BEFORE the first forward instead of meeting EKV_ERR_UNSUPPORTED in
 * the middle of a prompt (the reference has no such limit: easykv.py:459-499).  No GPU needed.  (Additive in ABI v8.) */
EKV_API int32_t ekv_chunk_entry_limit(const ekv_shape* shape, int32_t evict, int32_t kernel);

/* Fused forward for one layer: append -> QK^T/sqrt(d) -> softmax -> PV -> GQA fold ->
 * policy accumulate -> budgeted victim select -> in-place eviction.
 * Replaces: llama_patch.py:193-230 (cache append, repeat_kv, matmul, mask, softmax, matmul),
 * easykv.py:188-196 (GQA fold), :288-300/:443-457 (accumulate), :303-362/:459-499 (select),
 * :56-82 (KV compaction) and :315-333/:465-483 (state compaction).
 * `kernel`: 0 = automatic — q_len == 1: the persistent TMA-pipelined decode kernel (MHA) or the cluster-split
 * decode kernel (GQA, long caches, small batches); q_len > 1 in a 16-bit dtype with scratch: the tensor-core chunk
 * kernels; anything else: the general kernel.  1 = force the exact CUDA-core general kernel. */
EKV_API int ekv_attend_evict(const ekv_shape* shape, const ekv_layer_io* io, const ekv_step* step,
                     int32_t kernel, void* stream);

/* Standalone select over existing state (no attention, no append).  Same victim semantics as
 * the fused call (step->accumulate is ignored, counter_add IS applied); used when attention ran
 * elsewhere and by the unit tests.  `io` needs S, SQ, C, lidx, victim_slots, victim_lidx.
 * n = shape.n_before valid slots inside [0, n_phys); shape.q_len is ignored.
 * Replaces easykv.py:310-347 / :462-493 in isolation. */
EKV_API int ekv_select(const ekv_shape* shape, const ekv_layer_io* io, const ekv_step* step, void* stream);

/* Evict an explicit victim list.  `victims` [B, Hkv, evict] holds logical ids (any order, no
 * duplicates).  Renumbers lidx, frees the slots and writes victim_slots / victim_lidx
 * (ascending).  Replaces truncate_kv_cache_silo / _liso / truncate_kv_cache
 * (easykv.py:56-82,105-112) plus the matching state compaction. */
EKV_API int ekv_evict_explicit(const ekv_shape* shape, const ekv_layer_io* io, const int32_t* victims,
                       int32_t evict, void* stream);

/* Materialise the reference's arrival-ordered view: rows with lidx >= 0 are copied to
 * K_out/V_out[B, Hkv, n, d] (and S/SQ/C_out[B, Hkv, n] if non-NULL) at index lidx
 * (n = shape.n_before).  For export of a legacy `[layer][0|1] -> [B,Hkv,n,d]` cache
 * (easykv.py:251,302), for re-densifying the physical layout, and for parity checks. */
EKV_API int ekv_export_logical(const ekv_shape* shape, const ekv_layer_io* io, void* K_out, void* V_out,
                       float* S_out, float* SQ_out, float* C_out, void* stream);

/* RoPE of q and of the new k at explicit positions, fused with the re-layout of the projections' outputs
 * q_in [B, q_len, H, d], k_in / v_in [B, q_len, Hkv, d] into the head-major q_out [B, H, q_len, d], k_out / v_out
 * [B, Hkv, q_len, d] that ekv_attend_evict takes (v is only re-laid out; any of the three may be NULL).
 * cos / sin are [rows, d] tables in the model dtype; row of (b, i) = positions[b * q_len + i], or b * q_len + i
 * when positions is NULL (tables already gathered per token).  Arithmetic as the reference evaluates it in
 * the model dtype: rn(x*cos) + rn(rotate_half(x)*sin), rounded.  shape->n_before / n_phys / cap are ignored.
 * Replaces apply_rotary_pos_emb (easykv/llama_patch.py:47-72, mistral_patch.py:62-87) and the transposes at
 * llama_patch.py:169-171. */
EKV_API int ekv_rope_qk(const ekv_shape* shape, const void* q_in, const void* k_in, const void* v_in, const void* cos,
                const void* sin, const int32_t* positions, void* q_out, void* k_out, void* v_out, void* stream);

/* Streaming variant (generation_config['streaming'], llama_forward_stream / mistral_forward_stream,
 * easykv/llama_patch.py:251-379, mistral_patch.py:189-286): the cache keeps UN-rotated keys and every forward
 * rotates all of them at their cache-relative positions.  K_raw [B, Hkv, cap, d] holds the un-rotated rows in the
 * physical layout of io->K; this call writes io->K[slot] = rope(K_raw[slot], position = io->lidx[slot]) for every
 * valid slot in [0, n_phys) (cos / sin: [rows, d] tables, model dtype), after which ekv_attend_evict runs unchanged
 * with q and the new k rotated at positions n_before .. n_before + q_len - 1 (ekv_rope_qk).  The caller stores the
 * new tokens' un-rotated k rows into K_raw at the slots ekv_attend_evict appended them to. */
EKV_API int ekv_rope_cache(const ekv_shape* shape, const ekv_layer_io* io, const void* K_raw, const void* cos, const void* sin,
                   void* stream);

/* Sampling tail (replaces logits_adapter, easykv/easykv.py:115-134, and the torch.multinomial call at :258 / :509 /
 * :671): for each of `rows` rows of fp32 logits [rows, vocab] compute softmax(logits / temperature), keep in
 * descending order every token whose exclusive cumulative mass is <= top_p, renormalise -> prob [rows, vocab]
 * (nullable when only the token is wanted and vocab * 4 <= 200 KB).  raw_prob (nullable) receives softmax(logits).
 * With q_exp [rows, vocab] (one Exp(1) variate per logit — `torch.empty_like(prob).exponential_(1)` is exactly the draw
 * torch.multinomial makes) token[row] = argmax(prob / q_exp), first index on ties: the token torch.multinomial(prob, 1)
 * returns from the same generator state, without its two host syncs.  arith as in ekv_step.arith (how logits /
 * temperature is formed: 1 = multiply by the fp32 reciprocal, ATen's CUDA flavour; 0 = true division).
 * One launch, no scratch; floating point: agrees with the ATen composition to fp32 rounding (summation order). */
EKV_API int ekv_sample_top_p(const float* logits, int32_t rows, int32_t vocab, float temperature, float top_p, int32_t arith,
                     const float* q_exp, float* prob, float* raw_prob, int64_t* token, void* stream);

/* Perplexity tail (replaces CrossEntropyLoss(reduction='none') over the concatenated chunk logits, easykv/easykv.py:
 * 826-827, 896-899): nll[row] = -(log_softmax(logits[row])[targets[row]]), fp32 logits [rows, vocab], int64 targets
 * (NaN for a target outside [0, vocab)).  Called once per prompt chunk, so no [prompt, vocab] tensor is ever kept. */
EKV_API int ekv_token_nll(const float* logits, const int64_t* targets, int32_t rows, int32_t vocab, float* nll, void* stream);

/* Number of kernels launched by this library in the calling process since load (bench.py's
 * gpu_launches claim). */
EKV_API int64_t ekv_launch_count(void);

/* Profiling hook (development aid, not part of the data path): when set to a device buffer of
 * uint64 [grid][16 units][8], the decode kernel's consumer groups record SM-clock timestamps at their
 * phase boundaries (unit start, header, first tile, K phase, softmax, V phase, out, tail) plus the
 * CTA's start %globaltimer in slot [cta][15][7].  NULL (the default) disables it. */
EKV_API void ekv_debug_set_timeline(void* device_buffer);

/* Kernel-selection override (development / test hook, not part of the data path).
 * decode_variant: 0 = automatic, 1 = one consumer group per CTA, 2 = ping-pong groups (ekv_decode.cu),
 *                 3 = never / 4 = always use the tensor-core variant of the cluster kernel (g >= 4, 16-bit;
 *                 automatic: g = 8 only), 5 = always / 6 = never use the tcgen05 GQA decode kernel (ekv_decode_umma.cu;
 *                 automatic: 16-bit dtypes, g >= 2, at least 32 (sequence, kv head) units).
 * cluster_size:   0 = automatic, -1 = never use the cluster-split decode kernel, 1/2/4/8 = always use it with
 *                 this many CTAs per (sequence, kv head) (ekv_decode_cluster.cu; with decode_variant 5: the tcgen05
 *                 kernel's cluster size, 1/2/4). */
EKV_API void ekv_debug_set_dispatch(int32_t decode_variant, int32_t cluster_size);

/* Strided-chunk kernel selection (development / test hook): 0 = automatic (the tcgen05 cluster kernel, falling back to
 * the two-pass mma.sync kernels for key ranges beyond 8 x 12 tiles), 2 = always the mma.sync kernels, 3 = the tcgen05 kernel
 * with 8 instead of 16 softmax warps; bits 8-15: force this cluster size (0 = the planner's choice). */
EKV_API void ekv_debug_set_chunk_variant(int32_t chunk_variant);

/* Primitive-level probe of the tensor-core path the strided-prefill chunk kernel is built from (development / test
 * hook, not part of the data path): one CTA loads K, V [128, 128] (16-bit, row-major) by tensor-map TMA, stages
 * Q [64, 128] and Pt [128 keys, 64 rows] in shared memory and computes with tcgen05.mma into tensor memory
 *   St [128 keys, 64 rows] = K . Q^T      Ot [128 dims, 64 rows] = V^T . Pt      (fp32).
 * tests/test_gpu_umma.py pins both against torch. */
EKV_API int ekv_debug_umma_probe(int32_t dtype, const void* K, const void* V, const void* Q, const void* Pt, float* St, float* Ot,
                                 void* stream);

#ifdef __cplusplus
}
#endif
#endif /* EASYKV_B200_H */
