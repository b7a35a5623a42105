// Strided-prefill chunk (q_len = stride query rows per head) on the Blackwell tensor cores: tcgen05.mma with the
// accumulators in tensor memory, K / V tiles by tensor-map TMA (cp.async.bulk.tensor, 128-byte swizzle) through an
// mbarrier ring, warp-specialised (TMA producer / MMA issuer / 8 softmax warps), one thread-block CLUSTER per
// (sequence, kv head, block of 64 (query, head) rows) so that every K and V byte is read from HBM exactly once and
// the [H, q, n] probability tensor of the reference (easykv/llama_patch.py:244-246) never exists anywhere.
//
// Orientation.  Both contractions are issued TRANSPOSED so that a tensor-memory lane is a KEY:
//     S^T [128 keys x 64 rows] = K_tile [128 x 128] . Q^T            A = K tile (K-major, from TMA), B = Q (K-major)
//     O^T [128 dims x 64 rows] += V_tile^T [128 x 128 keys] . P^T     A = V tile (MN-major, from TMA), B = P^T (MN-major)
// A softmax thread (tcgen05.ld 32x32b: one lane = one key) therefore owns one key and sees that key's logit for every
// row of the block in its registers: the per-key column statistics the eviction policies need (GQA fold over the g
// heads, sum over the chunk's queries of p and p^2 — easykv/easykv.py:188-196, :443-457) are thread-local, and so is
// the write of a key's P^T row (one 128-byte swizzled row of the B operand).  Row statistics (max, sum) go across
// lanes once per CTA, not per tile.
//
// Exact two-phase softmax (the reference rounds p = dtype(exp(x - max) / sum) AFTER normalising with the global
// row statistics, so an online rescaling softmax cannot reproduce it bit for bit):
//   K phase   the CTA's slice of the keys (<= 12 tiles): S^T by tcgen05.mma into a double-buffered TMEM accumulator,
//             logits rounded at the reference's rounding points (llama_patch.py:201-202), masked, packed to the model
//             dtype and PARKED IN TENSOR MEMORY (the 384 columns the S^T double buffer leaves free hold 12 tiles;
//             the O^T accumulator of the V phase reuses the S^T columns), row
//             maxima kept as packed 16-bit maxima in registers;
//   exchange  per-row slice maxima all-to-all over distributed shared memory (st.shared::cluster + remote mbarrier
//             arrives: only the softmax warps take part, the TMA / MMA warps run ahead into the V stream);
//   L pass    sum of exp(x - max) from the parked logits; second exchange; every CTA adds the slices in rank order,
//             so the denominators are bit-identical across the cluster;
//   V phase   p = dtype(exp(x - max) / sum) exactly as softmax + .to(dtype) form it (llama_patch.py:218-219), column
//             statistics -> scratch, P^T tile -> shared memory (double-buffered), O^T += V^T P^T by tcgen05.mma;
//   epilogue  the CTA's partial O^T: TMEM -> registers -> scratch; chunk_out_kernel / chunk_tail_kernel
//             (ekv_chunk_tc.cu) sum the partials, fold the column statistics into the policy state, select and evict.
//
// Replaces: llama_patch.py:193-230 / mistral_patch.py:137-170 and easykv.py:439-457 / :599-618 / :830-848 for one
// layer of one strided forward (and the dense prefill issued as causal chunks, h2o_head_score :173-186).
#include "ekv_chunk_plan.h"
#include "ekv_mma.cuh"
#include "ekv_umma.cuh"

namespace ekv {

namespace cu {
constexpr int D = 128;
constexpr int TKEYS = 128;               // keys per tile = MMA M
constexpr int NROWS = 64;                // (query, head) rows per cluster = MMA N
// softmax warps: 4 TMEM lane quarters x NSPLIT column groups (NSPLIT = 2: 8 warps x 32 rows per thread,
// NSPLIT = 4: 16 warps x 16 rows per thread); + TMA producer warp + MMA warp
__host__ __device__ constexpr int nsoft(int nsplit) { return 128 * nsplit; }
__host__ __device__ constexpr int nthreads(int nsplit) { return nsoft(nsplit) + 64; }
constexpr int STAGE_BYTES = 32768;       // one K or V tile: 2 boxes of [128 keys][64 dims]
constexpr int NSTAGE = 5;
constexpr int MAX_TILES = 12;            // logits parked in TMEM: 12 tiles x 32 columns
constexpr int MAX_CLUSTER = 8;
constexpr float LOG2E = 1.4426950408889634f;
// tensor memory columns: S^T double buffer [0,128); O^T reuses [0,64) once the K phase is over; parked logits [128,512)
constexpr uint32_t TM_S = 0, TM_O = 0, TM_LOG = 128, TM_COLS = 512;
// shared memory (after 1024-byte alignment)
constexpr int OFF_RING = 0;
constexpr int OFF_Q = OFF_RING + NSTAGE * STAGE_BYTES;
constexpr int OFF_P = OFF_Q + 16384;
constexpr int OFF_BAR = OFF_P + 2 * 16384;                // 32 mbarriers
constexpr int OFF_TMEM = OFF_BAR + 32 * 8;
constexpr int OFF_REDMAX = OFF_TMEM + 16;                 // [4 lane quarters][64 rows] slice maxima
constexpr int OFF_REDSUM = OFF_REDMAX + 4 * 64 * 4;       // [4 lane quarters][64 rows] slice sums
constexpr int OFF_REDM0 = OFF_REDSUM + 4 * 64 * 4;        // [4 lane quarters][64 rows] reference points of the slice sums
constexpr int OFF_XMAX = OFF_REDM0 + 4 * 64 * 4;          // [MAX_CLUSTER][64]
constexpr int OFF_XSUM = OFF_XMAX + MAX_CLUSTER * 64 * 4;
constexpr int OFF_XM0 = OFF_XSUM + MAX_CLUSTER * 64 * 4;
constexpr int OFF_ROWM = OFF_XM0 + MAX_CLUSTER * 64 * 4;  // [64] max | [64] sum or 1/sum | [64] rcp(sum) | flag
constexpr int OFF_LIDX = OFF_ROWM + 3 * 64 * 4 + 16;        // [MAX_TILES][128] the CTA's slice of the slot map (bulk copy)
// the exact pass's sums get their own buffer (a fast peer may already be sending them while a slow CTA still reads the
// first exchange): the slot-map slice, which is dead once the validity masks are built, before the first exchange
constexpr int OFF_XSUM2 = OFF_LIDX;
constexpr int SMEM_BYTES = OFF_LIDX + MAX_TILES * TKEYS * 4;
static_assert(MAX_CLUSTER * 64 * 4 <= MAX_TILES * TKEYS * 4, "xsum2 aliases the slot-map slice");
constexpr int SMEM_ALLOC = SMEM_BYTES + 1024;
static_assert(SMEM_ALLOC <= 227 * 1024, "shared memory per CTA");
// barrier indices
constexpr int B_FULL = 0, B_EMPTY = NSTAGE, B_SFULL = 2 * NSTAGE, B_SEMPTY = B_SFULL + 2, B_PFULL = B_SEMPTY + 2,
              B_PEMPTY = B_PFULL + 2, B_OFULL = B_PEMPTY + 2, B_XCH = B_OFULL + 1, B_LIDX = B_XCH + 2;
static_assert(B_LIDX + 1 <= 32, "barrier block");
}  // namespace cu

int umma_force_cluster();      // ekv_api.cu (env EKV_CHUNK_CLUSTER / ekv_debug_set_chunk_variant): 0 = planner's choice

template <typename T, int G, bool ARITH, int NSPLIT>
__global__ void __launch_bounds__(cu::nthreads(NSPLIT), 1)
chunk_umma_kernel(const KernelArgs a, const ChunkPlan pl, const __grid_constant__ CUtensorMap mapK,
                  const __grid_constant__ CUtensorMap mapV, const __grid_constant__ CUtensorMap mapKn,
                  const __grid_constant__ CUtensorMap mapVn) {
  using namespace cu;
  constexpr int NSOFT = nsoft(NSPLIT), NT = nthreads(NSPLIT), NSW = NSOFT / 32;
  constexpr int CW = NROWS / NSPLIT;                             // rows (TMEM columns) per softmax thread
  constexpr int CW2 = CW / 2;                                    // ... as packed 16-bit pairs
  static_assert(CW % G == 0 && CW2 >= 8, "a thread's rows hold whole queries");
  pdl_trigger();            // programmatic dependent launch (ekv_common.cuh): the next kernel may be placed now;
  pdl_wait();               // this one touches global memory only once the previous kernel of the stream has completed
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  unsigned char* ring = smem + OFF_RING;
  unsigned char* Qs = smem + OFF_Q;
  unsigned char* Ps = smem + OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  float* redmax = reinterpret_cast<float*>(smem + OFF_REDMAX);
  float* redsum = reinterpret_cast<float*>(smem + OFF_REDSUM);
  float* xmax = reinterpret_cast<float*>(smem + OFF_XMAX);
  float* xsum = reinterpret_cast<float*>(smem + OFF_XSUM);
  float* redm0 = reinterpret_cast<float*>(smem + OFF_REDM0);
  float* xm0 = reinterpret_cast<float*>(smem + OFF_XM0);
  float* xsum2 = reinterpret_cast<float*>(smem + OFF_XSUM2);
  float* rowM = reinterpret_cast<float*>(smem + OFF_ROWM);
  float* rowL = rowM + 64;
  float* rowR = rowL + 64;
  int* slow_flag = reinterpret_cast<int*>(rowR + 64);
  int32_t* lidx_s = reinterpret_cast<int32_t*>(smem + OFF_LIDX);              // a row's one-pass denominator was unusable: exact L pass

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = pl.splits;
  const int rank = blockIdx.x % C;
  const int rb = (blockIdx.x / C) % pl.RB;
  const int unit = blockIdx.x / (C * pl.RB);
  const int b = unit / a.Hkv, h = unit % a.Hkv;
  const int QL = a.q_len, n_phys = a.n_phys;
  const int nct = pl.nct, nt = pl.nct + pl.nnt;
  const int t0 = (int)((long long)rank * nt / C), t1 = (int)((long long)(rank + 1) * nt / C);   // balanced: sizes differ by <= 1
  const int T_ = t1 - t0;
  unsigned long long* tl = a.timeline ? a.timeline + (size_t)a.B * a.Hkv * 8 + (size_t)blockIdx.x * 16 : nullptr;
  auto stamp = [&](int i) { if (tl && tid == 0) tl[i] = global_ns(); };
  stamp(0);

  // ---- setup -----------------------------------------------------------------------------------------------------------
  auto issue_tile = [&](int it, uint64_t pol) {                   // `it`-th tile of the producer's K-then-V stream
    const int slot = it % NSTAGE;
    const int i = it < T_ ? it : it - T_, t = t0 + i;
    const bool is_new = t >= nct;
    const CUtensorMap* map = it < T_ ? (is_new ? &mapKn : &mapK) : (is_new ? &mapVn : &mapV);
    const int row = is_new ? unit * QL + (t - nct) * TKEYS : unit * a.cap + t * TKEYS;
    unsigned char* dst = ring + (size_t)slot * STAGE_BYTES;
    mbar_arrive_expect_tx(&bars[B_FULL + slot], STAGE_BYTES);
    umma::tma_load_2d(dst, map, 0, row, &bars[B_FULL + slot], pol);
    umma::tma_load_2d(dst + 16384, map, 64, row, &bars[B_FULL + slot], pol);
  };
  // several row blocks (clusters) of a unit read the same K / V: keep those lines in L2; a single reader streams
  uint64_t tma_pol = 0;
  if (tid == NSOFT) {
    for (int s = 0; s < NSTAGE; ++s) { mbar_init(&bars[B_FULL + s], 1); mbar_init(&bars[B_EMPTY + s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_SFULL + s], 1); mbar_init(&bars[B_SEMPTY + s], NSW);
      mbar_init(&bars[B_PFULL + s], NSW); mbar_init(&bars[B_PEMPTY + s], 1);
      mbar_init(&bars[B_XCH + s], 64 * C);
    }
    mbar_init(&bars[B_OFULL], 1);
    mbar_init(&bars[B_LIDX], 1);
    *slow_flag = 0;
    mbar_fence_init();
    {
      // this CTA's slice of the slot map, one bulk copy ahead of the tiles (a per-thread load issued into the opening
      // burst of tile requests was measured at ~4 us): ints [t0*128, +ncached) of the unit, clipped to the capacity
      const int first = t0 * TKEYS;
      int cnt = min(t1, nct) * TKEYS - first;
      if (cnt > a.cap - first) cnt = a.cap - first;              // cap is a multiple of 8 ints: 16-byte granularity holds
      if (cnt > 0) {
        cnt = (cnt + 3) & ~3;
        if (cnt > a.cap - first) cnt = (a.cap - first) & ~3;
      }
      if (cnt > 0) {
        mbar_arrive_expect_tx(&bars[B_LIDX], (uint32_t)cnt * 4u);
        tma_bulk_g2s(lidx_s, a.lidx + (size_t)unit * a.cap + first, (uint32_t)cnt * 4u, &bars[B_LIDX], l2_policy_evict_first());
      } else {
        mbar_arrive(&bars[B_LIDX]);
      }
    }
    // the first ring-full of tiles needs nothing but these barriers: their HBM latency overlaps the rest of the setup
    tma_pol = pl.RB > 1 ? umma::l2_policy_evict_last() : l2_policy_evict_first();
    for (int it = 0; it < NSTAGE && it < 2 * T_; ++it) issue_tile(it, tma_pol);
    if (tl) tl[8] = global_ns();
  }
  if (warp == NSW + 1) umma::tmem_alloc(tmem_slot, TM_COLS);
  const int q4 = warp & 3, cgp = warp >> 2;                      // TMEM lane quarter; column group (softmax warps)
  const int kl = q4 * 32 + lane;                                 // key inside a tile (softmax) / output dim (epilogue)
  if (warp == NSW + 1) {
    // Q block as the K-major B operand, staged by the MMA warp itself with asynchronous 16-byte copies (nobody else
    // reads it; the copies overlap the first K tile's HBM latency): row `col` of the block is (query qi, head g)
    // with qi = (rb*64 + col) / G
    const T* qg = reinterpret_cast<const T*>(a.q) + (size_t)b * a.H * QL * D;
    for (int i = lane; i < NROWS * 16; i += 32) {
      const int col = i >> 4, c = i & 15;
      const int r = rb * NROWS + col, qi = r / G, g = r % G;
      unsigned char* dst = Qs + (c >> 3) * 8192 + umma::swz128(col, c & 7);
      if (qi < QL) cp_async16(dst, reinterpret_cast<const uint4*>(qg + ((size_t)(h * G + g) * QL + qi) * D) + c);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
  }
  umma::fence_before_sync();
  stamp(9);
  __syncthreads();                                               // barriers initialised, TMEM base published
  stamp(10);
  // every CTA's barriers exist before any REMOTE arrive: arrive now, wait only right before the exchange
  // (relaxed: the barrier initialisations were published by fence.mbarrier_init.release.cluster; a releasing arrive
  // would also drain this thread's other traffic — measured at ~3 us here)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  stamp(1);

  if (warp == NSW) {
    // ===== TMA producer ==================================================================================================
    if (lane == 0) {
      for (int it = NSTAGE; it < 2 * T_; ++it) {
        mbar_wait_backoff(&bars[B_EMPTY + it % NSTAGE], (it / NSTAGE - 1) & 1, 128);
        issue_tile(it, tma_pol);
      }
    }
  } else if (warp == NSW + 1) {
    // ===== MMA issuer ====================================================================================================
    cp_async_wait<0>();
    umma::fence_proxy_async_smem();                              // this lane's Q chunks -> visible to the tensor core
    __syncwarp();
    if (lane == 0) {
      const uint32_t id_qk = umma::instr_desc<T>(TKEYS, NROWS, false, false);
      const uint32_t id_pv = umma::instr_desc<T>(D, NROWS, true, true);
      const uint32_t ring_a = smem_u32(ring), q_a = smem_u32(Qs), p_a = smem_u32(Ps);
      int it = 0;
      for (int i = 0; i < T_; ++i, ++it) {                       // S^T(tile) = K_tile . Q^T
        const int slot = it % NSTAGE, sb = i & 1;
        mbar_wait_backoff(&bars[B_FULL + slot], (it / NSTAGE) & 1, 32);
        if (i >= 2) mbar_wait_backoff(&bars[B_SEMPTY + sb], ((i >> 1) - 1) & 1, 32);
        umma::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {                            // k-step: dims [16j, 16j+16)
          const uint64_t da = umma::smem_desc(ring_a + slot * STAGE_BYTES + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
          const uint64_t db = umma::smem_desc(q_a + (j >> 2) * 8192 + (j & 3) * 32, 16, 1024);
          umma::mma_ss(tmem + TM_S + sb * NROWS, da, db, id_qk, j > 0);
        }
        umma::commit(&bars[B_EMPTY + slot]);
        umma::commit(&bars[B_SFULL + sb]);
      }
      // O^T reuses the columns of the S^T double buffer: the softmax warps have drained every S tile before they
      // produce the first P tile (program order in those warps), which the first wait below covers
      for (int i = 0; i < T_; ++i, ++it) {                       // O^T += V_tile^T . P^T(tile)
        const int slot = it % NSTAGE, pb = i & 1;
        mbar_wait_backoff(&bars[B_FULL + slot], (it / NSTAGE) & 1, 32);
        mbar_wait_backoff(&bars[B_PFULL + pb], (i >> 1) & 1, 32);
        umma::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {                            // k-step: keys [16j, 16j+16)
          const uint64_t da = umma::smem_desc(ring_a + slot * STAGE_BYTES + j * 2048, 16384, 1024);
          const uint64_t db = umma::smem_desc(p_a + pb * 16384 + j * 2048, 1024, 1024);
          umma::mma_ss(tmem + TM_O, da, db, id_pv, (i > 0 || j > 0) ? 1u : 0u);
        }
        umma::commit(&bars[B_EMPTY + slot]);
        umma::commit(&bars[B_PEMPTY + pb]);
      }
      umma::commit(&bars[B_OFULL]);
    }
  } else {
    // ===== softmax warps: lane of TMEM = key ==================================================================================
    const uint32_t lane_base = (uint32_t)(q4 * 32) << 16;
    const int col0 = cgp * CW;                                   // first of this thread's rows inside the block
    const int qbase = rb * (NROWS / G) + col0 / G;               // the query of that row
    // which of this thread's keys (one per tile) are attended: cached slots that are valid (free slots inside
    // [0, n_phys) are streamed and masked), the chunk's own keys that exist
    uint32_t vmask = 0;
    mbar_wait(&bars[B_LIDX], 0);
#pragma unroll
    for (int i = 0; i < MAX_TILES; ++i) {
      const int t = t0 + i, key = t * TKEYS + kl;
      const bool valid = i < T_ && (t < nct ? (key < n_phys && key < a.cap && lidx_s[i * TKEYS + kl] >= 0) : (t - nct) * TKEYS + kl < QL);
      vmask |= (valid ? 1u : 0u) << i;
    }

    // ---- K phase: logits -> TMEM, running row maxima, one-pass denominators ---------------------------------------------
    // The denominator sum_k exp(x_k - M) needs the global row maximum M, which is only known after the K phase.  The
    // K phase is HBM-bound (the ALUs idle), so every warp already sums exp(x - m0) against ITS OWN reference point
    // m0 = the maximum of the row over the warp's 32 keys of its first tile; the slices are rescaled by exp(m0 - M)
    // when they are combined (one cluster exchange instead of two, and no second sweep over the logits).  Each term
    // carries one more rounding than exp(x - M) would — the same order as the freedom of the summation order itself
    // (ATen's CUDA softmax sums in yet another order).  A row whose logits run away from m0 by more than 2^100 (or whose
    // first tile is fully masked and whose logits are tiny) is detected and the whole cluster takes the exact L pass.
    uint32_t rmax[CW2];
    float negm0[CW], Ssum[CW];
#pragma unroll
    for (int j = 0; j < CW2; ++j) rmax[j] = neg_inf2<T>();
#pragma unroll
    for (int j = 0; j < CW; ++j) { negm0[j] = 0.f; Ssum[j] = 0.f; }
    for (int i = 0; i < T_; ++i) {
      const int sb = i & 1;
      mbar_wait(&bars[B_SFULL + sb], (i >> 1) & 1);
      umma::fence_after_sync();
      uint32_t r[CW];
      TmemIO<CW>::ld(tmem + lane_base + TM_S + sb * NROWS + col0, r);
      umma::tmem_wait_ld();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb]);
      uint32_t w[CW2];
#pragma unroll
      for (int j = 0; j < CW2; ++j) {
        float x0 = __uint_as_float(r[2 * j]), x1 = __uint_as_float(r[2 * j + 1]);
        round2<T>(x0, x1);                                       // llama_patch.py:201: the matmul's result is a model-dtype tensor
        x0 = ARITH ? __fmul_rn(x0, a.scale_mul) : __fdiv_rn(x0, a.scale_div);    // :202
        x1 = ARITH ? __fmul_rn(x1, a.scale_mul) : __fdiv_rn(x1, a.scale_div);
        w[j] = pack2<T>(x0, x1);
      }
      if (!((vmask >> i) & 1u)) {
#pragma unroll
        for (int j = 0; j < CW2; ++j) w[j] = neg_inf2<T>();
      } else if (t0 + i >= nct) {                                // causal among the chunk's own keys (:210-215)
        const int jn = (t0 + i - nct) * TKEYS + kl;
#pragma unroll
        for (int j = 0; j < CW2; ++j) {
          const int q0 = qbase + (2 * j) / G, q1 = qbase + (2 * j + 1) / G;
          if (jn > q0) w[j] = (w[j] & 0xffff0000u) | (neg_inf2<T>() & 0xffffu);
          if (jn > q1) w[j] = (w[j] & 0x0000ffffu) | (neg_inf2<T>() & 0xffff0000u);
        }
      }
      TmemIO<CW2>::st(tmem + lane_base + TM_LOG + i * 32 + cgp * CW2, w);
      if (i == 0) {                                              // the warp's reference points
#pragma unroll
        for (int j = 0; j < CW2; ++j) {
          uint32_t m = w[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = max2<T>(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float2 f = Tr<T>::to_f2(m);
          negm0[2 * j] = f.x == -INFINITY ? 0.f : -f.x * LOG2E;          // kept pre-scaled: exp(x - m0) = 2^(x*log2e - m0*log2e)
          negm0[2 * j + 1] = f.y == -INFINITY ? 0.f : -f.y * LOG2E;
        }
      }
#pragma unroll
      for (int j = 0; j < CW2; ++j) {
        rmax[j] = max2<T>(rmax[j], w[j]);
        const float2 x = Tr<T>::to_f2(w[j]);
        Ssum[2 * j] += ex2_approx(fmaf(x.x, LOG2E, negm0[2 * j]));
        Ssum[2 * j + 1] += ex2_approx(fmaf(x.y, LOG2E, negm0[2 * j + 1]));
      }
    }
    umma::tmem_wait_st();
    stamp(2);

    // ---- row statistics: lanes -> warps -> cluster (one exchange) ------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < CW2; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax[j] = max2<T>(rmax[j], __shfl_xor_sync(0xffffffffu, rmax[j], o));
    }
#pragma unroll
    for (int j = 0; j < CW; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Ssum[j] += __shfl_xor_sync(0xffffffffu, Ssum[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < CW2; ++j) {
        const float2 f = Tr<T>::to_f2(rmax[j]);
        *reinterpret_cast<float2*>(&redmax[q4 * 64 + col0 + 2 * j]) = f;
        // usable iff the slice's logits stayed within e^+-80 of its reference point (an all-masked slice sums to 0)
        const float d0 = fmaf(f.x, LOG2E, negm0[2 * j]), d1 = fmaf(f.y, LOG2E, negm0[2 * j + 1]);      // in powers of two
        const bool ok0 = f.x == -INFINITY || (d0 < 100.f && d0 > -100.f), ok1 = f.y == -INFINITY || (d1 < 100.f && d1 > -100.f);
        redsum[q4 * 64 + col0 + 2 * j] = ok0 ? Ssum[2 * j] : __int_as_float(0x7fc00000);
        redsum[q4 * 64 + col0 + 2 * j + 1] = ok1 ? Ssum[2 * j + 1] : __int_as_float(0x7fc00000);
        redm0[q4 * 64 + col0 + 2 * j] = -negm0[2 * j];               // reference points travel in log2 units
        redm0[q4 * 64 + col0 + 2 * j + 1] = -negm0[2 * j + 1];
      }
    }
    named_bar_sync(1, NSOFT);
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // (arrived during setup: every peer's barriers exist)
    for (int i = tid; i < 64 * C; i += NSOFT) {                   // (row, peer): three remote stores + one remote arrive each
      const int row = i & 63, p = i >> 6;
      const float m = fmaxf(fmaxf(redmax[row], redmax[64 + row]), fmaxf(redmax[128 + row], redmax[192 + row]));
      const float z = fmaxf(fmaxf(redm0[row], redm0[64 + row]), fmaxf(redm0[128 + row], redm0[192 + row]));
      float sacc = 0.f;
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) sacc += redsum[qq * 64 + row] * ex2_approx(redm0[qq * 64 + row] - z);   // factors <= 1
      st_cluster_f32(map_to_rank(&xmax[rank * 64 + row], p), m);
      st_cluster_f32(map_to_rank(&xm0[rank * 64 + row], p), z);
      st_cluster_f32(map_to_rank(&xsum[rank * 64 + row], p), sacc);
      umma::mbar_arrive_remote(map_to_rank(&bars[B_XCH + 0], p));
    }
    umma::mbar_wait_cluster(&bars[B_XCH + 0], 0);
    if (tid < 64) {
      float m = xmax[tid], z = xm0[tid];
      for (int p = 1; p < C; ++p) { m = fmaxf(m, xmax[p * 64 + tid]); z = fmaxf(z, xm0[p * 64 + tid]); }
      float sacc = 0.f;
      for (int p = 0; p < C; ++p) sacc += xsum[p * 64 + tid] * ex2_approx(xm0[p * 64 + tid] - z);   // rank order on every CTA
      float s = sacc * ex2_approx(z - m * LOG2E);                 // z <= m*log2e: the reference points are maxima of subsets
      if (m == -INFINITY) s = 1.f;                                // a row with no key at all (padding): p = 0 whatever s
      if (!(s > 0.f) || !(s < INFINITY)) atomicOr(slow_flag, 1);  // NaN / 0 / inf: this cluster sums exactly (below)
      rowM[tid] = m;
      rowL[tid] = ARITH ? s : __fdiv_rn(1.0f, s);
      rowR[tid] = __frcp_rn(s);
    }
    named_bar_sync(1, NSOFT);
    stamp(3);
    float negM[CW];
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(&rowM[col0 + j]);
      negM[j] = -v.x; negM[j + 1] = -v.y; negM[j + 2] = -v.z; negM[j + 3] = -v.w;
    }
    float L[CW];
    if (*slow_flag) {
    // ---- exact L pass (slow path): sum of exp(x - max) over the parked logits, second exchange -------------------------------
#pragma unroll
    for (int j = 0; j < CW; ++j) L[j] = 0.f;
    for (int i = 0; i < T_; ++i) {
      uint32_t w[CW2];
      TmemIO<CW2>::ld(tmem + lane_base + TM_LOG + i * 32 + cgp * CW2, w);
      umma::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < CW2; ++j) {
        const float2 x = Tr<T>::to_f2(w[j]);
        L[2 * j] += expf(x.x + negM[2 * j]);
        L[2 * j + 1] += expf(x.y + negM[2 * j + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < CW; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) L[j] += __shfl_xor_sync(0xffffffffu, L[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < CW; ++j) redsum[q4 * 64 + col0 + j] = L[j];
    }
    named_bar_sync(1, NSOFT);
    stamp(4);
    for (int i = tid; i < 64 * C; i += NSOFT) {
      const int row = i & 63, p = i >> 6;
      const float s = ((redsum[row] + redsum[64 + row]) + redsum[128 + row]) + redsum[192 + row];
      st_cluster_f32(map_to_rank(&xsum2[rank * 64 + row], p), s);
      umma::mbar_arrive_remote(map_to_rank(&bars[B_XCH + 1], p));
    }
    umma::mbar_wait_cluster(&bars[B_XCH + 1], 0);
    if (tid < 64) {
      float s = xsum2[tid];
      for (int p = 1; p < C; ++p) s += xsum2[p * 64 + tid];       // rank order on every CTA: identical denominators
      if (s == 0.f) s = 1.f;
      rowL[tid] = ARITH ? s : __fdiv_rn(1.0f, s);
      rowR[tid] = __frcp_rn(s);
    }
    named_bar_sync(1, NSOFT);
    stamp(5);
    }
    float Rc[ARITH ? CW : 1];
#pragma unroll
    for (int j = 0; j < CW; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(&rowL[col0 + j]);
      L[j] = v.x; L[j + 1] = v.y; L[j + 2] = v.z; L[j + 3] = v.w;
      if (ARITH) {
        const float4 u = *reinterpret_cast<const float4*>(&rowR[col0 + j]);
        Rc[j] = u.x; Rc[j + 1] = u.y; Rc[j + 2] = u.z; Rc[j + 3] = u.w;
      }
    }

    // ---- V phase: probabilities, column statistics, P^T tiles ---------------------------------------------------------------
    const bool want_stats = a.st.accumulate != 0;
    const bool tova = a.st.policy == EKV_POLICY_TOVA;
    const float inv_g = 1.0f / (float)G;
    float2* cg = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(a.scratch) + pl.off_cpart) +
                 ((size_t)unit * (pl.cparts * pl.RB) + pl.cparts * rb + cgp) * pl.NEpad;
    for (int i = 0; i < T_; ++i) {
      const int t = t0 + i, pb = i & 1;
      uint32_t w[CW2];
      TmemIO<CW2>::ld(tmem + lane_base + TM_LOG + i * 32 + cgp * CW2, w);
      umma::tmem_wait_ld();
      float pr[CW];                                               // the key's probabilities for this thread's rows
#pragma unroll
      for (int j = 0; j < CW2; ++j) {
        const float2 x = Tr<T>::to_f2(w[j]);
        const float e0 = expf(x.x + negM[2 * j]), e1 = expf(x.y + negM[2 * j + 1]);
        float p0 = ARITH ? div_rn_by(e0, L[2 * j], Rc[ARITH ? 2 * j : 0]) : __fmul_rn(e0, L[2 * j]);              // llama_patch.py:218
        float p1 = ARITH ? div_rn_by(e1, L[2 * j + 1], Rc[ARITH ? 2 * j + 1 : 0]) : __fmul_rn(e1, L[2 * j + 1]);
        round2<T>(p0, p1);                                        // :219 .to(dtype)
        pr[2 * j] = p0; pr[2 * j + 1] = p1;
        w[j] = pack2<T>(p0, p1);
      }
      if (want_stats) {
        // GQA fold in the model dtype (process_for_mqa_gqa, easykv.py:188-196), then each query's share of the
        // chunk's row sums of p and model-dtype(p^2) (:450-451); tova keeps the last query only (:454)
        float cs = 0.f, csq = 0.f;
#pragma unroll
        for (int u = 0; u < CW / G; ++u) {                        // one query: its G heads are adjacent rows
          float fsum = pr[u * G];
#pragma unroll
          for (int g = 1; g < G; ++g) fsum += pr[u * G + g];
          const int qi = qbase + u;
          if (qi < QL && (!tova || qi == QL - 1)) {
            const float pf = G == 1 ? fsum : Tr<T>::round_f(__fmul_rn(fsum, inv_g));
            cs += pf;
            csq += Tr<T>::round_f(__fmul_rn(pf, pf));
          }
        }
        const int e = t < nct ? t * TKEYS + kl : n_phys + (t - nct) * TKEYS + kl;
        if (t < nct ? e < n_phys : e < n_phys + QL) cg[e] = make_float2(cs, csq);
      }
      if (i >= 2) mbar_wait(&bars[B_PEMPTY + pb], ((i >> 1) - 1) & 1);
      unsigned char* prow = Ps + pb * 16384;
#pragma unroll
      for (int c = 0; c < CW / 8; ++c)
        *reinterpret_cast<uint4*>(prow + umma::swz128(kl, cgp * (CW / 8) + c)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL + pb]);
    }
    stamp(6);

    // ---- epilogue: this CTA's partial O^T -> scratch ---------------------------------------------------------------------------
    float* op = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.scratch) + pl.off_opart) +
                (((size_t)unit * pl.splits + rank) * pl.Rpad + rb * NROWS + col0) * D + kl;     // kl = output dim here
    if (T_ > 0) {
      mbar_wait(&bars[B_OFULL], 0);
      umma::fence_after_sync();
      uint32_t r[CW];
      TmemIO<CW>::ld(tmem + lane_base + TM_O + col0, r);
      umma::tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < CW; ++j) op[(size_t)j * D] = __uint_as_float(r[j]);
    } else {
#pragma unroll
      for (int j = 0; j < CW; ++j) op[(size_t)j * D] = 0.f;
    }
    stamp(7);
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == NSW + 1) umma::tmem_dealloc(tmem, TM_COLS);
}

// ---- plan ----------------------------------------------------------------------------------------------------------------
// How many clusters of `c` CTAs of this kernel the device can hold at once (one CTA per SM; a cluster must fit one GPC,
// so the answer is NOT sms / c: 8-CTA clusters pack 13-16 per B200, 6-CTA clusters ~24, pairs 74).  Asked of the driver
// once per size; a fixed table stands in where no device is present (ekv_scratch_bytes is host-only).
static int concurrent_clusters(int c, int sms) {
  static int cached[cu::MAX_CLUSTER + 1] = {0};
  if (c < 1 || c > cu::MAX_CLUSTER) return 1;
  if (cached[c]) return cached[c];
  static const int fallback[cu::MAX_CLUSTER + 1] = {0, 148, 74, 48, 33, 26, 24, 16, 14};
  int n = 0, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) {
    auto kern = chunk_umma_kernel<__half, 1, true, 4>;
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, cu::SMEM_ALLOC) == cudaSuccess) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3((unsigned)(c * 4096), 1, 1);
      cfg.blockDim = dim3(cu::nthreads(4), 1, 1);
      cfg.dynamicSmemBytes = cu::SMEM_ALLOC;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = (unsigned)c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) n = 0;
    }
    cudaGetLastError();
  }
  if (n <= 0) n = fallback[c] * sms / 148;
  if (n < 1) n = 1;
  cached[c] = n;
  return n;
}

bool make_umma_plan(int B, int Hkv, int G, int q_len, int n_phys, int sms, ChunkPlan& out) {
  using namespace cu;
  ChunkPlan p = make_chunk_plan(B, Hkv, G, q_len, n_phys);       // R, RB, Rpad, NE, NEpad and the scratch carve-up
  p.nct = (n_phys + TKEYS - 1) / TKEYS;
  p.nnt = (q_len + TKEYS - 1) / TKEYS;
  const int nt = p.nct + p.nnt;
  const int cmin = (nt + MAX_TILES - 1) / MAX_TILES;
  if (cmin > MAX_CLUSTER) return false;
  const long long groups = (long long)B * Hkv * p.RB;
  const int force = umma_force_cluster();
  int best = 0;
  double best_cost = 0;
  for (int c = cmin; c <= MAX_CLUSTER; ++c) {
    if (force && c != force) continue;
    const int tps = (nt + c - 1) / c;
    if (tps > MAX_TILES) continue;
    const double waves = (double)((groups + concurrent_clusters(c, sms) - 1) / concurrent_clusters(c, sms));
    // per-CTA fixed work (setup, two exchanges, epilogue) ~ 4 tile-times; more for larger clusters (exchange skew)
    const double cost = waves * (tps + 3.5 + 0.25 * c);
    if (!best || cost < best_cost - 1e-9) { best = c; best_cost = cost; }
  }
  if (!best) return false;
  p.splits = best;
  p.tps = (nt + best - 1) / best;
  // re-carve the scratch for `splits` partial outputs (the row statistics block of the two-pass path is unused)
  const long long U = (long long)B * Hkv;
  long long o = 0;
  p.off_stats = 0;
  p.off_opart = o; o += U * p.splits * p.Rpad * D * 4;
  o = (o + 255) / 256 * 256;
  p.cparts = 4;
  p.off_cpart = o; o += U * (p.cparts * p.RB) * p.NEpad * 2 * 4;
  o = (o + 255) / 256 * 256;
  p.off_klj = o; o += U * p.NEpad * 4;
  p.off_ka = o; o += U * p.NEpad * 4;
  p.off_kb = o; o += U * p.NEpad * 4;
  p.off_kf = o; o += U * p.NEpad;
  p.bytes = (o + 255) / 256 * 256;
  out = p;
  return true;
}

// ---- launch --------------------------------------------------------------------------------------------------------------
static int device_sms() {
  static thread_local int sm_count[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (!sm_count[dev] && cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count[dev] = 148;
  return sm_count[dev];
}
int umma_sm_count() { return device_sms(); }
// softmax column groups per CTA: chunk_variant 3 = 2 groups (8 warps x 32 rows per thread), otherwise 4 (16 warps x 16)
static int umma_nsplit() { return chunk_variant() == 3 ? 2 : 4; }

template <typename T, int G, bool ARITH, int NSPLIT>
static int launch_umma_k(const KernelArgs& a, const ChunkPlan& pl, const CUtensorMap* maps, cudaStream_t stream) {
  using namespace cu;
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  cudaError_t err;
  if (!configured[dev]) {
    err = cudaFuncSetAttribute(chunk_umma_kernel<T, G, ARITH, NSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_ALLOC);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(chunk_umma)", err);
    configured[dev] = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.Hkv * pl.RB * pl.splits), 1, 1);
  cfg.blockDim = dim3(nthreads(NSPLIT), 1, 1);
  cfg.dynamicSmemBytes = SMEM_ALLOC;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)pl.splits;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the kernel waits (pdl_wait) before its first global access
  attr[1].val.programmaticStreamSerializationAllowed = pdl_allowed();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  err = cudaLaunchKernelEx(&cfg, chunk_umma_kernel<T, G, ARITH, NSPLIT>, a, pl, maps[0], maps[1], maps[2], maps[3]);
  if (err != cudaSuccess) return set_cuda_error("chunk_umma_kernel launch", err);
  count_launch();
  return EKV_OK;
}

template <typename T, int G> static int launch_umma_tg(const KernelArgs& a, cudaStream_t stream) {
  ChunkPlan pl;
  if (!make_umma_plan(a.B, a.Hkv, G, a.q_len, a.n_phys, device_sms(), pl)) return EKV_ERR_UNSUPPORTED;
  // tensor maps over the caller's buffers viewed as [rows][128]: the cache (all units' slots) and the chunk's new rows
  CUtensorMap maps[4];
  const unsigned long long rows_c = (unsigned long long)a.B * a.Hkv * a.cap, rows_n = (unsigned long long)a.B * a.Hkv * a.q_len;
  int rc = make_tensor_map_rows128(&maps[0], a.K, rows_c, cu::TKEYS, a.dtype);
  if (!rc) rc = make_tensor_map_rows128(&maps[1], a.V, rows_c, cu::TKEYS, a.dtype);
  if (!rc) rc = make_tensor_map_rows128(&maps[2], a.k_new, rows_n, cu::TKEYS, a.dtype);
  if (!rc) rc = make_tensor_map_rows128(&maps[3], a.v_new, rows_n, cu::TKEYS, a.dtype);
  if (rc) return rc;
  pl.cparts = umma_nsplit();
  if (pl.cparts == 4)
    rc = a.st.arith ? launch_umma_k<T, G, true, 4>(a, pl, maps, stream) : launch_umma_k<T, G, false, 4>(a, pl, maps, stream);
  else {
    pl.cparts = 2;
    rc = a.st.arith ? launch_umma_k<T, G, true, 2>(a, pl, maps, stream) : launch_umma_k<T, G, false, 2>(a, pl, maps, stream);
  }
  if (rc) return rc;
  return launch_chunk_finish(a, pl, stream);
}

template <typename T> static int launch_umma_t(const KernelArgs& a, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_umma_tg<T, 1>(a, stream);
    case 2: return launch_umma_tg<T, 2>(a, stream);
    case 4: return launch_umma_tg<T, 4>(a, stream);
    case 8: return launch_umma_tg<T, 8>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

// 16-bit dtypes, head_dim 128, needs scratch; EKV_ERR_UNSUPPORTED when the key range does not fit 8 CTAs x 10 tiles
// (more than 12 288 cached slots + the chunk): the caller then takes the two-pass mma.sync path.
int launch_chunk_umma(const KernelArgs& a, cudaStream_t stream) {
  if (a.d != cu::D || !a.scratch || a.q_len < 1 || (a.cap & 3)) return EKV_ERR_UNSUPPORTED;   // (bulk copy of the slot map: 16-byte rows)
  // the tensor maps address rows of 256 bytes from 16-byte aligned bases
  if ((reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V) | reinterpret_cast<uintptr_t>(a.k_new) |
       reinterpret_cast<uintptr_t>(a.v_new)) & 15) return EKV_ERR_UNSUPPORTED;
  switch (a.dtype) {
    case EKV_F16: return launch_umma_t<__half>(a, stream);
    case EKV_BF16: return launch_umma_t<__nv_bfloat16>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

}  // namespace ekv
