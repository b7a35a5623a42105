// Fused decode step (q_len == 1) for grouped-query layouts on the Blackwell tensor cores.
//
// With g = H / Hkv query heads per kv head a decode step needs g mixed-precision FMAs per loaded K element (and again
// per V element): beyond g = 2 the FP32 pipes cannot keep up with HBM, and the round-1 kernels (mma.sync, 8-CTA
// clusters with seven cluster barriers per unit) reached 0.33 - 0.55 of the roofline on the Mistral / Llama-2-70B
// layouts.  Here both contractions run on tcgen05.mma with the g heads padded to a 16-row operand, transposed as in
// the chunk kernel (ekv_chunk_umma.cu) so that a tensor-memory lane is a KEY:
//     S^T [128 keys x 16] = K_tile [128 x 128] . Q^T          A = K tile (K-major, tensor-map TMA), B = Q (K-major)
//     O^T [128 dims x 16] += V_tile^T . P^T                    A = V tile (MN-major, TMA), B = P^T (MN-major, no swizzle)
// A softmax thread owns one key per tile and sees its logit for all g heads: the GQA fold, the policy accumulate
// (S += p, SQ += p^2, counters) and the key's selection keys are thread-local.  16 rows x 2 bytes of logits per key
// park in tensor memory at g/2 columns per tile, so ONE CTA holds up to 34 tiles (4352 keys) and a unit needs a cluster
// of only 1 - 4 CTAs (pairs pack the chip exactly); the only cluster-wide steps are one row-statistics exchange, the
// partial-output gather and the victim walk — all over distributed shared memory with remote mbarrier arrives, none a
// hardware cluster barrier.
//
// Phases per CTA (rank r of C, tiles [t0, t1) of the unit's cached keys; rank 0 also owns the appended token's key):
//   K phase   K tiles -> S^T in TMEM (double-buffered) -> logits at the reference's rounding points
//             (llama_patch.py:201-202), masked, parked in TMEM; running maxima; one-pass denominators against per-warp
//             reference points (exact fallback) exactly as in the chunk kernel;
//   exchange  (max, reference point, sum) per head, all-to-all; every CTA derives identical M and L;
//   V phase   p = dtype(exp(x - M) / L) (:218-219), P^T tile -> shared memory, O^T += V^T P^T; per key: GQA fold
//             (easykv.py:188-196), accumulate + counter + selection keys (entry_update, ekv_select.cuh), state written back;
//   output    partial O^T gathered at rank 0 over DSMEM, summed in rank order, + p_new * v_new, written as the model dtype;
//   select    the victim: candidates in (mean, std, logical index) order, four per round, accepted when the std rank is
//             below k_feasible (easykv.py:322-324, :722-724); h2o_head / tova: one window argmin (:311, :335); recency:
//             positional (:343-347).  Slot-map renumbering per slice, new row appended by rank 0.
//
// Replaces the same reference lines as ekv_decode.cu / ekv_decode_cluster.cu.
#include "ekv_decode_common.cuh"
#include "ekv_bucket.cuh"
#include "ekv_mma.cuh"
#include "ekv_umma.cuh"

namespace ekv {

namespace du {
constexpr int D = 128;
constexpr int TKEYS = 128;               // keys per tile = MMA M
constexpr int NR = 16;                   // MMA N: the g heads padded to 16 rows
constexpr int NSOFT = 128;               // 4 softmax warps: one thread per TMEM lane
constexpr int NHS = 128;                 // helper warps, one SET of 4: the per-key policy work of the V phase (state update, IEEE
constexpr int NHELP = 2 * NHS;           // div / sqrt for the roco keys, select histograms) is a long dependent chain per key — two sets
                                         // take alternate tiles, so each has two tile-times for its chain; set 0 joins the tail
constexpr int NTAIL = NHELP;             // the tail (select, renumbering) runs on BOTH helper sets, concurrently with the softmax warps' output gather
constexpr int NT = NSOFT + 64 + NHELP;   // softmax | TMA producer warp | MMA warp | helper set 0 | helper set 1
constexpr int STAGE_BYTES = 32768;
constexpr int MAX_STAGE = 5;             // ring depth: whatever shared memory is left after the per-entry arrays (3 .. 5)
constexpr int MAX_CLUSTER = 4;
constexpr int NCAND = 2;                 // candidates per walk round
constexpr int NTW = NTAIL / 32;          // warps taking part in the tail
static_assert(NTAIL == 256, "one radix bin per tail thread");
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t TM_S = 0, TM_O = 32, TM_LOG = 48, TM_COLS = 512;      // S^T 2 x 16 | O^T 16 | parked logits
// shared memory (after 1024-byte alignment): the operands, then the fixed small blocks, then what depends on the plan
constexpr int OFF_Q = 0;                                  // 2 x [16 rows][128 B]
constexpr int OFF_P = OFF_Q + 4096;                       // 2 x [128 keys][32 B]
constexpr int OFF_BAR = OFF_P + 2 * 4096;
constexpr int OFF_TMEM = OFF_BAR + 40 * 8;
constexpr int OFF_RED = OFF_TMEM + 16;                    // [3][4 warps][8 heads]: max | reference point | sum
constexpr int OFF_XST = OFF_RED + 3 * 4 * 8 * 4;          // [MAX_CLUSTER][3][8] exchanged statistics
constexpr int OFF_XST2 = OFF_XST + MAX_CLUSTER * 3 * 8 * 4;   // [MAX_CLUSTER][8] exact sums (slow path)
constexpr int OFF_ROW = OFF_XST2 + MAX_CLUSTER * 8 * 4;   // M[8] | L[8] | rcp[8] | xnew[8] | pnew[8] | flags
constexpr int OFF_CNT = OFF_ROW + 6 * 8 * 4;              // [2][MAX_CLUSTER][NCAND] ints
constexpr int OFF_WIN = OFF_CNT + 2 * MAX_CLUSTER * NCAND * 4;           // NCAND winner tuples + warp counts [NTW][NCAND] + radix misc
constexpr int OFF_HIST = OFF_WIN + NCAND * 16 + NTW * NCAND * 4 + 32;    // radix select: local 256-bin histogram
constexpr int OFF_CAND = (OFF_HIST + 1024 + 15) / 16 * 16;               // candidates [2 parities][MAX_CLUSTER][NTW][NCAND] 128-bit tuples,
constexpr int CAND_BYTES = 2 * MAX_CLUSTER * 256 * 4;                     // reused for the radix histograms [2][MAX_CLUSTER][256] u32
static_assert(2 * MAX_CLUSTER * NTW * NCAND * 16 <= CAND_BYTES, "candidate exchange fits");
constexpr int OFF_FS = OFF_CAND + CAND_BYTES;                             // [2][128] folded probabilities: softmax warps -> helper warps
constexpr int OFF_RING = (OFF_FS + 2 * TKEYS * 4 + 1023) / 1024 * 1024;
// barriers
constexpr int B_FULL = 0, B_EMPTY = MAX_STAGE, B_SFULL = 2 * MAX_STAGE, B_SEMPTY = B_SFULL + 2, B_PFULL = B_SEMPTY + 2,
              B_PEMPTY = B_PFULL + 2, B_OFULL = B_PEMPTY + 2, B_LIDX = B_OFULL + 1, B_XST = B_LIDX + 1, B_XST2 = B_XST + 1,
              B_XOUT = B_XST2 + 1, B_XHIST = B_XOUT + 1, B_HRDY = B_XHIST + 2, B_XG = B_HRDY + 1, B_FSFULL = B_XG + 1, B_FSEMPTY = B_FSFULL + 2, B_PNEW = B_FSEMPTY + 2;
static_assert(B_PNEW + 1 <= 40, "barrier block");
}  // namespace du

// What depends on the plan: ring depth, the gathered partial outputs (rank 0, clusters only), one logical index and one
// 64-bit selection key per entry of the CTA's slice.
// the tail's select scratch (bucket histograms for roco; the gather / reduction buffers of every single-victim select):
// carved only for steps that select a victim by score — no-policy steps (generation over a retained cache) keep the
// shared memory for one more ring stage
__host__ __device__ inline bool du_select_scratch(const ekv_step& st) {
  return st.evict > 0 && (st.policy == EKV_POLICY_ROCO || st.policy == EKV_POLICY_H2O || st.policy == EKV_POLICY_TOVA);
}

struct DuSmem {
  int nstage, off_obuf, off_lj, off_kk, off_bkt, total;
  // `bucket`: the step selects a roco victim — the bucket select's histograms / lists (ekv_bucket.cuh) are carved as well
  __host__ __device__ DuSmem(int tps, int C, bool bucket) {
    using namespace du;
    const int nent = tps * TKEYS + 8;
    const int obuf = C > 1 ? C * 8 * D * 4 : 0;
    const int bkt = bucket ? (BucketScratch::bytes(MAX_CLUSTER) + 15) / 16 * 16 : 0;
    const int arrays = obuf + (nent * 4 + 15) / 16 * 16 + nent * 8 + bkt;
    int ns = (227 * 1024 - 1024 - OFF_RING - arrays) / STAGE_BYTES;
    nstage = ns > MAX_STAGE ? MAX_STAGE : ns;
    int o = OFF_RING + (nstage > 0 ? nstage : 0) * STAGE_BYTES;
    off_obuf = o; o += obuf;
    off_lj = o; o += (nent * 4 + 15) / 16 * 16;
    off_kk = o; o += nent * 8;
    off_bkt = o; o += bkt;
    total = o + 1024;
  }
};

struct DecodeUmmaPlan {
  int C, nct, tps;     // cluster size, 128-key tiles over the cached slots, tiles per CTA (upper bound)
};

template <typename T, int G, bool ARITH>
__global__ void __launch_bounds__(du::NT, 1)
decode_umma_kernel(const KernelArgs a, const DecodeUmmaPlan pl, const __grid_constant__ CUtensorMap mapK,
                   const __grid_constant__ CUtensorMap mapV) {
  using namespace du;
  constexpr int GP = G < 2 ? 2 : G;                              // columns read per key (packed pairs)
  constexpr int GW = GP / 2;                                     // ... as 32-bit words = parked TMEM columns per tile
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const bool want_bucket = du_select_scratch(a.st);            // (uniform over the launch; a budget gate only skips it per unit)
  const DuSmem SL(pl.tps, pl.C, want_bucket);
  const int NSTAGE = SL.nstage;
  unsigned char* ring = smem + OFF_RING;
  unsigned char* Qs = smem + OFF_Q;
  unsigned char* Ps = smem + OFF_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  float* red = reinterpret_cast<float*>(smem + OFF_RED);
  float* xst = reinterpret_cast<float*>(smem + OFF_XST);
  float* xst2 = reinterpret_cast<float*>(smem + OFF_XST2);
  float* rowM = reinterpret_cast<float*>(smem + OFF_ROW);
  float* rowL = rowM + 8;
  float* rowR = rowL + 8;
  float* xnew_s = rowR + 8;
  float* pnew_s = xnew_s + 8;
  int* flags = reinterpret_cast<int*>(pnew_s + 8);               // [0] slow path, [1] walk decision, [2..] scratch
  int* hmisc = reinterpret_cast<int*>(smem + OFF_WIN);           // radix select: digit, count below, count equal
  // radix select: every CTA's 256-bin histogram, double-buffered [2][MAX_CLUSTER][256] — the candidate exchange buffer,
  // idle once the walk has given up
  uint32_t* xhist = reinterpret_cast<uint32_t*>(smem + OFF_CAND);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + OFF_HIST);
  BucketScratch bs;
  bs.carve(smem + SL.off_bkt, MAX_CLUSTER);                     // (only touched when want_bucket)
  float* fs = reinterpret_cast<float*>(smem + OFF_FS);
  int32_t* lj = reinterpret_cast<int32_t*>(smem + SL.off_lj);
  unsigned long long* kk = reinterpret_cast<unsigned long long*>(smem + SL.off_kk);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = pl.C;
  const int rank = blockIdx.x % C;
  const int unit = blockIdx.x / C;
  pdl_trigger();                                                 // the next launch may be placed while this one runs (it waits itself)
  const int n_phys = a.n_phys, nct = pl.nct;
  const int t0 = (int)((long long)rank * nct / C), t1 = (int)((long long)(rank + 1) * nct / C);
  const int T_ = t1 - t0;
  const int first = t0 * TKEYS;                                  // first physical slot of this CTA's slice
  const int NEl = T_ * TKEYS + (rank == 0 ? 1 : 0);              // entries: one per slot of the slice (+ the appended token)
  const int e_new = T_ * TKEYS;                                  // rank 0: the appended token's entry

  auto issue_tile = [&](int it, uint64_t pol) {                  // `it`-th tile of the K-then-V stream
    const int slot = it % NSTAGE;
    const int i = it < T_ ? it : it - T_;
    const CUtensorMap* map = it < T_ ? &mapK : &mapV;
    const int row = unit * a.cap + (t0 + i) * TKEYS;
    unsigned char* dst = ring + (size_t)slot * STAGE_BYTES;
    mbar_arrive_expect_tx(&bars[B_FULL + slot], STAGE_BYTES);
    umma::tma_load_2d(dst, map, 0, row, &bars[B_FULL + slot], pol);
    umma::tma_load_2d(dst + 16384, map, 64, row, &bars[B_FULL + slot], pol);
  };
  uint64_t tma_pol = 0;
  // ---- setup -----------------------------------------------------------------------------------------------------------
  if (tid == NSOFT) {
    for (int s = 0; s < MAX_STAGE; ++s) { mbar_init(&bars[B_FULL + s], 1); mbar_init(&bars[B_EMPTY + s], 1); }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars[B_SFULL + s], 1); mbar_init(&bars[B_SEMPTY + s], NSOFT / 32);
      mbar_init(&bars[B_PFULL + s], NSOFT / 32); mbar_init(&bars[B_PEMPTY + s], 1);
      mbar_init(&bars[B_XHIST + s], NTAIL * C);
    }
    mbar_init(&bars[B_HRDY], C);
    mbar_init(&bars[B_XG], NTW * C);
    for (int s = 0; s < 2; ++s) { mbar_init(&bars[B_FSFULL + s], NSOFT / 32); mbar_init(&bars[B_FSEMPTY + s], NHS / 32); }
    mbar_init(&bars[B_PNEW], 1);
    mbar_init(&bars[B_OFULL], 1);
    mbar_init(&bars[B_LIDX], 1);
    mbar_init(&bars[B_XST], 8 * C);
    mbar_init(&bars[B_XST2], 8 * C);
    mbar_init(&bars[B_XOUT], NSOFT * (C - 1) + 1);               // rank 0: every peer thread + one local arrive
    flags[0] = 0; flags[1] = 0;
    mbar_fence_init();
    pdl_wait();                                                  // the previous kernel of the stream has completed: global memory may be touched
    {
      // the slice of the slot map ahead of the tiles (one bulk copy): ints [first, first + cnt) of the unit
      int cnt = T_ * TKEYS;
      if (cnt > a.cap - first) cnt = (a.cap - first) & ~3;
      if (cnt > 0) {
        mbar_arrive_expect_tx(&bars[B_LIDX], (uint32_t)cnt * 4u);
        tma_bulk_g2s(lj, a.lidx + (size_t)unit * a.cap + first, (uint32_t)cnt * 4u, &bars[B_LIDX], l2_policy_evict_first());
      } else {
        mbar_arrive(&bars[B_LIDX]);
      }
    }
    tma_pol = l2_policy_evict_first();
    for (int it = 0; it < NSTAGE && it < 2 * T_; ++it) issue_tile(it, tma_pol);
  }
  if (warp == NSOFT / 32 + 1) {
    umma::tmem_alloc(tmem_slot, TM_COLS);
    pdl_wait();
    // Q as the K-major B operand: rows g < G are the group's query heads, rows >= G zero; P^T buffers zeroed once
    // (the softmax warps only ever write the rows < GP of a key)
    const T* qg = reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D;
    for (int i = lane; i < NR * 16; i += 32) {
      const int r = i >> 4, c = i & 15;
      unsigned char* dst = Qs + (c >> 3) * 2048 + umma::swz128(r, c & 7);
      if (r < G) cp_async16(dst, reinterpret_cast<const uint4*>(qg + (size_t)r * D) + c);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
    }
    cp_async_commit();
    for (int i = lane; i < 2 * 4096 / 16; i += 32) reinterpret_cast<uint4*>(Ps)[i] = make_uint4(0, 0, 0, 0);
  }
  if (want_bucket && a.st.policy == EKV_POLICY_ROCO && warp >= NSOFT / 32 + 2) bs.clear(tid - NSOFT - 64, NHELP);  // the helper warps zero the select histograms
  pdl_wait();                                                    // (every thread, before its first global access; the set-up above is on-chip)
  umma::fence_before_sync();
  __syncthreads();                                               // barriers initialised, TMEM base published
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");      // waited for right before the first remote access
  umma::fence_after_sync();
  const uint32_t tmem = *tmem_slot;
  // ragged batches: this sequence's own count of valid slots; it evicts only past the budget gate (easykv.py:303)
  const int nb = a.seq_n_before ? a.seq_n_before[unit / a.Hkv] : a.n_before;
  ekv_step stu = a.st;
  if (stu.budget_gate > 0 && nb + 1 - stu.score_offset <= stu.budget_gate) stu.evict = 0;
  const int lsh = bk::lidx_shift(nb + 1);                       // logical indices -> histogram buckets

  if (warp == NSOFT / 32) {
    // ===== TMA producer ==================================================================================================
    if (lane == 0) {
      for (int it = NSTAGE; it < 2 * T_; ++it) {
        mbar_wait_backoff(&bars[B_EMPTY + it % NSTAGE], (it / NSTAGE - 1) & 1, 128);
        issue_tile(it, tma_pol);
      }
    }
  } else if (warp == NSOFT / 32 + 1) {
    // ===== MMA issuer ====================================================================================================
    cp_async_wait<0>();
    umma::fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const uint32_t id_qk = umma::instr_desc<T>(TKEYS, NR, false, false);
      const uint32_t id_pv = umma::instr_desc<T>(D, NR, true, true);
      const uint32_t ring_a = smem_u32(ring), q_a = smem_u32(Qs), p_a = smem_u32(Ps);
      int it = 0;
      for (int i = 0; i < T_; ++i, ++it) {                       // S^T(tile) = K_tile . Q^T
        const int slot = it % NSTAGE, sb = i & 1;
        mbar_wait_backoff(&bars[B_FULL + slot], (it / NSTAGE) & 1, 32);
        if (i >= 2) mbar_wait_backoff(&bars[B_SEMPTY + sb], ((i >> 1) - 1) & 1, 32);
        umma::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint64_t da = umma::smem_desc(ring_a + slot * STAGE_BYTES + (j >> 2) * 16384 + (j & 3) * 32, 16, 1024);
          const uint64_t db = umma::smem_desc(q_a + (j >> 2) * 2048 + (j & 3) * 32, 16, 1024);
          umma::mma_ss(tmem + TM_S + sb * NR, da, db, id_qk, j > 0);
        }
        umma::commit(&bars[B_EMPTY + slot]);
        umma::commit(&bars[B_SFULL + sb]);
      }
      for (int i = 0; i < T_; ++i, ++it) {                       // O^T += V_tile^T . P^T(tile)
        const int slot = it % NSTAGE, pb = i & 1;
        mbar_wait_backoff(&bars[B_FULL + slot], (it / NSTAGE) & 1, 32);
        mbar_wait_backoff(&bars[B_PFULL + pb], (i >> 1) & 1, 32);
        umma::fence_after_sync();
#pragma unroll
        for (int j = 0; j < 8; ++j) {                            // keys [16j, 16j+16): two 8-key groups of the P^T operand
          const uint64_t da = umma::smem_desc(ring_a + slot * STAGE_BYTES + j * 2048, 16384, 1024);
          const uint64_t db = umma::smem_desc_noswizzle(p_a + pb * 4096 + j * 512, 256, 128);
          umma::mma_ss(tmem + TM_O, da, db, id_pv, (i > 0 || j > 0) ? 1u : 0u);
        }
        umma::commit(&bars[B_EMPTY + slot]);
        umma::commit(&bars[B_PEMPTY + pb]);
      }
      umma::commit(&bars[B_OFULL]);
    }
  } else if (warp < NSOFT / 32) {
    // ===== softmax warps: TMEM lane = key =======================================================================================
    const int kl = warp * 32 + lane;                             // key inside a tile; output dim in the epilogue
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    unsigned long long* tl = a.timeline ? a.timeline + (size_t)blockIdx.x * 16 : nullptr;     // profiling hook
    auto stamp = [&](int i) { if (tl && tid == 0) tl[i] = global_ns(); };
    stamp(0);

    auto finish_logit2 = [&](float x0, float x1) -> uint32_t {
      round2<T>(x0, x1);                                         // llama_patch.py:201
      x0 = ARITH ? __fmul_rn(x0, a.scale_mul) : __fdiv_rn(x0, a.scale_div);    // :202
      x1 = ARITH ? __fmul_rn(x1, a.scale_mul) : __fdiv_rn(x1, a.scale_div);
      return pack2<T>(x0, x1);
    };
    mbar_wait(&bars[B_LIDX], 0);                                 // the slot-map slice is in lj[0, T_*128)
    // slots of the last tile beyond n_phys (or beyond the copied range) are not entries
    for (int e = tid; e < T_ * TKEYS; e += NSOFT)
      if (first + e >= n_phys) lj[e] = -1;
    if (rank == 0 && tid == 0) lj[e_new] = nb;
    named_bar_sync(1, NSOFT);

    // the appended token's own key (the reference attends it, llama_patch.py:193-196): rank 0, warp 0, CUDA cores —
    // ahead of the K phase, while the first tiles are still in flight
    if (rank == 0 && warp == 0) {
      const T* qg = reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D;
      const T* kn = reinterpret_cast<const T*>(a.k_new) + (size_t)unit * D;
      float kx[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) kx[c] = Tr<T>::to_f(kn[lane * 4 + c]);
#pragma unroll
      for (int g = 0; g < G; ++g) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < 4; ++c) acc = fmaf(Tr<T>::to_f(qg[g * D + lane * 4 + c]), kx[c], acc);
        acc = warp_sum(acc);
        float x = Tr<T>::round_f(acc);
        x = Tr<T>::round_f(ARITH ? __fmul_rn(x, a.scale_mul) : __fdiv_rn(x, a.scale_div));
        if (lane == 0) xnew_s[g] = x;
      }
    }

    // ---- K phase -----------------------------------------------------------------------------------------------------------
    uint32_t rmax[GW];
    float negm0[GP], Ssum[GP];
#pragma unroll
    for (int j = 0; j < GW; ++j) rmax[j] = neg_inf2<T>();
#pragma unroll
    for (int j = 0; j < GP; ++j) { negm0[j] = 0.f; Ssum[j] = 0.f; }
    for (int i = 0; i < T_; ++i) {
      const int sb = i & 1;
      mbar_wait(&bars[B_SFULL + sb], (i >> 1) & 1);
      umma::fence_after_sync();
      uint32_t r[8];
      TmemIO<8>::ld(tmem + lane_base + TM_S + sb * NR, r);       // 8 columns: the heads (G <= 8)
      umma::tmem_wait_ld();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_SEMPTY + sb]);
      const bool valid = lj[i * TKEYS + kl] >= 0;
      uint32_t w[GW];
#pragma unroll
      for (int j = 0; j < GW; ++j) w[j] = valid ? finish_logit2(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1])) : neg_inf2<T>();
      if (G == 1) w[0] = (w[0] & 0x0000ffffu) | (neg_inf2<T>() & 0xffff0000u);      // the padding row never takes part
      if (GW == 1) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(tmem + lane_base + TM_LOG + i * GW), "r"(w[0]) : "memory");
      } else if (GW == 2) {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(tmem + lane_base + TM_LOG + i * GW), "r"(w[0]), "r"(w[GW > 1 ? 1 : 0]) : "memory");
      } else {
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(tmem + lane_base + TM_LOG + i * GW), "r"(w[0]),
                     "r"(w[GW > 1 ? 1 : 0]), "r"(w[GW > 2 ? 2 : 0]), "r"(w[GW > 3 ? 3 : 0]) : "memory");
      }
      if (i == 0) {
#pragma unroll
        for (int j = 0; j < GW; ++j) {
          uint32_t m = w[j];
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) m = max2<T>(m, __shfl_xor_sync(0xffffffffu, m, o));
          const float2 f = Tr<T>::to_f2(m);
          negm0[2 * j] = f.x == -INFINITY ? 0.f : -f.x * LOG2E;
          negm0[2 * j + 1] = f.y == -INFINITY ? 0.f : -f.y * LOG2E;
        }
      }
#pragma unroll
      for (int j = 0; j < GW; ++j) {
        rmax[j] = max2<T>(rmax[j], w[j]);
        const float2 x = Tr<T>::to_f2(w[j]);
        Ssum[2 * j] += ex2_approx(fmaf(x.x, LOG2E, negm0[2 * j]));
        Ssum[2 * j + 1] += ex2_approx(fmaf(x.y, LOG2E, negm0[2 * j + 1]));
      }
    }
    umma::tmem_wait_st();
    stamp(1);
    // ---- row statistics: lanes -> warps -> cluster -------------------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < GW; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rmax[j] = max2<T>(rmax[j], __shfl_xor_sync(0xffffffffu, rmax[j], o));
    }
#pragma unroll
    for (int j = 0; j < GP; ++j) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) Ssum[j] += __shfl_xor_sync(0xffffffffu, Ssum[j], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int j = 0; j < GW; ++j) {
        const float2 f = Tr<T>::to_f2(rmax[j]);
        const float mm[2] = {f.x, f.y};
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int g = 2 * j + h2;
          if (g < 8) {
            const float dlt = fmaf(mm[h2], LOG2E, negm0[g]);
            const bool ok = mm[h2] == -INFINITY || (dlt < 100.f && dlt > -100.f);
            red[(0 * 4 + warp) * 8 + g] = mm[h2];
            red[(1 * 4 + warp) * 8 + g] = -negm0[g];
            red[(2 * 4 + warp) * 8 + g] = ok ? Ssum[g] : __int_as_float(0x7fc00000);
          }
        }
      }
    }
    named_bar_sync(1, NSOFT);
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // (arrived during setup: every peer's barriers exist)
    if (tid < 8 * C) {                                           // (head, peer): three remote stores + one remote arrive each
      const int g = tid & 7, p = tid >> 3;
      float m = -INFINITY, z = -INFINITY, sacc = 0.f;
      if (g < G) {
#pragma unroll
        for (int wq = 0; wq < 4; ++wq) { m = fmaxf(m, red[(0 * 4 + wq) * 8 + g]); z = fmaxf(z, red[(1 * 4 + wq) * 8 + g]); }
        if (rank == 0) {                                         // the appended token's logit joins rank 0's slice
          const float xn = xnew_s[g];
          m = fmaxf(m, xn);
          const float zn = xn * LOG2E;
          const float z2 = fmaxf(z, zn);
#pragma unroll
          for (int wq = 0; wq < 4; ++wq) sacc += red[(2 * 4 + wq) * 8 + g] * ex2_approx(red[(1 * 4 + wq) * 8 + g] - z2);
          sacc += ex2_approx(zn - z2);
          z = z2;
        } else {
#pragma unroll
          for (int wq = 0; wq < 4; ++wq) sacc += red[(2 * 4 + wq) * 8 + g] * ex2_approx(red[(1 * 4 + wq) * 8 + g] - z);
        }
      }
      if (C == 1) {                                              // one CTA per unit: plain shared memory, no cluster traffic
        xst[g] = m; xst[8 + g] = z; xst[16 + g] = sacc;
      } else {
        st_cluster_f32(map_to_rank(&xst[(rank * 3 + 0) * 8 + g], p), m);
        st_cluster_f32(map_to_rank(&xst[(rank * 3 + 1) * 8 + g], p), z);
        st_cluster_f32(map_to_rank(&xst[(rank * 3 + 2) * 8 + g], p), sacc);
        umma::mbar_arrive_remote(map_to_rank(&bars[B_XST], p));
      }
    }
    if (C == 1) named_bar_sync(1, NSOFT);
    else umma::mbar_wait_cluster(&bars[B_XST], 0);
    if (tid < 8) {
      const int g = tid;
      float m = xst[g], z = xst[8 + g];
      for (int p = 1; p < C; ++p) { m = fmaxf(m, xst[(p * 3 + 0) * 8 + g]); z = fmaxf(z, xst[(p * 3 + 1) * 8 + g]); }
      float sacc = 0.f;
      for (int p = 0; p < C; ++p) sacc += xst[(p * 3 + 2) * 8 + g] * ex2_approx(xst[(p * 3 + 1) * 8 + g] - z);   // rank order
      float s = sacc * ex2_approx(z - m * LOG2E);
      if (g >= G || m == -INFINITY) s = 1.f;
      if (!(s > 0.f) || !(s < INFINITY)) atomicOr(&flags[0], 1);
      rowM[g] = g < G ? m : 0.f;
      rowL[g] = ARITH ? s : __fdiv_rn(1.0f, s);
      rowR[g] = __frcp_rn(s);
    }
    named_bar_sync(1, NSOFT);
    float negM[GP], L[GP], Rc[GP];
#pragma unroll
    for (int j = 0; j < GP; ++j) negM[j] = -rowM[j < 8 ? j : 0];
    if (flags[0]) {
      // ---- exact denominators (slow path): sum of exp(x - M) over the parked logits, second exchange --------------------------
      float Lx[GP];
#pragma unroll
      for (int j = 0; j < GP; ++j) Lx[j] = 0.f;
      for (int i = 0; i < T_; ++i) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        uint32_t r4[4];
        if (GW == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r4[0]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        else if (GW == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r4[0]), "=r"(r4[1]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        else asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r4[0]), "=r"(r4[1]), "=r"(r4[2]), "=r"(r4[3]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        umma::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < GW; ++j) w[j] = r4[j];
#pragma unroll
        for (int j = 0; j < GW; ++j) {
          const float2 x = Tr<T>::to_f2(w[j]);
          Lx[2 * j] += expf(x.x + negM[2 * j]);
          Lx[2 * j + 1] += expf(x.y + negM[2 * j + 1]);
        }
      }
#pragma unroll
      for (int j = 0; j < GP; ++j) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) Lx[j] += __shfl_xor_sync(0xffffffffu, Lx[j], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int j = 0; j < GP; ++j) if (j < 8) red[(2 * 4 + warp) * 8 + j] = Lx[j];
      }
      named_bar_sync(1, NSOFT);
      if (tid < 8 * C) {
        const int g = tid & 7, p = tid >> 3;
        float sx = 0.f;
        if (g < G) {
          sx = ((red[(2 * 4 + 0) * 8 + g] + red[(2 * 4 + 1) * 8 + g]) + red[(2 * 4 + 2) * 8 + g]) + red[(2 * 4 + 3) * 8 + g];
          if (rank == 0) sx += expf(xnew_s[g] - rowM[g]);
        }
        if (C == 1) xst2[g] = sx;
        else {
          st_cluster_f32(map_to_rank(&xst2[rank * 8 + g], p), sx);
          umma::mbar_arrive_remote(map_to_rank(&bars[B_XST2], p));
        }
      }
      if (C == 1) named_bar_sync(1, NSOFT);
      else umma::mbar_wait_cluster(&bars[B_XST2], 0);
      if (tid < 8) {
        float s = xst2[tid];
        for (int p = 1; p < C; ++p) s += xst2[p * 8 + tid];
        if (tid >= G || s == 0.f) s = 1.f;
        rowL[tid] = ARITH ? s : __fdiv_rn(1.0f, s);
        rowR[tid] = __frcp_rn(s);
      }
      named_bar_sync(1, NSOFT);
    }
#pragma unroll
    for (int j = 0; j < GP; ++j) { L[j] = rowL[j < 8 ? j : 0]; Rc[j] = rowR[j < 8 ? j : 0]; }
    auto prob = [&](float x, int j) -> float {                  // llama_patch.py:218-219
      const float ex = expf(x + negM[j]);
      return Tr<T>::round_f(ARITH ? div_rn_by(ex, L[j], Rc[j]) : __fmul_rn(ex, L[j]));
    };
    stamp(2);

    // the appended token's probabilities (rank 0), ahead of the V phase: the helper warps take its policy entry from here
    if (rank == 0 && warp == 0) {
      if (tid < 8) {
        const int g = tid;
        float pn = 0.f;
        if (g < G) {
          const float ex = expf(xnew_s[g] - rowM[g]);
          pn = Tr<T>::round_f(ARITH ? div_rn_by(ex, rowL[g], rowR[g]) : __fmul_rn(ex, rowL[g]));
        }
        pnew_s[g] = pn;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PNEW]);
    }
    // ---- V phase: probabilities, P^T tiles; the folded probability of every key goes to the helper warps, which keep the
    // policy state and the selection keys (below) --------------------------------------------------------------------------------
    for (int i = 0; i < T_; ++i) {
      const int pb = i & 1;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      {
        uint32_t r4[4];
        if (GW == 1) asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r4[0]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        else if (GW == 2) asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r4[0]), "=r"(r4[1]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        else asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r4[0]), "=r"(r4[1]), "=r"(r4[2]), "=r"(r4[3]) : "r"(tmem + lane_base + TM_LOG + i * GW) : "memory");
        umma::tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < GW; ++j) w[j] = r4[j];
      }
      float fsum = 0.f;
#pragma unroll
      for (int j = 0; j < GW; ++j) {
        const float2 x = Tr<T>::to_f2(w[j]);
        const float p0 = prob(x.x, 2 * j);
        const float p1 = (G == 1) ? 0.f : prob(x.y, 2 * j + 1);
        w[j] = pack2<T>(p0, p1);
        fsum += p0;
        fsum += p1;
      }
      if (i >= 2) mbar_wait(&bars[B_FSEMPTY + pb], ((i >> 1) - 1) & 1);
      fs[pb * TKEYS + kl] = fsum;
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FSFULL + pb]);
      if (i >= 2) mbar_wait(&bars[B_PEMPTY + pb], ((i >> 1) - 1) & 1);
      {
        // the key's row group 0 (rows 0..7) of the un-swizzled MN-major operand: 16 bytes; rows >= GP stay zero
        unsigned char* dst = Ps + pb * 4096 + (kl >> 3) * 256 + (kl & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
      }
      umma::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_PFULL + pb]);
    }
    stamp(3);
    named_bar_sync(1, NSOFT);                                    // (pnew_s, written by warp 0 ahead of the V phase, is read below)

    // ---- output: partial O^T -> rank 0 -> out ---------------------------------------------------------------------------------
    {
      float o[8];
#pragma unroll
      for (int g = 0; g < 8; ++g) o[g] = 0.f;
      if (T_ > 0) {
        mbar_wait(&bars[B_OFULL], 0);
        umma::fence_after_sync();
        uint32_t r[8];
        TmemIO<8>::ld(tmem + lane_base + TM_O, r);
        umma::tmem_wait_ld();
#pragma unroll
        for (int g = 0; g < 8; ++g) o[g] = __uint_as_float(r[g]);
      }
      float* obuf = reinterpret_cast<float*>(smem + SL.off_obuf);   // rank 0: [C][8][128] (its own buffer: a fast peer may send
                                                                 // while rank 0 still streams)
      if (rank != 0) {
#pragma unroll
        for (int g = 0; g < G; ++g) st_cluster_f32(map_to_rank(&obuf[(rank * 8 + g) * D + kl], 0), o[g]);
        umma::mbar_arrive_remote(map_to_rank(&bars[B_XOUT], 0));
      } else {
        if (tid == 0) mbar_arrive(&bars[B_XOUT]);
        if (C > 1) umma::mbar_wait_cluster(&bars[B_XOUT], 0);
        const T* vn = reinterpret_cast<const T*>(a.v_new) + (size_t)unit * D;
        const float vx = Tr<T>::to_f(vn[kl]);
        T* og = reinterpret_cast<T*>(a.out) + (size_t)unit * G * D;
#pragma unroll
        for (int g = 0; g < G; ++g) {
          float v = o[g];
          for (int p = 1; p < C; ++p) v += obuf[(p * 8 + g) * D + kl];          // rank order
          v = fmaf(pnew_s[g], vx, v);                                            // the appended token's own value row
          og[g * D + kl] = Tr<T>::from_f(v);                                     // llama_patch.py:222
        }
      }
    }
    stamp(4);
  }
  if (warp >= NSOFT / 32 + 2) {
    // ===== helper warps, V phase: policy state + selection keys per key (accumulate, counter: easykv.py:288-304; keys:
    // ekv_select.cuh) and the select's histograms, from the folded probabilities the softmax warps hand over ===================
    const int new_slot = a.new_slots ? a.new_slots[unit] : n_phys;
    const int hset = (tid - NSOFT - 64) / NHS;                   // set 0: even tiles, set 1: odd tiles (= the hand-over buffer)
    const int kl = (tid - NSOFT - 64) % NHS;                     // key inside a tile
    const ekv_step& st = stu;
    const int P = st.score_offset;
    const int n_s = nb + 1 - P;
    const bool evicting = st.evict > 0 && st.policy != EKV_POLICY_NONE;
    float* Sg = a.S + (size_t)unit * a.cap;
    float* SQg = a.SQ + (size_t)unit * a.cap;
    float* Cg = a.C + (size_t)unit * a.cap;
    const bool stateful = st.policy == EKV_POLICY_ROCO || st.policy == EKV_POLICY_H2O || st.policy == EKV_POLICY_TOVA;
    const uint8_t need_flag = !evicting ? 0 : st.policy == EKV_POLICY_ROCO ? F_CAND : (st.policy == EKV_POLICY_RANGE ? 0 : F_FEAS);
    const bool roco_sel = evicting && st.policy == EKV_POLICY_ROCO;
    const float inv_g = 1.0f / (float)G;
    if (stateful && hset == 0) {
      // this slice's policy state -> L2 while the K phase streams (the helpers idle until the V phase): the per-tile loads
      // below then cost an L2 hit instead of a DRAM round trip per tile
      int cnt = T_ * TKEYS;
      if (cnt > a.cap - first) cnt = a.cap - first;
      const int lines = (cnt * 4 + 127) / 128;
      for (int i = kl; i < 3 * lines; i += NHS) {
        const float* base = (i < lines ? Sg : (i < 2 * lines ? SQg : Cg)) + first;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + (size_t)(i % lines) * 128));
      }
    }
    float s_nx = 0.f, sq_nx = 0.f, c_nx = 1.f;                   // state of the NEXT tile's entry, loaded one tile ahead
    auto load_state = [&](int i, float& sv, float& sq, float& cc) {
      sv = 0.f; sq = 0.f; cc = 1.f;
      if (i < T_ && stateful) {
        const int rl = lj[i * TKEYS + kl];
        if (rl >= P) { const int ph = first + i * TKEYS + kl; sv = Sg[ph]; sq = SQg[ph]; cc = Cg[ph]; }
      }
    };
    for (int i = hset; i < T_; i += 2) {
      const int pb = i & 1, e = i * TKEYS + kl;
      mbar_wait(&bars[B_FSFULL + pb], (i >> 1) & 1);             // (first tile: the slot-map slice is final as well)
      const float fsum = fs[pb * TKEYS + kl];
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars[B_FSEMPTY + pb]);
      if (i == hset) load_state(i, s_nx, sq_nx, c_nx);
      float sv = s_nx, sq = sq_nx, cc = c_nx;
      load_state(i + 2, s_nx, sq_nx, c_nx);
      const int rl = lj[e];
      uint32_t ka = 0, kb = 0;
      uint8_t f = 0;
      bool dirty = false;
      if (rl >= 0 && rl >= P) {
        float ds = 0.f, dsq = 0.f;
        if (st.accumulate) {
          ds = G == 1 ? fsum : Tr<T>::round_f(__fmul_rn(fsum, inv_g));          // process_for_mqa_gqa, easykv.py:188-196
          dsq = Tr<T>::round_f(__fmul_rn(ds, ds));                              // p**2 in the model dtype, :296
        }
        entry_update(st, rl - P, n_s, false, ds, dsq, sv, sq, cc, ka, kb, f, dirty);
        if (dirty) { const int ph = first + e; Sg[ph] = sv; SQg[ph] = sq; Cg[ph] = cc; }
      }
      const bool keyed = (f & need_flag) == need_flag && need_flag;
      kk[e] = keyed ? (((unsigned long long)kb << 32) | ka) : ~0ull;
      if (roco_sel) {                                                           // the select's histograms, on the fly
        __syncwarp();
        bs.add_warp(keyed && (((unsigned long long)kb << 32) | ka) != ~0ull, ka, (uint32_t)rl, lsh, lane);
      }
    }
    if (rank == 0 && hset == 0 && kl == 0) {                     // the appended token's own entry
      mbar_wait(&bars[B_PNEW], 0);
      float fsum = 0.f;
#pragma unroll
      for (int g = 0; g < G; ++g) fsum += pnew_s[g];
      float sv = 0.f, sq = 0.f, cc = st.c_new0;
      uint32_t ka = 0, kb = 0;
      uint8_t f = 0;
      bool dirty = false;
      const int rl = nb;
      if (rl >= P) {
        float ds = 0.f, dsq = 0.f;
        if (st.accumulate) {
          ds = G == 1 ? fsum : Tr<T>::round_f(__fmul_rn(fsum, inv_g));
          dsq = Tr<T>::round_f(__fmul_rn(ds, ds));
        }
        entry_update(st, rl - P, n_s, true, ds, dsq, sv, sq, cc, ka, kb, f, dirty);
        if (dirty) { Sg[new_slot] = sv; SQg[new_slot] = sq; Cg[new_slot] = cc; }
      }
      const bool keyed = (f & need_flag) == need_flag && need_flag;
      kk[e_new] = keyed ? (((unsigned long long)kb << 32) | ka) : ~0ull;
      if (roco_sel && keyed && (((unsigned long long)kb << 32) | ka) != ~0ull) bs.add(ka, (uint32_t)rl, lsh);
    }
  }
  // ===== tail: victim walk, renumbering, append — the softmax warps and the helper warps (8 warps) ================================
  if (warp >= NSOFT / 32 + 2) {
    const int ttid = tid - NSOFT - 64;                           // 0 .. NTAIL-1
    const int tw = ttid >> 5;                                    // tail warp 0 .. NTW-1
    const ekv_step& st = stu;
    const int P = st.score_offset;
    const bool evicting = st.evict > 0 && st.policy != EKV_POLICY_NONE;
    int32_t* lidx_g = a.lidx + (size_t)unit * a.cap;
    const int new_slot = a.new_slots ? a.new_slots[unit] : n_phys;
    unsigned long long* tl = a.timeline ? a.timeline + (size_t)blockIdx.x * 16 : nullptr;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");   // (arrived during setup: every peer's barriers exist)
    named_bar_sync(2, NTAIL);                                    // keys of every entry are in shared memory
    if (tl && ttid == 0) tl[7] = global_ns();
    bool found = false;
    uint32_t l_c = 0;
    int owner = -1, e_c = -1;
    if (evicting && st.policy == EKV_POLICY_RANGE) {
      l_c = (uint32_t)(P + st.range_start);
      found = true;
    } else if (evicting) {
      // roco: the k_feasible-th smallest std from the histograms the V phase filled (ekv_bucket.cuh), then ONE pass over the
      // entries — argmin of (mean, std, logical index) over the feasible ones, the cut bucket's entries listed — one exchange
      // over the cluster, exact ranks of the listed entries (easykv.py:322-324).  h2o_head / tova: the pass and the exchange
      // only (the window's argmin, :311, :335).
      const bool roco = st.policy == EKV_POLICY_ROCO;
      auto sync = [&] { named_bar_sync(2, NTAIL); };
      Feasibility fz;
      fz.mode = 0; fz.lsh = lsh; fz.T1 = 0u; fz.jT = 0xffffffffu;
      fz.b.status = bk::OK; fz.b.bsel = 0; fz.b.rb = 0; fz.b.csel = 0; fz.b.kind = 0; fz.b.mtot = 0; fz.b.my_off = 0;
      const uint32_t kk_a = smem_u32(kk), lj_a = smem_u32(lj);      // (explicit ld.shared: see ekv_bucket.cuh)
      auto get = [&](int e, uint32_t& ka, uint32_t& kb, uint32_t& l) -> bool {
        const unsigned long long k = lds_u64(kk_a + 8u * (uint32_t)e);
        l = lds_u32(lj_a + 4u * (uint32_t)e);
        ka = (uint32_t)k; kb = (uint32_t)(k >> 32);
        return k != ~0ull;
      };
      if (roco) {
        int hphase = 0;
        auto hist_sync = [&] {
          if (C > 1) {
            // every thread's histogram adds precede a named barrier; one release-arrive per peer publishes them
            if (ttid < C) umma::mbar_arrive_remote(map_to_rank(&bars[B_HRDY], ttid));
            umma::mbar_wait_cluster(&bars[B_HRDY], hphase);
            hphase ^= 1;
          }
        };
        hist_sync();
        if (tl && ttid == 0) tl[9] = global_ns();
        auto ld = [&](const uint32_t* ptr, int peer) -> uint4 {
          if (C == 1 || peer == rank) return *reinterpret_cast<const uint4*>(ptr);
          uint4 v;
          asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(map_to_rank(ptr, peer)) : "memory");
          return v;
        };
        fz.b = bucket_scan<MAX_CLUSTER>(bs, st.k_feasible, C, rank, ttid, NTAIL, NEl, sync, ld, get, hist_sync);
        fz.mode = 1;
        if (tl && ttid == 0) { tl[8] = (unsigned long long)(1 + fz.b.status + 4 * fz.b.kind); tl[10] = global_ns(); tl[14] = (unsigned long long)fz.b.mtot; }
        if (fz.b.status == bk::FALLBACK) {
          // The cut falls into a bucket crowded with non-NaN keys (e.g. thousands of slots whose std is exactly 0): select in
          // bounded time.  Cluster-wide MSB-first radix select (8 bits per pass, every pass's 256-bin histogram all-gathered
          // over DSMEM) of the k_feasible-th smallest std key, ties at the cut broken by logical index.
          int xpass = 0;
          auto cluster_radix32 = [&](int m, auto key, auto pred, uint32_t& Tk, int& needk, int& tcount) {
            uint32_t prefix = 0u, maskp = 0u;
            int rem = m;
            tcount = 0;
#pragma unroll 1
            for (int shift = 24; shift >= 0; shift -= 8, ++xpass) {
              const int hb = xpass & 1;
              hist[ttid] = 0u;
              named_bar_sync(2, NTAIL);
              // warp-aggregated histogram: the keys of a pass crowd into a few bins (same sign / exponent), where one
              // shared-memory atomic per key would serialise; lanes with equal bins elect one adder
              for (int e0 = 0; e0 < NEl; e0 += NTAIL) {
                const int e = e0 + ttid;
                uint32_t bin = 0xffffffffu;
                if (e < NEl && pred(e)) {
                  const uint32_t k = key(e);
                  if ((k & maskp) == prefix) bin = (k >> shift) & 255u;
                }
                const uint32_t peers = __match_any_sync(0xffffffffu, bin);
                if (bin != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
              }
              named_bar_sync(2, NTAIL);
              {
                const uint32_t v = hist[ttid];
                if (C == 1) xhist[(hb * MAX_CLUSTER) * 256 + ttid] = v;
                else
                  for (int p2 = 0; p2 < C; ++p2) {
                    st_cluster_u32(map_to_rank(&xhist[(hb * MAX_CLUSTER + rank) * 256 + ttid], p2), v);
                    umma::mbar_arrive_remote(map_to_rank(&bars[B_XHIST + hb], p2));
                  }
              }
              if (C > 1 && ttid == 0) umma::mbar_wait_cluster(&bars[B_XHIST + hb], (xpass >> 1) & 1);
              named_bar_sync(2, NTAIL);
              if (ttid < 32) {
                uint32_t loc[8], sum = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  uint32_t t = 0;
                  for (int p2 = 0; p2 < C; ++p2) t += xhist[(hb * MAX_CLUSTER + p2) * 256 + ttid * 8 + j];
                  loc[j] = t; sum += t;
                }
                uint32_t inc = sum;
#pragma unroll
                for (int o2 = 1; o2 < 32; o2 <<= 1) {
                  const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o2);
                  if (lane >= o2) inc += t;
                }
                const uint32_t exc = inc - sum, total = __shfl_sync(0xffffffffu, inc, 31);
                uint32_t want = (uint32_t)rem;
                if (want > total) want = total;                  // fewer candidates than requested: all of them
                if (want == 0) want = 1;
                if (exc < want && want <= inc) {
                  uint32_t cum = exc;
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    if (want <= cum + loc[j]) { hmisc[0] = ttid * 8 + j; hmisc[1] = (int)cum; hmisc[2] = (int)loc[j]; break; }
                    cum += loc[j];
                  }
                }
                if (lane == 0 && total == 0) { hmisc[0] = 255; hmisc[1] = 0; hmisc[2] = 0; }
              }
              named_bar_sync(2, NTAIL);
              prefix |= (uint32_t)hmisc[0] << shift;
              maskp |= 255u << shift;
              rem -= hmisc[1];
              tcount = hmisc[2];
              named_bar_sync(2, NTAIL);                          // hmisc / hist are rewritten by the next pass
            }
            Tk = prefix;
            needk = rem < tcount ? rem : tcount;
            if (needk < 0) needk = 0;
          };
          uint32_t T1, jT = 0xffffffffu;
          int need1, tc1;
          cluster_radix32(st.k_feasible, [&](int e) { return (uint32_t)kk[e]; }, [&](int e) { return kk[e] != ~0ull; }, T1, need1, tc1);
          if (need1 < tc1) {                                     // the cut falls inside a run of equal std: lowest logical index first
            int nd, tc;
            cluster_radix32(need1, [&](int e) { return (uint32_t)lj[e]; },
                            [&](int e) { return kk[e] != ~0ull && (uint32_t)kk[e] == T1; }, jT, nd, tc);
          }
          fz.mode = 2; fz.T1 = T1; fz.jT = jT;
        }
      }
      auto push = [&](unsigned long long* slot, unsigned long long hi, unsigned long long lo) {       // to every OTHER CTA
        for (int p2 = 0; p2 < C; ++p2) {
          if (p2 == rank) continue;
          const uint32_t dst = map_to_rank(slot, p2);
          asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dst), "l"(hi) : "memory");
          asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(dst + 8), "l"(lo) : "memory");
        }
      };
      bucket_pass(bs, fz, NEl, rank, ttid, NTAIL, get, push, sync, tl ? tl + 13 : nullptr);
      if (tl && ttid == 0) tl[11] = global_ns();
      if (C > 1) {
        __syncwarp();                                            // the warp's remote stores precede its lanes' release-arrives
        if (lane < C) umma::mbar_arrive_remote(map_to_rank(&bars[B_XG], lane));
        umma::mbar_wait_cluster(&bars[B_XG], 0);
      } else {
        sync();
      }
      if (tl && ttid == 0) tl[12] = global_ns();
      Tuple128 wn;
      if (bucket_final(bs, fz, C, ttid, NTAIL, sync, wn)) {
        l_c = (uint32_t)(wn.lo >> 32); owner = (int)((wn.lo >> 24) & 0xffu); e_c = (int)(wn.lo & 0xffffffu);
        found = true;
      }
    }
    if (tl && ttid == 0) tl[5] = global_ns();

    // ---- apply: renumber this slice, free the victim's slot, publish the new slot ------------------------------------------------
    const bool is_range = evicting && st.policy == EKV_POLICY_RANGE;
    for (int e = ttid; e < NEl; e += NTAIL) {
      const int l = lj[e];
      if (l < 0) continue;
      const bool is_new = rank == 0 && e == e_new;
      const int phys = is_new ? new_slot : first + e;
      const bool victim = found && (is_range ? (uint32_t)l == l_c : (rank == owner && e == e_c));
      if (victim) {
        if (a.victim_lidx) a.victim_lidx[unit] = l;
        if (a.victim_slots) a.victim_slots[unit] = phys;
        if (st.apply) lidx_g[phys] = -1;
        else if (is_new) lidx_g[phys] = l;
      } else if (found && st.apply && (uint32_t)l > l_c) {
        lidx_g[phys] = l - 1;
      } else if (is_new) {
        lidx_g[phys] = l;
      }
    }
    if (((evicting && !found) || (a.st.evict > 0 && stu.evict == 0)) && rank == 0 && ttid == 0) {     // no candidate, or below the budget gate
      if (a.victim_lidx) a.victim_lidx[unit] = -1;
      if (a.victim_slots) a.victim_slots[unit] = -1;
    }
    if (tl && ttid == 0) tl[6] = global_ns();
    // rank 0 appends the new row once every CTA of the cluster is past its streams: the output gather has completed
    // (the tail runs beside it on the helper warps, so it waits for the gather's barrier itself)
    if (rank == 0 && tw == 0) {
      if (C > 1) umma::mbar_wait_cluster(&bars[B_XOUT], 0);
      else if (T_ > 0) mbar_wait(&bars[B_OFULL], 0);
      const uint4* kn = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.k_new) + (size_t)unit * D);
      const uint4* vn = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.v_new) + (size_t)unit * D);
      uint4* Kw = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.K) + ((size_t)unit * a.cap + new_slot) * D);
      uint4* Vw = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.V) + ((size_t)unit * a.cap + new_slot) * D);
      if (lane < 16) Kw[lane] = kn[lane];
      else Vw[lane - 16] = vn[lane - 16];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == NSOFT / 32 + 1) umma::tmem_dealloc(tmem, TM_COLS);
}

// ---- plan / launch ---------------------------------------------------------------------------------------------------------
int decode_variant();   // ekv_api.cu

template <typename T, int G, bool ARITH>
static int launch_du_k(const KernelArgs& a, const DecodeUmmaPlan& pl, const CUtensorMap* maps, cudaStream_t stream) {
  using namespace du;
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  cudaError_t err;
  if (!configured[dev]) {
    err = cudaFuncSetAttribute(decode_umma_kernel<T, G, ARITH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(decode_umma)", err);
    configured[dev] = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.Hkv * pl.C), 1, 1);
  cfg.blockDim = dim3(NT, 1, 1);
  cfg.dynamicSmemBytes = (size_t)DuSmem(pl.tps, pl.C, du_select_scratch(a.st)).total;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)pl.C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the kernel waits (pdl_wait) before its first global access
  attr[1].val.programmaticStreamSerializationAllowed = pdl_allowed();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  err = cudaLaunchKernelEx(&cfg, decode_umma_kernel<T, G, ARITH>, a, pl, maps[0], maps[1]);
  if (err != cudaSuccess) return set_cuda_error("decode_umma_kernel launch", err);
  count_launch();
  return EKV_OK;
}

static int du_sms() {
  static thread_local int sm_count[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (!sm_count[dev] && cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sm_count[dev] = 148;
  return sm_count[dev];
}
int decode_cluster_size();   // ekv_api.cu: forced cluster size (0 = planner)

template <typename T, int G> static int launch_du_tg(const KernelArgs& a, cudaStream_t stream) {
  using namespace du;
  DecodeUmmaPlan pl;
  pl.nct = (a.n_phys + TKEYS - 1) / TKEYS;
  const int U = a.B * a.Hkv, sms = du_sms();
  // how many clusters of c CTAs the chip holds at once (one CTA per SM, a cluster inside one GPC)
  static const int conc[9] = {0, 148, 74, 48, 33, 26, 24, 16, 14};
  const int force = decode_cluster_size();
  int best = 0;
  double best_cost = 0;
  for (int c = 1; c <= MAX_CLUSTER; c *= 2) {
    if (force > 0 && force <= MAX_CLUSTER && c != force) continue;
    const int tps = (pl.nct + c - 1) / c;
    constexpr int GW = (G < 2 ? 2 : G) / 2;
    if (tps > (int)(TM_COLS - TM_LOG) / GW) continue;            // parked logits: GW tensor-memory columns per tile
    if (DuSmem(tps, c, du_select_scratch(a.st)).nstage < 3) continue;                     // the per-entry arrays must leave a 3-stage ring
    const int cc = conc[c] * sms / 148 > 0 ? conc[c] * sms / 148 : 1;
    const double waves = (double)((U + cc - 1) / cc);
    const double cost = waves * (tps + 5.0 + 0.5 * c);           // fixed per-CTA work ~ 5 tile-times; exchanges grow with c
    if (!best || cost < best_cost - 1e-9) { best = c; best_cost = cost; }
  }
  if (!best) return EKV_ERR_UNSUPPORTED;
  pl.C = best;
  pl.tps = (pl.nct + best - 1) / best;
  CUtensorMap maps[2];
  const unsigned long long rows_c = (unsigned long long)U * a.cap;
  int rc = make_tensor_map_rows128(&maps[0], a.K, rows_c, TKEYS, a.dtype);
  if (!rc) rc = make_tensor_map_rows128(&maps[1], a.V, rows_c, TKEYS, a.dtype);
  if (rc) return rc;
  return a.st.arith ? launch_du_k<T, G, true>(a, pl, maps, stream) : launch_du_k<T, G, false>(a, pl, maps, stream);
}

template <typename T> static int launch_du_t(const KernelArgs& a, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_du_tg<T, 1>(a, stream);
    case 2: return launch_du_tg<T, 2>(a, stream);
    case 4: return launch_du_tg<T, 4>(a, stream);
    case 8: return launch_du_tg<T, 8>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

// q_len == 1, 16-bit dtypes, head_dim 128, at most one victim per unit; EKV_ERR_UNSUPPORTED otherwise (the caller falls
// back to the FMA / mma.sync decode kernels).
int launch_decode_umma(const KernelArgs& a, cudaStream_t stream) {
  if (a.q_len != 1 || a.d != du::D || a.st.tova_head_mean || a.st.evict > 1 || (a.cap & 3) || a.n_phys < 1 || a.rope_cos) return EKV_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(a.K) | reinterpret_cast<uintptr_t>(a.V)) & 15) return EKV_ERR_UNSUPPORTED;
  switch (a.dtype) {
    case EKV_F16: return launch_du_t<__half>(a, stream);
    case EKV_BF16: return launch_du_t<__nv_bfloat16>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

}  // namespace ekv
