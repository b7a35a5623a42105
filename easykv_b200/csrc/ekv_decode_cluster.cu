// Fused decode step (q_len == 1), cluster-split variant: the key range of ONE (sequence, kv head) unit is
// divided over the C CTAs of a thread-block cluster.  Used when a whole unit does not fit one CTA's shared
// memory (long retained caches with large GQA groups: Mistral / Llama-2-70B layouts at 8K slots) or when
// there are fewer units than SMs (small batches), where one CTA per unit would leave most of the chip idle.
//
// Per CTA (rank r of C): a TMA producer warp streams the CTA's slice of K rows, then of V rows, through an
// mbarrier ring exactly as in ekv_decode.cu; 8 consumer warps compute the slice's logits, probabilities and
// partial P·V.  What crosses CTAs goes through distributed shared memory (st.shared::cluster) bracketed by
// cluster barriers:
//   1. per-head slice maxima               (C x G floats, all-to-all)
//   2. per-head slice sums of exp(x - max) (C x G floats, all-to-all)
//   3. partial outputs, reduce-scattered by output dimension (each CTA finishes D/C dims of every head),
//      together with each slice's best victim candidate (one 128-bit tuple)
//   4. per attempt: the candidate's std rank (one int per CTA)              — roco only
// Sums across CTAs are always formed in rank order, so every CTA derives bit-identical softmax
// denominators and the result does not depend on arrival order.
//
// The victim is found exactly as in the single-CTA fast path (ekv_select.cuh): candidates are visited in
// (mean, std, logical index) order and the first whose std rank is below k_feasible is taken — the slot
// argmin-over-the-k-smallest-std picks (easykv/easykv.py:322-324, :722-724).  h2o_head / tova need one
// cluster argmin (:311, :335); recency is positional (:343-347, :741-742).  Eviction renumbers each slice's
// share of the slot map locally; rank 0 appends the new token.
//
// Replaces the same reference lines as ekv_decode.cu.
#include "ekv_decode_common.cuh"
#include "ekv_mma.cuh"

namespace ekv {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_unaligned() {      // for a single thread of a diverged warp
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ void st_cluster_u64(uint32_t addr, unsigned long long v) {
  asm volatile("st.shared::cluster.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory");
}

constexpr int TC_PITCH = 272;                     // tensor-core variant: bytes per shared-memory row (256 + 16 keeps
constexpr int TC_TILE_BYTES = 64 * TC_PITCH;      // ldmatrix conflict-free), 64 rows per tile

template <typename T> struct ClusterSmem {
  int off_bar, off_red, off_xnew, off_q, off_qp, off_k, off_v, off_ns, off_lj, off_plog, off_ka, off_kb, off_flag;
  int off_xmax, off_xsum, off_xout, off_xbest, off_xcnt, off_hist, off_xhist, off_bkt, off_ring, fixed, slp;
  __host__ __device__ ClusterSmem(int G, int slice, int C, bool tc = false) {
    using Cfg = DecodeCfg<T>;
    const int NEl = slice + 1;                    // rank 0 also owns the appended token's entry
    slp = align_up(NEl, 8);
    int o = 0;
    off_bar = o; o += 2 * Cfg::MAX_STAGES * 8;
    o = align_up(o, 128);
    off_red = o; o += 2 * 8 * Cfg::NWARP * 4 + 64 * 8 + 64 * 4;     // float maxima/sums | u64 tuples | int counts
    off_xnew = o; o += 16 * 4;                                      // the appended token's logits (tensor-core variant)
    off_qp = o; o += tc ? 16 * TC_PITCH : 0;                        // q as a 16-row A operand, rows >= G zero
    off_q = o; o += G * Cfg::ROW_BYTES;
    off_k = o; o += Cfg::ROW_BYTES;
    off_v = o; o += Cfg::ROW_BYTES;
    off_ns = o; o += 16;
    off_lj = o; o += align_up(NEl * 4, 16);
    off_plog = o; o += align_up(G * slp * (int)sizeof(T), 16);
    off_ka = o; o += align_up(NEl * 4, 16);
    off_kb = o; o += align_up(NEl * 4, 16);
    off_flag = o; o += align_up(NEl, 16);
    off_xmax = o; o += C * G * 4;
    off_xsum = o; o += C * G * 4;
    off_xout = o; o += G * Cfg::D * 4;            // [C][G][D/C]
    off_xbest = align_up(o, 16); o = off_xbest + 2 * C * 16;
    off_xcnt = o; o += 2 * C * 4;
    off_hist = align_up(o, 16); o = off_hist + 256 * 4 + 16;          // radix select: local histogram + 4 ints
    off_xhist = o; o += 2 * C * 256 * 4;                              // every CTA's histogram, double-buffered
    off_bkt = align_up(o, 16); o = off_bkt + BucketScratch::bytes(8);  // bucket select (ekv_bucket.cuh)
    o = align_up(o, 128);
    off_ring = o;                                 // the ring doubles as the cross-warp P·V scratch [NWARP][G][D] fp32
    fixed = o;
  }
  static __host__ __device__ int min_ring(int G) { return DecodeCfg<T>::NWARP * G * DecodeCfg<T>::D * 4; }
};

// TC: the contractions run on the tensor cores (mma.sync m16n8k16, the g query heads padded to a 16-row A
// operand).  With g = 8 the FP32 pipes cannot keep up with HBM (g mixed-precision FMAs per 2 loaded bytes for
// Q·K^T and again for P·V), the tensor cores can.
// The producer warp then lands the rows with 16-byte cp.async copies at a 272-byte pitch so that ldmatrix is
// conflict-free.  Only for 16-bit dtypes.
template <typename T, int G, bool TC>
__global__ void __launch_bounds__(DecodeCfg<T>::NCONS + 32, (TC || G <= 4) ? 2 : 1)
decode_cluster_kernel(const KernelArgs a, const int stages, const int slice) {
  using Cfg = DecodeCfg<T>;
  constexpr int D = Cfg::D, NWARP = Cfg::NWARP, NCONS = Cfg::NCONS, RPT = Cfg::RPT, TILE_ROWS = Cfg::TILE_ROWS;
  constexpr int TILEB = TC ? TC_TILE_BYTES : Cfg::TILE_BYTES;
  static_assert(!TC || (sizeof(T) == 2 && TILE_ROWS == 64), "tensor-core variant: 16-bit dtypes");
  pdl_trigger();            // programmatic dependent launch: the next kernel may be placed now; this one touches global memory
  const int C = (int)cluster_nctarank(), rank = (int)cluster_ctarank();
  {
    // ... only once the previous kernel of the stream has completed (pdl_wait below).  Until then — small batches: this
    // CTA is resident beside the previous layer's CTAs for most of their run — it pulls its slice of the cache and of
    // the policy state into L2.  Prefetches are hints and L2 is the point of coherence, so this is safe whatever the
    // previous kernel still writes; nothing is READ before the wait.  Bounded so that a launch never asks for more
    // than ~48 MB (the slices of a small batch fit L2 whole; a larger launch takes the first rows only).
    const int lo_p = min(rank * slice, a.n_phys), n_p = min(lo_p + slice, a.n_phys) - lo_p;
    const long long per_row = 2ll * Cfg::ROW_BYTES * (long long)gridDim.x;
    const int rows_p = (int)min((long long)n_p, (48ll << 20) / per_row);
    const size_t e0 = (size_t)(blockIdx.x / C) * a.cap + lo_p;
    const char* kp = reinterpret_cast<const char*>(a.K) + e0 * Cfg::ROW_BYTES;
    const char* vp = reinterpret_cast<const char*>(a.V) + e0 * Cfg::ROW_BYTES;
    constexpr int CH = 2048;                                     // bytes per bulk prefetch (rows are 16-byte multiples)
    const int kv_bytes = rows_p * Cfg::ROW_BYTES;
    for (int off = (int)threadIdx.x * CH; off < kv_bytes; off += (int)blockDim.x * CH) {
      const uint32_t nbytes = (uint32_t)min(CH, kv_bytes - off);
      bulk_prefetch_l2(kp + off, nbytes);
      bulk_prefetch_l2(vp + off, nbytes);
    }
    const int st_lines = (rows_p * 4 + 127) / 128;               // slot map + S, SQ, C: one 128-byte line per prefetch
    for (int i = (int)threadIdx.x; i < 4 * st_lines; i += (int)blockDim.x) {
      const int which = i / st_lines;
      const char* base = which == 0 ? reinterpret_cast<const char*>(a.lidx) : which == 1 ? reinterpret_cast<const char*>(a.S)
                         : which == 2 ? reinterpret_cast<const char*>(a.SQ) : reinterpret_cast<const char*>(a.C);
      if (base) asm volatile("prefetch.global.L2 [%0];" ::"l"(base + e0 * 4 + (size_t)(i % st_lines) * 128));
    }
  }
  pdl_wait();
  extern __shared__ __align__(128) unsigned char smem[];
  const ClusterSmem<T> L(G, slice, C, TC);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t* empty = full + Cfg::MAX_STAGES;
  unsigned char* ring = smem + L.off_ring;

  const int unit = blockIdx.x / C;
  // ragged batches: this sequence's own count of valid slots; it evicts only past the budget gate (easykv.py:303)
  const int nb = a.seq_n_before ? a.seq_n_before[unit / a.Hkv] : a.n_before;
  ekv_step stu = a.st;
  if (stu.budget_gate > 0 && nb + 1 - stu.score_offset <= stu.budget_gate) stu.evict = 0;
  const bool stream_rope = sizeof(T) == 2 && !TC && a.rope_cos != nullptr;     // fused streaming variant (FMA path)
  const int n_phys = a.n_phys;
  const int lo = min(rank * slice, n_phys), hi = min(lo + slice, n_phys);
  const int nloc = hi - lo;                                   // physical slots of this CTA
  const int NEl = nloc + (rank == 0 ? 1 : 0);                 // + the appended token on rank 0 (local index nloc)
  const int slp = L.slp;
  const int nt = (nloc + TILE_ROWS - 1) / TILE_ROWS;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { mbar_init(&full[s], TC ? 32 : 1); mbar_init(&empty[s], NWARP); }
    mbar_fence_init();
  }
  __syncthreads();
  // Distributed shared memory of a peer may only be touched once that CTA has started executing (racecheck: "block
  // that might not have entered yet"): every thread arrives on the cluster barrier here, and waits for this phase
  // right before its first remote access — by then every peer has long arrived, so the wait costs nothing.
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");

  if (warp == NWARP) {
    // ===== TMA producer ==============================================================================
    if (TC || lane == 0) {
      const uint64_t pol = l2_policy_evict_first();
      const T* Kg = reinterpret_cast<const T*>(a.K) + ((size_t)unit * a.cap + lo) * D;
      const T* Vg = reinterpret_cast<const T*>(a.V) + ((size_t)unit * a.cap + lo) * D;
      int s = 0, use = 0;
      bool synced = false;
      for (int i = 0; i < 2 * nt; ++i) {
        if (use > 0) {
          // A slot last filled with a V tile is only released after the consumers have passed the two
          // softmax cluster barriers.  A cluster barrier completes when every NON-EXITED thread of the
          // cluster has arrived, so the producer has to take part in those two phases before it may
          // block on such a slot (it exits before the later ones; exited threads are not waited for).
          if (!synced && i - stages >= nt) {
            asm volatile("barrier.cluster.wait.acquire;" ::: "memory");      // the start-up phase (arrived above)
            cluster_sync_unaligned();
            cluster_sync_unaligned();
            synced = true;
          }
          if (!TC || lane == 0) mbar_wait(&empty[s], (use - 1) & 1);
        }
        const int tt = i < nt ? i : i - nt;
        const int rows = min(TILE_ROWS, nloc - tt * TILE_ROWS);
        const T* src = (i < nt ? Kg : Vg) + (size_t)tt * TILE_ROWS * D;
        if (TC) {
          // rows land at a 272-byte pitch: 16-byte LDGSTS copies, two rows per warp instruction, and every lane's
          // copies arrive on the slot's barrier when they complete (cp.async.mbarrier.arrive.noinc; the barrier
          // expects the 32 lanes).  (One 256-byte bulk copy per row was measured at ~58 cycles per row.)
          __syncwarp();                                           // lane 0 has seen the slot released
          unsigned char* dst = ring + (size_t)s * TILEB;
          const unsigned char* sb = reinterpret_cast<const unsigned char*>(src);
#pragma unroll 8
          for (int j = 0; j < TILE_ROWS * 16 / 32; ++j) {
            const int idx = lane + 32 * j, r = idx >> 4, c = idx & 15;
            if (r < rows) cp_async16(dst + r * TC_PITCH + c * 16, sb + (size_t)r * Cfg::ROW_BYTES + c * 16);
          }
          asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(&full[s])) : "memory");
        } else {
          const uint32_t bytes = (uint32_t)rows * Cfg::ROW_BYTES;
          mbar_arrive_expect_tx(&full[s], bytes);
          tma_bulk_g2s(ring + (size_t)s * TILEB, src, bytes, &full[s], pol);
        }
        if (++s == stages) { s = 0; ++use; }
      }
    }
    return;     // exited threads do not take part in the cluster barriers below
  }

  // ===== consumers ===================================================================================
  const int tid = threadIdx.x;
  const Grp grp{tid, NCONS, 1};
  const int hw = tid >> 4, l16 = tid & 15;
  float* red = reinterpret_cast<float*>(smem + L.off_red);
  unsigned long long* red64 = reinterpret_cast<unsigned long long*>(smem + L.off_red + 2 * 8 * NWARP * 4);
  int* redi = reinterpret_cast<int*>(smem + L.off_red + 2 * 8 * NWARP * 4 + 64 * 8);
  T* qh = reinterpret_cast<T*>(smem + L.off_q);
  T* kh = reinterpret_cast<T*>(smem + L.off_k);
  T* vh = reinterpret_cast<T*>(smem + L.off_v);
  int32_t* ns = reinterpret_cast<int32_t*>(smem + L.off_ns);
  int32_t* lj = reinterpret_cast<int32_t*>(smem + L.off_lj);
  T* plog = reinterpret_cast<T*>(smem + L.off_plog);
  uint32_t* keyA = reinterpret_cast<uint32_t*>(smem + L.off_ka);
  uint32_t* keyB = reinterpret_cast<uint32_t*>(smem + L.off_kb);
  uint8_t* flag = smem + L.off_flag;
  float* xmax = reinterpret_cast<float*>(smem + L.off_xmax);
  float* xsum = reinterpret_cast<float*>(smem + L.off_xsum);
  float* xout = reinterpret_cast<float*>(smem + L.off_xout);
  unsigned long long* xbest = reinterpret_cast<unsigned long long*>(smem + L.off_xbest);
  int* xcnt = reinterpret_cast<int*>(smem + L.off_xcnt);
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem + L.off_hist);
  int* hmisc = reinterpret_cast<int*>(hist + 256);
  uint32_t* xhist = reinterpret_cast<uint32_t*>(smem + L.off_xhist);
  BucketScratch bs;
  bs.carve(smem + L.off_bkt, 8);
  const int lsh = bk::lidx_shift(nb + 1);
  const bool roco_sel = stu.evict > 0 && stu.policy == EKV_POLICY_ROCO;

  auto finish_logit = [&](float dot, bool valid) -> T {
    float x = Tr<T>::round_f(dot);                                           // llama_patch.py:201
    x = stu.arith ? __fmul_rn(x, a.scale_mul) : __fdiv_rn(x, a.scale_div);  // :202
    return valid ? Tr<T>::from_f(x) : neg_inf<T>();
  };

  unsigned long long* tl = a.timeline ? a.timeline + (size_t)blockIdx.x * 8 : nullptr;   // profiling hook
  auto stamp = [&](int i) { if (tl && tid == 0) tl[i] = clock64(); };
  stamp(0);
  // ---- header ----------------------------------------------------------------------------------------
  {
    const uint4* qg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D);
    const uint4* kg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.k_new) + (size_t)unit * D);
    const uint4* vg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.v_new) + (size_t)unit * D);
    constexpr int QCH = G * Cfg::ROW_BYTES / 16, RCH = Cfg::ROW_BYTES / 16;
    for (int i = tid; i < QCH + 2 * RCH; i += NCONS) {
      if (i < QCH) reinterpret_cast<uint4*>(qh)[i] = qg[i];
      else if (i < QCH + RCH) reinterpret_cast<uint4*>(kh)[i - QCH] = kg[i - QCH];
      else reinterpret_cast<uint4*>(vh)[i - QCH - RCH] = vg[i - QCH - RCH];
    }
    if (roco_sel) bs.clear(tid, NCONS);                       // the select's histograms (filled by the tail's keys pass)
    const int32_t* lg = a.lidx + (size_t)unit * a.cap + lo;
    for (int e = tid; e < nloc; e += NCONS) lj[e] = lg[e];
    if (tid == 0) {
      ns[0] = a.new_slots ? a.new_slots[unit] : n_phys;
      if (rank == 0) lj[nloc] = nb;
    }
    if (stu.policy != EKV_POLICY_NONE && stu.policy != EKV_POLICY_RANGE) {
      const int lines = (nloc * 4 + 127) / 128;
      for (int i = tid; i < 3 * lines; i += NCONS) {
        const float* base = (i < lines ? a.S : (i < 2 * lines ? a.SQ : a.C)) + (size_t)unit * a.cap + lo;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(base) + (size_t)(i % lines) * 128));
      }
    }
  }
  unsigned char* qp = smem + L.off_qp;
  float* xnew_s = reinterpret_cast<float*>(smem + L.off_xnew);
  if (TC) {
    // q again, as the 16-row A operand (rows >= G are zero)
    const uint4* qg = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.q) + (size_t)unit * G * D);
    for (int i = tid; i < 16 * 16; i += NCONS) {
      const int r = i >> 4, c = i & 15;
      *reinterpret_cast<uint4*>(qp + r * TC_PITCH + c * 16) = r < G ? qg[r * 16 + c] : make_uint4(0, 0, 0, 0);
    }
    // rows of a partially filled LAST tile that no earlier tile has ever written hold whatever the shared
    // memory held; P·V multiplies them by p == 0, so they must at least be finite
    const int rows_last = nloc - (nt - 1) * TILE_ROWS;
    if (nt > 0 && rows_last < TILE_ROWS) {
      for (int which = 0; which < 2; ++which) {
        const int idx = which * nt + nt - 1;                  // sequence index of the last K / V tile
        if (idx < stages) {
          unsigned char* base = ring + (size_t)idx * TILEB;
          for (int i = tid; i < (TILE_ROWS - rows_last) * 16; i += NCONS)
            *reinterpret_cast<uint4*>(base + (rows_last + (i >> 4)) * TC_PITCH + (i & 15) * 16) = make_uint4(0, 0, 0, 0);
        }
      }
    }
  }
  grp.sync();
  Row8<T> qr[TC ? 1 : G], knew;
  if (!TC) {
#pragma unroll
    for (int g = 0; g < (TC ? 1 : G); ++g) qr[g].load(qh + g * D, l16);
  }
  knew.load(kh, l16);

  stamp(1);
  // ---- K phase -------------------------------------------------------------------------------------------------
  float mloc = -INFINITY;
  int s = 0;
  uint32_t par = 0;
  float xnew[G];
  const int gw = warp;
  if constexpr (TC) {
    // tensor cores: warp w owns keys [8w, 8w+8) of every 64-key tile; A = q (16 x 128, rows >= G zero)
    uint32_t aq[8][4];
    {
      const int row = (lane & 7) + ((lane >> 3) & 1) * 8, colb = (lane >> 4) * 16;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) ldsm_x4(smem_u32(qp + row * TC_PITCH + ks * 32 + colb), aq[ks]);
    }
    const int g_t = lane >> 2;                                  // the head (A row) this thread's accumulators belong to
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&full[s], (par >> s) & 1u);
      par ^= 1u << s;
      const unsigned char* tile = ring + (size_t)s * TILEB;
      float c[4] = {0.f, 0.f, 0.f, 0.f};
      const uint32_t kaddr = smem_u32(tile + (warp * 8 + (lane & 7)) * TC_PITCH + (lane >> 3) * 16);
#pragma unroll
      for (int ksp = 0; ksp < 4; ++ksp) {                       // two k-steps (32 dims) per ldmatrix.x4
        uint32_t bk[4];
        ldsm_x4(kaddr + ksp * 64, bk);
        mma16816<T>(c, aq[2 * ksp], bk[0], bk[1]);
        mma16816<T>(c, aq[2 * ksp + 1], bk[2], bk[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      s = s + 1 == stages ? 0 : s + 1;
      if (g_t < G) {
#pragma unroll
        for (int cc = 0; cc < 2; ++cc) {
          const int e = i * TILE_ROWS + warp * 8 + 2 * (lane & 3) + cc;
          if (e < nloc) {
            const T x = finish_logit(c[cc], lj[e] >= 0);
            plog[g_t * slp + e] = x;
            mloc = fmaxf(mloc, Tr<T>::to_f(x));
          }
        }
      }
    }
    // the appended token's own key (rank 0 owns it): one half-warp per head, FMA path
    if (tid < 16 * G) {
      const int g = tid >> 4;
      Row8<T> qg8;
      qg8.load(qh + g * D, l16);
      float v = dot8(knew, qg8, 0.f);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      const T x = finish_logit(v, true);
      if (l16 == 0) {
        if (rank == 0) plog[g * slp + nloc] = x;
        xnew_s[g] = rank == 0 ? Tr<T>::to_f(x) : -INFINITY;
      }
    }
    // this thread's running max covers head g_t only: combine the four lanes of the row
    mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 1));
    mloc = fmaxf(mloc, __shfl_xor_sync(0xffffffffu, mloc, 2));
    if ((lane & 3) == 0 && g_t < G) red[g_t * NWARP + gw] = mloc;
  } else {
  constexpr int NVT = RPT * G;
  constexpr int TB = NVT >= 16 ? 1 : 16 / NVT;
  constexpr int NB = NVT >= 16 ? NVT / 16 : 1;
  static_assert(NVT * TB == 16 * NB, "batch must be a whole number of 16-value reductions");
  const int vi0 = bitrev_idx<16>(l16);
  for (int i0 = 0; i0 < nt; i0 += TB) {
    float part[NVT * TB];
#pragma unroll
    for (int tb = 0; tb < TB; ++tb) {
      if (i0 + tb < nt) {
        mbar_wait(&full[s], (par >> s) & 1u);
        par ^= 1u << s;
        const T* tile = reinterpret_cast<const T*>(ring + (size_t)s * Cfg::TILE_BYTES);
        Row8<T> x[RPT];
#pragma unroll
        for (int k = 0; k < RPT; ++k) x[k].load(tile + (hw * RPT + k) * D, l16);
        if (stream_rope) {                                     // fused streaming variant: rotate at the cache-relative position
#pragma unroll
          for (int k = 0; k < RPT; ++k) {
            const int e = (i0 + tb) * TILE_ROWS + hw * RPT + k;
            const int pos = e < nloc ? max(lj[e], 0) : 0;
            rope_row8<T>(x[k], l16, a.rope_cos, a.rope_sin, pos);
          }
        }
#pragma unroll
        for (int k = 0; k < RPT; ++k)
#pragma unroll
          for (int g = 0; g < G; ++g) part[tb * NVT + k * G + g] = dot8(x[k], qr[g], 0.f);
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[s]);
        s = s + 1 == stages ? 0 : s + 1;
      } else {
#pragma unroll
        for (int j = 0; j < NVT; ++j) part[tb * NVT + j] = 0.f;
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = part[b * 16 + j];
      const float r = transpose_reduce16<16>(v, l16);
      const int vi = b * 16 + vi0;
      const int tb = vi / NVT, k = (vi % NVT) / G, g = vi % G;
      const int e = (i0 + tb) * TILE_ROWS + hw * RPT + k;
      if (e < nloc) {
        const T x = finish_logit(r, lj[e] >= 0);
        plog[g * slp + e] = x;
        mloc = fmaxf(mloc, Tr<T>::to_f(x));
      }
    }
  }
  // the appended token's own key: rank 0 owns it (every half-warp computes it, the shuffles need all lanes)
  {
    float v[G];
#pragma unroll
    for (int g = 0; g < G; ++g) v[g] = dot8(knew, qr[g], 0.f);
    const float r = transpose_reduce16<G>(v, l16);
    const T x = finish_logit(r, true);
    if (rank == 0 && hw == 0 && l16 < G) plog[bitrev_idx<G>(l16) * slp + nloc] = x;
#pragma unroll
    for (int g = 0; g < G; ++g) {
      const float xv = __shfl_sync(0xffffffffu, Tr<T>::to_f(x), (lane & 16) | bitrev_idx<G>(g));
      xnew[g] = rank == 0 ? xv : -INFINITY;
    }
  }

  }

  stamp(2);
  // ---- softmax across the cluster ------------------------------------------------------------------------
  float mx[G], inv[G], rcp[G];
  {
    if constexpr (!TC) {
      const int my_g = bitrev_idx<16>(l16) % G;
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float m = fmaxf(warp_max(my_g == g ? mloc : -INFINITY), xnew[g]);
        if (lane == 0) red[g * NWARP + gw] = m;
      }
    }
    grp.sync();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");      // start-up phase: every peer CTA is running
    if (tid < G * C) {                                        // thread (g, p): this slice's max of head g -> peer p
      const int g = tid / C, p = tid % C;
      float v = red[g * NWARP];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) v = fmaxf(v, red[g * NWARP + w]);
      if (TC) v = fmaxf(v, xnew_s[g]);
      st_cluster_f32(map_to_rank(&xmax[rank * G + g], p), v);
    }
    cluster_sync_all();                                                             // (1)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float v = xmax[g];
      for (int p = 1; p < C; ++p) v = fmaxf(v, xmax[p * G + g]);
      mx[g] = v;
    }
    float* red2 = red + 8 * NWARP;
    float sacc[G];
#pragma unroll
    for (int g = 0; g < G; ++g) sacc[g] = 0.f;
    for (int e = tid; e < NEl; e += NCONS)
#pragma unroll
      for (int g = 0; g < G; ++g) sacc[g] += expf(Tr<T>::to_f(plog[g * slp + e]) - mx[g]);
#pragma unroll
    for (int g = 0; g < G; ++g) {
      sacc[g] = warp_sum(sacc[g]);
      if (lane == 0) red2[g * NWARP + gw] = sacc[g];
    }
    grp.sync();
    if (tid < G * C) {
      const int g = tid / C, p = tid % C;
      float v = red2[g * NWARP];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) v += red2[g * NWARP + w];
      st_cluster_f32(map_to_rank(&xsum[rank * G + g], p), v);
    }
    cluster_sync_all();                                                             // (2)
#pragma unroll
    for (int g = 0; g < G; ++g) {
      float v = xsum[g];
      for (int p = 1; p < C; ++p) v += xsum[p * G + g];                            // rank order on every CTA
      inv[g] = stu.arith ? v : __fdiv_rn(1.0f, v);
      rcp[g] = __frcp_rn(v);
    }
    for (int e = tid; e < NEl; e += NCONS)
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const float ex = expf(Tr<T>::to_f(plog[g * slp + e]) - mx[g]);
        plog[g * slp + e] = Tr<T>::from_f(stu.arith ? div_rn_by(ex, inv[g], rcp[g]) : __fmul_rn(ex, inv[g]));   // llama_patch.py:218-219
      }
  }
  grp.sync();

  stamp(3);
  // ---- V phase ---------------------------------------------------------------------------------------------
  if constexpr (TC) {
    // warp w: keys [16*(w%4), +16) of every tile x dims [64*(w/4), +64); A = P (rows = heads) from shared memory
    const int kq = warp & 3, dh = warp >> 2, g_t = lane >> 2;
    float o[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[nb][c] = 0.f;
    const int key_l = (lane & 7) + ((lane >> 3) & 1) * 8, dim_l = (lane >> 4) * 16;
    for (int i = 0; i < nt; ++i) {
      mbar_wait(&full[s], (par >> s) & 1u);
      par ^= 1u << s;
      const unsigned char* tile = ring + (size_t)s * TILEB;
      const int e0 = i * TILE_ROWS + kq * 16 + 2 * (lane & 3);
      uint32_t pa[4] = {0u, 0u, 0u, 0u};
      if (g_t < G) {                                            // keys >= nloc (incl. the appended token) contribute nothing here
        const T* pr = plog + g_t * slp;
        const uint32_t lo0 = e0 < nloc ? (uint32_t)(*reinterpret_cast<const uint16_t*>(&pr[e0])) : 0u;
        const uint32_t hi0 = e0 + 1 < nloc ? (uint32_t)(*reinterpret_cast<const uint16_t*>(&pr[e0 + 1])) : 0u;
        const uint32_t lo1 = e0 + 8 < nloc ? (uint32_t)(*reinterpret_cast<const uint16_t*>(&pr[e0 + 8])) : 0u;
        const uint32_t hi1 = e0 + 9 < nloc ? (uint32_t)(*reinterpret_cast<const uint16_t*>(&pr[e0 + 9])) : 0u;
        pa[0] = lo0 | (hi0 << 16);
        pa[2] = lo1 | (hi1 << 16);
      }
      const uint32_t vaddr = smem_u32(tile + (kq * 16 + key_l) * TC_PITCH + dh * 128 + dim_l);
#pragma unroll
      for (int dbp = 0; dbp < 4; ++dbp) {
        uint32_t bv[4];
        ldsm_x4_trans(vaddr + dbp * 32, bv);
        mma16816<T>(o[2 * dbp], pa, bv[0], bv[1]);
        mma16816<T>(o[2 * dbp + 1], pa, bv[2], bv[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      s = s + 1 == stages ? 0 : s + 1;
    }
    grp.sync();                     // every tile of this CTA has been consumed: the ring is free scratch now
    float* part = reinterpret_cast<float*>(ring);               // [4 key quarters][G][D]
    if (g_t < G) {
#pragma unroll
      for (int nb = 0; nb < 8; ++nb)
        *reinterpret_cast<float2*>(&part[(kq * G + g_t) * D + dh * 64 + nb * 8 + 2 * (lane & 3)]) = make_float2(o[nb][0], o[nb][1]);
    }
    grp.sync();
    const int DPC = D / C;                                      // output dims finished by each CTA
    for (int i = tid; i < G * D; i += NCONS) {
      const int g = i / D, dim = i % D;
      float v = part[i];
#pragma unroll
      for (int w = 1; w < 4; ++w) v += part[w * G * D + i];
      if (rank == 0) v = fmaf(Tr<T>::to_f(plog[g * slp + nloc]), Tr<T>::to_f(vh[dim]), v);     // the appended token's own value row
      st_cluster_f32(map_to_rank(&xout[(rank * G + g) * DPC + dim % DPC], dim / DPC), v);
    }
  } else {
  float oacc[G][8];
#pragma unroll
  for (int g = 0; g < G; ++g)
#pragma unroll
    for (int j = 0; j < 8; ++j) oacc[g][j] = 0.f;
  for (int i = 0; i < nt; ++i) {
    mbar_wait(&full[s], (par >> s) & 1u);
    par ^= 1u << s;
    const T* tile = reinterpret_cast<const T*>(ring + (size_t)s * Cfg::TILE_BYTES);
    const int e0 = i * TILE_ROWS + hw * RPT;
    Row8<T> x[RPT];
    T pv[RPT][G];
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
      x[k].load(tile + (hw * RPT + k) * D, l16);
#pragma unroll
      for (int g = 0; g < G; ++g) pv[k][g] = plog[g * slp + min(e0 + k, slp - 1)];
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty[s]);
    s = s + 1 == stages ? 0 : s + 1;
#pragma unroll
    for (int k = 0; k < RPT; ++k)
      if (e0 + k < nloc) {
#pragma unroll
        for (int g = 0; g < G; ++g) axpy8(pv[k][g], x[k], oacc[g]);
      }
  }
  if (rank == 0 && hw == 0) {
    Row8<T> vnew;
    vnew.load(vh, l16);
#pragma unroll
    for (int g = 0; g < G; ++g) axpy8(plog[g * slp + nloc], vnew, oacc[g]);
  }
  grp.sync();                       // every tile of this CTA has been consumed: the ring is free scratch now
  {
    float* part = reinterpret_cast<float*>(ring);             // [NWARP][G][D]
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float v = oacc[g][j] + __shfl_xor_sync(0xffffffffu, oacc[g][j], 16);
        if (lane < 16) part[(gw * G + g) * D + dim_of<T>(l16, j)] = v;
      }
    grp.sync();
    const int DPC = D / C;                                    // output dims finished by each CTA
    for (int i = tid; i < G * D; i += NCONS) {
      float v = part[i];
#pragma unroll
      for (int w = 1; w < NWARP; ++w) v += part[w * G * D + i];
      const int g = i / D, dim = i % D;
      st_cluster_f32(map_to_rank(&xout[(rank * G + g) * DPC + dim % DPC], dim / DPC), v);
    }
  }

  }

  stamp(4);
  // ---- tail, pass 1: this slice's policy state and selection keys ------------------------------------------------
  const ekv_step& st = stu;
  const int P = st.score_offset;
  const int n_after = nb + 1, n_s = n_after - P;
  const bool evicting = st.evict > 0 && st.policy != EKV_POLICY_NONE;
  float* Sg = a.S + (size_t)unit * a.cap;
  float* SQg = a.SQ + (size_t)unit * a.cap;
  float* Cg = a.C + (size_t)unit * a.cap;
  int32_t* lidx_g = a.lidx + (size_t)unit * a.cap;
  const float inv_g = 1.0f / (float)G;
  auto phys_of = [&](int e) { return e >= nloc ? ns[0] : lo + e; };
  {
    constexpr int FCH = 5;
    for (int base = 0; base < NEl; base += FCH * NCONS) {
      float sv[FCH], sq[FCH], cc[FCH], ds[FCH], dsq[FCH];
      int rl[FCH];
#pragma unroll
      for (int k = 0; k < FCH; ++k) {
        const int e = base + k * NCONS + tid;
        rl[k] = -1; sv[k] = 0.f; sq[k] = 0.f; cc[k] = 1.f; ds[k] = 0.f; dsq[k] = 0.f;
        if (e < NEl) {
          if (e >= nloc) {
            rl[k] = nb;
            cc[k] = st.c_new0;
          } else {
            rl[k] = lj[e];
            if (rl[k] >= P) { sv[k] = Sg[lo + e]; sq[k] = SQg[lo + e]; cc[k] = Cg[lo + e]; }
          }
          if (st.accumulate) {
            float pf;
            if (G == 1) pf = Tr<T>::to_f(plog[e]);
            else {                                                  // process_for_mqa_gqa, easykv.py:188-196
              float sum = 0.f;
#pragma unroll
              for (int g = 0; g < G; ++g) sum += Tr<T>::to_f(plog[g * slp + e]);
              pf = Tr<T>::round_f(__fmul_rn(sum, inv_g));
            }
            ds[k] = pf;
            dsq[k] = Tr<T>::round_f(__fmul_rn(pf, pf));             // p**2 in the model dtype, easykv.py:296
          }
        }
      }
#pragma unroll
      for (int k = 0; k < FCH; ++k) {
        const int e = base + k * NCONS + tid;
        uint32_t ka = 0, kb = 0;
        uint8_t f = 0;
        if (e < NEl) {
          bool dirty = false;
          if (rl[k] >= 0 && rl[k] >= P)
            entry_update(st, rl[k] - P, n_s, e >= nloc, ds[k], dsq[k], sv[k], sq[k], cc[k], ka, kb, f, dirty);
          if (dirty) { const int ph = phys_of(e); Sg[ph] = sv[k]; SQg[ph] = sq[k]; Cg[ph] = cc[k]; }
          keyA[e] = ka; keyB[e] = kb; flag[e] = f;
        }
        if (roco_sel) bs.add_warp(e < NEl && (f & F_CAND), ka, (uint32_t)rl[k], lsh, lane);
      }
    }
  }
  grp.sync();

  // ---- select across the cluster ------------------------------------------------------------------------------------
  const bool single = evicting && st.policy != EKV_POLICY_RANGE;
  const uint8_t need_flag = st.policy == EKV_POLICY_ROCO ? F_CAND : F_FEAS;
  bool found = false;
  uint32_t l_c = 0;
  int owner = -1, e_c = -1;
  auto local_best = [&](uint8_t need) {
    Tuple128 best; best.hi = ~0ull; best.lo = ~0ull;
    for (int e = tid; e < NEl; e += NCONS) {
      if ((flag[e] & (need | F_REJ)) == need) {
        Tuple128 t;
        t.hi = ((unsigned long long)keyB[e] << 32) | keyA[e];
        t.lo = ((unsigned long long)(uint32_t)lj[e] << 32) | ((uint32_t)rank << 24) | (uint32_t)e;
        if (tuple_less(t, best)) best = t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      Tuple128 t;
      t.hi = __shfl_xor_sync(0xffffffffu, best.hi, o);
      t.lo = __shfl_xor_sync(0xffffffffu, best.lo, o);
      if (tuple_less(t, best)) best = t;
    }
    if (lane == 0) { red64[2 * gw] = best.hi; red64[2 * gw + 1] = best.lo; }
    grp.sync();
    if (tid < C) {                                            // thread p forwards the CTA's best to peer p
      Tuple128 b; b.hi = red64[0]; b.lo = red64[1];
      for (int w = 1; w < NWARP; ++w) {
        Tuple128 t; t.hi = red64[2 * w]; t.lo = red64[2 * w + 1];
        if (tuple_less(t, b)) b = t;
      }
      return b;
    }
    Tuple128 none; none.hi = ~0ull; none.lo = ~0ull;
    return none;
  };
  int attempt = 0;
  constexpr int MAX_TRY = 1;                                    // candidate walks before the radix select (2 cluster barriers each)
  // roco: one candidate attempt (the lowest-mean slot is usually among the k_feasible lowest std), then the bucket select
  // (ekv_bucket.cuh; barrier (3) publishes its histograms) instead of walking on
  const bool bucket = single && roco_sel && NEl < 65535;
  const bool walk = single;
  if (walk) {
    const Tuple128 b = local_best(need_flag);
    if (tid < C) {
      const uint32_t dst = map_to_rank(&xbest[(0 * C + rank) * 2], tid);
      st_cluster_u64(dst, b.hi);
      st_cluster_u64(dst + 8, b.lo);
    }
  }
  stamp(5);
  cluster_sync_all();                                                               // (3) partial outputs + candidates
  {
    // finish this CTA's share of the output: dims [rank*DPC, (rank+1)*DPC) of every head (llama_patch.py:222)
    const int DPC = D / C;
    T* og = reinterpret_cast<T*>(a.out) + (size_t)unit * G * D;
    for (int i = tid; i < G * DPC; i += NCONS) {
      const int g = i / DPC, dd = i % DPC;
      float v = xout[(0 * G + g) * DPC + dd];
      for (int p = 1; p < C; ++p) v += xout[(p * G + g) * DPC + dd];
      og[g * D + rank * DPC + dd] = Tr<T>::from_f(v);
    }
    // rank 0 appends the new row: every CTA of the cluster is past its V stream
    if (rank == 0 && hw == 0) {
      const int slot = ns[0];
      float x[8];
      if (stream_rope) load_row8<T>(reinterpret_cast<const T*>(a.k_new_raw) + (size_t)unit * D, l16, x);     // the cache keeps un-rotated keys
      else load_row8<T>(kh, l16, x);
      store_row8<T>(reinterpret_cast<T*>(a.K) + ((size_t)unit * a.cap + slot) * D, l16, x);
      load_row8<T>(vh, l16, x);
      store_row8<T>(reinterpret_cast<T*>(a.V) + ((size_t)unit * a.cap + slot) * D, l16, x);
    }
  }
  if (walk) {
    while (true) {
      const int pb = attempt & 1;
      Tuple128 best; best.hi = xbest[(pb * C) * 2]; best.lo = xbest[(pb * C) * 2 + 1];
      for (int p = 1; p < C; ++p) {
        Tuple128 t; t.hi = xbest[(pb * C + p) * 2]; t.lo = xbest[(pb * C + p) * 2 + 1];
        if (tuple_less(t, best)) best = t;
      }
      if (best.lo == ~0ull) break;                            // no candidate left
      const uint32_t ka_c = (uint32_t)(best.hi & 0xffffffffu);
      l_c = (uint32_t)(best.lo >> 32);
      owner = (int)((best.lo >> 24) & 0xffu);
      e_c = (int)(best.lo & 0xffffffu);
      if (st.policy != EKV_POLICY_ROCO) { found = true; break; }
      int cnt = 0;                                            // this slice's share of the candidate's std rank
      for (int e = tid; e < NEl; e += NCONS)
        cnt += ((flag[e] & F_CAND) && (keyA[e] < ka_c || (keyA[e] == ka_c && (uint32_t)lj[e] < l_c))) ? 1 : 0;
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) redi[pb * 32 + gw] = cnt;
      grp.sync();
      if (tid < C) {
        int tot = 0;
        for (int w = 0; w < NWARP; ++w) tot += redi[pb * 32 + w];
        st_cluster_u32(map_to_rank(&xcnt[pb * C + rank], tid), (uint32_t)tot);
      }
      cluster_sync_all();                                                           // (4a)
      int rk = 0;
      for (int p = 0; p < C; ++p) rk += xcnt[pb * C + p];
      if (rk < st.k_feasible) { found = true; break; }
      // rejected: the owner drops it and every CTA publishes its next best
      if (rank == owner && tid == 0) flag[e_c] |= F_REJ;
      grp.sync();
      ++attempt;
      if (attempt >= MAX_TRY) {
        bool bucket_done = false;
        if (bucket) {
          // the k_feasible-th smallest std from the histograms of every CTA (read over DSMEM), one pass over the entries, one
          // exchange (per-warp argmins + the cut bucket's entries), exact ranks of the listed entries (easykv.py:322-324)
          auto sync = [&] { grp.sync(); };
          auto ld = [&](const uint32_t* ptr, int peer) -> uint4 {
            if (peer == rank) return *reinterpret_cast<const uint4*>(ptr);
            uint4 v;
            asm volatile("ld.shared::cluster.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(map_to_rank(ptr, peer)) : "memory");
            return v;
          };
          const uint32_t ka_a = smem_u32(keyA), kb_a = smem_u32(keyB), lj_a = smem_u32(lj), fl_a = smem_u32(flag);   // (explicit ld.shared)
          auto get = [&](int e, uint32_t& ka, uint32_t& kb, uint32_t& l) -> bool {
            ka = lds_u32(ka_a + 4u * (uint32_t)e); kb = lds_u32(kb_a + 4u * (uint32_t)e); l = lds_u32(lj_a + 4u * (uint32_t)e);
            return (lds_u8(fl_a + (uint32_t)e) & F_CAND) != 0;
          };
          auto push = [&](unsigned long long* slot, unsigned long long hi, unsigned long long lo2) {     // to every OTHER CTA
            for (int p = 0; p < C; ++p) {
              if (p == rank) continue;
              const uint32_t dst = map_to_rank(slot, p);
              st_cluster_u64(dst, hi);
              st_cluster_u64(dst + 8, lo2);
            }
          };
          Feasibility fz;
          fz.mode = 1; fz.lsh = lsh; fz.T1 = 0u; fz.jT = 0xffffffffu;
          fz.b = bucket_scan<8>(bs, st.k_feasible, C, rank, tid, NCONS, NEl, sync, ld, get, [&] { cluster_sync_all(); });
          if (fz.b.status != bk::FALLBACK) {
            bucket_pass(bs, fz, NEl, rank, tid, NCONS, get, push, sync);
            cluster_sync_all();                                                             // (4) argmins + boundary entries
            Tuple128 wn;
            if (bucket_final(bs, fz, C, tid, NCONS, sync, wn)) {
              l_c = (uint32_t)(wn.lo >> 32); owner = (int)((wn.lo >> 24) & 0xffu); e_c = (int)(wn.lo & 0xffffffu);
              found = true;
            }
            bucket_done = true;
          }
        }
        if (bucket_done) break;
        // (a cut bucket crowded with bit-identical keys falls through to here)
        // The low-mean slots keep falling outside the k_feasible lowest std: stop walking and select properly.
        // Cluster-wide MSB-first radix select (8 bits per pass) of the k-th smallest 64-bit key (std key,
        // logical index) — unique keys, so no tie handling — then one cluster argmin over the feasible set.
        int xpass = 0;                                             // exchange-buffer parity across both selects
        // m-th smallest (1-based) 32-bit key over {e : pred(e)}: threshold T, how many entries equal to T belong
        // to the m smallest (need) and how many equal T (tcount) — ekv_select.cuh's radix_select, with the
        // histogram of every pass summed over the cluster
        auto cluster_radix32 = [&](int m, auto key, auto pred, uint32_t& Tk, int& need, int& tcount) {
          uint32_t prefix = 0u, maskp = 0u;
          int rem = m;
          tcount = 0;
#pragma unroll 1
          for (int shift = 24; shift >= 0; shift -= 8, ++xpass) {
            const int hb = xpass & 1;
            hist[tid] = 0u;                                        // NCONS == 256 bins
            grp.sync();
            for (int e = tid; e < NEl; e += NCONS) {
              if (pred(e)) {
                const uint32_t k = key(e);
                if ((k & maskp) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
              }
            }
            grp.sync();
            {
              const uint32_t v = hist[tid];
              for (int p2 = 0; p2 < C; ++p2) st_cluster_u32(map_to_rank(&xhist[(hb * C + rank) * 256 + tid], p2), v);
            }
            cluster_sync_all();                                                     // (5) one per radix pass
            if (tid < 32) {
              uint32_t loc[8], sum = 0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                uint32_t t = 0;
                for (int p2 = 0; p2 < C; ++p2) t += xhist[(hb * C + p2) * 256 + tid * 8 + j];
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
              if (want > total) want = total;                      // fewer candidates than requested: all of them
              if (want == 0) want = 1;
              if (exc < want && want <= inc) {
                uint32_t cum = exc;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  if (want <= cum + loc[j]) { hmisc[0] = tid * 8 + j; hmisc[1] = (int)cum; hmisc[2] = (int)loc[j]; break; }
                  cum += loc[j];
                }
              }
              if (lane == 0 && total == 0) { hmisc[0] = 255; hmisc[1] = 0; hmisc[2] = 0; }
            }
            grp.sync();
            prefix |= (uint32_t)hmisc[0] << shift;
            maskp |= 255u << shift;
            rem -= hmisc[1];
            tcount = hmisc[2];
            grp.sync();                                            // hmisc / hist are rewritten by the next pass
          }
          Tk = prefix;
          need = rem < tcount ? rem : tcount;
          if (need < 0) need = 0;
        };
        uint32_t T1, jT = 0xffffffffu;
        int need1, tc1;
        cluster_radix32(st.k_feasible, [&](int e) { return keyA[e]; }, [&](int e) { return (flag[e] & F_CAND) != 0; }, T1, need1, tc1);
        if (need1 < tc1) {                                         // the cut falls inside a run of equal std: lowest logical index first
          int nd, tc;
          cluster_radix32(need1, [&](int e) { return (uint32_t)lj[e]; },
                          [&](int e) { return (flag[e] & F_CAND) && keyA[e] == T1; }, jT, nd, tc);
        }
        for (int e = tid; e < NEl; e += NCONS) {
          if (flag[e] & F_CAND) {
            const uint32_t ka = keyA[e];
            const bool feas = ka < T1 || (ka == T1 && (uint32_t)lj[e] <= jT);
            flag[e] = (uint8_t)((flag[e] & ~F_REJ) | (feas ? F_FEAS : 0));
          }
        }
        grp.sync();
        const Tuple128 b = local_best((uint8_t)F_FEAS);
        if (tid < C) {
          const uint32_t dst = map_to_rank(&xbest[((attempt & 1) * C + rank) * 2], tid);
          st_cluster_u64(dst, b.hi);
          st_cluster_u64(dst + 8, b.lo);
        }
        cluster_sync_all();                                                         // (6)
        const int pb2 = attempt & 1;
        Tuple128 best2; best2.hi = xbest[(pb2 * C) * 2]; best2.lo = xbest[(pb2 * C) * 2 + 1];
        for (int p2 = 1; p2 < C; ++p2) {
          Tuple128 t; t.hi = xbest[(pb2 * C + p2) * 2]; t.lo = xbest[(pb2 * C + p2) * 2 + 1];
          if (tuple_less(t, best2)) best2 = t;
        }
        if (best2.lo != ~0ull) {
          l_c = (uint32_t)(best2.lo >> 32);
          owner = (int)((best2.lo >> 24) & 0xffu);
          e_c = (int)(best2.lo & 0xffffffu);
          found = true;
        }
        break;
      }
      const Tuple128 b = local_best(need_flag);
      if (tid < C) {
        const uint32_t dst = map_to_rank(&xbest[((attempt & 1) * C + rank) * 2], tid);
        st_cluster_u64(dst, b.hi);
        st_cluster_u64(dst + 8, b.lo);
      }
      cluster_sync_all();                                                           // (4b)
    }
  } else if (evicting && !single) {                           // RANGE with one victim: positional
    l_c = (uint32_t)(P + st.range_start);
    found = true;
  }

  stamp(6);
  if (tl && tid == 0) tl[7] = (unsigned long long)attempt;
  // ---- apply: renumber this slice, free the victim's slot, publish the new slot ----------------------------------------
  const bool is_range = evicting && st.policy == EKV_POLICY_RANGE;
  for (int e = tid; e < NEl; e += NCONS) {
    const int l = lj[e];
    if (l < 0) continue;
    const int phys = phys_of(e);
    const bool victim = found && (is_range ? (uint32_t)l == l_c : (rank == owner && e == e_c));
    if (victim) {
      if (a.victim_lidx) a.victim_lidx[unit] = l;
      if (a.victim_slots) a.victim_slots[unit] = phys;
      if (st.apply) lidx_g[phys] = -1;
      else if (e >= nloc) lidx_g[phys] = l;
    } else if (found && st.apply && (uint32_t)l > l_c) {
      lidx_g[phys] = l - 1;
    } else if (e >= nloc) {
      lidx_g[phys] = l;
    }
  }
  if (((evicting && !found) || (a.st.evict > 0 && stu.evict == 0)) && rank == 0 && tid == 0) {   // no candidate (degenerate), or below the budget gate
    if (a.victim_lidx) a.victim_lidx[unit] = -1;
    if (a.victim_slots) a.victim_slots[unit] = -1;
  }
  // peers may still be reading what this CTA was sent, never what it owns: no remote access follows, but a
  // CTA must not exit while others can still write into it
  cluster_sync_all();
}

// ---------------------------------------------------------------------------------------------------
template <typename T, int G, bool TC>
static int launch_cluster_tg(const KernelArgs& a, int C, int slice, int stages, int smem_bytes, cudaStream_t stream) {
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  cudaError_t err;
  if (!configured[dev]) {
    err = cudaFuncSetAttribute(decode_cluster_kernel<T, G, TC>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(decode_cluster)", err);
    configured[dev] = 1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(a.B * a.Hkv * C), 1, 1);
  cfg.blockDim = dim3(DecodeCfg<T>::NCONS + 32, 1, 1);
  cfg.dynamicSmemBytes = (size_t)smem_bytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;      // the kernel waits (pdl_wait) before its first global access
  attr[1].val.programmaticStreamSerializationAllowed = pdl_allowed();
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  err = cudaLaunchKernelEx(&cfg, decode_cluster_kernel<T, G, TC>, a, stages, slice);
  if (err != cudaSuccess) return set_cuda_error("decode_cluster_kernel launch", err);
  count_launch();
  return EKV_OK;
}

// Chooses the cluster size.  FMA variant: the smallest power of two for which a slice fits one CTA, doubled
// while the grid still fits the chip in one wave (more CTAs = more of the chip's HBM bandwidth in flight).
// Tensor-core variant: slices of at most ~1.5K slots so that a CTA needs less than half an SM's shared memory
// and two CTAs (of different clusters) overlap each other's barrier and tail latencies.
template <typename T, int G, bool TC>
static int plan_cluster(const KernelArgs& a, int sms, int force_c, int& C, int& slice, int& stages, int& smem) {
  using Cfg = DecodeCfg<T>;
  constexpr int TILEB = TC ? TC_TILE_BYTES : Cfg::TILE_BYTES;
  const int U = a.B * a.Hkv, sm_total = 227 * 1024, half_sm = 112 * 1024;
  auto fits = [&](int c, int& sl, int& stg, int& bytes) {
    sl = align_up((a.n_phys + c - 1) / c, 8);
    if (sl < 8) sl = 8;
    const ClusterSmem<T> L(G, sl, c, TC);
    const int min_ring = ClusterSmem<T>::min_ring(G) > 3 * TILEB ? ClusterSmem<T>::min_ring(G) : 3 * TILEB;
    int ring = sm_total - L.fixed;
    if (ring < min_ring) return false;
    // more CTAs than SMs: leave room for a second CTA per SM so that one CTA's barrier / tail latencies overlap
    // the other's streaming (the FMA variant with g = 8 needs too many registers for that)
    if ((TC || (G <= 4 && U * c > sms)) && half_sm - L.fixed >= min_ring) ring = half_sm - L.fixed;
    stg = ring / TILEB;
    if (stg > Cfg::MAX_STAGES) stg = Cfg::MAX_STAGES;
    const int tiles = 2 * ((sl + Cfg::TILE_ROWS - 1) / Cfg::TILE_ROWS);
    if (stg > tiles) stg = tiles < 3 ? 3 : tiles;
    bytes = L.fixed + stg * TILEB;
    if (bytes < L.fixed + ClusterSmem<T>::min_ring(G)) bytes = L.fixed + ClusterSmem<T>::min_ring(G);
    return true;
  };
  int best = 0, best_fixed = 0;
  for (int c = 1; c <= 8; c *= 2) {
    if (force_c && c != force_c) continue;
    int sl, stg, bytes;
    if (!fits(c, sl, stg, bytes)) continue;
    if (best && !force_c) {
      if (TC) {
        if (slice <= 1536 && (U * c > 2 * sms || sl < 4 * Cfg::TILE_ROWS)) break;   // small enough already; keep slices non-trivial
      } else {
        // keep one wave and non-trivial slices — unless a CTA still needs more than half an SM's shared memory
        // while there are more CTAs than SMs: then halving the slice again lets two CTAs share an SM
        const bool crowded = G <= 4 && U * best * 2 > sms && best_fixed > 64 * 1024;
        if (!crowded && (U * c > sms || sl < 2 * Cfg::TILE_ROWS)) break;
      }
    }
    best = c; C = c; slice = sl; stages = stg; smem = bytes;
    best_fixed = ClusterSmem<T>(G, sl, c, TC).fixed;
  }
  return best ? EKV_OK : EKV_ERR_UNSUPPORTED;
}

// largest n_phys a decode step can hold on this kernel (clusters of 8, FMA variant): ekv_chunk_entry_limit
template <typename T> static int cluster_limit_t(int G) {
  using Cfg = DecodeCfg<T>;
  int lo = 0, hi = 1 << 22;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    int sl = align_up((mid + 7) / 8, 8);
    if (sl < 8) sl = 8;
    const ClusterSmem<T> L(G, sl, 8, false);
    const int min_ring = ClusterSmem<T>::min_ring(G) > 3 * Cfg::TILE_BYTES ? ClusterSmem<T>::min_ring(G) : 3 * Cfg::TILE_BYTES;
    if (227 * 1024 - L.fixed >= min_ring) lo = mid; else hi = mid - 1;
  }
  return lo;
}
int decode_cluster_entry_limit(int dtype, int G) {
  switch (dtype) {
    case EKV_F16: return cluster_limit_t<__half>(G);
    case EKV_BF16: return cluster_limit_t<__nv_bfloat16>(G);
    case EKV_F32: return cluster_limit_t<float>(G);
    default: return 0;
  }
}

int decode_cluster_size();   // ekv_api.cu (env EKV_DECODE_CLUSTER): 0 = automatic, else forced cluster size
int decode_variant();        // ekv_api.cu: 3 = never / 4 = always use the tensor-core variant (g >= 4, 16-bit)

template <typename T, int G, bool TC>
static int launch_cluster_plan_v(const KernelArgs& a, bool only_if_better, cudaStream_t stream) {
  static thread_local int sm_count[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (!sm_count[dev]) {
    cudaError_t err = cudaDeviceGetAttribute(&sm_count[dev], cudaDevAttrMultiProcessorCount, dev);
    if (err != cudaSuccess) return set_cuda_error("cudaDeviceGetAttribute", err);
  }
  int C = 0, slice = 0, stages = 0, smem = 0;
  const int rc = plan_cluster<T, G, TC>(a, sm_count[dev], decode_cluster_size(), C, slice, stages, smem);
  if (rc) return rc;
  if (only_if_better && C == 1 && !TC) return EKV_ERR_UNSUPPORTED;   // the single-CTA kernels serve this shape
  return launch_cluster_tg<T, G, TC>(a, C, slice, stages, smem, stream);
}

template <typename T, int G> static int launch_cluster_plan(const KernelArgs& a, bool only_if_better, cudaStream_t stream) {
  // tensor-core variant: by default for g = 8 (measured 1.15-1.85x over the FMA variant, which is FP32-bound there);
  // for g = 4 the FMA variant keeps up with HBM and is kept (decode_variant 4 forces the tensor cores, 3 forbids them)
  if constexpr (sizeof(T) == 2 && G >= 4) {
    const int v = decode_variant() >= 5 ? 0 : decode_variant();
    if (!a.rope_cos && (v == 4 || (v != 3 && G >= 8))) return launch_cluster_plan_v<T, G, true>(a, only_if_better, stream);   // (fused streaming: FMA path)
  }
  return launch_cluster_plan_v<T, G, false>(a, only_if_better, stream);
}

template <typename T> static int launch_cluster_t(const KernelArgs& a, bool only_if_better, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_cluster_plan<T, 1>(a, only_if_better, stream);
    case 2: return launch_cluster_plan<T, 2>(a, only_if_better, stream);
    case 4: return launch_cluster_plan<T, 4>(a, only_if_better, stream);
    case 8: return launch_cluster_plan<T, 8>(a, only_if_better, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

// `only_if_better`: decline (EKV_ERR_UNSUPPORTED) when the plan degenerates to one CTA per unit.
int launch_decode_cluster(const KernelArgs& a, bool only_if_better, cudaStream_t stream) {
  if (a.q_len != 1 || a.d != 128 || a.st.tova_head_mean || a.st.evict > 1) return EKV_ERR_UNSUPPORTED;
  if (a.rope_cos && a.dtype == EKV_F32) return EKV_ERR_UNSUPPORTED;
  switch (a.dtype) {
    case EKV_F16: return launch_cluster_t<__half>(a, only_if_better, stream);
    case EKV_BF16: return launch_cluster_t<__nv_bfloat16>(a, only_if_better, stream);
    case EKV_F32: return launch_cluster_t<float>(a, only_if_better, stream);
    default: return EKV_ERR_INVALID;
  }
}

}  // namespace ekv
