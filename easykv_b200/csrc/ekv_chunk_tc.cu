// Strided-prefill chunk (q_len = stride query rows per head) for 16-bit dtypes on the tensor cores.
//
// Here Q·K^T is a real dense contraction: R = q_len * g (query, head) rows share one pass over a kv head's
// keys, arithmetic intensity g*q_len flop/B (64 for Mistral stride 16 or 7B stride 64, 512 for 70B stride
// 64) — far beyond what the FP32 pipes sustain at HBM speed (the CUDA-core general kernel in ekv_chunk.cu
// is FP32-issue-bound by three orders of magnitude on these shapes), so the contraction runs as
// mma.sync.m16n8k16 (HMMA, fp32 accumulate) out of shared memory.
//
// A chunk forward of one layer is four launches:
//   1. chunk_tc_kernel<PASS=1>  grid (unit, row block of 64 rows, key split): streams the split's K tiles
//      (cp.async, 3 stages), S = Q K^T, logits rounded at the reference's rounding points, per-row online
//      (max, sum of exp) -> scratch.
//   2. chunk_tc_kernel<PASS=2>  same grid: combines the splits' row statistics, recomputes S (K now comes
//      from L2), p = model-dtype(exp(x - max) / sum) exactly as softmax forms it, O_partial = P V with P fed
//      to the tensor cores straight from the accumulator registers, and the per-key column statistics
//      (GQA fold in the model dtype, p and p^2 summed over the chunk's queries) -> scratch.
//   3. chunk_out_kernel         sums the splits' partial outputs in split order -> out; and, one thread per
//      (unit, slot): folds the column statistics into the policy state and prepares the selection keys.
//   4. chunk_tail_kernel        one CTA per unit: appends the chunk's K/V rows and runs the budgeted select /
//      eviction (ekv_select.cuh) on those keys.
// The [H, q, n] probability tensor the reference materialises per layer (easykv/llama_patch.py:244-246)
// never exists; nothing is synchronised with the host.
//
// Replaces: llama_patch.py:193-230 / mistral_patch.py:137-170 and easykv.py:439-499 / :599-661 / :830-892
// for one layer of one strided forward (and the dense prefill issued as causal chunks).
#include "ekv_mma.cuh"
#include "ekv_select.cuh"
#include "ekv_kernels.h"
#include "ekv_chunk_plan.h"

namespace ekv {

namespace tc {
constexpr int D = 128;
constexpr int TK = 64;               // keys per tile
constexpr int MR = 64;               // (query, head) rows per CTA = 4 warps x 16
constexpr int NT = 128;
constexpr int PITCH = 272;           // bytes per shared-memory row: 256 + 16 keeps ldmatrix conflict-free
constexpr int TILE_BYTES = TK * PITCH;
constexpr int STAGES = 3;            // pass 1; pass 2 (K and V tiles) uses 2 so that two CTAs fit an SM
constexpr float NEG = -1.0e30f;      // masked logit: finite, exp(NEG - max) == 0 exactly
constexpr int TARGET_CTAS_PASS1 = 444;  // three lighter CTAs per SM in pass 1
constexpr int TARGET_CTAS = 296;      // two CTAs per SM: their dependency stalls overlap
}  // namespace tc

ChunkPlan make_chunk_plan(int B, int Hkv, int G, int q_len, int n_phys) {
  using namespace tc;
  ChunkPlan p = {};
  p.R = q_len * G;
  p.RB = (p.R + MR - 1) / MR;
  p.Rpad = p.RB * MR;
  p.NE = n_phys + q_len;
  p.ntiles = (p.NE + TK - 1) / TK;
  p.NEpad = p.ntiles * TK;
  const int U = B * Hkv;
  int want = TARGET_CTAS / (U * p.RB);
  if (want < 1) want = 1;
  if (want > p.ntiles) want = p.ntiles;
  p.tps = (p.ntiles + want - 1) / want;
  if (p.tps < 2 && p.ntiles >= 2) p.tps = 2;             // at least two tiles per CTA: keep the pipeline worth its prologue
  p.splits = (p.ntiles + p.tps - 1) / p.tps;
  int want1 = TARGET_CTAS_PASS1 / (U * p.RB);
  if (want1 < 1) want1 = 1;
  if (want1 > p.ntiles) want1 = p.ntiles;
  p.tps1 = (p.ntiles + want1 - 1) / want1;
  if (p.tps1 < 2 && p.ntiles >= 2) p.tps1 = 2;
  p.splits1 = (p.ntiles + p.tps1 - 1) / p.tps1;
  long long o = 0;
  p.off_stats = o; o += (long long)U * p.splits1 * p.Rpad * 2 * 4;
  o = (o + 255) / 256 * 256;
  p.off_opart = o; o += (long long)U * p.splits * p.Rpad * D * 4;
  o = (o + 255) / 256 * 256;
  p.cparts = 2;
  p.off_cpart = o; o += (long long)U * (2 * p.RB) * p.NEpad * 2 * 4;      // two query halves per row block
  o = (o + 255) / 256 * 256;
  p.off_klj = o; o += (long long)U * p.NEpad * 4;                          // per entry: logical index, the two selection
  p.off_ka = o; o += (long long)U * p.NEpad * 4;                           // keys and the candidate flags (written by the
  p.off_kb = o; o += (long long)U * p.NEpad * 4;                           // chip-wide state update, read by the tail)
  p.off_kf = o; o += (long long)U * p.NEpad;
  p.bytes = (o + 255) / 256 * 256;
  return p;
}

long long chunk_tc_scratch_bytes(int B, int Hkv, int G, int q_len, int n_phys) {
  return make_chunk_plan(B, Hkv, G, q_len, n_phys).bytes;
}

// ---- passes 1 and 2 ---------------------------------------------------------------------------------------------------
// ARITH: which ATen flavour's two non-associative spots to reproduce (ekv_step.arith) — a template parameter so
// that the per-element code is straight-line (as a run-time switch it cost a branch per element and the
// IEEE-division slow-path calls of the other flavour in the instruction stream).
template <typename T, int G, int PASS, bool ARITH>
__global__ void __launch_bounds__(tc::NT, PASS == 1 ? 3 : 2) chunk_tc_kernel(const KernelArgs a, const ChunkPlan pl) {
  using namespace tc;
  constexpr int STAGES = PASS == 2 ? 2 : tc::STAGES;
  extern __shared__ __align__(128) unsigned char smem[];
  unsigned char* Qs = smem;                                       // [MR][PITCH]
  unsigned char* Ks = Qs + MR * PITCH;                            // [STAGES][TK][PITCH]
  unsigned char* Vs = Ks + STAGES * TILE_BYTES;                   // [STAGES][TK][PITCH]   (pass 2)
  int32_t* ljs = reinterpret_cast<int32_t*>(Vs + (PASS == 2 ? STAGES * TILE_BYTES : 0));   // [STAGES][TK]
  T* pfs = reinterpret_cast<T*>(ljs + STAGES * TK);               // [MR / G][PFP] folded probabilities (pass 2)
  constexpr int PFP = TK + 2;                                     // row pitch in elements: conflict-free both ways

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nsplit = PASS == 1 ? pl.splits1 : pl.splits, tps = PASS == 1 ? pl.tps1 : pl.tps;
  const int split = blockIdx.x % nsplit;
  const int rb = (blockIdx.x / nsplit) % pl.RB;
  const int unit = blockIdx.x / (nsplit * pl.RB);
  const int b = unit / a.Hkv, h = unit % a.Hkv;
  const int QL = a.q_len, n_phys = a.n_phys, NE = pl.NE, R = pl.R;
  const int t_begin = split * tps, t_end = min(pl.ntiles, t_begin + tps);
  const T* Kg = reinterpret_cast<const T*>(a.K) + (size_t)unit * a.cap * D;
  const T* Vg = reinterpret_cast<const T*>(a.V) + (size_t)unit * a.cap * D;
  const T* kn = reinterpret_cast<const T*>(a.k_new) + (size_t)unit * QL * D;
  const T* vn = reinterpret_cast<const T*>(a.v_new) + (size_t)unit * QL * D;
  const T* qg = reinterpret_cast<const T*>(a.q) + (size_t)b * a.H * QL * D;
  const int32_t* lg = a.lidx + (size_t)unit * a.cap;

  auto issue_tile = [&](int tile, int stage) {
    unsigned char* kd = Ks + stage * TILE_BYTES;
    unsigned char* vd = Vs + stage * TILE_BYTES;
    if ((tile + 1) * TK <= n_phys) {                              // a tile of cached rows: straight copies
      const int row = tid >> 4, c = tid & 15;
      const unsigned char* ks = reinterpret_cast<const unsigned char*>(Kg + (size_t)(tile * TK + row) * D) + c * 16;
      const unsigned char* vs = reinterpret_cast<const unsigned char*>(Vg + (size_t)(tile * TK + row) * D) + c * 16;
      unsigned char* kdd = kd + row * PITCH + c * 16;
      unsigned char* vdd = vd + row * PITCH + c * 16;
#pragma unroll
      for (int j = 0; j < TK * 16 / NT; ++j) {                    // 8 rows further per step
        cp_async16(kdd + j * 8 * PITCH, ks + (size_t)j * 8 * D * sizeof(T));
        if (PASS == 2) cp_async16(vdd + j * 8 * PITCH, vs + (size_t)j * 8 * D * sizeof(T));
      }
      if (tid < TK) cp_async4(&ljs[stage * TK + tid], lg + tile * TK + tid);
      return;
    }
#pragma unroll
    for (int j = 0; j < TK * 16 / NT; ++j) {
      const int idx = tid + j * NT, row = idx >> 4, c = idx & 15;
      const int e = tile * TK + row;
      if (e < NE) {
        const T* ks = e < n_phys ? Kg + (size_t)e * D : kn + (size_t)(e - n_phys) * D;
        cp_async16(kd + row * PITCH + c * 16, reinterpret_cast<const unsigned char*>(ks) + c * 16);
        if (PASS == 2) {
          const T* vs = e < n_phys ? Vg + (size_t)e * D : vn + (size_t)(e - n_phys) * D;
          cp_async16(vd + row * PITCH + c * 16, reinterpret_cast<const unsigned char*>(vs) + c * 16);
        }
      } else {
        *reinterpret_cast<uint4*>(kd + row * PITCH + c * 16) = make_uint4(0, 0, 0, 0);
        if (PASS == 2) *reinterpret_cast<uint4*>(vd + row * PITCH + c * 16) = make_uint4(0, 0, 0, 0);
      }
    }
    if (tid < TK) {
      const int e = tile * TK + tid;
      if (e < n_phys) cp_async4(&ljs[stage * TK + tid], lg + e);
      else ljs[stage * TK + tid] = e < NE ? a.n_before + (e - n_phys) : -1;
    }
  };

  // ---- Q block -> shared memory -> A fragments ---------------------------------------------------------------------
#pragma unroll
  for (int j = 0; j < MR * 16 / NT; ++j) {
    const int idx = tid + j * NT, rr = idx >> 4, c = idx & 15;
    const int r = rb * MR + rr;
    if (r < R) {
      const int i = r / G, g = r % G;                             // query-major rows: the g heads of a query are adjacent
      cp_async16(Qs + rr * PITCH + c * 16,
                 reinterpret_cast<const unsigned char*>(qg + ((size_t)(h * G + g) * QL + i) * D) + c * 16);
    } else {
      *reinterpret_cast<uint4*>(Qs + rr * PITCH + c * 16) = make_uint4(0, 0, 0, 0);
    }
  }
  cp_async_commit();
  for (int s = 0; s < STAGES - 1; ++s) {
    if (t_begin + s < t_end) issue_tile(t_begin + s, s);
    cp_async_commit();
  }
  cp_async_wait<STAGES - 1>();                                    // the Q group
  __syncthreads();
  uint32_t aq[8][4];
  {
    const int row = warp * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
    const int colb = (lane >> 4) * 16;                            // second pair of matrices: dims +8
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) ldsm_x4(smem_u32(Qs + row * PITCH + ks * 32 + colb), aq[ks]);
  }

  // the two rows of this thread (C-fragment layout): r0 = warp*16 + lane/4, r1 = r0 + 8
  const int rr0 = warp * 16 + (lane >> 2), rr1 = rr0 + 8;
  const int row0 = rb * MR + rr0, row1 = rb * MR + rr1;
  const int qi0 = row0 / G, qi1 = row1 / G;                       // query index inside the chunk

  // ---- pass 2: the rows' softmax statistics over ALL keys (combine the splits of pass 1) ---------------------------------
  const float2* stats = reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(a.scratch) + pl.off_stats);
  float M0 = 0.f, M1 = 0.f, L0 = 1.f, L1 = 1.f, R0 = 1.f, R1 = 1.f;
  if (PASS == 2) {
    float m0 = NEG, m1 = NEG;
    for (int s = 0; s < pl.splits1; ++s) {
      const float2* st = stats + ((size_t)unit * pl.splits1 + s) * pl.Rpad;
      m0 = fmaxf(m0, st[rb * MR + rr0].x); m1 = fmaxf(m1, st[rb * MR + rr1].x);
    }
    float l0 = 0.f, l1 = 0.f;
    for (int s = 0; s < pl.splits1; ++s) {                        // split order on every CTA
      const float2* st = stats + ((size_t)unit * pl.splits1 + s) * pl.Rpad;
      const float2 x0 = st[rb * MR + rr0], x1 = st[rb * MR + rr1];
      l0 += x0.y * expf(x0.x - m0);
      l1 += x1.y * expf(x1.x - m1);
    }
    M0 = m0; M1 = m1;
    if (l0 == 0.f) l0 = 1.f;                                      // padding rows (row >= R): p = 0
    if (l1 == 0.f) l1 = 1.f;
    L0 = ARITH ? l0 : __fdiv_rn(1.0f, l0);
    L1 = ARITH ? l1 : __fdiv_rn(1.0f, l1);
    R0 = __frcp_rn(l0); R1 = __frcp_rn(l1);
  }

  float mrun0 = NEG, mrun1 = NEG, lrun0 = 0.f, lrun1 = 0.f;      // pass 1: this thread's columns only
  float o[16][4];
  if (PASS == 2) {
#pragma unroll
    for (int nb = 0; nb < 16; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[nb][c] = 0.f;
  }
  const bool tova = a.st.policy == EKV_POLICY_TOVA;
  const bool want_stats = PASS == 2 && a.st.accumulate != 0;
  const float inv_g = 1.0f / (float)G;

  for (int tile = t_begin; tile < t_end; ++tile) {
    const int it = tile - t_begin, stage = it % STAGES;
    cp_async_wait<STAGES - 2>();
    __syncthreads();                                              // tile `tile` has landed; slot (it-1)%STAGES is free
    {
      const int nxt = tile + STAGES - 1;
      if (nxt < t_end) issue_tile(nxt, (it + STAGES - 1) % STAGES);
      cp_async_commit();
    }
    const unsigned char* kt = Ks + stage * TILE_BYTES;
    const int32_t* lt = ljs + stage * TK;

    // ---- S = Q K^T ------------------------------------------------------------------------------------------------
    float s[8][4];
#pragma unroll
    for (int nb = 0; nb < 8; ++nb)
#pragma unroll
      for (int c = 0; c < 4; ++c) s[nb][c] = 0.f;
    {
      const int key_l = (lane & 7) + (lane >> 4) * 8;             // matrices 2,3: next 8 keys
      const int dim_l = ((lane >> 3) & 1) * 16;                   // matrices 1,3: dims +8 (bytes)
#pragma unroll
      for (int ks = 0; ks < 8; ++ks)
#pragma unroll
        for (int nbp = 0; nbp < 4; ++nbp) {
          uint32_t bk[4];
          ldsm_x4(smem_u32(kt + (nbp * 16 + key_l) * PITCH + ks * 32 + dim_l), bk);
          mma16816<T>(s[2 * nbp], aq[ks], bk[0], bk[1]);
          mma16816<T>(s[2 * nbp + 1], aq[ks], bk[2], bk[3]);
        }
    }
    // ---- logits at the reference's rounding points (llama_patch.py:201-202) ---------------------------------------------
#pragma unroll
    for (int nb = 0; nb < 8; ++nb) {
      round2<T>(s[nb][0], s[nb][1]);
      round2<T>(s[nb][2], s[nb][3]);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s[nb][c] = ARITH ? __fmul_rn(s[nb][c], a.scale_mul) : __fdiv_rn(s[nb][c], a.scale_div);
      round2<T>(s[nb][0], s[nb][1]);
      round2<T>(s[nb][2], s[nb][3]);
    }
    // ---- mask (:210-215): free slots, and causality among the chunk's own keys.  Masked logits become a huge
    // negative FINITE value: exp(x - max) is exactly 0 as with the reference's finfo.min, and no inf - inf can
    // arise.  Padding rows (row >= R) are left alone: nothing reads them.
    {
      const bool cached_tile = (tile + 1) * TK <= n_phys;
      const bool clean = cached_tile && !__any_sync(0xffffffffu, (lt[lane] | lt[lane + 32]) < 0);
      if (!clean) {
        const int new0 = tile * TK - n_phys;                      // index of the tile's first key among the chunk's keys
#pragma unroll
        for (int nb = 0; nb < 8; ++nb)
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            const int col = nb * 8 + 2 * (lane & 3) + cc;
            const int jn = new0 + col;
            const bool ok = jn < 0 ? lt[col] >= 0 : tile * TK + col < NE;
            if (!(ok && (jn < 0 || jn <= qi0))) s[nb][cc] = NEG;
            if (!(ok && (jn < 0 || jn <= qi1))) s[nb][2 + cc] = NEG;
          }
      }
    }

    if (PASS == 1) {
      float tm0 = NEG, tm1 = NEG;
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        tm0 = fmaxf(tm0, fmaxf(s[nb][0], s[nb][1]));
        tm1 = fmaxf(tm1, fmaxf(s[nb][2], s[nb][3]));
      }
      const float n0 = fmaxf(mrun0, tm0), n1 = fmaxf(mrun1, tm1);
      float acc0 = lrun0 * expf(mrun0 - n0), acc1 = lrun1 * expf(mrun1 - n1);
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
        acc0 += expf(s[nb][0] - n0) + expf(s[nb][1] - n0);
        acc1 += expf(s[nb][2] - n1) + expf(s[nb][3] - n1);
      }
      lrun0 = acc0; mrun0 = n0; lrun1 = acc1; mrun1 = n1;
    } else {
      // ---- probabilities (llama_patch.py:218-219) -------------------------------------------------------------------
#pragma unroll
      for (int nb = 0; nb < 8; ++nb) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const bool hi = c >= 2;
          const float x = s[nb][c];
          const float ex = expf(x - (hi ? M1 : M0));
          s[nb][c] = ARITH ? div_rn_by(ex, hi ? L1 : L0, hi ? R1 : R0) : __fmul_rn(ex, hi ? L1 : L0);
        }
        round2<T>(s[nb][0], s[nb][1]);
        round2<T>(s[nb][2], s[nb][3]);
      }
      // ---- O += P V: the C fragments of S are the A fragments of P ------------------------------------------------------
      const unsigned char* vt = Vs + stage * TILE_BYTES;
      {
        const int key_l = (lane & 7) + ((lane >> 3) & 1) * 8;     // matrices 1,3: keys +8
        const int dim_l = (lane >> 4) * 16;                       // matrices 2,3: dims +8 (bytes)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pa[4];
          pa[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
          pa[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
          pa[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
          pa[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
          for (int dbp = 0; dbp < 8; ++dbp) {
            uint32_t bv[4];
            ldsm_x4_trans(smem_u32(vt + (kk * 16 + key_l) * PITCH + dbp * 32 + dim_l), bv);
            mma16816<T>(o[2 * dbp], pa, bv[0], bv[1]);
            mma16816<T>(o[2 * dbp + 1], pa, bv[2], bv[3]);
          }
        }
      }
      // ---- per-key column statistics of this tile -------------------------------------------------------------------------
      if (want_stats) {
        // fold the g heads of a query (adjacent rows = lanes 4 apart) and round to the model dtype
        // (process_for_mqa_gqa, easykv.py:188-196); the lead lane of each query parks the folded row in shared
        // memory, then every thread sums one key column over half of the CTA's queries: p and
        // model-dtype(p^2), the chunk's row sums of easykv.py:450-451.
        constexpr int QW = 16 / G;                                // queries per warp
        const bool lead = ((lane >> 2) % G) == 0;
        const int qlo = warp * QW + (lane >> 2) / G, qhi = qlo + 8 / G;
#pragma unroll
        for (int nb = 0; nb < 8; ++nb) {
          float l0 = s[nb][0], l1 = s[nb][1], h0 = s[nb][2], h1 = s[nb][3];
          if (G > 1) {
#pragma unroll
            for (int off = 4; off < 4 * G; off <<= 1) {
              l0 += __shfl_xor_sync(0xffffffffu, l0, off); l1 += __shfl_xor_sync(0xffffffffu, l1, off);
              h0 += __shfl_xor_sync(0xffffffffu, h0, off); h1 += __shfl_xor_sync(0xffffffffu, h1, off);
            }
            l0 = __fmul_rn(l0, inv_g); l1 = __fmul_rn(l1, inv_g); h0 = __fmul_rn(h0, inv_g); h1 = __fmul_rn(h1, inv_g);
          }
          if (lead) {                                             // pack2 rounds to the model dtype
            const int col = nb * 8 + 2 * (lane & 3);
            *reinterpret_cast<uint32_t*>(&pfs[qlo * PFP + col]) = pack2<T>(l0, l1);
            if (G < 16) *reinterpret_cast<uint32_t*>(&pfs[qhi * PFP + col]) = pack2<T>(h0, h1);
          }
        }
        __syncthreads();
        {
          constexpr int QR = MR / G;                              // queries of this CTA
          constexpr int QH = QR >= 2 ? QR / 2 : 1;                // per thread: half of them
          const int col = tid & (TK - 1), part = tid >> 6;
          const int e = tile * TK + col;
          float cs = 0.f, csq = 0.f;
          if (part * QH < QR) {
#pragma unroll 8
            for (int qq = 0; qq < QH; ++qq) {
              const int ql = part * QH + qq;                      // query row inside the CTA
              const int qi = rb * QR + ql;                        // query index inside the chunk
              if (qi < QL && (!tova || qi == QL - 1)) {
                const float v = Tr<T>::to_f(pfs[ql * PFP + col]);
                cs += v;
                csq += Tr<T>::round_f(__fmul_rn(v, v));
              }
            }
          }
          if (e < NE) {
            float2* cg = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(a.scratch) + pl.off_cpart);
            cg[((size_t)unit * (2 * pl.RB) + 2 * rb + part) * pl.NEpad + e] = make_float2(cs, csq);
          }
        }
      }
    }
  }
  cp_async_wait<0>();

  if (PASS == 1) {
    // combine the four lanes that share a row, then one lane per row writes (max, sum)
#pragma unroll
    for (int off = 1; off < 4; off <<= 1) {
      const float om0 = __shfl_xor_sync(0xffffffffu, mrun0, off), ol0 = __shfl_xor_sync(0xffffffffu, lrun0, off);
      const float om1 = __shfl_xor_sync(0xffffffffu, mrun1, off), ol1 = __shfl_xor_sync(0xffffffffu, lrun1, off);
      const float n0 = fmaxf(mrun0, om0), n1 = fmaxf(mrun1, om1);
      lrun0 = lrun0 * expf(mrun0 - n0) + ol0 * expf(om0 - n0);
      lrun1 = lrun1 * expf(mrun1 - n1) + ol1 * expf(om1 - n1);
      mrun0 = n0; mrun1 = n1;
    }
    if ((lane & 3) == 0) {
      float2* st = reinterpret_cast<float2*>(reinterpret_cast<unsigned char*>(a.scratch) + pl.off_stats) +
                   ((size_t)unit * pl.splits1 + split) * pl.Rpad;
      st[rb * MR + rr0] = make_float2(mrun0, lrun0);
      st[rb * MR + rr1] = make_float2(mrun1, lrun1);
    }
  } else {
    float* op = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.scratch) + pl.off_opart) +
                ((size_t)unit * pl.splits + split) * pl.Rpad * D;
#pragma unroll
    for (int nb = 0; nb < 16; ++nb) {
      const int dim = nb * 8 + 2 * (lane & 3);
      *reinterpret_cast<float2*>(op + (size_t)(rb * MR + rr0) * D + dim) = make_float2(o[nb][0], o[nb][1]);
      *reinterpret_cast<float2*>(op + (size_t)(rb * MR + rr1) * D + dim) = make_float2(o[nb][2], o[nb][3]);
    }
  }
}

// ---- 3. out = sum over splits of the partial outputs (llama_patch.py:222) ------------------------------------------------------
template <typename T, int G> __global__ void chunk_out_kernel(const KernelArgs a, const ChunkPlan pl, const int out_blocks) {
  using namespace tc;
  pdl_trigger();            // programmatic dependent launch (ekv_common.cuh)
  pdl_wait();
  if ((int)blockIdx.x >= out_blocks) {
    // the second part of the grid, one thread per (unit, entry): column statistics summed over the row blocks
    // (in row-block order), folded into the policy state (accumulate, counter), selection keys for the tail
    const int idx = (blockIdx.x - out_blocks) * blockDim.x + threadIdx.x;
    const int e = idx % pl.NEpad, unit = idx / pl.NEpad;
    if (unit >= a.B * a.Hkv || e >= pl.NE) return;
    const ekv_step& st = a.st;
    const int n_phys = a.n_phys, P = st.score_offset, n_s = a.n_before + a.q_len - P;
    const bool is_new = e >= n_phys;
    int rl, phys = e;
    float sv = 0.f, sq = 0.f, cc = 1.f;
    if (is_new) {
      rl = a.n_before + (e - n_phys);
      cc = __fsub_rn(st.c_new0, __fmul_rn((float)(e - n_phys), st.c_new_step));
      phys = a.new_slots ? a.new_slots[(size_t)unit * a.q_len + (e - n_phys)] : e;
    } else {
      rl = a.lidx[(size_t)unit * a.cap + e];
      if (rl >= P) { sv = a.S[(size_t)unit * a.cap + e]; sq = a.SQ[(size_t)unit * a.cap + e]; cc = a.C[(size_t)unit * a.cap + e]; }
    }
    float ds = 0.f, dsq = 0.f;
    if (st.accumulate) {
      const float2* cg = reinterpret_cast<const float2*>(reinterpret_cast<const unsigned char*>(a.scratch) + pl.off_cpart) +
                         (size_t)unit * (pl.cparts * pl.RB) * pl.NEpad;
      float s1 = 0.f, s2 = 0.f;
      for (int r = 0; r < pl.cparts * pl.RB; ++r) { const float2 v = cg[(size_t)r * pl.NEpad + e]; s1 += v.x; s2 += v.y; }
      ds = st.raw_colsum ? s1 : Tr<T>::round_f(s1);         // p.sum(dim=1) is a model-dtype result (easykv.py:450)
      dsq = st.raw_colsum ? s2 : Tr<T>::round_f(s2);        // (p**2).sum(dim=1) likewise (:451)
    }
    uint32_t ka = 0, kb = 0;
    uint8_t f = 0;
    bool dirty = false;
    if (rl >= 0 && rl >= P) entry_update(st, rl - P, n_s, is_new, ds, dsq, sv, sq, cc, ka, kb, f, dirty);
    if (dirty) {
      a.S[(size_t)unit * a.cap + phys] = sv; a.SQ[(size_t)unit * a.cap + phys] = sq; a.C[(size_t)unit * a.cap + phys] = cc;
    }
    unsigned char* sb = reinterpret_cast<unsigned char*>(a.scratch);
    reinterpret_cast<int32_t*>(sb + pl.off_klj)[(size_t)unit * pl.NEpad + e] = rl;
    reinterpret_cast<uint32_t*>(sb + pl.off_ka)[(size_t)unit * pl.NEpad + e] = ka;
    reinterpret_cast<uint32_t*>(sb + pl.off_kb)[(size_t)unit * pl.NEpad + e] = kb;
    (sb + pl.off_kf)[(size_t)unit * pl.NEpad + e] = f;
    return;
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;          // (unit, row, dim4)
  const int d4 = idx % (D / 4), r = (idx / (D / 4)) % pl.R, unit = idx / ((D / 4) * pl.R);
  if (unit >= a.B * a.Hkv) return;
  const float* op = reinterpret_cast<const float*>(reinterpret_cast<const unsigned char*>(a.scratch) + pl.off_opart);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < pl.splits; ++s) {
    const float4 v = *reinterpret_cast<const float4*>(op + (((size_t)unit * pl.splits + s) * pl.Rpad + r) * D + d4 * 4);
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  const int b = unit / a.Hkv, h = unit % a.Hkv, i = r / G, g = r % G;
  T* o = reinterpret_cast<T*>(a.out) + (((size_t)b * a.H + h * G + g) * a.q_len + i) * D + d4 * 4;
  o[0] = Tr<T>::from_f(acc.x); o[1] = Tr<T>::from_f(acc.y); o[2] = Tr<T>::from_f(acc.z); o[3] = Tr<T>::from_f(acc.w);
}

// ---- 4. tail: append, accumulate, select, evict --------------------------------------------------------------------------------------
constexpr int TAIL_NT = 1024;
struct TailSmem {
  int off_ns, off_lj, off_pool, total;
  // evicting == false: the tail only appends the rows and publishes the new slots' logical indices — no per-entry
  // select scratch, hence no ceiling on the cache length (dense prefill / policy 'full' / 'decoding' prompts)
  __host__ __device__ TailSmem(int NE, int q_len, int evict, bool evicting) {
    int o = 0;
    off_ns = o; o += (q_len * 4 + 15) / 16 * 16;
    off_lj = o; o += evicting ? (NE * 4 + 15) / 16 * 16 : 0;
    o = (o + 127) / 128 * 128;
    off_pool = o;
    total = o + (evicting ? (int)SelScratch::bytes(NE, evict) : 0);
  }
};

// largest NE whose evicting tail fits one CTA's shared memory (ekv_chunk_entry_limit)
int chunk_tc_entry_limit(int q_len, int evict) {
  int lo = 0, hi = 1 << 22;
  while (lo < hi) {
    const int mid = (lo + hi + 1) / 2;
    if (TailSmem(mid, q_len, evict, true).total <= 227 * 1024) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <typename T> __global__ void __launch_bounds__(TAIL_NT) chunk_tail_kernel(const KernelArgs a, const ChunkPlan pl) {
  using namespace tc;
  pdl_trigger();            // programmatic dependent launch (ekv_common.cuh)
  pdl_wait();
  extern __shared__ __align__(128) unsigned char smem[];
  const int n_phys = a.n_phys, QL = a.q_len, NE = pl.NE;
  const bool evicting = a.st.evict > 0 && a.st.policy != EKV_POLICY_NONE;
  const TailSmem L(NE, QL, a.st.evict, evicting);
  int32_t* ns = reinterpret_cast<int32_t*>(smem + L.off_ns);
  int32_t* lj = reinterpret_cast<int32_t*>(smem + L.off_lj);
  const int unit = blockIdx.x, tid = threadIdx.x;
  const Grp grp{tid, TAIL_NT, 0};
  for (int i = tid; i < QL; i += TAIL_NT) ns[i] = a.new_slots ? a.new_slots[(size_t)unit * QL + i] : n_phys + i;
  __syncthreads();
  {                                                               // append the chunk's K/V rows (16-byte pieces)
    const uint4* kn = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.k_new) + (size_t)unit * QL * D);
    const uint4* vn = reinterpret_cast<const uint4*>(reinterpret_cast<const T*>(a.v_new) + (size_t)unit * QL * D);
    uint4* Kw = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.K) + (size_t)unit * a.cap * D);
    uint4* Vw = reinterpret_cast<uint4*>(reinterpret_cast<T*>(a.V) + (size_t)unit * a.cap * D);
    constexpr int CPR = D * (int)sizeof(T) / 16;                  // 16-byte pieces per row
    for (int idx = tid; idx < QL * CPR; idx += TAIL_NT) {
      const int i = idx / CPR, c = idx % CPR;
      Kw[(size_t)ns[i] * CPR + c] = kn[(size_t)i * CPR + c];
      Vw[(size_t)ns[i] * CPR + c] = vn[(size_t)i * CPR + c];
    }
  }
  if (!evicting) {                                                // the state was written by the chip-wide update
    int32_t* lidx = a.lidx + (size_t)unit * a.cap;
    for (int i = tid; i < QL; i += TAIL_NT) lidx[ns[i]] = a.n_before + i;
    return;
  }
  SelScratch sc;
  sc.lj = lj;
  sc.carve(smem + L.off_pool, NE, a.st.evict);
  if (a.timeline) {                                               // profiling hook: [unit][8] clock64 stamps
    sc.dbg = a.timeline + (size_t)unit * 8;
    if (tid == 0) sc.dbg[6] = clock64();                          // kernel start .. (stamps 0-5 are state_select_apply's)
  }
  UnitState u;
  u.S = a.S + (size_t)unit * a.cap; u.SQ = a.SQ + (size_t)unit * a.cap; u.C = a.C + (size_t)unit * a.cap;
  u.lidx = a.lidx + (size_t)unit * a.cap;
  u.new_slots = ns;
  u.victim_slots = a.victim_slots ? a.victim_slots + (size_t)unit * a.st.evict : nullptr;
  u.victim_lidx = a.victim_lidx ? a.victim_lidx + (size_t)unit * a.st.evict : nullptr;
  {                                                               // the keys the chip-wide state update left in scratch
    const unsigned char* sb = reinterpret_cast<const unsigned char*>(a.scratch);
    const int32_t* klj = reinterpret_cast<const int32_t*>(sb + pl.off_klj) + (size_t)unit * pl.NEpad;
    const uint32_t* ka = reinterpret_cast<const uint32_t*>(sb + pl.off_ka) + (size_t)unit * pl.NEpad;
    const uint32_t* kb = reinterpret_cast<const uint32_t*>(sb + pl.off_kb) + (size_t)unit * pl.NEpad;
    const unsigned char* kf = sb + pl.off_kf + (size_t)unit * pl.NEpad;
    for (int e = tid; e < NE; e += TAIL_NT) { sc.lj[e] = klj[e]; sc.keyA[e] = ka[e]; sc.keyB[e] = kb[e]; sc.flag[e] = kf[e]; }
  }
  __syncthreads();
  auto none = [](int, float& ds, float& dsq) { ds = 0.f; dsq = 0.f; };
  state_select_apply(a.st, u, a.n_before, n_phys, QL, /*lj_preloaded=*/true, none, sc, grp, /*keys_ready=*/true);
}

// ---- launch ---------------------------------------------------------------------------------------------------------------------------
static bool tail_evicting(const KernelArgs& a) { return a.st.evict > 0 && a.st.policy != EKV_POLICY_NONE; }

template <typename T, int G> static int configure_chunk_tc() {
  using namespace tc;
  static thread_local int configured[16] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16) dev = 15;
  if (configured[dev]) return EKV_OK;
  const int smem1 = MR * PITCH + STAGES * TILE_BYTES + STAGES * TK * 4;
  const int smem2 = MR * PITCH + 2 * 2 * TILE_BYTES + 2 * TK * 4 + MR * (TK + 2) * 2;
  cudaError_t err = cudaFuncSetAttribute(chunk_tc_kernel<T, G, 1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(chunk_tc_kernel<T, G, 1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem1);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(chunk_tc_kernel<T, G, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(chunk_tc_kernel<T, G, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem2);
  if (err == cudaSuccess) err = cudaFuncSetAttribute(chunk_tail_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
  if (err != cudaSuccess) return set_cuda_error("cudaFuncSetAttribute(chunk_tc)", err);
  configured[dev] = 1;
  return EKV_OK;
}

// 3 + 4: partial outputs -> out, chip-wide policy-state update, per-unit tail.  Shared by the tcgen05 path.
template <typename T, int G> static int launch_finish_tg(const KernelArgs& a, const ChunkPlan& pl, cudaStream_t stream) {
  using namespace tc;
  const TailSmem TL(pl.NE, a.q_len, a.st.evict, tail_evicting(a));
  if (TL.total > 227 * 1024)
    return set_error(EKV_ERR_UNSUPPORTED, "chunk tail: selecting victims among %d entries needs %d bytes of shared memory (limit 227 KB, "
                     "about 17.6K retained slots); larger caches are served only without eviction", pl.NE, TL.total);
  int rc = configure_chunk_tc<T, G>();
  if (rc) return rc;
  cudaError_t err;
  const int U = a.B * a.Hkv;
  const int out_blocks = (U * pl.R * (D / 4) + 255) / 256;
  const int col_blocks = (U * pl.NEpad + 255) / 256;               // state update: always (new slots need their state)
  cudaLaunchAttribute pdl[1];
  pdl[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;       // both kernels wait (pdl_wait) before their first global access
  pdl[0].val.programmaticStreamSerializationAllowed = pdl_allowed();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(out_blocks + col_blocks), 1, 1);
  cfg.blockDim = dim3(256, 1, 1);
  cfg.stream = stream;
  cfg.attrs = pdl;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, chunk_out_kernel<T, G>, a, pl, out_blocks);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("chunk_out_kernel launch", err);
  count_launch();
  cfg.gridDim = dim3((unsigned)U, 1, 1);
  cfg.blockDim = dim3(TAIL_NT, 1, 1);
  cfg.dynamicSmemBytes = (size_t)TL.total;
  cudaLaunchKernelEx(&cfg, chunk_tail_kernel<T>, a, pl);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("chunk_tail_kernel launch", err);
  count_launch();
  return EKV_OK;
}

template <typename T, int G> static int launch_chunk_tc_tg(const KernelArgs& a, cudaStream_t stream) {
  using namespace tc;
  const ChunkPlan pl = make_chunk_plan(a.B, a.Hkv, G, a.q_len, a.n_phys);
  const TailSmem TL(pl.NE, a.q_len, a.st.evict, tail_evicting(a));
  if (TL.total > 227 * 1024) return set_error(EKV_ERR_UNSUPPORTED, "chunk tail: %d bytes of shared memory needed", TL.total);
  const int smem1 = MR * PITCH + STAGES * TILE_BYTES + STAGES * TK * 4;
  const int smem2 = MR * PITCH + 2 * 2 * TILE_BYTES + 2 * TK * 4 + MR * (TK + 2) * 2;
  int rc = configure_chunk_tc<T, G>();
  if (rc) return rc;
  cudaError_t err;
  const int U = a.B * a.Hkv;
  const int grid = U * pl.RB * pl.splits, grid1 = U * pl.RB * pl.splits1;
  if (a.st.arith) chunk_tc_kernel<T, G, 1, true><<<grid1, NT, smem1, stream>>>(a, pl);
  else chunk_tc_kernel<T, G, 1, false><<<grid1, NT, smem1, stream>>>(a, pl);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("chunk_tc_kernel<1> launch", err);
  count_launch();
  if (a.st.arith) chunk_tc_kernel<T, G, 2, true><<<grid, NT, smem2, stream>>>(a, pl);
  else chunk_tc_kernel<T, G, 2, false><<<grid, NT, smem2, stream>>>(a, pl);
  if ((err = cudaGetLastError()) != cudaSuccess) return set_cuda_error("chunk_tc_kernel<2> launch", err);
  count_launch();
  return launch_finish_tg<T, G>(a, pl, stream);
}

template <typename T> static int launch_chunk_tc_t(const KernelArgs& a, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_chunk_tc_tg<T, 1>(a, stream);
    case 2: return launch_chunk_tc_tg<T, 2>(a, stream);
    case 4: return launch_chunk_tc_tg<T, 4>(a, stream);
    case 8: return launch_chunk_tc_tg<T, 8>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}
template <typename T> static int launch_finish_t(const KernelArgs& a, const ChunkPlan& pl, cudaStream_t stream) {
  switch (a.H / a.Hkv) {
    case 1: return launch_finish_tg<T, 1>(a, pl, stream);
    case 2: return launch_finish_tg<T, 2>(a, pl, stream);
    case 4: return launch_finish_tg<T, 4>(a, pl, stream);
    case 8: return launch_finish_tg<T, 8>(a, pl, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

int launch_chunk_finish(const KernelArgs& a, const ChunkPlan& pl, cudaStream_t stream) {
  switch (a.dtype) {
    case EKV_F16: return launch_finish_t<__half>(a, pl, stream);
    case EKV_BF16: return launch_finish_t<__nv_bfloat16>(a, pl, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

// 16-bit dtypes, head_dim 128, any q_len; needs `scratch` of chunk_tc_scratch_bytes().
int launch_chunk_tc(const KernelArgs& a, cudaStream_t stream) {
  if (a.d != tc::D || !a.scratch) return EKV_ERR_UNSUPPORTED;
  switch (a.dtype) {
    case EKV_F16: return launch_chunk_tc_t<__half>(a, stream);
    case EKV_BF16: return launch_chunk_tc_t<__nv_bfloat16>(a, stream);
    default: return EKV_ERR_UNSUPPORTED;
  }
}

}  // namespace ekv
