// hfb_stats_tc.cuh -- K4 on the 5th-generation tensor cores.
//
// UpMixParms (HTKLib/HFB.c:1426-1744) as the two contractions north_star asks for, per tied state, over the frames of
// the whole wave in which one of the state's positions survives the alpha beam (stats_pre_kernel's dense lists):
//
//   (1) component log-likelihoods   V[frames x components] = A[frames x (1 | x'^2, x' | 1)] x B_state^T
//       -- the very product gmm_tc3_kernel forms for the state log-likelihood (same operands, same 3xFP16 split, same
//       MMA order, so the same numbers), for 128 gathered frames at a time: tcgen05.mma M = 128, N = 16 / 32, K = 80;
//   (2) occupancy-weighted sums     S[(1 | x'^2, x') x components] = A^T[columns x frames] x Lr[frames x components]
//       with Lr = exp(initx + log weight + log N) under the minimum-occupancy rule (HFB.c:1581-1612).  A^T is the SAME
//       shared-memory tile read as an MN-major operand (a 128-byte-swizzled K-major tile of [frames][64 columns] IS the
//       canonical MN-major tile with MN = 64, K = frames), Lr is written by the epilogue of (1) as the K-major B
//       operand, hi / lo FP16 halves: M = 128 (operand columns), N = components, K = 128 frames.
//
// The accumulators of (2) live in TMEM for one tile only (24 truncating tensor-core accumulations, ~1e-6 relative); the
// running sums of a state stay in FP32 registers of the thread that owns the operand column and are moved once per
// state -- to HTK's centred form about each component mean (HFB.c:1673-1678) -- into the FP64 accumulators:
//     sum Lr (o - mu)   = S1 / s - d S0,      sum Lr (o - mu)^2 = S2 / s^2 - 2 d S1 / s + d^2 S0,   d = mu - offset,
// with s the per-dimension power-of-two scale of the operands.  Replaces stats5_kernel (mma.sync 3xTF32 sums behind an
// FP32 recomputation of the posteriors: 25 % of its instructions, 12 warps / SM); stats_pre_kernel and the position sort
// are unchanged.  Frames outside the FP16 operand range (gmm_tc3's flags) get their posteriors from an FP32 evaluation.
//
// One CTA = 4 worker warps (thread = tile row: gather + expansion, epilogue of (1), TMEM lane of (2)) + 1 MMA / TMA
// warp; it owns ST_CAP consecutive entries of the state-sorted position list.  The phases of a tile are sequential
// inside a CTA; two CTAs per SM overlap them.
#pragma once
#include "gmm_tc3.cuh"
#include "hfb_stats_mma.cuh"

#define ST_CAP 128                   // sorted positions per CTA
#define ST_THREADS 160
#define ST_LR_SCALE 4096.f           // Lr in (4.5e-5, 1] -> FP16 normal range; removed again at the flush

struct StatsTcParams {
   const PosRec *list;
   const int *listEnd;               // number of sorted positions
   const ValidFrame *vbuf;
   const int *vcnt;
   const unsigned char *flag;        // per frame of the wave: FP16 operands out of range (gmm_tc3)
   const int *overflow;              // != 0: some position did not fit the frame lists -> stats5_kernel does the wave
   const float *offset, *scale;
   const uint4 *expA;                // PRE: the expanded operand rows gmm_tc3_kernel wrote, [frame of the wave][hi units | lo units]
   float C0;
   int kSteps, N, MP;
};

__device__ __forceinline__ void st_cp_async16(void *smemDst, const void *gsrc)
{
   asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(tc_smem_u32(smemDst)), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void st_mma_f16(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
   asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major, 128-byte swizzle: LBO = stride between 64-element groups along MN, SBO = stride between 8-element groups along K
__device__ __forceinline__ uint64_t st_desc_mn(uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes)
{
   return (uint64_t)((smemAddr >> 4) & 0x3FFF) | ((uint64_t)((lboBytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sboBytes >> 4) & 0x3FFF) << 32) |
          ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// operand tile WITHOUT swizzle ("interleaved" canonical layout): core matrices of 8 rows x 16 bytes stored as 128 contiguous
// bytes; `lbo` / `sbo` = byte strides between core matrices along the two tile dimensions (which is which depends on the
// major-ness, see the call sites); bit 46 = descriptor version, swizzle field 0
__device__ __forceinline__ uint64_t st_desc_ns(uint32_t smemAddr, uint32_t lboBytes, uint32_t sboBytes, bool mnMajor)
{
   (void)mnMajor;
   return (uint64_t)((smemAddr >> 4) & 0x3FFF) | ((uint64_t)((lboBytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sboBytes >> 4) & 0x3FFF) << 32) |
          ((uint64_t)1 << 46);
}
template <int NC>
__device__ __forceinline__ void st_tmem_ld(uint32_t taddr, float *v)
{
   uint32_t *r = reinterpret_cast<uint32_t *>(v);
   if (NC == 16)
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                   : "r"(taddr));
   else {
      tc_tmem_ld32(taddr, v);
      return;
   }
   asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// PRE = the rows of a tile are COPIED (cp.async, 16 bytes at a time, straight into the swizzled tile) from the expanded
// operand gmm_tc3_kernel left in global memory -- bit-identical to what the expansion below produces -- instead of being
// gathered as raw features and expanded again: the expansion was 35 % of this kernel's instructions and sat at the head
// of every tile's phase chain.  The copy of the NEXT tile is issued as soon as contraction (2) has released the tile.
// NSW (with PRE) = the tile is stored WITHOUT swizzle and without the padding of the 128-byte-swizzled form: per group of 8
// rows the 2 kSteps core matrices (8 rows x 8 columns, 128 contiguous bytes) that exist, one after the other -- 40 KB
// instead of 64 KB for D = 39, so that THREE CTAs fit an SM (the kernel is bound by the latency of its per-tile phase chain).
// The same bytes serve both contractions: a core matrix read K-major is 8 frames x 8 columns, read MN-major it is 8 columns
// (MN) x 8 frames (K).
template <int N, int DP, bool PRE, bool NSW>
__global__ void __launch_bounds__(ST_THREADS, NSW ? 3 : 2)
stats_tc_kernel(const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, DevModel M, Wave W, StatsTcParams p)
{
   extern __shared__ uint8_t st_smem_raw[];
   if (*p.overflow) return;
   const int nSorted = *p.listEnd;
   const int i0 = blockIdx.x * ST_CAP, i1 = min(nSorted, i0 + ST_CAP);
   if (i0 >= i1) return;
   uint8_t *base = (uint8_t *)(((uintptr_t)st_smem_raw + 1023) & ~(uintptr_t)1023);
   uint8_t *sA = base;                                  // [hi c0 | hi c1 | lo c0 | lo c1] x 16 KB: 128 rows x 128 columns
   const uint32_t GS = 2u * (uint32_t)p.kSteps * 128u;  // NSW: bytes of one group of 8 rows (2 kSteps core matrices)
   const uint32_t HALF = NSW ? 16u * GS : 32768u;       // bytes of the hi (or lo) part of the tile
   const uint32_t TILE = NSW ? ((2u * HALF + 1023u) & ~1023u) : 65536u;
   uint8_t *sB1 = sA + TILE;                            // the state's Gaussians: [hi c0 | hi c1 | lo c0 | lo c1] x N x 128 B
   uint8_t *sB2 = sB1 + 4 * N * 128;                    // Lr: [hi f0-63 | hi f64-127 | lo .. | lo ..] x N x 128 B
   float *stage = (float *)sB2;                         // [128][N] flush staging: Lr is dead when a state is flushed
   int *pre = (int *)(sB2 + 4 * N * 128);               // [ST_CAP + 1] prefix sums of the positions' frame counts
   int *stt = pre + ST_CAP + 1;                         // [ST_CAP] tied state of each position
   float2 *sXf = (float2 *)(((uintptr_t)(stt + ST_CAP) + 7) & ~(uintptr_t)7);   // per dimension: (scale, -offset * scale): x' = fma(x, s, c)
   float *sScl = (float *)(sXf + 64);                   // (padding kept for the layout below)
   int *segEnd = (int *)sScl;                           // [ST_CAP] end of the run of equal states that starts at a position
   long long *pV = (long long *)(((uintptr_t)(segEnd + ST_CAP) + 15) & ~(uintptr_t)15);   // per position: first entry of its frame list,
   long long *pF = pV + ST_CAP, *pB = pF + ST_CAP;      // first frame of its utterance in the feature matrix / in the flag array
   uint64_t *bars = (uint64_t *)(pB + ST_CAP);
   uint64_t *barB = bars, *bar1 = bars + 1, *bar2 = bars + 2;
   uint32_t *tmemSlot = (uint32_t *)(bars + 3);
   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const bool worker = warp < 4;
   const int D = M.D, Dp = M.Dp;
   constexpr uint32_t TCOLS = (2 * N <= 32) ? 32 : 64;

   if (tid == 0) {
      tc_mbar_init(barB, 1); tc_mbar_init(bar1, 1); tc_mbar_init(bar2, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == 4) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmemSlot)), "r"(TCOLS) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
   }
   // frame counts / states of my slice; the A tile starts out finite (stale rows are multiplied by Lr = 0 in (2))
   for (int k = tid; k < ST_CAP; k += ST_THREADS) {
      const int it = i0 + k;
      pre[k + 1] = (it < i1 && p.list[it].vOff >= 0) ? p.vcnt[it] : 0;
      stt[k] = (it < i1) ? p.list[it].s : -1;
      if (it < i1) { pV[k] = p.list[it].vOff; pF[k] = p.list[it].featOff; pB[k] = p.list[it].frameBase; }
   }
   if (tid < 64) sXf[tid] = (tid < D) ? make_float2(p.scale[tid], -p.offset[tid] * p.scale[tid]) : make_float2(0.f, 0.f);
   for (uint32_t o = tid * 16; o < TILE; o += ST_THREADS * 16) *reinterpret_cast<uint4 *>(sA + o) = make_uint4(0u, 0u, 0u, 0u);
   tc_fence_before();
   __syncthreads();
   tc_fence_after();
   if (tid == 0) { pre[0] = 0; for (int k = 0; k < ST_CAP; k++) pre[k + 1] += pre[k]; }
   if (tid == 32) {
      const int np = i1 - i0;
      int e = np;
      for (int k = np - 1; k >= 0; k--) { if (k + 1 < np && stt[k + 1] != stt[k]) e = k + 1; segEnd[k] = e; }
   }
   __syncthreads();
   const uint32_t tmem = *tmemSlot;
   const uint32_t tD1 = tmem, tD2 = tmem + N;
   const int nPos = i1 - i0;
   const int uf = W.uFlags;
   const bool upM = (uf & HFB_UPMEANS) != 0, upV = (uf & HFB_UPVARS) != 0, upW = (uf & HFB_UPMIXES) != 0;
   const double minF = W.minFrwdP;
   const uint32_t idesc1 = tc_idesc(TC_BM, N, 0u);                     // K-major A and B
   const uint32_t idesc2 = tc_idesc(TC_BM, N, 0u) | (1u << 15);        // A MN-major (the transposed tile), B K-major
   uint32_t phB = 0, ph1 = 0, ph2 = 0;
   float run[N];                                        // running sums of operand column `tid` of the current state

   // ---- the tiles of my slice: consecutive blocks of 128 frames of one state segment [a, b) of the position list.
   //      The rows of the NEXT tile are fetched (list entry -> frame record -> feature row: three dependent global
   //      loads) while the tensor core and the epilogue work on the current one.
   struct TileAt { int a, b, t0, g1, s; bool have; };
   auto next_tile = [&](TileAt c) -> TileAt {
      if (c.have && c.t0 + TC_BM < c.g1) { c.t0 += TC_BM; return c; }
      int na = c.have ? c.b : 0;
      for (;;) {
         if (na >= nPos) { c.have = false; return c; }
         const int nb = segEnd[na];                     // first position of the slice with another state
         if (pre[nb] > pre[na]) { c.a = na; c.b = nb; c.t0 = pre[na]; c.g1 = pre[nb]; c.s = stt[na]; c.have = true; return c; }
         na = nb;
      }
   };
   float x[PRE ? 1 : DP];                               // my row of the tile in flight: RAW features (transformed when stored)
   float px0 = 0.f, px0l = 0.f;                         // initx / log occupancy as hi + lo floats (it can be ~1e6 on outlier frames)
   bool pvalid = false, pfar = false;
   const float *pfrow = nullptr;
   // two steps, so that neither of the two dependent global loads (frame record, then feature row) is waited for:
   // fetch_index issues the record load BEFORE the current tile's rows are converted and stored, fetch_row uses it after
   ValidFrame nvf; nvf.x0 = 0.0; nvf.t = 0; nvf.pad = 0;
   int nlo = 0;
   bool nvalid = false;
   auto fetch_index = [&](const TileAt &c) {
      nvalid = false;
      if (!worker || !c.have) return;
      const int g = c.t0 + tid;
      nvalid = g < c.g1;
      if (nvalid) {
         int lo = c.a, hi = c.b;                        // position with pre[it] <= g < pre[it + 1]
         while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= g) lo = mid; else hi = mid; }
         nvf = p.vbuf[pV[lo] + (g - pre[lo])];
         nlo = lo;
      }
   };
   auto fetch_row = [&]() {
      pvalid = nvalid; pfar = false; pfrow = nullptr; px0 = 0.f; px0l = 0.f;
      if (!PRE) {
#pragma unroll
         for (int d = 0; d < (PRE ? 1 : DP); d++) x[d] = 0.f;
      }
      if (pvalid) {
         const ValidFrame vf = nvf;
         const int lo = nlo;
         px0 = (float)vf.x0; px0l = (float)(vf.x0 - (double)px0);
         pfrow = W.feat + ((size_t)pF[lo] + vf.t) * D;
         pfar = p.flag != nullptr && p.flag[pB[lo] + vf.t] != 0;
         if (PRE) {
            // rows of invalid tile positions keep whatever finite halfs they held: their Lr is zero in (2), their row of V
            // is never read; the same holds for a frame outside the FP16 range (its expanded row is clamped, finite)
            const int nUn = 2 * p.kSteps;
            const uint4 *src = p.expA + (size_t)(pB[lo] + vf.t) * (size_t)(2 * nUn);
#pragma unroll
            for (int un = 0; un < 16; un++) {
               if (un >= nUn) break;
               const uint32_t off = NSW ? (uint32_t)(tid >> 3) * GS + (uint32_t)un * 128u + (uint32_t)(tid & 7) * 16u : tc3_unit_off(tid, un);
               st_cp_async16(sA + off, src + un);
               st_cp_async16(sA + HALF + off, src + nUn + un);
            }
         } else {
#pragma unroll
            for (int d = 0; d < (PRE ? 1 : DP); d++)
               if (d < D) x[d] = pfrow[d];              // no arithmetic here: the loads stay in flight behind the MMAs
         }
      }
      if (PRE) asm volatile("cp.async.commit_group;" ::: "memory");
   };
   TileAt cur; cur.a = cur.b = cur.t0 = cur.g1 = 0; cur.s = -1; cur.have = false;
   cur = next_tile(cur);
   fetch_index(cur);
   fetch_row();
   const float minFf = (float)minF;
   while (cur.have) {
      const int s = cur.s, mo = M.stateMixOff[s], Mn = M.stateMixOff[s + 1] - mo;
      const bool first = cur.t0 == pre[cur.a], last = cur.t0 + TC_BM >= cur.g1;
      // single-Gaussian sets (MP = 1): Lr is the state occupancy itself, contraction (1) is not needed at all
      const bool needV = p.MP > 1;
      if (first) {
         // ---- the state's Gaussians: rows TC3_ROW0 + s MP .. + N of the tensor-core B operand
         if (needV && warp == 4 && lane == 0) {
            tc_mbar_expect_tx(barB, 4 * N * 128);
            constexpr int BOXR = (N < 64) ? ((N == 16) ? 16 : 32) : 64;
            const int boxR = (p.MP < BOXR) ? p.MP : BOXR;            // rows per TMA box as the maps were built
            for (int c = 0; c < 2; c++)
               for (int r0 = 0; r0 < N; r0 += boxR) {
                  tc_tma_load_2d(sB1 + c * (N * 128) + r0 * 128, &mapBhi, barB, c * 64, TC3_ROW0 + s * p.MP + r0);
                  tc_tma_load_2d(sB1 + (2 + c) * (N * 128) + r0 * 128, &mapBlo, barB, c * 64, TC3_ROW0 + s * p.MP + r0);
               }
         }
#pragma unroll
         for (int m = 0; m < N; m++) run[m] = 0.f;
      }
      // ================= phase A: my row of the tile -> shared memory (thread = row) =================
      const bool valid = pvalid, far = pfar;
      const float x0 = px0, x0l = px0l;
      const float *frow = pfrow;
      const TileAt nxt = next_tile(cur);
      fetch_index(nxt);                                 // frame records of the next tile: in flight during phase A
      if (PRE) {
         if (worker) {
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
         }
      } else if (worker) {
#pragma unroll
         for (int d = 0; d < (PRE ? 1 : DP); d++)
            if (d < D) {
               // rows of frames outside the FP16 range are zeroed (they are accumulated apart, see the epilogue); every
               // other row is within TC_FAR of the centre by construction of the flags -- without flags (FP32 output
               // probabilities were requested) the clamp keeps the operands finite
               const float2 sc = sXf[d];
               float xv = fmaf(x[d], sc.x, sc.y);
               if (p.flag == nullptr) xv = fminf(fmaxf(xv, -250.f), 250.f);
               x[d] = (valid && !far) ? xv : 0.f;
            }
#pragma unroll
         for (int un = 0; un < 16; un++) {
            if (un >= 2 * p.kSteps) break;
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
               const int k = un * 8 + e;
               if (k == 0) v[e] = valid ? 1.f : 0.f;
               else if (k & 1) { const int d = (k - 1) >> 1; v[e] = (!PRE && d < DP && d < D) ? x[(!PRE && d < DP) ? d : 0] * x[(!PRE && d < DP) ? d : 0] : ((d == D && valid) ? 1.f : 0.f); }
               else { const int d = (k - 2) >> 1; v[e] = (!PRE && d < DP && d < D) ? x[(!PRE && d < DP) ? d : 0] : 0.f; }
            }
            uint32_t h4[4], l4[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
               const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
               const float2 hf = __half22float2(h);
               const __half2 l = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
               h4[e] = *reinterpret_cast<const uint32_t *>(&h);
               l4[e] = *reinterpret_cast<const uint32_t *>(&l);
            }
            const uint32_t off = tc3_unit_off(tid, un);
            *reinterpret_cast<uint4 *>(sA + off) = make_uint4(h4[0], h4[1], h4[2], h4[3]);
            *reinterpret_cast<uint4 *>(sA + 32768 + off) = make_uint4(l4[0], l4[1], l4[2], l4[3]);
         }
         asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      if (!PRE) fetch_row();                            // feature rows of the next tile: in flight during (1), its epilogue and (2)
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ================= (1): V = A x B_state^T =================
      if (warp == 4 && needV) {
         if (first) { tc_mbar_wait(barB, phB); phB ^= 1; tc_fence_after(); }
         if (lane == 0) {
            const uint32_t aB = tc_smem_u32(sA), bB = tc_smem_u32(sB1);
            // A slice of K step ks: swizzled tile = 32 bytes further inside the 128-byte rows of chunk ks / 4; compact tile =
            // two core matrices further (K-major, no swizzle: 128 bytes between core matrices along K, GS between row groups)
            auto adesc1 = [&](uint32_t part, int ks) -> uint64_t {
               if (NSW) return st_desc_ns(aB + part * HALF + ks * 256, 128, GS, false);
               return tc_smem_desc(aB + part * 32768 + (ks >> 2) * 16384 + (ks & 3) * 32);
            };
            for (int ks = 0; ks < p.kSteps; ks++) {              // corrections first (see gmm_tc3_kernel)
               const uint32_t ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
               st_mma_f16(tD1, adesc1(0, ks), tc_smem_desc(bB + 2 * N * 128 + ob), idesc1, ks ? 1u : 0u);
               st_mma_f16(tD1, adesc1(1, ks), tc_smem_desc(bB + ob), idesc1, 1u);
            }
            for (int ks = 0; ks < p.kSteps; ks++) {
               const uint32_t ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
               st_mma_f16(tD1, adesc1(0, ks), tc_smem_desc(bB + ob), idesc1, 1u);
            }
            tc_commit(bar1);
         }
         __syncwarp();
      }
      // ================= epilogue of (1): Lr -> B operand of (2) =================
      if (worker) {
         float v[N];
         if (needV) {
            tc_mbar_wait(bar1, ph1);
            tc_fence_after();
            st_tmem_ld<N>(tD1 + ((uint32_t)(warp * 32) << 16), v);
            ph1 ^= 1;
         } else {
#pragma unroll
            for (int m = 0; m < N; m++) v[m] = 0.f;
         }
#pragma unroll
         for (int m = 0; m < N; m++) v[m] -= p.C0;      // log weight + log N_m
         if (valid && far) {
            // A frame outside the FP16 operand range: its row of the tile is zero and its Lr below is zero, i.e. the tensor
            // core never sees it.  It goes straight to the accumulators the way the reference forms them: x = initx + log
            // weight + log N_m summed in DOUBLE from the float log N_m of IDOutP (HFB.c:1581-1599, HModel.c:5420-5431).  On
            // such a frame |log N_m| ~ 1e6, where the float log-sum b_j(o_t) that went into alpha is a fraction of an ulp
            // (0.5) off the exact log-sum of the components: the reference's component occupancies then do not add up to
            // the state occupancy, and neither may ours (parity is with the reference, not with the arithmetic identity).
            const double xi = (double)x0 + (double)x0l;        // initx (log occupancy for a single-Gaussian state)
#pragma unroll 1
            for (int m = 0; m < Mn; m++) {
               const float wt = M.mixLogWt[mo + m];
               if (Mn > 1 && !(wt > LMINMIX_F)) continue;      // :1573
               const int g = M.mixGauss[mo + m], mId = M.meanId[g], vId = M.varId[g];
               const float *mu = M.mean + (size_t)g * Dp, *iv = M.ivar + (size_t)g * Dp;
               double xx = xi;                                 // :1575-1576
               if (Mn > 1) {
                  float acc = M.gconst[g];
                  for (int k = 0; k < D; k++) { const float dd = frow[k] - mu[k]; acc = fmaf(dd * dd, iv[k], acc); }
                  xx = (xi + (double)wt) + (double)(-0.5f * acc);
               }
               if (!(-xx < minF)) continue;                    // :1606
               const double L = exp(xx);
               for (int k = 0; k < D; k++) {
                  const double z = (double)frow[k] - (double)mu[k];
                  if (upM) atomicAdd(&W.acc[M.L.muSum + (size_t)mId * D + k], z * L);
                  if (upV) atomicAdd(&W.acc[M.L.vaSum + (size_t)vId * D + k], z * z * L);
               }
               if (upM) atomicAdd(&W.acc[M.L.muOcc + mId], L);
               if (upV) atomicAdd(&W.acc[M.L.vaOcc + vId], L);
               if (upW) atomicAdd(&W.acc[M.L.wtC + mo + m], L);
               atomicAdd(&W.acc[M.L.wtOcc + s], L);
            }
         }
         const uint32_t offT = (uint32_t)((tid >> 6) * (N * 128) + (tid & 7) * 2), unit = (uint32_t)((tid & 63) >> 3);
#pragma unroll
         for (int m = 0; m < N; m++) {
            float Lr = 0.f;
            if (valid && !far && m < Mn) {
               // x = initx + log weight + log N_m (:1581-1599); single-Gaussian states: x = log occupancy (:1575-1576)
               // (on an outlier frame initx and log N_m are both ~1e6 with opposite signs: their float sum is exact, the
               // low part of initx restores what its own rounding to float lost)
               const float xx = (Mn > 1) ? (x0 + v[m]) + x0l : x0 + x0l;
               if (-xx < minFf && (Mn > 1 || m == 0)) Lr = tc_ex2(xx * 1.4426950408889634f) * ST_LR_SCALE;   // :1606, :1612
            }
            const __half h = __float2half_rn(Lr), l = __float2half_rn(Lr - __half2float(h));
            const uint32_t off = offT + (uint32_t)(m * 128) + ((unit ^ (uint32_t)(m & 7)) << 4);
            *reinterpret_cast<__half *>(sB2 + off) = h;
            *reinterpret_cast<__half *>(sB2 + 2 * N * 128 + off) = l;
         }
         asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      }
      tc_fence_before();
      __syncthreads();
      tc_fence_after();
      // ================= (2): S = A^T x Lr =================
      if (warp == 4) {
         if (lane == 0) {
            const uint32_t aB = tc_smem_u32(sA), bB = tc_smem_u32(sB2);
            // the transposed tile, 16 frames (K) per step: swizzled = 16 rows of 128 bytes further; compact = two row groups
            // further (MN-major, no swizzle: the descriptor's "leading" stride is the one along K = frames here, GS, and its
            // "stride" field the 128 bytes between core matrices along MN = operand columns -- found by trying both on the
            // golden fixtures; operand columns past 16 kSteps read the next group's bytes -- finite halfs -- into rows of S
            // nobody uses)
            auto adesc2 = [&](uint32_t part, int ks) -> uint64_t {
               if (NSW) return st_desc_ns(aB + part * HALF + ks * 2 * GS, GS, 128, true);
               return st_desc_mn(aB + part * 32768 + ks * 2048, 16384, 1024);
            };
            for (int ks = 0; ks < 8; ks++) {                    // 16 frames per step
               const uint32_t ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
               st_mma_f16(tD2, adesc2(0, ks), tc_smem_desc(bB + 2 * N * 128 + ob), idesc2, ks ? 1u : 0u);
               st_mma_f16(tD2, adesc2(1, ks), tc_smem_desc(bB + ob), idesc2, 1u);
            }
            for (int ks = 0; ks < 8; ks++) {
               const uint32_t ob = (ks >> 2) * (N * 128) + (ks & 3) * 32;
               st_mma_f16(tD2, adesc2(0, ks), tc_smem_desc(bB + ob), idesc2, 1u);
            }
            tc_commit(bar2);
         }
         __syncwarp();
      }
      if (worker) {
         tc_mbar_wait(bar2, ph2);
         ph2 ^= 1;
         tc_fence_after();
         if (PRE) fetch_row();                          // (2) has read the tile: the next tile's rows start to arrive
         float v[N];
         st_tmem_ld<N>(tD2 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
         for (int m = 0; m < N; m++) run[m] += v[m];
      }
      tc_fence_before();
      __syncthreads();                                  // the tile, Lr and both accumulators may be overwritten
      tc_fence_after();
      if (last) {
         // ================= flush: the state's sums -> FP64 accumulators, centred on the component means =================
         if (worker) {
#pragma unroll
            for (int m = 0; m < N; m++) stage[tid * N + m] = run[m];
         }
         __syncthreads();
         if (worker) {
            const double inv = 1.0 / (double)ST_LR_SCALE;
            double wsum = 0.0;
            for (int e = tid; e < Mn * (D + 1); e += 128) {
               const int m = e / (D + 1), k = e - m * (D + 1);
               const double S0 = (double)stage[0 * N + m] * inv;
               if (!(S0 > 0.0)) continue;               // component never passed the minimum-occupancy rule
               const int g = M.mixGauss[mo + m];
               if (k == D) {
                  if (upM) atomicAdd(&W.acc[M.L.muOcc + M.meanId[g]], S0);
                  if (upV) atomicAdd(&W.acc[M.L.vaOcc + M.varId[g]], S0);
                  if (upW) atomicAdd(&W.acc[M.L.wtC + mo + m], S0);
               } else {
                  const double sk = (double)p.scale[k], d = ((double)M.mean[(size_t)g * Dp + k] - (double)p.offset[k]) * sk;
                  const double S2 = (double)stage[(2 * k + 1) * N + m] * inv, S1 = (double)stage[(2 * k + 2) * N + m] * inv;
                  if (upM) atomicAdd(&W.acc[M.L.muSum + (size_t)M.meanId[g] * D + k], (S1 - d * S0) / sk);
                  if (upV) atomicAdd(&W.acc[M.L.vaSum + (size_t)M.varId[g] * D + k], (S2 - 2.0 * d * S1 + d * d * S0) / (sk * sk));
               }
            }
            if (tid < Mn) wsum = (double)stage[tid] * inv;
            for (int o = 16; o > 0; o >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
            if (tid == 0 && wsum > 0.0) atomicAdd(&W.acc[M.L.wtOcc + s], wsum);
         }
         __syncthreads();
      }
      cur = nxt;
   }
   tc_fence_before();
   __syncthreads();
   if (warp == 4) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TCOLS) : "memory");
   }
}

template <int N>
static inline size_t stats_tc_smem_bytes(int kSteps = 0)   // kSteps > 0: the compact (no-swizzle) tile
{
   const size_t tile = kSteps > 0 ? (((size_t)2 * 16 * 2 * kSteps * 128 + 1023) & ~(size_t)1023) : 65536;
   return 1024 + tile + 8 * N * 128 + sizeof(int) * (3 * ST_CAP + 1) + 8 + 512 + 16 + 3 * 8 * ST_CAP + 128;
}

// Launch: returns false when the model is outside what the kernel covers (the caller keeps stats5_kernel).
static inline bool stats_tc_supported(const GmmTc3Model &t, int D) { return t.ready && (t.MP == 1 || t.MP == 8 || t.MP == 16 || t.MP == 32) && D <= 63; }

static inline void stats_tc_set_attributes()
{
#define ST_SET(NV, DPV) \
   cudaFuncSetAttribute(stats_tc_kernel<NV, DPV, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stats_tc_smem_bytes<NV>()); \
   cudaFuncSetAttribute(stats_tc_kernel<NV, DPV, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stats_tc_smem_bytes<NV>()); \
   cudaFuncSetAttribute(stats_tc_kernel<NV, DPV, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stats_tc_smem_bytes<NV>())
   ST_SET(16, 40); ST_SET(16, 64); ST_SET(32, 40); ST_SET(32, 64);
#undef ST_SET
}

// expA: the expanded operand rows of the wave (gmm_tc3_launch with wantExp), or nullptr = gather raw features and expand
static inline void stats_tc_launch(const GmmTc3Model &t, const DevModel &dm, const Wave &W, const PosRec *list, const int *listEnd,
                                   const ValidFrame *vbuf, const int *vcnt, const unsigned char *flag, const int *overflow,
                                   long long totalP, cudaStream_t st, const uint4 *expA = nullptr)
{
   StatsTcParams p;
   p.list = list; p.listEnd = listEnd; p.vbuf = vbuf; p.vcnt = vcnt; p.flag = flag; p.overflow = overflow;
   p.offset = t.dOffset; p.scale = t.dScale; p.C0 = t.C0; p.kSteps = (2 * dm.D + 2 + 15) / 16; p.MP = t.MP;
   p.N = (t.MP <= 16) ? 16 : 32;
   p.expA = expA;
   const unsigned grid = (unsigned)((totalP + ST_CAP - 1) / ST_CAP);
   if (grid == 0) return;
   // the compact tile pays when it lets a third CTA onto the SM: K steps <= 6 and 16 components per tile (with 32 the two
   // B operands take 32 KB and three CTAs miss the SM's shared memory by 1 KB: measured 2.16 against 2.03 ms on config #4)
   const bool ns = expA != nullptr && p.kSteps <= 6 && p.N == 16 && !getenv("HFBGPU_ST_SWZ");
#define ST_GO(NV, DPV) do { \
      if (ns) stats_tc_kernel<NV, DPV, true, true><<<grid, ST_THREADS, stats_tc_smem_bytes<NV>(p.kSteps), st>>>(t.mapBhi, t.mapBlo, dm, W, p); \
      else if (expA) stats_tc_kernel<NV, DPV, true, false><<<grid, ST_THREADS, stats_tc_smem_bytes<NV>(), st>>>(t.mapBhi, t.mapBlo, dm, W, p); \
      else stats_tc_kernel<NV, DPV, false, false><<<grid, ST_THREADS, stats_tc_smem_bytes<NV>(), st>>>(t.mapBhi, t.mapBlo, dm, W, p); } while (0)
   if (p.N == 16) { if (dm.D <= 40) ST_GO(16, 40); else ST_GO(16, 64); }
   else { if (dm.D <= 40) ST_GO(32, 40); else ST_GO(32, 64); }
#undef ST_GO
}
