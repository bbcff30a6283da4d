// gmm_tc3.cuh -- K1, third generation: the CTA-pair tcgen05 kernel of gmm_tc.cuh (3xFP16 split, FP32 accumulation in
// TMEM, fused log-sum-exp) with three changes that came out of the round-1 profile:
//
//   * The A operand [1 | x'^2, x' interleaved | 1] (hi and lo halves) is built IN SHARED MEMORY from the raw FP32
//     features by four expander warps, in the 128-byte-swizzled K-major layout the UMMA descriptors expect.  The
//     pre-expanded operand in HBM (0.5 GB written and 0.5 GB read per step on config #3) and the kernel that wrote it
//     (0.16 ms) are gone; the kernel reads 156 B per frame instead of 512.
//   * (tile, frame block) combinations outside the beam taper are never issued -- the reference's Setotprob only
//     evaluates models qLo-1 .. qHi of a frame (HFB.c:1014-1016).  prep_kernel gives every tile of eight (128 / MP)
//     slots the frame interval in which any of its states can be needed; the TMA producer, the MMA issuer and the
//     epilogue all skip a tile whose interval misses the work item's 512 frames, the MMA issuer skips the 256-frame
//     block it misses, and every epilogue warp skips its own 32 frames.
//   * Single-Gaussian sets (MP = 1, config #2) run here too: with "global slots" (slot = tied state) a tile is 128
//     consecutive rows of B, the epilogue has no log-sum-exp and the kernel is bound by writing b.
//
// Replaces, for every (frame, distinct tied state) inside the taper, MOutP/IDOutP (HTKLib/HModel.c:5484-5499,
// :5420-5431) and the mixture log-add of ShStrP (HTKLib/HFB.c:949-960).  Numerics are those of gmm_tc2_kernel<MP, true>
// (same split, same MMA order: corrections first, the five large hi x hi products last into one accumulator); only
// the column order of the operands differs (x'^2 and x' interleaved per dimension so that the expander indexes its
// registers statically).  Frames / states outside the FP16 range go to gmm_fixup_kernel as before.
#pragma once
#include "gmm_tc.cuh"

#define TC3_THREADS 448        // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue, warps 10-13 expanders
#define TC3_ROW0 128           // first real row of the B operand: rows [0, 128) are the dummy tile
#define TC3_DMAX 63            // 2 D + 2 <= 128 columns (kernel variants for D <= 40 and D <= 64)
#define TC3_SMEM_BYTES (2 * 65536 + 3 * 32768 + 512 + 1024)

struct GmmTc3Model {
   bool ready = false;
   int MP = 0;
   long long rows = 0;
   __half *dBhi = nullptr, *dBlo = nullptr;
   size_t bBytes = 0;          // bytes of the operand (hi + lo), starting at dBhi
   float *dOffset = nullptr, *dScale = nullptr;
   CUtensorMap mapBhi, mapBlo;
   float C0 = 0.f;             // what the epilogue subtracts: C0H - C1H
   int globalSlots = 0;        // MP == 1: number of tied states (slot = state)
};

struct Tc3Params {
   const int2 *items;          // (utterance in wave, first frame of the 512-frame work item)
   int nItems;
   const UttDesc *utt;
   const int *slotState;
   const int2 *tileIv;         // per tile: (first, last) frame in which one of its states can be needed
   const float *feat;          // raw features of the wave [frames][D]
   const float *featPad;       // or (single-Gaussian sets): scaled features [frames][DP], 16-byte aligned rows, see gmm_tc3_pad_kernel
   const float *offset, *scale;
   float *b;
   unsigned char *flag;
   uint4 *expA;                // optional output: every frame's expanded operand row, [frames][hi units | lo units] of 16 bytes
                               // (2 * 2 * kSteps per row) -- the tensor-core statistics kernel gathers them instead of
                               // expanding the rows a second time (hfb_stats_tc.cuh)
   float C0;
   int D;
   int kSteps;                 // 16-half K steps that hold data: ceil((2D+2)/16)
   float deadBelow;
   int dbg;
   long long *trace;           // HFBGPU_TC_TRACE: clock64 stamps of pair 0 (gmm_tc4_kernel), else nullptr
};

// element (row r, column k) of one 128-row, 128-column FP16 operand block [chunk 0 | chunk 1], each chunk
// 128 rows x 128 bytes with the 128-byte swizzle (16-byte unit index XOR row mod 8)
__device__ __forceinline__ uint32_t tc3_unit_off(int r, int unit)       // unit = 16-byte unit 0..15 across both chunks
{
   return (uint32_t)((unit >> 3) * 16384 + r * 128 + (((unit & 7) ^ (r & 7)) << 4));
}

// cluster-scope release / acquire around the hand-written A blocks: the expanders of BOTH CTAs arrive on the leader's
// barrier after their generic-proxy stores (made visible to the tensor core by fence.proxy.async), the leader's MMA
// thread acquires at cluster scope before issuing MMAs that read both CTAs' shared memory
__device__ __forceinline__ void tc3_arrive_leader_release(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(tc_smem_u32(bar) & TC_PEER_MASK) : "memory");
}
__device__ __forceinline__ void tc3_wait_acquire_cluster(uint64_t *bar, uint32_t parity)
{
   uint32_t done, addr = tc_smem_u32(bar);
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
   } while (!done);
}

template <int MP, int DP>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC3_THREADS, 1)
gmm_tc3_kernel(const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, Tc3Params p)
{
   extern __shared__ uint8_t tc_smem_raw[];
   uint8_t *base = (uint8_t *)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
   constexpr uint32_t A_BLK = 65536;                    // one 128-frame block: [hi c0 | hi c1 | lo c0 | lo c1] x 16 KB
   constexpr int NST = 3;
   constexpr uint32_t ST_BYTES = 32768;                 // this CTA's half of a B tile: 2 chunks x [hi 8 KB | lo 8 KB]
   uint8_t *sA = base;
   uint8_t *sB = base + 2 * A_BLK;
   uint64_t *bars = (uint64_t *)(sB + NST * ST_BYTES);
   uint64_t *fullA = bars, *emptyA = bars + 1, *fullB = bars + 2, *emptyB = bars + 2 + NST;
   uint64_t *tmemFull = bars + 2 + 2 * NST, *tmemEmpty = tmemFull + 2;
   uint32_t *tmemSlot = (uint32_t *)(tmemEmpty + 2);
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const uint32_t rank = tc_cluster_ctarank();          // 0 = leader (issues the MMAs)
   const int pair = blockIdx.x >> 1, nPairs = gridDim.x >> 1;
   constexpr int SPT = TC_BN / MP;                      // states per tile
   constexpr int EPW = 8;
   constexpr int HB = (SPT >= 2) ? ((MP == 1) ? 1 : SPT / 2) : 1;   // B boxes per operand half held by one CTA
   constexpr int BOXR = (MP == 1) ? 64 : ((MP < 64) ? MP : 64);     // rows per box

   if (warp == 0 && lane == 0) {
      // fullA: both CTAs' expanders (4 warps each) arrive on the LEADER's barrier; emptyA: one commit, multicast
      tc_mbar_init(fullA, 8); tc_mbar_init(emptyA, 1);
      for (int s = 0; s < NST; s++) { tc_mbar_init(&fullB[s], 1); tc_mbar_init(&emptyB[s], 1); }
      for (int s = 0; s < 2; s++) { tc_mbar_init(&tmemFull[s], 1); tc_mbar_init(&tmemEmpty[s], 2 * EPW); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmemSlot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
   }
   tc_fence_before();
   tc_cluster_sync();
   tc_fence_after();
   const uint32_t tmem = *tmemSlot;
   const int nChunks = (p.kSteps + 3) >> 2;             // 64-half chunks that hold data (1 or 2)

   // which of the item's two 256-frame blocks need tile n: bit b set = frames [y + 256 b, y + 256 (b+1)) ∩ [0, T)
   // intersect the tile's interval.  Identical in every role of both CTAs (same inputs), so they walk the same tiles.
   // The intervals are read ONE TILE AHEAD (an L2 round trip per tile on the critical path of the epilogue warps cost
   // 0.3 ms per step on config #3).
   auto tile_need = [&](int f, int l, int y, int T) -> int {
      int need = 0;
      if (f < min(T, y + 2 * TC_BM) && l >= y) need |= 1;
      if (T - y > 2 * TC_BM && f < min(T, y + 4 * TC_BM) && l >= y + 2 * TC_BM) need |= 2;
      return need;
   };

   if (warp == 0) {
      // ================= TMA producer (both CTAs): own half of every needed B tile =================
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
      uint32_t stage = 0, phB = 0;
      for (int it = pair; it < p.nItems; it += nPairs) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         const int nTiles = (u.Jt + SPT - 1) / SPT;
         const int *ss = p.slotState + u.slotOff;
         const int2 *tiv = p.tileIv + u.slotOff;
         const int bi = lane % HB;                              // lanes [0,HB): hi boxes, [HB,2HB): lo boxes
         int2 ivN = tiv[0];
         for (int n = 0; n < nTiles; n++) {
            const int2 iv = ivN;
            if (n + 1 < nTiles) ivN = tiv[n + 1];
            if (!tile_need(iv.x, iv.y, item.y, u.T)) continue;
            int row = 0;                                        // rows [0, 128) = dummy state ("log zero")
            if (MP == 1) row = TC3_ROW0 + n * TC_BN + (int)rank * 64;  // global slots: 64 consecutive states per CTA
            else if (lane < 2 * HB) {
               if (SPT >= 2) { const int slot = n * SPT + (int)rank * HB + bi; if (slot < u.Jt) row = TC3_ROW0 + ss[slot] * MP; }
               else row = TC3_ROW0 + ss[n] * MP + (int)rank * 64;      // MP = 128: each CTA takes 64 rows of the state
            }
            tc_mbar_wait(&emptyB[stage], phB ^ 1);
            if (elected && rank == 0) tc_mbar_expect_tx(&fullB[stage], 2 * nChunks * 16384);
            for (int k = 0; k < nChunks; k++) {
               uint8_t *dst = sB + stage * ST_BYTES + k * 16384;
#pragma unroll
               for (int j = 0; j < 2 * HB; j++) {
                  const int rj = __shfl_sync(0xffffffffu, row, j);
                  if (elected)
                     tc_tma_load_2d_pair(dst + (j / HB) * 8192 + (j % HB) * (BOXR * 128), (j / HB) ? &mapBlo : &mapBhi,
                                         &fullB[stage], k * 64, rj);
               }
            }
            __syncwarp();
            if (++stage == NST) { stage = 0; phB ^= 1; }
         }
      }
   } else if (warp == 1) {
      // ================= MMA issuer: one elected thread of the leader CTA =================
      if (rank == 0) {
         uint32_t elected;
         asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
         const uint32_t idesc = tc_idesc(2 * TC_BM, TC_BN, 0u);
         const uint32_t aBase = tc_smem_u32(sA), bBase = tc_smem_u32(sB);
         uint32_t stage = 0, phB = 0, phA = 0, tile = 0;
         for (int it = pair; it < p.nItems; it += nPairs) {
            const int2 item = p.items[it];
            const UttDesc u = p.utt[item.x];
            const int nTiles = (u.Jt + SPT - 1) / SPT;
            const int2 *tiv = p.tileIv + u.slotOff;
            int2 ivN = tiv[0];
            tc3_wait_acquire_cluster(fullA, phA);               // both CTAs' A blocks are in shared memory
            phA ^= 1;
            tc_fence_after();
            for (int n = 0; n < nTiles; n++) {
               const int2 iv = ivN;
               if (n + 1 < nTiles) ivN = tiv[n + 1];
               const int need = tile_need(iv.x, iv.y, item.y, u.T);
               if (!need) continue;
               const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
               tc_mbar_wait(&tmemEmpty[as], phT ^ 1);
               tc_fence_after();
               const uint32_t dMain = tmem + as * (2 * TC_BN);
               tc_mbar_wait(&fullB[stage], phB);
               tc_fence_after();
               const uint32_t bSt = bBase + stage * ST_BYTES;
               for (int b = 0; b < 2; b++) {
                  if (!(need & (1 << b))) continue;
                  const uint32_t aB = aBase + b * A_BLK, dAcc = dMain + b * TC_BN;
#pragma unroll
                  for (int ks = 0; ks < 8; ks++) {
                     if (ks >= p.kSteps) break;
                     const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
                     const uint64_t dAhi = tc_smem_desc(aB + o), dAlo = tc_smem_desc(aB + 32768 + o);
                     const uint64_t dBhi = tc_smem_desc(bSt + o), dBlo = tc_smem_desc(bSt + o + 8192);
                     if (elected) {
                        tc_mma_pair<true>(dAcc, dAhi, dBlo, idesc, ks ? 1u : 0u);
                        tc_mma_pair<true>(dAcc, dAlo, dBhi, idesc, 1u);
                     }
                  }
#pragma unroll
                  for (int ks = 0; ks < 8; ks++) {
                     if (ks >= p.kSteps) break;
                     const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
                     const uint64_t dAhi = tc_smem_desc(aB + o), dBhi = tc_smem_desc(bSt + o);
                     if (elected) tc_mma_pair<true>(dAcc, dAhi, dBhi, idesc, 1u);
                  }
               }
               if (elected) tc_commit_pair(&emptyB[stage]);
               __syncwarp();
               if (++stage == NST) { stage = 0; phB ^= 1; }
               if (elected) tc_commit_pair(&tmemFull[as]);
               __syncwarp();
               tile++;
            }
            if (elected) tc_commit_pair(emptyA);                 // A blocks reusable (arrives in both CTAs)
            __syncwarp();
         }
      }
   } else if (warp < 2 + EPW) {
      // ================= epilogue (both CTAs): own frames x 128 components per block =================
      // (Measured and dropped, round 2: accumulator hand-over per 128-frame block -- "full" and "empty" barriers per
      // (buffer, block), so that block 0's epilogue overlaps block 1's MMAs -- and handing a block back as soon as its two
      // chunks are in registers, before the log-sum-exp.  Both made the kernel SLOWER, 1.86 -> 2.03 / 2.00 ms on config #3,
      // and the no-epilogue-math run went 1.41 -> 1.57: the second commit / second cluster-wide wait per tile costs the
      // MMA-issuing thread more than the finer hand-over gains.)
      const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
      constexpr int CPW = 2;                            // 32-column chunks per warp (8 warps: two per quadrant)
      const int c0 = ((warp - 2) >> 2) * CPW;
      const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
      uint32_t tile = 0;
      for (int it = pair; it < p.nItems; it += nPairs) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         const int nTiles = (u.Jt + SPT - 1) / SPT;
         const int2 *tiv = p.tileIv + u.slotOff;
         int2 ivN = tiv[0];
         for (int n = 0; n < nTiles; n++) {
            const int2 iv = ivN;
            if (n + 1 < nTiles) ivN = tiv[n + 1];
            const int need = tile_need(iv.x, iv.y, item.y, u.T);
            if (!need) continue;
            const int f = iv.x, l = iv.y;
            const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
            tc_mbar_wait(&tmemFull[as], phT);
            tc_fence_after();
            for (int blk = 0; blk < 2; blk++) {
               const int w0 = item.y + (2 * blk + (int)rank) * TC_BM + quad * 32;      // first frame of this warp
               // nothing of this warp's 32 frames lies inside the tile's interval (or inside the utterance)
               if (!(need & (1 << blk)) || w0 >= u.T || f >= w0 + 32 || l < w0 || (p.dbg & 2)) continue;
               const int t = w0 + lane;
               float *brow = p.b + u.bOff + (size_t)t * u.J;
               const uint32_t taddr = tmem + as * (2 * TC_BN) + blk * TC_BN + ((uint32_t)(quad * 32) << 16);
               const float C0 = p.C0;
               if (MP == 1) {
                  // one column = one state: no log-sum-exp, 32 consecutive slots per chunk and lane
#pragma unroll
                  for (int cc = 0; cc < CPW; cc++) {
                     const int c = c0 + cc;
                     float v[32];
                     tc_tmem_ld32(taddr + c * 32, v);
                     const int slot0 = n * SPT + c * 32;
                     bool far = false;
#pragma unroll
                     for (int i = 0; i < 32; i++) { far |= (v[i] < p.deadBelow) && (slot0 + i < u.Jt); v[i] -= C0; }
                     if (t < u.T) {
                        if (far) p.flag[u.frameBase + t] = 1;
#pragma unroll
                        for (int g = 0; g < 8; g++)
                           if (slot0 + 4 * g < u.J)
                              *reinterpret_cast<float4 *>(brow + slot0 + 4 * g) = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
                     }
                  }
                  continue;
               }
               float cmx = -INFINITY, csum = 0.f;       // carry for states wider than one 32-column chunk
               constexpr int NOUT = (MP <= 32) ? CPW * 32 / MP : 0;
               float outv[NOUT > 0 ? NOUT : 1];
               int no = 0;
#pragma unroll
               for (int cc = 0; cc < CPW; cc++) {
                  const int c = c0 + cc;
                  float v[32];
                  tc_tmem_ld32(taddr + c * 32, v);
                  constexpr int G = (MP < 32) ? MP : 32;   // columns of one state inside this chunk
#pragma unroll
                  for (int s0 = 0; s0 < 32; s0 += G) {
                     float mx = v[s0];
#pragma unroll
                     for (int i = 1; i < G; i++) mx = fmaxf(mx, v[s0 + i]);
                     float sum = 0.f;
                     const float mb = mx * LOG2E;
#pragma unroll
                     for (int i = 0; i < G; i++) sum += tc_ex2(fmaf(v[s0 + i], LOG2E, -mb));
                     if (MP > 32) {                        // merge into the carry
                        float nm = fmaxf(cmx, mx);
                        csum = csum * tc_ex2((cmx - nm) * LOG2E) + sum * tc_ex2((mx - nm) * LOG2E);
                        cmx = nm; mx = cmx; sum = csum;
                     }
                     const int colEnd = c * 32 + s0 + G;   // columns consumed so far
                     if (colEnd % MP == 0) {
                        const int slot = n * SPT + colEnd / MP - 1;
                        float val = (mx < p.deadBelow) ? (float)HFB_LZERO : fmaf(tc_lg2(sum), LN2, mx - C0);
                        // "no live component" and "every live component far away" look alike here: gmm_fixup_kernel decides
                        if (mx < p.deadBelow && t < u.T && slot < u.Jt) p.flag[u.frameBase + t] = 1;
                        if (NOUT > 0) outv[no++] = val;
                        else if (t < u.T && slot < u.J) brow[slot] = val;
                        cmx = -INFINITY; csum = 0.f;
                     }
                  }
               }
               if (NOUT > 0 && t < u.T) {
                  const int slot0 = n * SPT + c0 * 32 / MP;
                  if (NOUT >= 4) {
#pragma unroll
                     for (int g = 0; g < NOUT / 4; g++)
                        if (slot0 + 4 * g < u.J)
                           *reinterpret_cast<float4 *>(brow + slot0 + 4 * g) = make_float4(outv[4 * g], outv[4 * g + 1], outv[4 * g + 2], outv[4 * g + 3]);
                  } else if (slot0 < u.J)
                     *reinterpret_cast<float2 *>(brow + slot0) = make_float2(outv[0], outv[NOUT > 1 ? 1 : 0]);
               }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive_leader(&tmemEmpty[as]);
            tile++;
         }
      }
   } else {
      // ================= expanders (both CTAs, 4 warps): raw FP32 features -> this CTA's two A blocks =================
      // Thread r owns row r of both blocks.  Columns: 0 = 1 (pairs with the symmetrising constant), 2d+1 = x'_d^2,
      // 2d+2 = x'_d (d < D), 2D+1 = 1 (pairs with the Gaussian constant), the rest 0; x' = (x - offset) * scale.
      // The rows of the NEXT item travel into registers while the tensor core works on the current one (the loads are a
      // row per thread, i.e. 32 cache lines per instruction: slow, but hidden there); what is exposed between two items
      // is only the conversion and ten 16-byte stores per row and half.  Measured the other way round -- lane =
      // dimension, coalesced loads, 2-byte stores -- the instruction count (70 per row and warp) cost 17 us per item.
      const int r = threadIdx.x - (2 + EPW) * 32;       // 0..127
      const int D = p.D;
      const int nUnits = 2 * p.kSteps;                  // 16-byte units (8 columns) the MMAs read
      constexpr int NPRE = (DP <= 40) ? 2 : 1;          // rows held in registers ahead of time
      uint32_t phA = 0;
      float x[NPRE][DP];
      auto load_row = [&](float (&xr)[DP], const int2 item, const UttDesc &u, int b) {
         const int t = item.y + (2 * b + (int)rank) * TC_BM + r;
         const bool inside = t < u.T;
         bool far = false;
         if (p.featPad != nullptr) {
            // pre-scaled rows of DP floats: ten 16-byte loads per row instead of 39 scalar ones (with one tile per item --
            // single-Gaussian sets -- the row loads are NOT hidden behind the previous item's MMAs and were what the
            // kernel waited for: 32 cache lines per load instruction)
            const float4 *s4 = reinterpret_cast<const float4 *>(p.featPad + ((size_t)u.featOff + (inside ? t : 0)) * DP);
#pragma unroll
            for (int i = 0; i < DP / 4; i++) {
               const float4 q = inside ? s4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
               xr[4 * i] = q.x; xr[4 * i + 1] = q.y; xr[4 * i + 2] = q.z; xr[4 * i + 3] = q.w;
            }
            return;
         }
         const float *src = p.feat + ((size_t)u.featOff + (inside ? t : 0)) * D;
#pragma unroll
         for (int d = 0; d < DP; d++) {
            float v = 0.f;
            if (d < D && inside) {
               v = (src[d] - p.offset[d]) * p.scale[d];
               if (!(fabsf(v) <= TC_FAR)) far = true;
               v = fminf(fmaxf(v, -250.f), 250.f);             // keeps inf / NaN out of the tensor core; a far row is recomputed
            }
            xr[d] = v;
         }
         if (inside) p.flag[u.frameBase + t] = far ? 1 : 0;
      };
      auto store_row = [&](const float (&xr)[DP], int b, uint4 *grow, const bool toSmem) {
         uint8_t *blk = sA + b * A_BLK;
#pragma unroll
         for (int un = 0; un < 16; un++) {
            if (un >= nUnits) break;                   // every unit the MMAs read is rewritten for every item
            float v[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
               const int k = un * 8 + e;                // compile-time column
               if (k == 0) v[e] = 1.f;
               else if (k & 1) { const int d = (k - 1) >> 1; v[e] = (d < DP && d < D) ? xr[d < DP ? d : 0] * xr[d < DP ? d : 0] : ((d == D) ? 1.f : 0.f); }
               else { const int d = (k - 2) >> 1; v[e] = (d < DP && d < D) ? xr[d < DP ? d : 0] : 0.f; }
            }
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int e = 0; e < 4; e++) {
               const __half2 h = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
               const float2 hf = __half22float2(h);
               const __half2 l = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
               hi[e] = *reinterpret_cast<const uint32_t *>(&h);
               lo[e] = *reinterpret_cast<const uint32_t *>(&l);
            }
            if (toSmem) {
               const uint32_t off = tc3_unit_off(r, un);
               *reinterpret_cast<uint4 *>(blk + off) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
               *reinterpret_cast<uint4 *>(blk + 32768 + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
            if (grow != nullptr) {
               grow[un] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
               grow[nUnits + un] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
            }
         }
      };
      // where block b's row of this thread goes in the expanded-operand output (nullptr: not wanted / past the utterance)
      auto exp_row = [&](const int2 item, const UttDesc &u, int b) -> uint4 * {
         const int t = item.y + (2 * b + (int)rank) * TC_BM + r;
         return (p.expA != nullptr && t < u.T) ? p.expA + (size_t)(u.frameBase + t) * (size_t)(2 * nUnits) : nullptr;
      };
      int2 item = make_int2(0, 0);
      UttDesc u;
      if (pair < p.nItems) {
         item = p.items[pair]; u = p.utt[item.x];
         load_row(x[0], item, u, 0);
         if (NPRE == 2) load_row(x[NPRE - 1], item, u, 1);
      }
      for (int it = pair; it < p.nItems; it += nPairs) {
         const bool skipX = (p.dbg & 32) && it != pair;  // timing experiment: reuse the first item's A blocks
         tc_mbar_wait(emptyA, phA ^ 1);
         phA ^= 1;
         if (!skipX) {
            // (the copy of the rows to global memory waits until the MMA thread has been told, below: with both rows in
            // registers the conversion is simply done a second time, outside the window the tensor core idles in)
            store_row(x[0], 0, NPRE == 1 ? exp_row(item, u, 0) : nullptr, true);
            if (NPRE == 1) load_row(x[0], item, u, 1);
            store_row(x[NPRE - 1], 1, NPRE == 1 ? exp_row(item, u, 1) : nullptr, true);
         }
         // generic-proxy writes -> visible to the tensor core (async proxy), then tell the leader's MMA thread
         asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
         __syncwarp();
         if (lane == 0) tc3_arrive_leader_release(fullA);
         if (NPRE == 2 && p.expA != nullptr && !skipX) {
            store_row(x[0], 0, exp_row(item, u, 0), false);
            store_row(x[NPRE - 1], 1, exp_row(item, u, 1), false);
         }
         // the next item's rows: in flight during this item's MMAs
         if (it + nPairs < p.nItems) {
            item = p.items[it + nPairs]; u = p.utt[item.x];
            load_row(x[0], item, u, 0);
            if (NPRE == 2) load_row(x[NPRE - 1], item, u, 1);
         }
      }
   }
   tc_fence_before();
   tc_cluster_sync();                                   // neither CTA leaves while the other may still signal it
   if (warp == 1) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
   }
}

// Scaled, clamped copy of the wave's features with rows padded to DP floats (16-byte aligned) + the per-frame flags;
// only for sets with one K1 tile per work item (single-Gaussian sets), see load_row above.
// expA != nullptr: also every frame's expanded operand row ([hi units | lo units], as the expanders of the GMM kernels
// build it) for the statistics kernel -- written here, by a streaming kernel, because the single-Gaussian GMM kernel is
// bound by its load / store queue.
__global__ void __launch_bounds__(256)
gmm_tc3_pad_kernel(const float *__restrict__ feat, const float *__restrict__ offset, const float *__restrict__ scale,
                   int D, int DPad, long long nFrames, float *__restrict__ out, unsigned char *__restrict__ flag,
                   uint4 *__restrict__ expA, int kSteps)
{
   const int lane = threadIdx.x & 31;
   const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);     // one warp per frame
   if (f >= nFrames) return;
   bool far = false;
   float v01[2] = {0.f, 0.f};                                              // dimensions lane and lane + 32
   for (int d = lane, i = 0; d < DPad; d += 32, i++) {
      float v = 0.f;
      if (d < D) {
         v = (feat[f * D + d] - offset[d]) * scale[d];
         if (!(fabsf(v) <= TC_FAR)) far = true;
         v = fminf(fmaxf(v, -250.f), 250.f);
      }
      out[f * DPad + d] = v;
      v01[i] = v;
   }
   far = __any_sync(0xffffffffu, far);
   if (lane == 0) flag[f] = far ? 1 : 0;
   if (expA != nullptr) {
      // lane un < 2 kSteps builds unit un = operand columns 8 un .. 8 un + 7: column 0 = 1, 2d+1 = x'_d^2, 2d+2 = x'_d, 2D+1 = 1
      const int nUn = 2 * kSteps, un = (lane < nUn) ? lane : 0;
      float c[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
         const int k = 8 * un + e;
         const int d = (k & 1) ? (k - 1) >> 1 : (k - 2) >> 1;
         const int dd = (d >= 0 && d < D) ? d : 0;
         const float a = __shfl_sync(0xffffffffu, v01[0], dd & 31), b = __shfl_sync(0xffffffffu, v01[1], dd & 31);
         const float x = (dd >= 32) ? b : a;
         if (k == 0) c[e] = 1.f;
         else if (k & 1) c[e] = (d < D) ? x * x : ((d == D) ? 1.f : 0.f);
         else c[e] = (d >= 0 && d < D) ? x : 0.f;
      }
      if (lane < nUn) {
         uint32_t hi[4], lo[4];
#pragma unroll
         for (int e = 0; e < 4; e++) {
            const __half2 h = __floats2half2_rn(c[2 * e], c[2 * e + 1]);
            const float2 hf = __half22float2(h);
            const __half2 l = __floats2half2_rn(c[2 * e] - hf.x, c[2 * e + 1] - hf.y);
            hi[e] = *reinterpret_cast<const uint32_t *>(&h);
            lo[e] = *reinterpret_cast<const uint32_t *>(&l);
         }
         uint4 *row = expA + (size_t)f * (size_t)(2 * nUn);
         row[lane] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
         row[nUn + lane] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      }
   }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
static inline void gmm_tc3_release(GmmTc3Model &t)
{
   if (t.dBhi) cudaFree(t.dBhi);                         // dBlo lives in the same allocation
   if (t.dOffset) cudaFree(t.dOffset);
   if (t.dScale) cudaFree(t.dScale);
   t = GmmTc3Model();
}

// B operand, one row of 128 halfs per mixture component, states padded to MP rows, rows [0, 128) = dummy ("log zero"):
//   [ C0H | -ivar_d / (2 s_d^2), (mu_d - o_d) ivar_d / s_d  interleaved over d | c - C1 at column 2D+1 | 0... ]
// Same quantities as gmm_tc_prepare's FP16 operands (gmm_tc.cuh), other column order.
static inline int gmm_tc3_prepare(GmmTc3Model &t, const hfb_model *m, cudaStream_t st, void *encodeFn)
{
   t = GmmTc3Model();
   const int D = m->vecSize, J = m->numStates;
   if (!encodeFn || 2 * D + 2 > TC_KH || D > TC3_DMAX) return HFB_OK;
   int maxM = 0;
   for (int s = 0; s < J; s++) maxM = std::max(maxM, m->stateMixOff[s + 1] - m->stateMixOff[s]);
   int MP;
   if (maxM == 1) {
      if (J > 4096) return HFB_OK;                       // big single-Gaussian sets keep the FP32 kernel
      MP = 1;
   } else {
      MP = 8;
      while (MP < maxM) MP *= 2;
      if (MP > 64) return HFB_OK;                         // eight epilogue warps = two 32-column chunks each: a state spans at most one warp's share
   }
   int dev = 0, major = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
   if (major != 10) return HFB_OK;
   t.MP = MP;
   t.rows = TC3_ROW0 + (((long long)J * MP + TC_BN - 1) / TC_BN) * TC_BN;   // whole tiles: no box starts out of bounds
   std::vector<double> off(D, 0.0);
   for (int g = 0; g < m->numGauss; g++)
      for (int k = 0; k < D; k++) off[k] += m->mean[(size_t)g * D + k];
   std::vector<float> offF(D), sc(D);
   for (int k = 0; k < D; k++) offF[k] = (float)(off[k] / m->numGauss);
   for (int k = 0; k < D; k++) {
      double a = 0.0;
      for (int g = 0; g < m->numGauss; g++) a += m->ivar[(size_t)g * D + k];
      a /= m->numGauss;
      sc[k] = (float)ldexp(1.0, (int)lrint(0.5 * log2(a > 1e-30 ? a : 1e-30)));
   }
   std::vector<double> cRow((size_t)t.rows, 0.0);
   std::vector<char> live((size_t)t.rows, 0);
   double c1 = 0.0; long long nLive = 0;
   for (int s2 = 0; s2 < J; s2++) {
      const int mo = m->stateMixOff[s2], Mn = m->stateMixOff[s2 + 1] - mo;
      for (int k2 = 0; k2 < Mn; k2++) {
         const long long r = TC3_ROW0 + (long long)s2 * MP + k2;
         const float wt = m->mixLogWt[mo + k2];
         if (Mn > 1 && !(wt > (float)HFB_LMINMIX)) continue;
         const int g = m->mixGauss[mo + k2];
         double c = m->gConst[g];
         for (int k = 0; k < D; k++) {
            const double iv = m->ivar[(size_t)g * D + k], mu = (double)m->mean[(size_t)g * D + k] - (double)offF[k];
            c += mu * mu * iv;
         }
         cRow[r] = -0.5 * c + (Mn > 1 ? (double)wt : 0.0);
         live[r] = 1; c1 += cRow[r]; nLive++;
      }
   }
   c1 = nLive ? c1 / nLive : 0.0;
   std::vector<__half> hh((size_t)t.rows * TC_KH, __float2half_rn(0.f)), hl((size_t)t.rows * TC_KH, __float2half_rn(0.f));
   bool inRange = true;
   auto putH = [&](long long r, int k, double v, double lim) {
      if (!(fabs(v) < lim)) inRange = false;
      const float f = (float)v;
      const __half h = __float2half_rn(f);
      hh[(size_t)r * TC_KH + k] = h;
      hl[(size_t)r * TC_KH + k] = __float2half_rn(f - __half2float(h));
   };
   for (long long r = 0; r < t.rows; r++) if (!live[r]) putH(r, 2 * D + 1, -60000.0, 65000.0);     // dead rows: log zero
   for (int s2 = 0; s2 < J; s2++) {
      const int mo = m->stateMixOff[s2], Mn = m->stateMixOff[s2 + 1] - mo;
      for (int k2 = 0; k2 < Mn; k2++) {
         const long long r = TC3_ROW0 + (long long)s2 * MP + k2;
         if (!live[r]) continue;
         const int g = m->mixGauss[mo + k2];
         for (int k = 0; k < D; k++) {
            const double iv = m->ivar[(size_t)g * D + k], mu = (double)m->mean[(size_t)g * D + k] - (double)offF[k], sk = sc[k];
            putH(r, 2 * k + 1, -0.5 * iv / (sk * sk), 30000.0);
            putH(r, 2 * k + 2, mu * iv / sk, 30000.0);
         }
         putH(r, 2 * D + 1, cRow[r] - c1, 20000.0);
      }
   }
   if (!inRange) return HFB_OK;                          // scaled parameters leave the FP16 range: older kernels
   const float C0H = __half2float(__float2half_rn((float)(0.25 * D)));        // -half of the expected product part of log b
   for (long long r = 0; r < t.rows; r++) hh[(size_t)r * TC_KH] = __float2half_rn(C0H);
   t.C0 = C0H - (float)c1;
   const size_t hb = hh.size() * sizeof(__half);
   // hi and lo halves in ONE allocation: a single L2 access-policy window can cover the whole operand (hfbgpu.cu)
   t.bBytes = 2 * hb;
   if (cudaMalloc(&t.dBhi, 2 * hb) != cudaSuccess || (t.dBlo = (__half *)((char *)t.dBhi + hb), false) ||
       cudaMalloc(&t.dOffset, D * sizeof(float)) != cudaSuccess || cudaMalloc(&t.dScale, D * sizeof(float)) != cudaSuccess) {
      cudaGetLastError(); gmm_tc3_release(t); return HFB_ENOMEM;
   }
   cudaMemcpyAsync(t.dBhi, hh.data(), hb, cudaMemcpyHostToDevice, st);
   cudaMemcpyAsync(t.dBlo, hl.data(), hb, cudaMemcpyHostToDevice, st);
   cudaMemcpyAsync(t.dOffset, offF.data(), D * sizeof(float), cudaMemcpyHostToDevice, st);
   cudaMemcpyAsync(t.dScale, sc.data(), D * sizeof(float), cudaMemcpyHostToDevice, st);
   cudaStreamSynchronize(st);
   const int boxR = (MP == 1) ? 64 : std::min(MP, 64);
   if (tc_make_map_f16(encodeFn, &t.mapBhi, t.dBhi, t.rows, boxR) || tc_make_map_f16(encodeFn, &t.mapBlo, t.dBlo, t.rows, boxR)) {
      gmm_tc3_release(t); return HFB_OK;
   }
#define TC3_SET(MPV) cudaFuncSetAttribute(gmm_tc3_kernel<MPV, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES); \
                     cudaFuncSetAttribute(gmm_tc3_kernel<MPV, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM_BYTES)
   TC3_SET(1); TC3_SET(8); TC3_SET(16); TC3_SET(32); TC3_SET(64); TC3_SET(128);
#undef TC3_SET
   t.ready = (cudaGetLastError() == cudaSuccess);
   t.globalSlots = (MP == 1) ? J : 0;
   return HFB_OK;
}

// gmm_tc4.cuh: the same work with the A operand in tensor memory (the default; HFBGPU_GMM_V3 keeps this file's kernel)
static inline void gmm_tc4_go(const GmmTc3Model &t, const Tc3Params &p, int D, int grid2, cudaStream_t st, bool lean);

static inline int gmm_tc3_launch(GmmTc3Model &t, GmmTcWork &wk, const DevModel &dm, const Wave &W, long long waveFrames,
                                 const int2 *dItems128, int nItems128, const int2 *dItems4, int nItems4, int smCount,
                                 cudaStream_t st, int *launches, bool wantExp = false, bool lean = false)
{
   if (!t.ready) return HFB_EUNSUPPORTED;
   if (nItems4 == 0) return HFB_OK;
   const size_t need = (size_t)waveFrames + TC_BM;
   if (need > wk.flagCap) {
      if (wk.dFlag3) cudaFree(wk.dFlag3);
      wk.dFlag3 = nullptr; wk.flagCap = 0;
      if (cudaMalloc(&wk.dFlag3, need + need / 8) != cudaSuccess) { cudaGetLastError(); return HFB_ENOMEM; }
      wk.flagCap = need + need / 8;
   }
   Tc3Params p;
   p.items = dItems4; p.nItems = nItems4; p.utt = W.utt; p.slotState = W.slotState;
   p.tileIv = W.tileIv;
   p.feat = W.feat; p.featPad = nullptr; p.offset = t.dOffset; p.scale = t.dScale; p.b = W.b; p.flag = wk.dFlag3;
   p.expA = nullptr;
   if (wantExp) {
      const size_t units = (size_t)4 * ((2 * dm.D + 2 + 15) / 16);          // 16-byte units per row: hi + lo
      const size_t needU = ((size_t)waveFrames + TC_BM) * units;
      if (needU > wk.expCap) {
         if (wk.dExpA) cudaFree(wk.dExpA);
         wk.dExpA = nullptr; wk.expCap = 0;
         if (cudaMalloc(&wk.dExpA, (needU + needU / 8) * sizeof(uint4)) != cudaSuccess) { cudaGetLastError(); return HFB_ENOMEM; }
         wk.expCap = needU + needU / 8;
      }
      p.expA = wk.dExpA;
   }
   int nl = 0;
   // Single-Gaussian sets: one tile per work item, so the row loads of the expanders are NOT hidden behind MMAs and the
   // kernel waits for its load / store queue (`lg_throttle`, profiles/r2_gmm_tc3_cfg2_raw.txt).  A pre-pass that scales the
   // features into 16-byte-aligned rows turns 39 scalar loads per row into ten vector loads: config #2, A/B on one box,
   // 0.70 ms without it, 0.47-0.48 ms with it (pre-pass included).  HFBGPU_NO_PAD switches it off.
   if (t.MP == 1 && !getenv("HFBGPU_NO_PAD")) {
      const int DPad = (dm.D <= 40) ? 40 : 64;
      const size_t needPad = ((size_t)waveFrames + TC_BM) * DPad;
      if (needPad > wk.padCap) {
         if (wk.dPad) cudaFree(wk.dPad);
         wk.dPad = nullptr; wk.padCap = 0;
         if (cudaMalloc(&wk.dPad, (needPad + needPad / 8) * sizeof(float)) != cudaSuccess) { cudaGetLastError(); return HFB_ENOMEM; }
         wk.padCap = needPad + needPad / 8;
      }
      gmm_tc3_pad_kernel<<<(unsigned)((waveFrames + 7) / 8), 256, 0, st>>>(W.feat, t.dOffset, t.dScale, dm.D, DPad, waveFrames, wk.dPad, wk.dFlag3,
                                                                            p.expA, (2 * dm.D + 2 + 15) / 16);
      p.featPad = wk.dPad;
      p.expA = nullptr;                                  // written by the pre-pass, not by the GMM kernel
      nl++;
   }
   p.C0 = t.C0; p.D = dm.D; p.kSteps = (2 * dm.D + 2 + 15) / 16; p.deadBelow = TC_DEAD_BELOW;
   { const char *e = getenv("HFBGPU_TC_DEBUG"); p.dbg = e ? atoi(e) : 0; }
   p.trace = nullptr;
   static long long *dTrace = nullptr;
   if (getenv("HFBGPU_TC_TRACE")) {
      if (!dTrace) cudaMalloc(&dTrace, 8 * 4096 * sizeof(long long));
      cudaMemsetAsync(dTrace, 0, 8 * 4096 * sizeof(long long), st);
      p.trace = dTrace;
   }
   const int grid2 = 2 * std::min(nItems4, smCount / 2);
#define TC3_GO(MPV) do { if (dm.D <= 40) gmm_tc3_kernel<MPV, 40><<<grid2, TC3_THREADS, TC3_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p); \
                          else gmm_tc3_kernel<MPV, 64><<<grid2, TC3_THREADS, TC3_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p); } while (0)
   const bool ssMode = getenv("HFBGPU_GMM_V3") != nullptr;
   if (!ssMode) gmm_tc4_go(t, p, dm.D, grid2, st, lean);
   else
   switch (t.MP) {
   case 1: TC3_GO(1); break;
   case 8: TC3_GO(8); break;
   case 16: TC3_GO(16); break;
   case 32: TC3_GO(32); break;
   case 64: TC3_GO(64); break;
   default: TC3_GO(128); break;
   }
#undef TC3_GO
   nl += 1;
   if (p.trace) {                                        // diagnostic: per-block stamps of the MMA thread and of one epilogue warp
      static int printed = 0;
      cudaStreamSynchronize(st);
      if (printed++ == 3) {
         std::vector<long long> h(8 * 4096);
         cudaMemcpy(h.data(), dTrace, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
         const long long t0 = h[0];
         for (int i = 40; i < 72; i++)
            fprintf(stderr, "[tc trace] blk %3d  mma: waitB %6lld  gotEmpty %6lld  issued %6lld | epi(cta0 w2): gotFull %6lld arrived %6lld | epi(cta1 w2): gotFull %6lld arrived %6lld\n", i,
                    h[8 * i] - t0, h[8 * i + 1] - t0, h[8 * i + 2] - t0, h[8 * i + 3] - t0, h[8 * i + 4] - t0, h[8 * i + 5] - t0, h[8 * i + 6] - t0);
      }
   }
   if (!getenv("HFBGPU_NO_FIXUP")) { gmm_fixup_kernel<<<nItems128, 128, 0, st>>>(dm, W, dItems128, wk.dFlag3); nl++; }
   if (launches) *launches = nl;
   return HFB_OK;
}
