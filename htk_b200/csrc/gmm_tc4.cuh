// gmm_tc4.cuh -- K1, fourth version: the A operand of the state log-likelihood product lives in TENSOR MEMORY.
//
// gmm_tc3_kernel keeps the expanded feature rows (A) in shared memory and every tcgen05.mma reads both operands from
// there.  Measured with tools/mma_rate.cu on B200: a kind::f16 MMA of the shape this path uses (cta_group::2, M = 256,
// N = 128, K = 16) costs 104 cycles with A in shared memory (61 % of the tensor pipe's rate; 127 inside the kernel, next to
// the TMA writes) and 67 cycles with A in TMEM (95 %) -- the operand fetch, not the tensor pipe, paced gmm_tc3.
// Here the expander warps write their rows with tcgen05.st (thread = row = TMEM lane; one 32-bit column holds two
// consecutive halfs of K), hi and lo parts of both 128-frame blocks side by side: 4 x 8 kSteps <= 256 columns.  That leaves
// 256 columns for accumulators: ONE 128-column accumulator per block, and the two blocks of a work item alternate --
// while the epilogue warps take block 0's log-sum-exp the tensor core works on block 1 (15 MMAs ~ 1000 cycles against
// ~1700 for the epilogue of a block: the kernel is now paced by its epilogue, the tensor core has slack).
// Everything else is gmm_tc3's: work items of 512 frames per CTA pair, B tiles by TMA (each CTA its half), taper tile
// skipping with intervals read one tile ahead, flags + gmm_fixup_kernel for frames / states outside the FP16 range, the
// optional expanded-row output for the statistics kernel.  Shared memory now only holds the B ring (4 x 32 KB).
#pragma once
#include "gmm_tc3.cuh"

#define TC4_NST 4
// Launch bound of the kernel (template parameter LB): TC3_THREADS (448) lets the compiler use up to 144 registers per thread
// (it takes 127); 576 makes it budget five warps per register sub-partition = 96 registers (400 bytes of spills): the "lean"
// build for SMALL waves, whose recursion warps should run beside the next wave's K1 ("small waves" in launch_wave, hfbgpu.cu).
// A CTA's warps go to the four sub-partitions of the register file by warp number, so the 14 warps load them 4 / 4 / 3 / 3:
// at 128 registers two partitions are full and nothing else fits the SM; at 96 every partition keeps >= 4 096 registers
// and two of them 7 168 -- room for a ring-window beta warp (165 registers = 5 376) placed AFTER the CTA.
#define TC4_LB_LEAN 576
#define TC4_SMEM_BYTES (TC4_NST * 32768 + 512 + 1024)

__device__ __forceinline__ void tc4_mma_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
   asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                ::"r"(tmemD), "r"(tmemA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc4_tmem_st8(uint32_t taddr, const uint32_t (&w)[8])
{
   asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
}

// The 3 x kSteps MMAs of one (B tile, block): corrections first (hi x lo, lo x hi per K step), then the large hi x hi terms
// (see gmm_tc2_kernel).  KS > 0 = K steps known at compile time: ONE basic block, so that the moves of the operands into
// uniform registers (R2UR, a handful per MMA) are scheduled ahead instead of being waited for between two MMAs -- with a
// run-time trip count every pair of MMAs was its own basic block and the issuing thread, not the tensor pipe, set the
// pace (120 cycles per MMA against 67 in tools/mma_rate.cu).
// a wait that is expected to last long (the expanders wait a whole work item, the producer a B stage): polling warps
// share issue slots -- and, in a kernel that runs into the power cap, energy -- with the epilogue
__device__ __forceinline__ void tc4_mbar_wait_idle(uint64_t *bar, uint32_t parity, unsigned ns)
{
   uint32_t done, addr = tc_smem_u32(bar);
   for (;;) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
      if (done) break;
      __nanosleep(ns);
   }
}
// packed FP32 pairs (Blackwell): one instruction, two lanes of the FMA pipe
__device__ __forceinline__ float2 tc4_fma2(float2 a, float2 b, float2 c)
{
   unsigned long long d;
   asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)),
       "l"(*reinterpret_cast<unsigned long long *>(&c)));
   return *reinterpret_cast<float2 *>(&d);
}
__device__ __forceinline__ float2 tc4_add2(float2 a, float2 b)
{
   unsigned long long d;
   asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
   return *reinterpret_cast<float2 *>(&d);
}

template <int KS>
__device__ __forceinline__ void tc4_issue_block(uint32_t dAcc, uint32_t aHi, uint32_t aLo, uint32_t bLo32, uint32_t idesc, int kSteps,
                                                uint32_t elected)
{
   // shared-memory descriptor of the B stage: low word = (address >> 4) | LBO, high word constant; operand slices differ
   // by a compile-time number of 16-byte units
   const uint32_t hiW = (uint32_t)((1024 >> 4)) | (1u << 14) | (2u << 29);
   auto bdesc = [&](uint32_t byteOff) -> uint64_t { return ((uint64_t)hiW << 32) | (uint64_t)(bLo32 + (byteOff >> 4)); };
#pragma unroll
   for (int ks = 0; ks < (KS ? KS : 8); ks++) {
      if (!KS && ks >= kSteps) break;
      const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
      if (elected) {
         tc4_mma_ts(dAcc, aHi + ks * 8, bdesc(o + 8192), idesc, ks ? 1u : 0u);
         tc4_mma_ts(dAcc, aLo + ks * 8, bdesc(o), idesc, 1u);
      }
   }
#pragma unroll
   for (int ks = 0; ks < (KS ? KS : 8); ks++) {
      if (!KS && ks >= kSteps) break;
      const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
      if (elected) tc4_mma_ts(dAcc, aHi + ks * 8, bdesc(o), idesc, 1u);
   }
}

template <int MP, int DP, int LB = TC3_THREADS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(LB, 1)
gmm_tc4_kernel(const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, Tc3Params p)
{
   extern __shared__ uint8_t tc_smem_raw[];
   uint8_t *base = (uint8_t *)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
   constexpr int NST = TC4_NST;
   constexpr uint32_t ST_BYTES = 32768;                 // this CTA's half of a B tile: 2 chunks x [hi 8 KB | lo 8 KB]
   uint8_t *sB = base;
   uint64_t *bars = (uint64_t *)(sB + NST * ST_BYTES);
   uint64_t *fullA = bars, *emptyA = bars + 1, *fullB = bars + 2, *emptyB = bars + 2 + NST;
   uint64_t *tmemFull = bars + 2 + 2 * NST, *tmemEmpty = tmemFull + 2;   // one pair per 128-frame BLOCK
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
   const int nChunks = (p.kSteps + 3) >> 2;             // 64-half chunks of B that hold data (1 or 2)
   // tensor-memory map: accumulator of block b = columns [128 b, 128 b + 128); A of block b, part (0 = hi, 1 = lo) =
   // columns 256 + (2 b + part) KC .. + KC, KC = 8 columns per K step of 16 halfs
   const uint32_t KC = 8u * (uint32_t)p.kSteps;
   const uint32_t tmemA = tmem + 2 * TC_BN;

   auto tile_need = [&](int f, int l, int y, int T) -> int {          // see gmm_tc3_kernel
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
            tc4_mbar_wait_idle(&emptyB[stage], phB ^ 1, 100);
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
         const uint32_t bBase = tc_smem_u32(sB);
         uint32_t stage = 0, phB = 0, phA = 0, phE = 0;      // phE: one phase bit per block's "accumulator empty" barrier
         int trB = 0;                                        // trace: blocks issued so far (pair 0 only)
         const bool tr = p.trace != nullptr && pair == 0;
         for (int it = pair; it < p.nItems; it += nPairs) {
            const int2 item = p.items[it];
            const UttDesc u = p.utt[item.x];
            const int nTiles = (u.Jt + SPT - 1) / SPT;
            const int2 *tiv = p.tileIv + u.slotOff;
            int2 ivN = tiv[0];
            tc3_wait_acquire_cluster(fullA, phA);               // both CTAs' A blocks are in their tensor memories
            phA ^= 1;
            tc_fence_after();
            for (int n = 0; n < nTiles; n++) {
               const int2 iv = ivN;
               if (n + 1 < nTiles) ivN = tiv[n + 1];
               const int need = tile_need(iv.x, iv.y, item.y, u.T);
               if (!need) continue;
               tc_mbar_wait(&fullB[stage], phB);
               tc_fence_after();
               const uint32_t bSt = bBase + stage * ST_BYTES;
               for (int b = 0; b < 2; b++) {
                  if (!(need & (1 << b))) continue;
                  if (tr && elected && trB < 4096) p.trace[8 * trB] = clock64();
                  tc_mbar_wait(&tmemEmpty[b], ((phE >> b) & 1u) ^ 1u);     // both CTAs' epilogues have read this block's accumulator
                  tc_fence_after();
                  if (tr && elected && trB < 4096) p.trace[8 * trB + 1] = clock64();
                  const uint32_t dAcc = tmem + b * TC_BN, aHi = tmemA + (2 * b) * KC, aLo = aHi + KC;
                  const uint32_t bLo32 = ((bSt >> 4) & 0x3FFFu) | (1u << 16);
                  if (p.kSteps == 5) tc4_issue_block<5>(dAcc, aHi, aLo, bLo32, idesc, 5, elected);        // D = 39
                  else tc4_issue_block<0>(dAcc, aHi, aLo, bLo32, idesc, p.kSteps, elected);
                  if (elected) tc_commit_pair(&tmemFull[b]);     // this block's accumulator: ready for both CTAs' epilogues
                  if (tr && elected && trB < 4096) p.trace[8 * trB + 2] = clock64();
                  trB++;
                  __syncwarp();
                  phE ^= 1u << b;
               }
               if (elected) tc_commit_pair(&emptyB[stage]);
               __syncwarp();
               if (++stage == NST) { stage = 0; phB ^= 1; }
            }
            if (elected) tc_commit_pair(emptyA);                 // A blocks reusable (arrives in both CTAs)
            __syncwarp();
         }
      }
   } else if (warp < 2 + EPW) {
      // ================= epilogue (both CTAs): own frames x 128 components per block =================
      const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
      constexpr int CPW = 2;                            // 32-column chunks per warp (8 warps: two per quadrant)
      const int c0 = ((warp - 2) >> 2) * CPW;
      const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
      uint32_t phF = 0;                                 // one phase bit per block's "accumulator full" barrier
      int trB = 0;
      const bool tr = p.trace != nullptr && pair == 0 && warp == 2 && lane == 0;
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
            for (int blk = 0; blk < 2; blk++) {
               if (!(need & (1 << blk))) continue;
               tc_mbar_wait(&tmemFull[blk], (phF >> blk) & 1u);
               phF ^= 1u << blk;
               tc_fence_after();
               if (tr && trB < 4096) p.trace[8 * trB + 3 + 2 * rank] = clock64();
               const int w0 = item.y + (2 * blk + (int)rank) * TC_BM + quad * 32;      // first frame of this warp
               // something of this warp's 32 frames lies inside the tile's interval (and inside the utterance)
               const bool active = !(w0 >= u.T || f >= w0 + 32 || l < w0 || (p.dbg & 2));
               if (active) {
                  const int t = w0 + lane;
                  float *brow = p.b + u.bOff + (size_t)t * u.J;
                  const uint32_t taddr = tmem + blk * TC_BN + ((uint32_t)(quad * 32) << 16);
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
                  } else {
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
                           float sum;
                           const float mb = mx * LOG2E;
                           if (G % 2 == 0) {                     // packed pairs: half the FFMA / FADD instructions
                              float2 s2 = make_float2(0.f, 0.f);
                              const float2 l2 = make_float2(LOG2E, LOG2E), m2 = make_float2(-mb, -mb);
#pragma unroll
                              for (int i = 0; i < G; i += 2) {
                                 const float2 a = tc4_fma2(make_float2(v[s0 + i], v[s0 + i + 1]), l2, m2);
                                 s2 = tc4_add2(s2, make_float2(tc_ex2(a.x), tc_ex2(a.y)));
                              }
                              sum = s2.x + s2.y;
                           } else {
                              sum = 0.f;
#pragma unroll
                              for (int i = 0; i < G; i++) sum += tc_ex2(fmaf(v[s0 + i], LOG2E, -mb));
                           }
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
               }
               tc_fence_before();
               __syncwarp();
               if (lane == 0) tc_mbar_arrive_leader(&tmemEmpty[blk]);
               if (tr && trB < 4096) p.trace[8 * trB + 4 + 2 * rank] = clock64();
               trB++;
            }
         }
      }
   } else {
      // ================= expanders (both CTAs, 4 warps): raw FP32 features -> this CTA's two A blocks in TMEM =================
      // Thread = row = TMEM lane (a warp reaches the 32 lanes of quadrant warp % 4).  Columns of the operand: 0 = 1 (pairs
      // with the symmetrising constant), 2d+1 = x'_d^2, 2d+2 = x'_d (d < D), 2D+1 = 1 (pairs with the Gaussian constant),
      // the rest 0; x' = (x - offset) * scale.  The rows of the NEXT item travel into registers while the tensor core works
      // on the current one; what is exposed between two items is the conversion and ten tcgen05.st per row.
      const int quad = warp & 3;
      const int r = quad * 32 + lane;                   // 0..127
      const int D = p.D;
      const int nUnits = 2 * p.kSteps;                  // 16-byte units (8 columns) per half of an expanded row
      constexpr int NPRE = (DP <= 40) ? 2 : 1;          // rows held in registers ahead of time
      const uint32_t tLane = tmemA + ((uint32_t)(quad * 32) << 16);
      uint32_t phA = 0;
      float x[NPRE][DP];
      auto load_row = [&](float (&xr)[DP], const int2 item, const UttDesc &u, int b) {
         const int t = item.y + (2 * b + (int)rank) * TC_BM + r;
         const bool inside = t < u.T;
         bool far = false;
         if (p.featPad != nullptr) {
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
      auto store_row = [&](const float (&xr)[DP], int b, uint4 *grow, const bool toTmem) {
#pragma unroll
         for (int ks = 0; ks < 8; ks++) {
            if (ks >= p.kSteps) break;                 // every K step the MMAs read is rewritten for every item
            uint32_t hi[8], lo[8];
#pragma unroll
            for (int e = 0; e < 8; e++) {
               float v2[2];
#pragma unroll
               for (int h = 0; h < 2; h++) {
                  const int k = ks * 16 + 2 * e + h;    // compile-time column
                  if (k == 0) v2[h] = 1.f;
                  else if (k & 1) { const int d = (k - 1) >> 1; v2[h] = (d < DP && d < D) ? xr[d < DP ? d : 0] * xr[d < DP ? d : 0] : ((d == D) ? 1.f : 0.f); }
                  else { const int d = (k - 2) >> 1; v2[h] = (d < DP && d < D) ? xr[d < DP ? d : 0] : 0.f; }
               }
               const __half2 hh = __floats2half2_rn(v2[0], v2[1]);
               const float2 hf = __half22float2(hh);
               const __half2 ll = __floats2half2_rn(v2[0] - hf.x, v2[1] - hf.y);
               hi[e] = *reinterpret_cast<const uint32_t *>(&hh);
               lo[e] = *reinterpret_cast<const uint32_t *>(&ll);
            }
            if (toTmem) {
               tc4_tmem_st8(tLane + (2 * b) * KC + ks * 8, hi);
               tc4_tmem_st8(tLane + (2 * b + 1) * KC + ks * 8, lo);
            }
            if (grow != nullptr) {
               grow[2 * ks] = make_uint4(hi[0], hi[1], hi[2], hi[3]);
               grow[2 * ks + 1] = make_uint4(hi[4], hi[5], hi[6], hi[7]);
               grow[nUnits + 2 * ks] = make_uint4(lo[0], lo[1], lo[2], lo[3]);
               grow[nUnits + 2 * ks + 1] = make_uint4(lo[4], lo[5], lo[6], lo[7]);
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
         tc4_mbar_wait_idle(emptyA, phA ^ 1, 400);        // the previous item's MMAs have read the A blocks
         phA ^= 1;
         tc_fence_after();
         store_row(x[0], 0, NPRE == 1 ? exp_row(item, u, 0) : nullptr, true);
         if (NPRE == 1) load_row(x[0], item, u, 1);
         store_row(x[NPRE - 1], 1, NPRE == 1 ? exp_row(item, u, 1) : nullptr, true);
         asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
         tc_fence_before();
         __syncwarp();
         if (lane == 0) tc3_arrive_leader_release(fullA);
         if (NPRE == 2 && p.expA != nullptr) {          // the copy for the statistics kernel: outside the window the tensor core idles in
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

static inline void gmm_tc4_set_attributes()
{
#define TC4_SET(MPV) cudaFuncSetAttribute(gmm_tc4_kernel<MPV, 40>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES); \
                     cudaFuncSetAttribute(gmm_tc4_kernel<MPV, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES)
   TC4_SET(1); TC4_SET(8); TC4_SET(16); TC4_SET(32); TC4_SET(64); TC4_SET(128);
#undef TC4_SET
   cudaFuncSetAttribute(gmm_tc4_kernel<8, 40, TC4_LB_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES);
   cudaFuncSetAttribute(gmm_tc4_kernel<16, 40, TC4_LB_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES);
   cudaFuncSetAttribute(gmm_tc4_kernel<32, 40, TC4_LB_LEAN>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC4_SMEM_BYTES);
}

// the launch of the kernel alone (gmm_tc3_launch prepares Tc3Params, the flags, the optional outputs and the fix-up)
static inline void gmm_tc4_go(const GmmTc3Model &t, const Tc3Params &p, int D, int grid2, cudaStream_t st, bool lean)
{
   if (lean && D <= 40 && (t.MP == 8 || t.MP == 16 || t.MP == 32)) {      // the 96-register build (mixture sets, D <= 40)
      if (t.MP == 8) gmm_tc4_kernel<8, 40, TC4_LB_LEAN><<<grid2, TC3_THREADS, TC4_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p);
      else if (t.MP == 16) gmm_tc4_kernel<16, 40, TC4_LB_LEAN><<<grid2, TC3_THREADS, TC4_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p);
      else gmm_tc4_kernel<32, 40, TC4_LB_LEAN><<<grid2, TC3_THREADS, TC4_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p);
      return;
   }
#define TC4_GO(MPV) do { if (D <= 40) gmm_tc4_kernel<MPV, 40><<<grid2, TC3_THREADS, TC4_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p); \
                          else gmm_tc4_kernel<MPV, 64><<<grid2, TC3_THREADS, TC4_SMEM_BYTES, st>>>(t.mapBhi, t.mapBlo, p); } while (0)
   switch (t.MP) {
   case 1: TC4_GO(1); break;
   case 8: TC4_GO(8); break;
   case 16: TC4_GO(16); break;
   case 32: TC4_GO(32); break;
   case 64: TC4_GO(64); break;
   default: TC4_GO(128); break;
   }
#undef TC4_GO
}
