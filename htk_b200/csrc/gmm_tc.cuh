// gmm_tc.cuh -- K1 on the 5th-generation tensor cores: GMM log-likelihoods as a dense
// contraction  [x'^2, x', 1] (frames)  x  [-ivar/2, mu'*ivar, c] (mixture components)
// with a 3xTF32 split and FP32 accumulation in TMEM, operands staged by TMA, and the
// log-sum-exp over each state's mixtures fused into the TMEM epilogue.
//
// Replaces, for every (frame, distinct tied state) of an utterance, MOutP/IDOutP
// (HTKLib/HModel.c:5484-5499, :5420-5431) and the mixture log-add of ShStrP
// (HTKLib/HFB.c:949-960).
//
// Precision (stated choice): each FP32 operand v is split v = hi + lo with hi, lo exactly
// representable in TF32 (cvt.rna.tf32.f32); the product is evaluated as
// hi*hi + hi*lo + lo*hi -- three tcgen05.mma.kind::tf32 per K step -- and accumulated in
// FP32.  The dropped lo*lo term is ~2^-22 relative.  Features and means are shifted by the
// mean of the Gaussian means (x' = x - o, mu' = mu - o) to keep the x'^2 terms small.
// The tensor core truncates its FP32 accumulator after every MMA; measured on B200 that is a
// +2e-5 common-mode bias on log b ~ -60 when everything goes through one accumulator.  Two
// measures remove it: (1) the large hi*hi terms and the small hi*lo + lo*hi corrections use
// SEPARATE TMEM accumulators, so only 12 roundings touch the large one; (2) a constant column
// +C0 ~ -E[log b]/2 is contracted FIRST, so the running sum crosses zero half way and the
// truncation errors of the two halves cancel; the epilogue subtracts C0 again.
//
// Tiling: one work item = (utterance, 128 consecutive frames).  The item's A operand
// (128 x 96 floats, hi and lo = 96 KB) stays resident in shared memory while the B operand
// of the utterance's states streams through a 3-stage TMA ring in [128 components x 32
// floats] x {hi, lo} blocks (32 KB per stage).  Accumulators are double-buffered in TMEM
// (2 x (128 main + 128 correction) columns = all 512) so the epilogue of tile n overlaps the
// MMAs of tile n+1.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+ TMEM allocation), warps 2-5 epilogue.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <vector>
#include "hfb_common.h"

#define TC_KE 96            // expanded K, 2D+1 padded to 3 swizzle atoms of 32 floats
#define TC_BM 128           // frames per work item (UMMA M)
#define TC_BN 128           // mixture components per tile (UMMA N)
#define TC_STAGES 3
#define TC_A_BYTES (6 * 16384)
#define TC_B_STAGE_BYTES 32768
#define TC_SMEM_BYTES (TC_A_BYTES + TC_STAGES * TC_B_STAGE_BYTES + 256 + 1024)
#define TC_NEG_BIG (-1.0e30f)

struct GmmTcModel {
   bool ready = false;
   int MP = 0;                 // mixture rows per state, padded to a power of two >= 8
   int GPS = 0;                // 8-row groups per state = MP / 8
   long long rows = 0;         // rows in Bhi/Blo (row group 0 is the all-"-inf" dummy)
   float *dBhi = nullptr, *dBlo = nullptr, *dOffset = nullptr;
   CUtensorMap mapBhi, mapBlo;         // boxes of MP rows (one state)
   CUtensorMap mapBhiP, mapBloP;       // boxes of min(MP, 64) rows for the CTA-pair kernel
   bool pairReady = false;
   // FP16 operands of the CTA-pair kernel (3xFP16 split, see gmm_tc2_kernel): rows of TC_KH halfs
   __half *dBhiH = nullptr, *dBloH = nullptr;
   float *dScale = nullptr;            // per-dimension power-of-two feature scale
   CUtensorMap mapBhiH, mapBloH;
   bool f16Ready = false;
   float C0H = 0.f, C1H = 0.f;         // symmetrising constant / common part of the Gaussian constants
   void *encodeFn = nullptr;
   float C0 = 0.f;             // symmetrising constant contracted first, subtracted in the epilogue
};

struct GmmTcWork {             // per-stream expanded feature operand (floats: TF32 split; halfs: FP16 split)
   float *dAhi = nullptr, *dAlo = nullptr;
   unsigned char *dFlag = nullptr;   // [frames] 1 = this frame's row of b is recomputed in FP32 (gmm_fixup_kernel)
   unsigned char *dFlag3 = nullptr;  // the same for gmm_tc3_kernel (which needs no expanded operand)
   size_t flagCap = 0;
   float *dPad = nullptr;            // gmm_tc3_pad_kernel: scaled features with padded rows (single-Gaussian sets)
   size_t padCap = 0;
   uint4 *dExpA = nullptr;           // gmm_tc3_kernel's optional output: every frame's expanded operand row (hfb_stats_tc.cuh)
   size_t expCap = 0;                // in 16-byte units
   size_t aCapFrames = 0;
   bool f16Init = false;       // constant / padding columns of the FP16 layout are in place
   void release()
   {
      if (dAhi) cudaFree(dAhi);
      if (dAlo) cudaFree(dAlo);
      if (dFlag) cudaFree(dFlag);
      if (dFlag3) cudaFree(dFlag3);
      if (dPad) cudaFree(dPad);
      if (dExpA) cudaFree(dExpA);
      dExpA = nullptr; expCap = 0;
      dAhi = dAlo = nullptr; dFlag = dFlag3 = nullptr; dPad = nullptr; aCapFrames = 0; flagCap = 0; padCap = 0; f16Init = false;
   }
};
#define TC_KH 128           // expanded K of the FP16 operands: 2 swizzle atoms of 64 halfs

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tc_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tc_mbar_init(uint64_t *bar, uint32_t count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void tc_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tc_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_mbar_arrive(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t *bar, uint32_t parity)
{
   uint32_t done, addr = tc_smem_u32(bar);
   do {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                   "selp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done) : "r"(addr), "r"(parity) : "memory");
   } while (!done);
}
__device__ __forceinline__ void tc_tma_load_2d(void *smemDst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
   asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(tc_smem_u32(smemDst)), "l"((uint64_t)map), "r"(tc_smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
   asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar)
{
   asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tc_tmem_ld32(uint32_t taddr, float *v)
{
   uint32_t *r = reinterpret_cast<uint32_t *>(v);
   asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                  "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                  "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                  "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr));
   asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128-byte-swizzled operand block of [rows][32 floats]: 8-row groups 1024 B apart.
// Field layout: cute::UMMA::SmemDescriptor (start>>4 | LBO<<16 | SBO<<32 | version=1<<46 | SWIZZLE_128B=2<<61).
__device__ __forceinline__ uint64_t tc_smem_desc(uint32_t smemAddr)
{
   return (uint64_t)((smemAddr >> 4) & 0x3FFF) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
          ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1<<4), a/b format TF32 (2<<7, 2<<10), K-major A and B,
// n_dim = N>>3 at bit 17, m_dim = M>>4 at bit 24.
__device__ __forceinline__ constexpr uint32_t tc_idesc(int M, int N, uint32_t fmt = 2u /* TF32; F16 = 0, BF16 = 1 */)
{
   return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ float tc_ex2(float x)
{
   float r;
   asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}
__device__ __forceinline__ float tc_lg2(float x)
{
   float r;
   asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}
__device__ __forceinline__ float tc_tf32(float x)
{
   uint32_t r;
   asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
   return __uint_as_float(r);
}

// ------------------------------------------------------------------------------------------
// feature expansion: A'hi / A'lo [frames][96] = split of [ 1 | (x-o)^2 | (x-o) | 1 | 0... ]
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_KE * 4)
gmm_tc_expand_kernel(const float *__restrict__ feat, const float *__restrict__ off, int D, long long nFrames,
                     float *__restrict__ Ahi, float *__restrict__ Alo)
{
   // thread (k, r): column k of frame 4 * blockIdx + r; rows are written as whole 384-byte lines
   const int k = threadIdx.x;
   const long long f = (long long)blockIdx.x * 4 + threadIdx.y;
   if (f >= nFrames) return;
   float v = 0.f;
   if (k == 0) v = 1.f;                                   // pairs with the constant column C0
   else if (k <= D) { float x = feat[f * D + k - 1] - off[k - 1]; v = x * x; }
   else if (k <= 2 * D) v = feat[f * D + (k - D - 1)] - off[k - D - 1];
   else if (k == 2 * D + 1) v = 1.f;                      // pairs with c
   const float hi = tc_tf32(v);
   Ahi[f * TC_KE + k] = hi;
   Alo[f * TC_KE + k] = tc_tf32(v - hi);
}

// FP16 variant: A'hi / A'lo [frames][128] halfs = split of [ 1 | ((x-o) s)^2 | (x-o) s | 1 | 0... ], s = per-dimension
// power-of-two scale that brings every dimension to unit-order variance (exact; B carries 1/s, 1/s^2).
// The constant and padding columns never change: gmm_tc_init_f16_kernel writes them once per buffer, the per-wave
// kernel only writes the 2 D data columns (thread = (dimension, frame)).
__global__ void __launch_bounds__(256)
gmm_tc_init_f16_kernel(int D, long long nRows, __half *__restrict__ Ahi, __half *__restrict__ Alo)
{
   const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
   if (idx >= nRows * TC_KH) return;
   const int k = (int)(idx % TC_KH);
   Ahi[idx] = __float2half_rn((k == 0 || k == 2 * D + 1) ? 1.f : 0.f);
   Alo[idx] = __float2half_rn(0.f);
}

// 32 frames per block: coalesced feature reads, the split rows assembled in shared memory, then 16-byte stores of
// the used part of every row ([1 | x'^2 | x' | 1 | 0..] hi and lo; the columns past it stay as gmm_tc_init_f16_kernel
// left them)
#define TC_XF_FR 32
#define TC_FAR 64.f         // scaled |x - o| beyond which a frame goes to the FP32 fix-up
#define TC_DEAD_BELOW (-30000.f)   // FP16 path: a state whose best component is below this goes to the FP32 fix-up
__global__ void __launch_bounds__(256)
gmm_tc_expand_f16_kernel(const float *__restrict__ feat, const float *__restrict__ off, const float *__restrict__ scale,
                         int D, long long nFrames, __half *__restrict__ Ahi, __half *__restrict__ Alo,
                         unsigned char *__restrict__ flag)
{
   __shared__ __align__(16) __half sh[2][TC_XF_FR][TC_KH];
   __shared__ int far[TC_XF_FR];
   if (threadIdx.x < TC_XF_FR) far[threadIdx.x] = 0;
   __syncthreads();
   const long long f0 = (long long)blockIdx.x * TC_XF_FR;
   const int nf = (int)((nFrames - f0 < TC_XF_FR) ? nFrames - f0 : TC_XF_FR);
   const int nk8 = (2 * D + 2 + 7) >> 3, kUsed = nk8 * 8, nConst = kUsed - 2 * D;   // k = 0 and k = 2D+1 .. kUsed-1
   for (int e = threadIdx.x; e < nf * nConst; e += 256) {
      const int fr = e / nConst, c = e - fr * nConst, k = (c == 0) ? 0 : 2 * D + c;
      sh[0][fr][k] = __float2half_rn((k == 0 || k == 2 * D + 1) ? 1.f : 0.f);
      sh[1][fr][k] = __float2half_rn(0.f);
   }
   const float *src = feat + f0 * D;
   for (int e = threadIdx.x; e < nf * D; e += 256) {
      const int fr = e / D, d = e - fr * D;
      const float x = (src[e] - off[d]) * scale[d];
      // A frame with a coordinate beyond TC_FAR typical standard deviations from the global mean (or a non-finite one)
      // leaves the range where the FP16 split is as accurate as the reference's float: its whole row of b is
      // recomputed by gmm_fixup_kernel in FP32, as IDOutP does (HModel.c:5420-5431).  The clamp only keeps inf / NaN
      // out of the tensor core; what it produces for such a frame is overwritten.
      if (!(fabsf(x) <= TC_FAR)) far[fr] = 1;
      const float v1 = fminf(fmaxf(x, -65000.f), 65000.f), v2 = fminf(x * x, 65000.f);
      const __half h1 = __float2half_rn(v1), h2 = __float2half_rn(v2);
      sh[0][fr][1 + d] = h2;     sh[1][fr][1 + d] = __float2half_rn(v2 - __half2float(h2));
      sh[0][fr][1 + D + d] = h1; sh[1][fr][1 + D + d] = __float2half_rn(v1 - __half2float(h1));
   }
   __syncthreads();
   if (threadIdx.x < nf) flag[f0 + threadIdx.x] = (unsigned char)far[threadIdx.x];
   const int per = nf * nk8;
   for (int e = threadIdx.x; e < 2 * per; e += 256) {
      const int arr = e >= per, r = e - arr * per, fr = r / nk8, c = r - fr * nk8;
      __half *dst = (arr ? Alo : Ahi) + (f0 + fr) * TC_KH + c * 8;
      *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(&sh[arr][fr][c * 8]);
   }
}

// ------------------------------------------------------------------------------------------
// the GEMM + log-sum-exp kernel
// ------------------------------------------------------------------------------------------
struct TcParams {
   const int2 *items;          // (utterance in wave, first frame of the 128-frame block)
   int nItems;
   const UttDesc *utt;
   const int *slotState;
   float *b;
   int GPS;                    // 8-row groups per state
   float C0;
   int kSteps;                 // 8-float K steps that hold data: ceil((2D+2)/8) <= 12; the zero padding is skipped
   float deadBelow;            // a column maximum below this means "no live component": output log zero
   int dbg;                    // timing experiments only (HFBGPU_TC_DEBUG): 1 = no A_lo x B_hi, 2 = no epilogue math
   long long *trace;           // HFBGPU_TC_TRACE: clock64() timeline of the first pair (tools/tc_trace.py); else null
   unsigned char *flag;        // FP16 path: per-frame "recompute in FP32" flags (see gmm_fixup_kernel); else null
};
#define TC_TRACE_TILES 512
#define TC_TR(role, tile, slot) do { if (p.trace && pair == 0 && (tile) < TC_TRACE_TILES) \
      p.trace[((size_t)((role) * 2 + rank) * TC_TRACE_TILES + (tile)) * 16 + (slot)] = clock64(); } while (0)

template <int MP>
__global__ void __launch_bounds__(192, 1)
gmm_tc_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
              const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, TcParams p)
{
   extern __shared__ uint8_t tc_smem_raw[];
   uint8_t *base = (uint8_t *)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
   uint8_t *sA = base;                                  // [hi k0,k1,k2 | lo k0,k1,k2] x 16 KB
   uint8_t *sB = base + TC_A_BYTES;                     // stages x [hi 16 KB | lo 16 KB]
   uint64_t *bars = (uint64_t *)(sB + TC_STAGES * TC_B_STAGE_BYTES);
   uint64_t *fullA = bars, *emptyA = bars + 1, *fullB = bars + 2, *emptyB = bars + 2 + TC_STAGES;
   uint64_t *tmemFull = bars + 2 + 2 * TC_STAGES, *tmemEmpty = tmemFull + 2;
   uint32_t *tmemSlot = (uint32_t *)(tmemEmpty + 2);
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

   if (warp == 0 && lane == 0) {
      tc_mbar_init(fullA, 1); tc_mbar_init(emptyA, 1);
      for (int s = 0; s < TC_STAGES; s++) { tc_mbar_init(&fullB[s], 1); tc_mbar_init(&emptyB[s], 1); }
      for (int s = 0; s < 2; s++) { tc_mbar_init(&tmemFull[s], 1); tc_mbar_init(&tmemEmpty[s], 4); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmemSlot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
   }
   tc_fence_before();
   __syncthreads();
   tc_fence_after();
   const uint32_t tmem = *tmemSlot;
   constexpr int SPT = TC_BN / MP;                      // states per tile
   const int nChunks = (p.kSteps + 3) >> 2;             // 32-float K chunks that hold data

   if (warp == 0) {
      // ================= TMA producer: one box of MP rows per lane =================
      uint32_t stage = 0, phB = 0, phA = 0;
      for (int it = blockIdx.x; it < p.nItems; it += gridDim.x) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         const int row0 = (int)u.featOff + item.y;
         if (lane == 0) { tc_mbar_wait(emptyA, phA ^ 1); tc_mbar_expect_tx(fullA, TC_A_BYTES); }
         __syncwarp();
         if (lane < 6) tc_tma_load_2d(sA + lane * 16384, lane < 3 ? &mapAhi : &mapAlo, fullA, (lane % 3) * 32, row0);
         phA ^= 1;
         const int nTiles = (u.J + SPT - 1) / SPT;
         const int *ss = p.slotState + u.slotOff;
         const int half = lane / SPT, bi = lane % SPT;          // lanes [0,SPT): hi boxes, [SPT,2SPT): lo boxes
         for (int n = 0; n < nTiles; n++) {
            const int slot = n * SPT + bi;
            const int row = (lane < 2 * SPT && slot < u.Jt) ? (1 + ss[slot]) * MP : 0;   // rows [0,MP) = dummy state
            for (int k = 0; k < nChunks; k++) {
               if (lane == 0) { tc_mbar_wait(&emptyB[stage], phB ^ 1); tc_mbar_expect_tx(&fullB[stage], (p.dbg & 16) ? 0 : TC_B_STAGE_BYTES); }
               __syncwarp();
               if (lane < 2 * SPT && !(p.dbg & 16))
                  tc_tma_load_2d(sB + stage * TC_B_STAGE_BYTES + half * 16384 + bi * (MP * 128),
                                 half ? &mapBlo : &mapBhi, &fullB[stage], k * 32, row);
               if (++stage == TC_STAGES) { stage = 0; phB ^= 1; }
            }
         }
      }
   } else if (warp == 1) {
      // ================= MMA issuer =================
      if (lane == 0) {
         const uint32_t idesc = tc_idesc(TC_BM, TC_BN), idesc2 = tc_idesc(TC_BM, 2 * TC_BN);
         const uint32_t aBase = tc_smem_u32(sA), bBase = tc_smem_u32(sB);
         uint32_t stage = 0, phB = 0, phA = 0, tile = 0;
         for (int it = blockIdx.x; it < p.nItems; it += gridDim.x) {
            const int2 item = p.items[it];
            const UttDesc u = p.utt[item.x];
            const int nTiles = (u.J + SPT - 1) / SPT;
            tc_mbar_wait(fullA, phA);
            phA ^= 1;
            for (int n = 0; n < nTiles; n++, tile++) {
               const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
               tc_mbar_wait(&tmemEmpty[as], phT ^ 1);
               tc_fence_after();
               const uint32_t dMain = tmem + as * (2 * TC_BN), dCorr = dMain + TC_BN;
               for (int k = 0; k < nChunks; k++) {
                  tc_mbar_wait(&fullB[stage], phB);
                  tc_fence_after();
                  const uint32_t bHi = bBase + stage * TC_B_STAGE_BYTES;
                  const uint32_t aHi = aBase + k * 16384, aLo = aBase + (3 + k) * 16384;
#pragma unroll
                  for (int kk = 0; kk < 4; kk++) {
                     if (k * 4 + kk >= p.kSteps) break;  // columns >= 2D+2 are zero padding
                     const uint64_t dAhi = tc_smem_desc(aHi + kk * 32), dAlo = tc_smem_desc(aLo + kk * 32);
                     const uint64_t dBhi = tc_smem_desc(bHi + kk * 32);
                     // A_hi x [B_hi ; B_lo]: the stage holds hi and lo back to back = one N=256 operand,
                     // columns [0,128) -> main accumulator, [128,256) -> correction accumulator
                     if (!(p.dbg & 8)) tc_mma_tf32(dMain, dAhi, dBhi, idesc2, (k | kk) ? 1u : 0u);
                     if (!(p.dbg & 9)) tc_mma_tf32(dCorr, dAlo, dBhi, idesc, 1u);
                  }
                  tc_commit(&emptyB[stage]);            // stage reusable once these MMAs retire
                  if (++stage == TC_STAGES) { stage = 0; phB ^= 1; }
               }
               tc_commit(&tmemFull[as]);                // accumulator ready for the epilogue
            }
            tc_commit(emptyA);                          // A block reusable
         }
      }
   } else {
      // ================= epilogue: TMEM -> log-sum-exp over mixtures -> b[t][slot] =================
      const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
      const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
      uint32_t tile = 0;
      for (int it = blockIdx.x; it < p.nItems; it += gridDim.x) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         const int nTiles = (u.J + SPT - 1) / SPT;
         const int t = item.y + quad * 32 + lane;
         float *brow = p.b + u.bOff + (size_t)t * u.J;
         for (int n = 0; n < nTiles; n++, tile++) {
            const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
            tc_mbar_wait(&tmemFull[as], phT);
            tc_fence_after();
            const uint32_t taddr = tmem + as * (2 * TC_BN) + ((uint32_t)(quad * 32) << 16);
            const float C0 = p.C0;
            float cmx = -INFINITY, csum = 0.f;          // carry for states wider than one 32-column chunk
#pragma unroll
            for (int c = 0; c < TC_BN / 32; c++) {
               if (p.dbg & 2) break;
               float v[32], vc[32];
               tc_tmem_ld32(taddr + c * 32, v);
               tc_tmem_ld32(taddr + TC_BN + c * 32, vc);
#pragma unroll
               for (int i = 0; i < 32; i++) v[i] += vc[i];
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
                     if (t < u.T && slot < u.J) brow[slot] = val;
                     cmx = -INFINITY; csum = 0.f;
                  }
               }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) tc_mbar_arrive(&tmemEmpty[as]);
         }
      }
   }
   tc_fence_before();
   __syncthreads();
   if (warp == 1) {
      tc_fence_after();
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
   }
}


// ------------------------------------------------------------------------------------------
// CTA-pair version (tcgen05 cta_group::2, UMMA M = 256)
//
// Measured on B200 (tools/tc_experiments.sh): the single-CTA kernel above is bound by the B
// operand stream -- 96 KB from L2 per 128 x 128 tile, 6.6 TB/s chip-wide -- not by the MMAs
// (dropping 17 % or 33 % of them does not change its time) nor by the epilogue.  A CTA pair shares
// one B tile between two 128-frame blocks of the same utterance: each CTA keeps its own A block
// and loads only HALF of every B stage (64 of the 128 mixture components, hi and lo = 16 KB),
// the leader CTA issues M = 256 MMAs that read both halves, and each CTA's TMEM receives the
// accumulators of its own 128 frames.  L2 -> SM bytes and shared-memory fill per SM halve, and the
// smaller stage allows an 8-deep TMA ring.
//
// Three N = 128 MMAs per K step (instead of N = 256 + N = 128): A_hi x B_hi -> main accumulator,
// A_hi x B_lo and A_lo x B_hi -> correction accumulator; in cta_group::2 mode each CTA supplies
// N/2 = 64 rows of B at the descriptor address, so the stage is laid out [hi half | lo half].
// Barrier protocol (CUTLASS PipelineTmaUmmaAsync): TMA loads of both CTAs complete on the LEADER's
// full barriers (which expect the bytes of both), the leader's tcgen05.commit multicasts the
// "stage empty" / "accumulator full" arrivals to both CTAs, and the epilogue warps of both CTAs
// arrive on the leader's "accumulator empty" barrier.
// ------------------------------------------------------------------------------------------
#define TC2_STAGES 6
#define TC2_B_STAGE_BYTES 16384
#define TC2_THREADS 320        // warp 0 TMA producer, warp 1 MMA issuer, warps 2-9 epilogue
#define TC2_SMEM_BYTES (TC_A_BYTES + 8 * TC2_B_STAGE_BYTES + 512 + 1024)   // = 64 KB A + 5 x 32 KB tile stages for 3xFP16
#define TC_PEER_MASK 0xFEFFFFFFu      // clears the CTA-rank bit of a shared::cluster address (cute::Sm100MmaPeerBitMask)

__device__ __forceinline__ uint32_t tc_cluster_ctarank()
{
   uint32_t r;
   asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
   return r;
}
__device__ __forceinline__ void tc_cluster_sync()
{
   asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
   asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the barrier of the pair's leader CTA
__device__ __forceinline__ void tc_tma_load_2d_pair(void *smemDst, const CUtensorMap *map, uint64_t *bar, int c0, int c1)
{
   asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                ::"r"(tc_smem_u32(smemDst)), "l"((uint64_t)map), "r"(tc_smem_u32(bar) & TC_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
template <bool F16>
__device__ __forceinline__ void tc_mma_pair(uint32_t tmemD, uint64_t descA, uint64_t descB, uint32_t idesc, uint32_t accumulate)
{
   if (F16)      // K = 16 halfs per instruction: twice the contraction depth of a TF32 MMA in the same time
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
   else
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                   "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                   ::"r"(tmemD), "l"(descA), "l"(descB), "r"(idesc), "r"(accumulate) : "memory");
}
// arrives (once all prior MMAs of the pair have retired) on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void tc_commit_pair(uint64_t *bar)
{
   asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                ::"r"(tc_smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
// arrive on the leader CTA's copy of the barrier (works from either CTA of the pair)
__device__ __forceinline__ void tc_mbar_arrive_leader(uint64_t *bar)
{
   asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(tc_smem_u32(bar) & TC_PEER_MASK) : "memory");
}

template <int MP, bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(TC2_THREADS, 1)
gmm_tc2_kernel(const __grid_constant__ CUtensorMap mapAhi, const __grid_constant__ CUtensorMap mapAlo,
               const __grid_constant__ CUtensorMap mapBhi, const __grid_constant__ CUtensorMap mapBlo, TcParams p)
{
   extern __shared__ uint8_t tc_smem_raw[];
   uint8_t *base = (uint8_t *)(((uintptr_t)tc_smem_raw + 1023) & ~(uintptr_t)1023);
   constexpr int NCH = F16 ? 2 : 3;                     // 128-byte K chunks of the A block: 2 x 64 halfs or 3 x 32 floats
   constexpr int KCH = F16 ? 64 : 32;                   // elements per chunk
   // 3xFP16: TWO 128-frame blocks per CTA (work item = 512 frames per pair) share every B stage -- half the L2 -> SM
   // bytes per frame again (the B stream, ~6 TB/s chip-wide, is what bounds the kernel); with one accumulator per
   // block, 2 blocks x 2 buffers x 128 columns = all 512 TMEM columns.
   constexpr int NBLK = F16 ? 2 : 1;
   constexpr uint32_t A_BLK = 2 * NCH * 16384;          // one block: [hi chunks | lo chunks]
   constexpr uint32_t A_BYTES = NBLK * A_BLK;
   uint8_t *sA = base;                                  // [hi chunks | lo chunks] x 16 KB: this CTA's 128 frames
   // B ring.  3xTF32: one stage = one 128-byte K chunk [hi 8 KB | lo 8 KB] of this CTA's 64 components, separate main /
   // correction accumulators.  3xFP16: one stage = the WHOLE tile (both K chunks), because the MMAs of a tile are issued
   // corrections first, main products last, into ONE accumulator (see the MMA issuer) -- the epilogue then reads half
   // as much TMEM, which at ~64 B/clk was the bound of the two-accumulator version (131 KB per tile = 2048 cycles).
   constexpr int NST = F16 ? 3 : TC2_STAGES;
   constexpr uint32_t ST_BYTES = F16 ? NCH * 16384 : TC2_B_STAGE_BYTES;
   uint8_t *sB = base + A_BYTES;
   uint64_t *bars = (uint64_t *)(sB + NST * ST_BYTES);
   uint64_t *fullA = bars, *emptyA = bars + 1, *fullB = bars + 2, *emptyB = bars + 2 + NST;
   uint64_t *tmemFull = bars + 2 + 2 * NST, *tmemEmpty = tmemFull + 2;
   uint32_t *tmemSlot = (uint32_t *)(tmemEmpty + 2);
   const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const uint32_t rank = tc_cluster_ctarank();          // 0 = leader (issues the MMAs)
   const int pair = blockIdx.x >> 1, nPairs = gridDim.x >> 1;

   if (warp == 0 && lane == 0) {
      tc_mbar_init(fullA, 1); tc_mbar_init(emptyA, 1);
      for (int s = 0; s < NST; s++) { tc_mbar_init(&fullB[s], 1); tc_mbar_init(&emptyB[s], 1); }
      for (int s = 0; s < 2; s++) { tc_mbar_init(&tmemFull[s], 1); tc_mbar_init(&tmemEmpty[s], 2 * ((MP <= 64) ? 8 : 4)); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   }
   if (warp == 1) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(tmemSlot)), "r"(512) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
   }
   tc_fence_before();
   tc_cluster_sync();                                   // barriers initialised and TMEM allocated in both CTAs
   tc_fence_after();
   const uint32_t tmem = *tmemSlot;
   constexpr int SPT = TC_BN / MP;                      // states per tile
   // epilogue warps: 8 when a state's columns fit one 64-column half (two warps per TMEM lane quadrant, each
   // taking half of the tile's columns: the epilogue, not the MMAs, bounds the FP16 kernel otherwise), else 4
   constexpr int EPW = (MP <= 64) ? 8 : 4;
   constexpr int HB = (SPT >= 2) ? SPT / 2 : 1;         // B boxes per operand half held by one CTA
   constexpr int BOXR = (MP < 64) ? MP : 64;            // rows per box (the B tensor maps are built with this)
   const int nChunks = (p.kSteps + 3) >> 2;

   if (warp == 0) {
      // ================= TMA producer (both CTAs): own A block, own half of every B stage =================
      // The warp stays converged: every lane knows the row of "its" box, the rows travel by shuffle and ONE
      // elected lane issues all copies of a stage back to back with warp-uniform operands.  (Per-lane issue
      // from a divergent warp costs ~100 cycles per cp.async.bulk.tensor -- the compiler serialises the
      // lanes through ELECT + R2UR.BROADCAST -- and made the producer, not the MMAs, the bottleneck.)
      uint32_t elected;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
      uint32_t stage = 0, phB = 0, phA = 0, ptile = 0;
      for (int it = pair; it < p.nItems; it += nPairs) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         // block b of CTA r holds frames item.y + (2 b + r) 128 ..; a block that starts beyond T is neither loaded nor used
         const int nBlk = (NBLK == 2 && u.T - item.y > 2 * TC_BM) ? 2 : 1;
         tc_mbar_wait(emptyA, phA ^ 1);
         if (elected) {
            if (rank == 0) tc_mbar_expect_tx(fullA, 2 * nBlk * A_BLK);
            for (int b = 0; b < nBlk; b++) {
               const int row0 = (int)u.featOff + item.y + (2 * b + (int)rank) * TC_BM;
#pragma unroll
               for (int j = 0; j < 2 * NCH; j++)
                  tc_tma_load_2d_pair(sA + b * A_BLK + j * 16384, j < NCH ? &mapAhi : &mapAlo, fullA, (j % NCH) * KCH, row0);
            }
         }
         __syncwarp();
         phA ^= 1;
         const int nTiles = (u.J + SPT - 1) / SPT;
         const int *ss = p.slotState + u.slotOff;
         const int bi = lane % HB;                              // lanes [0,HB): hi boxes, [HB,2HB): lo boxes
         auto box_row = [&](int n) {
            int row = 0;                                        // rows [0,MP) = dummy state
            if (lane < 2 * HB && n < nTiles) {
               if (SPT >= 2) { const int slot = n * SPT + (int)rank * HB + bi; if (slot < u.Jt) row = (1 + ss[slot]) * MP; }
               else row = (1 + ss[n]) * MP + (int)rank * 64;    // MP = 128: each CTA takes 64 rows of the state
            }
            return row;
         };
         int rowNext = box_row(0);
         for (int n = 0; n < nTiles; n++, ptile++) {
            const int row = rowNext;
            rowNext = box_row(n + 1);                           // the state lookup of the next tile is in flight meanwhile
            if (F16) {
               // one stage per tile: all K chunks of this CTA's half of the B tile
               if (elected) TC_TR(2, ptile, 0);
               tc_mbar_wait(&emptyB[stage], phB ^ 1);
               if (elected) TC_TR(2, ptile, 1);
               if (elected && rank == 0) tc_mbar_expect_tx(&fullB[stage], 2 * nChunks * 16384);
               for (int k = 0; k < nChunks; k++) {
                  uint8_t *dst = sB + stage * ST_BYTES + k * 16384;
#pragma unroll
                  for (int j = 0; j < 2 * HB; j++) {
                     const int rj = __shfl_sync(0xffffffffu, row, j);
                     if (elected)
                        tc_tma_load_2d_pair(dst + (j / HB) * 8192 + (j % HB) * (BOXR * 128), (j / HB) ? &mapBlo : &mapBhi,
                                            &fullB[stage], k * KCH, rj);
                  }
               }
               __syncwarp();
               if (++stage == NST) { stage = 0; phB ^= 1; }
            } else
            for (int k = 0; k < nChunks; k++) {
               if (elected) TC_TR(2, ptile, k * 2);
               tc_mbar_wait(&emptyB[stage], phB ^ 1);
               if (elected) TC_TR(2, ptile, k * 2 + 1);
               if (elected && rank == 0) tc_mbar_expect_tx(&fullB[stage], 2 * TC2_B_STAGE_BYTES);
               uint8_t *dst = sB + stage * TC2_B_STAGE_BYTES;
#pragma unroll
               for (int j = 0; j < 2 * HB; j++) {
                  const int rj = __shfl_sync(0xffffffffu, row, j);
                  if (elected)
                     tc_tma_load_2d_pair(dst + (j / HB) * 8192 + (j % HB) * (BOXR * 128), (j / HB) ? &mapBlo : &mapBhi,
                                         &fullB[stage], k * KCH, rj);
               }
               __syncwarp();
               if (++stage == NST) { stage = 0; phB ^= 1; }
            }
         }
      }
   } else if (warp == 1) {
      // ================= MMA issuer: one thread of the leader CTA =================
      // The whole warp runs the loop so that the descriptor arithmetic stays warp-uniform (uniform
      // datapath, no per-MMA R2UR chains: measured 130 -> ~103 cycles per MMA); one elected lane issues.
      if (rank == 0) {
         uint32_t elected;
         asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(elected));
         const uint32_t idesc = tc_idesc(2 * TC_BM, TC_BN, F16 ? 0u : 2u);
         const uint32_t aBase = tc_smem_u32(sA), bBase = tc_smem_u32(sB);
         uint32_t stage = 0, phB = 0, phA = 0, tile = 0;
         for (int it = pair; it < p.nItems; it += nPairs) {
            const int2 item = p.items[it];
            const UttDesc u = p.utt[item.x];
            const int nTiles = (u.J + SPT - 1) / SPT;
            const int nBlk = (NBLK == 2 && u.T - item.y > 2 * TC_BM) ? 2 : 1;
            tc_mbar_wait(fullA, phA);
            phA ^= 1;
            for (int n = 0; n < nTiles; n++, tile++) {
               const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
               if (elected) TC_TR(0, tile, 0);
               tc_mbar_wait(&tmemEmpty[as], phT ^ 1);
               if (elected) TC_TR(0, tile, 1);
               tc_fence_after();
               const uint32_t dMain = tmem + as * (2 * TC_BN), dCorr = dMain + TC_BN;
               if (F16) {
                  // 3xFP16, one accumulator: the small correction products (A_hi x B_lo, A_lo x B_hi) of ALL K steps
                  // first -- the accumulator is still ~1e-2, so the tensor core's truncating accumulation costs
                  // nothing -- then the five large A_hi x B_hi products, starting with the symmetrising constant column
                  if (elected) TC_TR(0, tile, 2);
                  tc_mbar_wait(&fullB[stage], phB);
                  if (elected) TC_TR(0, tile, 3);
                  tc_fence_after();
                  const uint32_t bSt = bBase + stage * ST_BYTES;
                  for (int b = 0; b < nBlk; b++) {
                     const uint32_t aB = aBase + b * A_BLK, dAcc = dMain + b * TC_BN;
#pragma unroll
                     for (int ks = 0; ks < 4 * NCH; ks++) {
                        if (ks >= p.kSteps) break;
                        const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
                        const uint64_t dAhi = tc_smem_desc(aB + o), dAlo = tc_smem_desc(aB + NCH * 16384 + o);
                        const uint64_t dBhi = tc_smem_desc(bSt + o), dBlo = tc_smem_desc(bSt + o + 8192);
                        if (elected) {
                           tc_mma_pair<F16>(dAcc, dAhi, dBlo, idesc, ks ? 1u : 0u);
                           tc_mma_pair<F16>(dAcc, dAlo, dBhi, idesc, 1u);
                        }
                     }
#pragma unroll
                     for (int ks = 0; ks < 4 * NCH; ks++) {
                        if (ks >= p.kSteps) break;
                        const uint32_t o = (ks >> 2) * 16384 + (ks & 3) * 32;
                        const uint64_t dAhi = tc_smem_desc(aB + o), dBhi = tc_smem_desc(bSt + o);
                        if (elected) tc_mma_pair<F16>(dAcc, dAhi, dBhi, idesc, 1u);
                     }
                  }
                  if (elected) tc_commit_pair(&emptyB[stage]);
                  __syncwarp();
                  if (++stage == NST) { stage = 0; phB ^= 1; }
               } else
               for (int k = 0; k < nChunks; k++) {
                  if (elected) TC_TR(0, tile, 2 + 2 * k);
                  tc_mbar_wait(&fullB[stage], phB);
                  if (elected) TC_TR(0, tile, 3 + 2 * k);
                  tc_fence_after();
                  const uint32_t bHi = bBase + stage * TC2_B_STAGE_BYTES, bLo = bHi + 8192;
                  const uint32_t aHi = aBase + k * 16384, aLo = aBase + (NCH + k) * 16384;
#pragma unroll
                  for (int kk = 0; kk < 4; kk++) {
                     if (k * 4 + kk >= p.kSteps) break;  // columns >= 2D+2 are zero padding
                     const uint64_t dAhi = tc_smem_desc(aHi + kk * 32), dAlo = tc_smem_desc(aLo + kk * 32);
                     const uint64_t dBhi = tc_smem_desc(bHi + kk * 32), dBlo = tc_smem_desc(bLo + kk * 32);
                     if (elected) {
                        tc_mma_pair<F16>(dMain, dAhi, dBhi, idesc, (k | kk) ? 1u : 0u);
                        tc_mma_pair<F16>(dCorr, dAhi, dBlo, idesc, (k | kk) ? 1u : 0u);
                        tc_mma_pair<F16>(dCorr, dAlo, dBhi, idesc, 1u);
                     }
                  }
                  if (elected) tc_commit_pair(&emptyB[stage]);   // stage reusable (in both CTAs) once these MMAs retire
                  __syncwarp();
                  if (++stage == NST) { stage = 0; phB ^= 1; }
               }
               if (elected) tc_commit_pair(&tmemFull[as]);       // accumulators ready for both CTAs' epilogues
               __syncwarp();
               if (elected) TC_TR(0, tile, 8);
            }
            if (elected) tc_commit_pair(emptyA);                 // A blocks reusable
            __syncwarp();
         }
      }
   } else if (warp - 2 < EPW) {
      // ================= epilogue (both CTAs): own 128 frames x 128 components =================
      const int quad = warp & 3;                        // TMEM lane quadrant this warp may read
      constexpr int CPW = (TC_BN / 32) * 4 / EPW;       // 32-column chunks per warp: 4, or 2 with eight warps
      const int c0 = ((warp - 2) >> 2) * CPW;           // first chunk of this warp
      const float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
      uint32_t tile = 0;
      for (int it = pair; it < p.nItems; it += nPairs) {
         const int2 item = p.items[it];
         const UttDesc u = p.utt[item.x];
         const int nTiles = (u.J + SPT - 1) / SPT;
         const int nBlk = (NBLK == 2 && u.T - item.y > 2 * TC_BM) ? 2 : 1;
         for (int n = 0; n < nTiles; n++, tile++) {
            const uint32_t as = tile & 1, phT = (tile >> 1) & 1;
            if (warp == 2 && lane == 0) TC_TR(1, tile, 0);
            tc_mbar_wait(&tmemFull[as], phT);
            if (warp == 2 && lane == 0) TC_TR(1, tile, 1);
            tc_fence_after();
            for (int blk = 0; blk < nBlk; blk++) {
            const int t = item.y + (2 * blk + (int)rank) * TC_BM + quad * 32 + lane;
            float *brow = p.b + u.bOff + (size_t)t * u.J;
            const uint32_t taddr = tmem + as * (2 * TC_BN) + blk * TC_BN + ((uint32_t)(quad * 32) << 16);
            const float C0 = p.C0;
            float cmx = -INFINITY, csum = 0.f;          // carry for states wider than one 32-column chunk
            // states of at most 32 components: the warp's columns are NOUT consecutive slots of this frame's row, stored
            // with 16-byte (8-byte) stores -- one store per lane and state cost 32 sector writes per instruction
            constexpr int NOUT = (MP <= 32) ? CPW * 32 / MP : 0;
            float outv[NOUT > 0 ? NOUT : 1];
            int no = 0;
#pragma unroll
            for (int cc = 0; cc < CPW; cc++) {
               if (p.dbg & 2) break;
               const int c = c0 + cc;
               float v[32];
               tc_tmem_ld32(taddr + c * 32, v);
               if (!F16) {                              // 3xTF32: main + correction accumulator
                  float vc[32];
                  tc_tmem_ld32(taddr + TC_BN + c * 32, vc);
#pragma unroll
                  for (int i = 0; i < 32; i++) v[i] += vc[i];
               }
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
                     // FP16 operands: "no live component" and "every live component far away" look alike here (dead rows
                     // carry -60000); the FP32 fix-up decides, so that a far-away state gets its true value like the
                     // reference's (HModel.c:5420-5431) instead of log zero
                     if (p.flag && mx < p.deadBelow && t < u.T && slot < u.Jt) p.flag[u.frameBase + t] = 1;
                     if (NOUT > 0) outv[no++] = val;
                     else if (t < u.T && slot < u.J) brow[slot] = val;
                     cmx = -INFINITY; csum = 0.f;
                  }
               }
            }
            if (NOUT > 0 && !(p.dbg & 2) && t < u.T) {
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
            if (warp == 2 && lane == 0) TC_TR(1, tile, 2);
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


// ------------------------------------------------------------------------------------------
// FP32 fix-up of the FP16 tensor-core path.  Rows of b flagged by gmm_tc_expand_f16_kernel (a coordinate further
// than TC_FAR typical standard deviations from the global mean, or not finite) or by the epilogue (a state whose
// best component came out below TC_DEAD_BELOW) are recomputed the way the reference does: IDOutP per component
// (HModel.c:5420-5431), log-add over the components with weight > LMINMIX (ShStrP, HFB.c:949-960), log zero
// only when no component is live.  One CTA per (utterance, 128 frames); it leaves at once when no frame of
// its block is flagged -- on clean data the kernel costs one byte read per frame.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gmm_fixup_kernel(DevModel M, Wave W, const int2 *__restrict__ items, const unsigned char *__restrict__ flag)
{
   __shared__ float xs[64];
   __shared__ int list[128];
   __shared__ int nList;
   const int2 item = items[blockIdx.x];
   const UttDesc u = W.utt[item.x];
   const int t = item.y + (int)threadIdx.x;
   const int mine = (t < u.T) ? (int)flag[u.frameBase + t] : 0;
   if (threadIdx.x == 0) nList = 0;
   if (!__syncthreads_or(mine)) return;
   if (mine) list[atomicAdd(&nList, 1)] = t;
   __syncthreads();
   const int D = M.D, Dp = M.Dp, n = nList;
   const int *ss = W.slotState + u.slotOff;
   for (int i = 0; i < n; i++) {
      const int tt = list[i];
      if ((int)threadIdx.x < D) xs[threadIdx.x] = W.feat[((size_t)u.featOff + tt) * D + threadIdx.x];
      __syncthreads();
      float *brow = W.b + u.bOff + (size_t)tt * u.J;
      for (int slot = threadIdx.x; slot < u.Jt; slot += 128) {
         const int st = ss[slot];
         const int mo = M.stateMixOff[st], Mn = M.stateMixOff[st + 1] - mo;
         float mx = -INFINITY, sm = 0.f;
         bool any = false;
         for (int m = 0; m < Mn; m++) {
            const float wt = M.mixLogWt[mo + m];
            if (Mn > 1 && !(wt > LMINMIX_F)) continue;
            any = true;
            const int g = M.mixGauss[mo + m];
            const float *mu = M.mean + (size_t)g * Dp, *iv = M.ivar + (size_t)g * Dp;
            float a = M.gconst[g];
            for (int k = 0; k < D; k++) { const float d = xs[k] - __ldg(mu + k); a = fmaf(d * d, __ldg(iv + k), a); }
            float v = -0.5f * a;
            if (Mn == 1) { mx = v; sm = 1.f; break; }
            v += wt;
            if (v > mx) { sm = sm * expf(mx - v) + 1.f; mx = v; } else sm += expf(v - mx);
         }
         brow[slot] = any ? mx + logf(sm) : (float)HFB_LZERO;
      }
      __syncthreads();
   }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*TcEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                               const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline int tc_make_map(void *fn, CUtensorMap *map, float *basePtr, long long rows, int boxRows)
{
   cuuint64_t dims[2] = {(cuuint64_t)TC_KE, (cuuint64_t)rows};
   cuuint64_t strides[1] = {(cuuint64_t)TC_KE * sizeof(float)};
   cuuint32_t box[2] = {32, (cuuint32_t)boxRows};
   cuuint32_t estr[2] = {1, 1};
   CUresult r = ((TcEncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, basePtr, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   return r == CUDA_SUCCESS ? HFB_OK : HFB_ECUDA;
}

static inline float tc_host_tf32(float x)
{
   uint32_t u;
   memcpy(&u, &x, 4);
   u += 0x1000u; u &= 0xffffe000u;
   float r;
   memcpy(&r, &u, 4);
   return r;
}

static inline int tc_make_map_f16(void *fn, CUtensorMap *map, __half *basePtr, long long rows, int boxRows)
{
   cuuint64_t dims[2] = {(cuuint64_t)TC_KH, (cuuint64_t)rows};
   cuuint64_t strides[1] = {(cuuint64_t)TC_KH * sizeof(__half)};
   cuuint32_t box[2] = {64, (cuuint32_t)boxRows};
   cuuint32_t estr[2] = {1, 1};
   CUresult r = ((TcEncodeFn)fn)(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, basePtr, dims, strides, box, estr,
                                CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
   return r == CUDA_SUCCESS ? HFB_OK : HFB_ECUDA;
}

static inline void gmm_tc_release(GmmTcModel &t)
{
   if (t.dBhi) cudaFree(t.dBhi);
   if (t.dBlo) cudaFree(t.dBlo);
   if (t.dOffset) cudaFree(t.dOffset);
   if (t.dBhiH) cudaFree(t.dBhiH);
   if (t.dBloH) cudaFree(t.dBloH);
   if (t.dScale) cudaFree(t.dScale);
   t = GmmTcModel();
}

static inline bool gmm_tc_available(const GmmTcModel &t) { return t.ready; }

// Builds the expanded, split B operand: one row per mixture component of each tied state, each
// state padded to MP rows, row group 0 = dummy.  Row = [-ivar/2 | (mu-o)*ivar | c | 0...] with
// c = -(gConst + sum (mu-o)^2 ivar)/2 + log weight (weight omitted for single-mixture states,
// HFB.c:917-928; components with weight <= LMINMIX get c = -1e30, HFB.c:953).
static inline int gmm_tc_prepare(GmmTcModel &t, const hfb_model *m, cudaStream_t st)
{
   t = GmmTcModel();
   const int D = m->vecSize, J = m->numStates;
   int maxM = 0;
   for (int s = 0; s < J; s++) maxM = std::max(maxM, m->stateMixOff[s + 1] - m->stateMixOff[s]);
   if (maxM < 2 || 2 * D + 2 > TC_KE) return HFB_OK;          // FP32 kernel handles these
   int MP = 8;
   while (MP < maxM) MP *= 2;
   if (MP > TC_BN) return HFB_OK;
   int dev = 0, major = 0;
   cudaGetDevice(&dev);
   cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
   if (major != 10) return HFB_OK;                              // tcgen05 is sm_100 family only
   void *fn = nullptr;
   cudaDriverEntryPointQueryResult qr;
   if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn) {
      cudaGetLastError();
      return HFB_OK;
   }
   t.encodeFn = fn;
   t.MP = MP; t.GPS = MP / 8;
   t.rows = (long long)(1 + (long long)J) * MP;                 // rows [0, MP) = dummy state
   std::vector<double> off(D, 0.0);
   for (int g = 0; g < m->numGauss; g++)
      for (int k = 0; k < D; k++) off[k] += m->mean[(size_t)g * D + k];
   std::vector<float> offF(D);
   for (int k = 0; k < D; k++) offF[k] = (float)(off[k] / m->numGauss);
   std::vector<float> hi((size_t)t.rows * TC_KE, 0.f), lo((size_t)t.rows * TC_KE, 0.f);
   auto put = [&](long long r, int k, double v) {
      float f = (float)v, h = tc_host_tf32(f);
      hi[(size_t)r * TC_KE + k] = h;
      lo[(size_t)r * TC_KE + k] = tc_host_tf32(f - h);
   };
   for (long long r = 0; r < t.rows; r++) put(r, 2 * D + 1, TC_NEG_BIG);
   double cSum = 0.0; long long cCnt = 0;
   for (int s = 0; s < J; s++) {
      int mo = m->stateMixOff[s], Mn = m->stateMixOff[s + 1] - mo;
      for (int k2 = 0; k2 < Mn; k2++) {
         long long r = (long long)(1 + (long long)s) * MP + k2;
         float wt = m->mixLogWt[mo + k2];
         if (Mn > 1 && !(wt > (float)HFB_LMINMIX)) continue;
         int g = m->mixGauss[mo + k2];
         double c = m->gConst[g];
         for (int k = 0; k < D; k++) {
            double iv = m->ivar[(size_t)g * D + k], mu = (double)m->mean[(size_t)g * D + k] - (double)offF[k];
            put(r, 1 + k, -0.5 * iv);
            put(r, 1 + D + k, mu * iv);
            c += mu * mu * iv;
         }
         const double cc = -0.5 * c + (Mn > 1 ? (double)wt : 0.0);
         put(r, 2 * D + 1, cc);
         cSum += cc; cCnt++;
      }
   }
   // expected log b ~ mean(c) - D/2 (chi-square term of matched data): start the running sum at -half of it
   {
      double eb = (cCnt ? cSum / cCnt : 0.0) - 0.5 * D;
      t.C0 = tc_host_tf32((float)(-0.5 * eb));
      for (long long r = 0; r < t.rows; r++) { hi[(size_t)r * TC_KE] = t.C0; lo[(size_t)r * TC_KE] = 0.f; }
   }
   size_t bytes = hi.size() * sizeof(float);
   if (cudaMalloc(&t.dBhi, bytes) != cudaSuccess || cudaMalloc(&t.dBlo, bytes) != cudaSuccess ||
       cudaMalloc(&t.dOffset, D * sizeof(float)) != cudaSuccess) {
      cudaGetLastError(); gmm_tc_release(t); return HFB_ENOMEM;
   }
   cudaMemcpyAsync(t.dBhi, hi.data(), bytes, cudaMemcpyHostToDevice, st);
   cudaMemcpyAsync(t.dBlo, lo.data(), bytes, cudaMemcpyHostToDevice, st);
   cudaMemcpyAsync(t.dOffset, offF.data(), D * sizeof(float), cudaMemcpyHostToDevice, st);
   cudaStreamSynchronize(st);
   if (tc_make_map(fn, &t.mapBhi, t.dBhi, t.rows, MP) || tc_make_map(fn, &t.mapBlo, t.dBlo, t.rows, MP)) {
      gmm_tc_release(t); return HFB_OK;
   }
#define TC_SET_SMEM(MPV) cudaFuncSetAttribute(gmm_tc_kernel<MPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES)
   TC_SET_SMEM(8); TC_SET_SMEM(16); TC_SET_SMEM(32); TC_SET_SMEM(64); TC_SET_SMEM(128);
#undef TC_SET_SMEM
   t.ready = (cudaGetLastError() == cudaSuccess);
   if (t.ready && tc_make_map(fn, &t.mapBhiP, t.dBhi, t.rows, std::min(MP, 64)) == HFB_OK &&
       tc_make_map(fn, &t.mapBloP, t.dBlo, t.rows, std::min(MP, 64)) == HFB_OK) {
#define TC_SET_SMEM(MPV) cudaFuncSetAttribute(gmm_tc2_kernel<MPV, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES); \
                         cudaFuncSetAttribute(gmm_tc2_kernel<MPV, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC2_SMEM_BYTES)
      TC_SET_SMEM(8); TC_SET_SMEM(16); TC_SET_SMEM(32); TC_SET_SMEM(64); TC_SET_SMEM(128);
#undef TC_SET_SMEM
      t.pairReady = (cudaGetLastError() == cudaSuccess);
   }
   if (t.pairReady && 2 * D + 2 <= TC_KH) {
      // ---- FP16 operands (3xFP16 split): per-dimension power-of-two scale s_k ~ sqrt(mean inverse variance), so
      //      that (x-o) s, -ivar / (2 s^2) and (mu-o) ivar / s are all of order one; the common part C1 of the
      //      Gaussian constants is kept out of the product (added back exactly in the epilogue)
      std::vector<float> sc(D);
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
         int mo = m->stateMixOff[s2], Mn = m->stateMixOff[s2 + 1] - mo;
         for (int k2 = 0; k2 < Mn; k2++) {
            long long r = (long long)(1 + (long long)s2) * MP + k2;
            float wt = m->mixLogWt[mo + k2];
            if (Mn > 1 && !(wt > (float)HFB_LMINMIX)) continue;
            int g = m->mixGauss[mo + k2];
            double c = m->gConst[g];
            for (int k = 0; k < D; k++) {
               double iv = m->ivar[(size_t)g * D + k], mu = (double)m->mean[(size_t)g * D + k] - (double)offF[k];
               c += mu * mu * iv;
            }
            cRow[r] = -0.5 * c + (Mn > 1 ? (double)wt : 0.0);
            live[r] = 1; c1 += cRow[r]; nLive++;
         }
      }
      c1 = nLive ? c1 / nLive : 0.0;
      t.C1H = (float)c1;
      std::vector<__half> hh((size_t)t.rows * TC_KH, __float2half_rn(0.f)), hl((size_t)t.rows * TC_KH, __float2half_rn(0.f));
      bool inRange = true;             // a model whose scaled parameters leave the FP16 range keeps the 3xTF32 kernel
      auto putH = [&](long long r, int k, double v) {
         if (!(fabs(v) < 30000.0) && k != 2 * D + 1) inRange = false;
         if (k == 2 * D + 1 && live[r] && !(fabs(v) < 20000.0)) inRange = false;
         float f = (float)v;
         __half h = __float2half_rn(f);
         hh[(size_t)r * TC_KH + k] = h;
         hl[(size_t)r * TC_KH + k] = __float2half_rn(f - __half2float(h));
      };
      for (long long r = 0; r < t.rows; r++) if (!live[r]) putH(r, 2 * D + 1, -60000.0);     // dead rows: log zero
      for (int s2 = 0; s2 < J; s2++) {
         int mo = m->stateMixOff[s2], Mn = m->stateMixOff[s2 + 1] - mo;
         for (int k2 = 0; k2 < Mn; k2++) {
            long long r = (long long)(1 + (long long)s2) * MP + k2;
            if (!live[r]) continue;
            int g = m->mixGauss[mo + k2];
            for (int k = 0; k < D; k++) {
               double iv = m->ivar[(size_t)g * D + k], mu = (double)m->mean[(size_t)g * D + k] - (double)offF[k], sk = sc[k];
               putH(r, 1 + k, -0.5 * iv / (sk * sk));
               putH(r, 1 + D + k, mu * iv / sk);
            }
            putH(r, 2 * D + 1, cRow[r] - c1);
         }
      }
      {
         double eb = -0.5 * D;                                   // expected value of the product part of log b
         t.C0H = __half2float(__float2half_rn((float)(-0.5 * eb)));
         for (long long r = 0; r < t.rows; r++) { hh[(size_t)r * TC_KH] = __float2half_rn(t.C0H); hl[(size_t)r * TC_KH] = __float2half_rn(0.f); }
      }
      size_t hb = hh.size() * sizeof(__half);
      if (!inRange) { /* stay on the 3xTF32 operands */ }
      else if (cudaMalloc(&t.dBhiH, hb) == cudaSuccess && cudaMalloc(&t.dBloH, hb) == cudaSuccess &&
          cudaMalloc(&t.dScale, D * sizeof(float)) == cudaSuccess) {
         cudaMemcpyAsync(t.dBhiH, hh.data(), hb, cudaMemcpyHostToDevice, st);
         cudaMemcpyAsync(t.dBloH, hl.data(), hb, cudaMemcpyHostToDevice, st);
         cudaMemcpyAsync(t.dScale, sc.data(), D * sizeof(float), cudaMemcpyHostToDevice, st);
         cudaStreamSynchronize(st);
         t.f16Ready = tc_make_map_f16(fn, &t.mapBhiH, t.dBhiH, t.rows, std::min(MP, 64)) == HFB_OK &&
                      tc_make_map_f16(fn, &t.mapBloH, t.dBloH, t.rows, std::min(MP, 64)) == HFB_OK;
      } else cudaGetLastError();
   }
   return HFB_OK;
}

// Launches expansion + GEMM for every utterance of the wave.  `items` lives in the wave blob.
static inline int gmm_tc_launch(GmmTcModel &t, GmmTcWork &wk, const DevModel &dm, const Wave &W, long long waveFrames,
                                const int2 *dItems, int nItems, const int2 *dItems2, int nItems2,
                                const int2 *dItems4, int nItems4, int smCount,
                                cudaStream_t st, int *launches, cudaEvent_t afterExpand = nullptr)
{
   if (!t.ready) return HFB_EUNSUPPORTED;
   if (nItems == 0) return HFB_OK;
   const size_t need = (size_t)waveFrames + TC_BM;
   if (need > wk.aCapFrames) {
      wk.release();
      size_t cap = need + need / 8;
      if (cudaMalloc(&wk.dAhi, cap * TC_KE * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&wk.dAlo, cap * TC_KE * sizeof(float)) != cudaSuccess ||
          cudaMalloc(&wk.dFlag, cap) != cudaSuccess) { cudaGetLastError(); wk.release(); return HFB_ENOMEM; }
      wk.aCapFrames = cap;
   }
   TcParams p;
   p.items = dItems; p.nItems = nItems; p.utt = W.utt; p.slotState = W.slotState; p.b = W.b; p.GPS = t.GPS; p.C0 = t.C0;
   p.kSteps = (2 * dm.D + 2 + 7) / 8;
   p.deadBelow = -1.0e29f;
   { const char *e = getenv("HFBGPU_TC_DEBUG"); p.dbg = e ? atoi(e) : 0; if (p.dbg & 4) p.kSteps = 12; }
   p.trace = nullptr; p.flag = nullptr;
   const char *traceFile = getenv("HFBGPU_TC_TRACE");
   const size_t traceN = (size_t)6 * TC_TRACE_TILES * 16;
   if (traceFile) { cudaMalloc(&p.trace, traceN * sizeof(long long)); cudaMemsetAsync(p.trace, 0, traceN * sizeof(long long), st); }
   if (launches) *launches = 2;
   const bool pair = t.pairReady && smCount >= 2 && !getenv("HFBGPU_NO_PAIR");
   const bool f16 = pair && t.f16Ready && !getenv("HFBGPU_TC_TF32");
   CUtensorMap mapAhi, mapAlo;
   // rows beyond the wave (at most 2 TC_BM - 1) are allocated but stale or out of bounds (zero fill): every output
   // row depends on its own A row only and rows with t >= T are never stored
   if (f16) {
      // ---- 3xFP16 split: same buffers, rows of TC_KH halfs (256 B) instead of TC_KE floats (384 B)
      __half *ahi = (__half *)wk.dAhi, *alo = (__half *)wk.dAlo;
      if (tc_make_map_f16(t.encodeFn, &mapAhi, ahi, waveFrames + TC_BM, TC_BM) ||
          tc_make_map_f16(t.encodeFn, &mapAlo, alo, waveFrames + TC_BM, TC_BM))
         return HFB_ECUDA;
      if (!wk.f16Init) {
         const long long n = (long long)wk.aCapFrames * TC_KH;
         gmm_tc_init_f16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(dm.D, (long long)wk.aCapFrames, ahi, alo);
         wk.f16Init = true;
      }
      gmm_tc_expand_f16_kernel<<<(unsigned)((waveFrames + TC_XF_FR - 1) / TC_XF_FR), 256, 0, st>>>(W.feat, t.dOffset, t.dScale, dm.D, waveFrames, ahi, alo, wk.dFlag);
      p.flag = wk.dFlag;
      if (afterExpand) cudaEventRecord(afterExpand, st);
      p.items = dItems4; p.nItems = nItems4;            // work items of 4 x 128 frames: two blocks per CTA
      p.C0 = t.C0H - t.C1H;                              // the epilogue subtracts C0 and adds the common constant C1 back
      p.kSteps = (2 * dm.D + 2 + 15) / 16;               // 16 halfs per MMA
      p.deadBelow = TC_DEAD_BELOW;
      const int grid2 = 2 * std::min(nItems4, smCount / 2);
      switch (t.MP) {
      case 8: gmm_tc2_kernel<8, true><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiH, t.mapBloH, p); break;
      case 16: gmm_tc2_kernel<16, true><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiH, t.mapBloH, p); break;
      case 32: gmm_tc2_kernel<32, true><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiH, t.mapBloH, p); break;
      case 64: gmm_tc2_kernel<64, true><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiH, t.mapBloH, p); break;
      default: gmm_tc2_kernel<128, true><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiH, t.mapBloH, p); break;
      }
      if (!getenv("HFBGPU_NO_FIXUP")) {
         gmm_fixup_kernel<<<nItems, 128, 0, st>>>(dm, W, dItems, wk.dFlag);
         if (launches) *launches = 3;
      }
      if (traceFile) {                                  // diagnostics only: synchronous dump of the timeline
         std::vector<long long> h(traceN);
         cudaStreamSynchronize(st);
         cudaMemcpy(h.data(), p.trace, traceN * sizeof(long long), cudaMemcpyDeviceToHost);
         cudaFree(p.trace);
         if (FILE *f = fopen(traceFile, "wb")) { fwrite(h.data(), sizeof(long long), traceN, f); fclose(f); }
      }
      return HFB_OK;
   }
   if (tc_make_map(t.encodeFn, &mapAhi, wk.dAhi, waveFrames + TC_BM, TC_BM) ||
       tc_make_map(t.encodeFn, &mapAlo, wk.dAlo, waveFrames + TC_BM, TC_BM))
      return HFB_ECUDA;
   wk.f16Init = false;                                  // the float layout overwrites the FP16 constant columns
   gmm_tc_expand_kernel<<<(unsigned)((waveFrames + 3) / 4), dim3(TC_KE, 4), 0, st>>>(W.feat, t.dOffset, dm.D, waveFrames, wk.dAhi, wk.dAlo);
   if (afterExpand) cudaEventRecord(afterExpand, st);
   if (pair) {
      // CTA pairs: one work item = (utterance, 256 frames), 128 per CTA
      p.items = dItems2; p.nItems = nItems2;
      const int grid2 = 2 * std::min(nItems2, smCount / 2);
      switch (t.MP) {
      case 8: gmm_tc2_kernel<8, false><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiP, t.mapBloP, p); break;
      case 16: gmm_tc2_kernel<16, false><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiP, t.mapBloP, p); break;
      case 32: gmm_tc2_kernel<32, false><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiP, t.mapBloP, p); break;
      case 64: gmm_tc2_kernel<64, false><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiP, t.mapBloP, p); break;
      default: gmm_tc2_kernel<128, false><<<grid2, TC2_THREADS, TC2_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhiP, t.mapBloP, p); break;
      }
      if (traceFile) {                                  // diagnostics only: synchronous dump of the timeline
         std::vector<long long> h(traceN);
         cudaStreamSynchronize(st);
         cudaMemcpy(h.data(), p.trace, traceN * sizeof(long long), cudaMemcpyDeviceToHost);
         cudaFree(p.trace);
         if (FILE *f = fopen(traceFile, "wb")) { fwrite(h.data(), sizeof(long long), traceN, f); fclose(f); }
      }
      return HFB_OK;
   }
   int grid = std::min(nItems, smCount);
   switch (t.MP) {
   case 8: gmm_tc_kernel<8><<<grid, 192, TC_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhi, t.mapBlo, p); break;
   case 16: gmm_tc_kernel<16><<<grid, 192, TC_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhi, t.mapBlo, p); break;
   case 32: gmm_tc_kernel<32><<<grid, 192, TC_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhi, t.mapBlo, p); break;
   case 64: gmm_tc_kernel<64><<<grid, 192, TC_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhi, t.mapBlo, p); break;
   default: gmm_tc_kernel<128><<<grid, 192, TC_SMEM_BYTES, st>>>(mapAhi, mapAlo, t.mapBhi, t.mapBlo, p); break;
   }
   return HFB_OK;
}
