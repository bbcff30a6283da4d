// hfb_feat.cuh -- parameter-kind qualifiers on the device: delta / acceleration / third
// differential coefficients and per-utterance cepstral mean normalisation, so that a caller whose files hold
// only the static coefficients (the usual HTK set-up: MFCC_0 on disk, TARGETKIND = MFCC_0_D_A[_Z] in the
// training configuration) uploads a third of the bytes and leaves HParm's AddQualifiers to the GPU.
//
// Restates, for tables (whole utterances: hdValid = tlValid = 0), HTKLib/HParm.c:1618-1722 AddQualifiers ->
// :1552-1599 AddDiffs -> HTKLib/HSigP.c:827-857 Regress and HSigP.c:803-823 FZeroMean:
//   * order o+1 coefficients = regression over order o coefficients, all static columns (cepstra, c0 / energy):
//       sum_{th=1..W} th * (c[min(t+th, T-1)] - c[max(t-th, 0)]) / (2 sum th^2)          (:842-852)
//     or (c[t+W] - c[t-W]) / (2W) with SIMPLEDIFFS (:849-850); the first / last frame is replicated at the
//     ends (:844-845).  Every operation is a separately rounded FP32 operation in the reference's order, so the
//     result is bit-identical to HCopy / HERest's own loader;
//   * _Z: the mean of the leading `zeroMeanCols` static columns (cepstra and c0, not the energy) over the
//     utterance is accumulated in double, rounded to float and subtracted (:809-821), AFTER the
//     differentials were formed (AddQualifiers' order).  The reference adds the T values sequentially; here 32
//     lanes add strided partial sums: the double sum can differ in the last bit, the float mean only when that
//     bit decides a rounding (never seen on the golden vectors).
//   * _N (absolute energy suppressed, HParm.c:2882, :4655-4656): the last static column (energy or c0) feeds the
//     differentials like any other but is not part of the observation; every later column moves one to the left.
// Not covered (hfbgpu_set_qualifiers rejects what it cannot express): _V / global mean / variance files, MatTran
// input transforms, V1COMPAT differences, HIGHDIFF fourth order.
#pragma once
#include "hfb_common.h"

struct FeatQual {
   int numStatic;                   // columns of the source matrix
   int win[3];                      // DELTAWINDOW, ACCWINDOW, THIRDWINDOW; 0 = order absent
   int simpleDiffs;
   int zeroMeanCols;
   int suppressEnergy;              // _N: static column numStatic-1 is not written
   int enabled;
};

// One regression pass: dst[t][dstCol + c] = Regress(src[.][srcCol + c]), c < d.  copyStatic: also
// dst[t][c] = src[t][c] (first pass, source = the caller's static matrix).
__global__ void __launch_bounds__(256) feat_regress_kernel(const UttDesc *__restrict__ utt, const float *__restrict__ src,
                                                           int srcStride, int srcCol, float *__restrict__ dst, int dstStride,
                                                           int dstCol, int d, int win, int simple, int copyStatic)
{
   // copyStatic = how many leading static columns go to dst as they are (0 = none; numStatic - 1 with _N)
   const UttDesc &u = utt[blockIdx.y];
   const int T = u.T;
   const int FR = 256 / 16;                              // frames per block pass (a 16 x 16 tile of (frame, column))
   const float *s0 = src + (size_t)u.featOff * srcStride + srcCol;
   float *d0 = dst + (size_t)u.featOff * dstStride;
   float sigmaT2 = 0.f;
   for (int th = 1; th <= win; th++) sigmaT2 = __fadd_rn(sigmaT2, (float)(th * th));
   sigmaT2 = __fmul_rn(sigmaT2, 2.f);
   for (int t = blockIdx.x * FR + (threadIdx.x >> 4); t < T; t += gridDim.x * FR)
      for (int c = threadIdx.x & 15; c < d; c += 16) {
         float sum = 0.f, fw = 0.f, bk = 0.f;
         for (int th = 1; th <= win; th++) {
            fw = s0[(size_t)min(t + th, T - 1) * srcStride + c];
            bk = s0[(size_t)max(t - th, 0) * srcStride + c];
            if (!simple) sum = __fadd_rn(sum, __fmul_rn((float)th, __fsub_rn(fw, bk)));
         }
         const float v = simple ? __fdiv_rn(__fsub_rn(fw, bk), (float)(2 * win)) : __fdiv_rn(sum, sigmaT2);
         d0[(size_t)t * dstStride + dstCol + c] = v;
         if (c < copyStatic) d0[(size_t)t * dstStride + c] = s0[(size_t)t * srcStride + c];
      }
}

// plain widening copy for a qualifier set without differentials (only _Z)
__global__ void __launch_bounds__(256) feat_copy_kernel(const UttDesc *__restrict__ utt, const float *__restrict__ src, int srcStride,
                                                        float *__restrict__ dst, int dstStride, int d)
{
   const UttDesc &u = utt[blockIdx.y];
   const float *s0 = src + (size_t)u.featOff * srcStride;
   float *d0 = dst + (size_t)u.featOff * dstStride;
   for (long long e = (long long)blockIdx.x * 256 + threadIdx.x; e < (long long)u.T * d; e += (long long)gridDim.x * 256) {
      const int t = (int)(e / d), c = (int)(e - (long long)t * d);
      d0[(size_t)t * dstStride + c] = s0[(size_t)t * srcStride + c];
   }
}

// FZeroMean over the leading zCols columns of one utterance per block, one warp per column
__global__ void __launch_bounds__(256) feat_zeromean_kernel(const UttDesc *__restrict__ utt, float *__restrict__ dst, int dstStride,
                                                            int zCols)
{
   const UttDesc &u = utt[blockIdx.x];
   const int T = u.T, lane = threadIdx.x & 31;
   float *d0 = dst + (size_t)u.featOff * dstStride;
   for (int c = threadIdx.x >> 5; c < zCols; c += 8) {
      double s = 0.0;
      for (int t = lane; t < T; t += 32) s += (double)d0[(size_t)t * dstStride + c];
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = (float)(s / (double)T);
      for (int t = lane; t < T; t += 32) d0[(size_t)t * dstStride + c] = __fsub_rn(d0[(size_t)t * dstStride + c], mean);
   }
}

// Enqueues the passes for nU utterances described by utt[] (device; only featOff and T are read).
// HTK compressed parameter files (`_C`, HASCOMPX): the file holds 16-bit integers and two float vectors A, B per file;
// the reference's loader turns every value back into a float as  v[j] = ((float)s[j] + B[j]) / A[j]
// (HTKLib/HParm.c:3489-3494; A, B read at :3683-3694; written by CalcCompress / the save path, :4892-4960).  Here the
// caller uploads the integers as they are (half the bytes of the float table) and this kernel does the same two
// separately rounded FP32 operations per value: bit-identical to what ReadAsTable yields.  A, B: [utterance][cols].
__global__ void __launch_bounds__(256) feat_decompress_kernel(const UttDesc *__restrict__ utt, const short *__restrict__ src,
                                                              const float *__restrict__ A, const float *__restrict__ B,
                                                              float *__restrict__ dst, int cols)
{
   const UttDesc &u = utt[blockIdx.y];
   const short *s0 = src + (size_t)u.featOff * cols;
   float *d0 = dst + (size_t)u.featOff * cols;
   const float *a = A + (size_t)blockIdx.y * cols, *b = B + (size_t)blockIdx.y * cols;
   const int n = u.T * cols;                             // T <= HFB_MAX_FRAMES: fits 31 bits for any real vector size
   const int stride = (int)gridDim.x * 256, dc = stride % cols;
   int e = (int)blockIdx.x * 256 + (int)threadIdx.x, c = e % cols;
   for (; e < n; e += stride) {                          // the column follows incrementally: no division in the loop
      d0[e] = __fdiv_rn(__fadd_rn((float)s0[e], __ldg(b + c)), __ldg(a + c));
      c += dc; if (c >= cols) c -= cols;
   }
}

static inline void feat_decompress_launch(const UttDesc *utt, int nU, int maxT, const short *src, const float *A, const float *B,
                                          float *dst, int cols, cudaStream_t st)
{
   if (nU <= 0) return;
   const dim3 grid((unsigned)std::max(1, std::min(32, (int)(((long long)maxT * cols + 2047) / 2048))), (unsigned)nU);
   feat_decompress_kernel<<<grid, 256, 0, st>>>(utt, src, A, B, dst, cols);
}

static inline int feat_expand_launch(const FeatQual &q, const UttDesc *utt, int nU, int maxT, const float *src, float *dst, int D,
                                     cudaStream_t st, int *launches)
{
   if (nU <= 0) return 0;
   const int ns = q.numStatic, sh = q.suppressEnergy ? 1 : 0;   // columns after the statics sit `sh` to the left
   const dim3 grid((unsigned)std::max(1, std::min(64, (maxT + 15) / 16)), (unsigned)nU);
   int n = 0;
   if (q.win[0] > 0) {
      feat_regress_kernel<<<grid, 256, 0, st>>>(utt, src, ns, 0, dst, D, ns - sh, ns, q.win[0], q.simpleDiffs, ns - sh);
      n++;
      for (int o = 1; o < 3 && q.win[o] > 0; o++) {
         feat_regress_kernel<<<grid, 256, 0, st>>>(utt, dst, D, o * ns - sh, dst, D, (o + 1) * ns - sh, ns, q.win[o], q.simpleDiffs, 0);
         n++;
      }
   } else {
      feat_copy_kernel<<<grid, 256, 0, st>>>(utt, src, ns, dst, D, ns);
      n++;
   }
   if (q.zeroMeanCols > 0) { feat_zeromean_kernel<<<nU, 256, 0, st>>>(utt, dst, D, q.zeroMeanCols); n++; }
   if (launches) *launches += n;
   return 0;
}
