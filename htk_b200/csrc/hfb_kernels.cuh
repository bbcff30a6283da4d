// hfb_kernels.cuh -- the sm_100a kernels of the E-step (first correct path).
//
//   gmm_fp32_kernel   log b_j(o_t) for every (frame, distinct tied state) of an utterance:
//                     diagonal Gaussians (IDOutP, HTKLib/HModel.c:5420-5431) + log-sum-exp
//                     over mixtures (ShStrP, HTKLib/HFB.c:898-988).  FP32 on CUDA cores; the
//                     tcgen05 3xTF32 contraction in gmm_tc.cuh replaces it on the hot path.
//   beta_kernel       SetBeamTaper + SetBeta + the StepBack retry loop
//                     (HFB.c:1116-1145, :1149-1296, :1321-1366).
//   alpha_kernel      InitAlpha / StepAlpha with the alpha beam, SetOcct, UpTranParms and the
//                     per-state part of UpMixParms (HFB.c:616-784, :399-418, :1371-1423,
//                     :1480-1489).
//   stats_kernel      the per-mixture part of UpMixParms (HFB.c:1549-1736): minimum-occupancy
//                     rule and the centred mean / variance / weight sums, FP64 accumulators.
//
// One CTA per utterance for the two recursions, one thread per model of the transcription:
// the time-sliding window (two beta or alpha columns) lives in shared memory, the T-step
// serial chain costs two block barriers per frame, and many utterances share an SM.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include "hfb_common.h"

#define LZERO_D   (-1.0e10)
#define LSMALL_D  (-0.5e10)
#define MINEARG_D (-708.3)
#define LMINMIX_F (-11.5129254649702f)
#define MINLOGEXP (-23.025850929940457)

// LAdd, HTKLib/HMath.c:1576-1590.  EXACT=false keeps the magnitude in FP64 and evaluates the
// bounded correction log(1+exp(d)), d in [-23.03, 0], in FP32 (<= 1e-7 absolute).
template <bool EXACT>
__device__ __forceinline__ double ladd(double x, double y)
{
   if (x < y) { double t = x; x = y; y = t; }
   double d = y - x;
   if (d < MINLOGEXP) return (x < LSMALL_D) ? LZERO_D : x;
   if (EXACT) return x + log(1.0 + exp(d));
   return x + (double)log1pf(expf((float)d));
}

__device__ __forceinline__ double warp_max(double v)
{
   for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
   return v;
}
__device__ __forceinline__ int warp_maxi(int v)
{
   for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
   return v;
}
__device__ __forceinline__ int warp_mini(int v)
{
   for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
   return v;
}
__device__ __forceinline__ float warp_sumf(float v)
{
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}

// ------------------------------------------------------------------------------------------
// K1 (FP32 path): [64 frames x 32 slots] tile per CTA
// ------------------------------------------------------------------------------------------
#define GT_FR 64
#define GT_SL 32

__global__ void __launch_bounds__(128)
gmm_fp32_kernel(DevModel M, Wave W, const GmmTile *__restrict__ tiles)
{
   extern __shared__ float gsm[];
   float *xs = gsm;                              // [D][GT_FR]
   float *outT = gsm + M.D * GT_FR;              // [GT_FR][GT_SL+1]
   const GmmTile tl = tiles[blockIdx.x];
   const UttDesc u = W.utt[tl.utt];
   const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
   const int D = M.D, Dp = M.Dp;
   const float *feat = W.feat + (size_t)u.featOff * D;

   for (int idx = tid; idx < D * GT_FR; idx += 128) {
      int f = idx / D, k = idx - f * D, t = tl.t0 + f;
      xs[k * GT_FR + f] = (t < u.T) ? feat[(size_t)t * D + k] : 0.f;
   }
   __syncthreads();

   for (int sl = wid; sl < GT_SL; sl += 4) {
      int slot = tl.s0 + sl;
      if (slot >= u.J) break;
      int s = W.slotState[u.slotOff + slot];
      int mo = M.stateMixOff[s], Mn = M.stateMixOff[s + 1] - mo;
      float mx0 = -INFINITY, mx1 = -INFINITY, sm0 = 0.f, sm1 = 0.f;
      bool any = false;
      for (int m = 0; m < Mn; m++) {
         float wt = M.mixLogWt[mo + m];
         if (Mn > 1 && !(wt > LMINMIX_F)) continue;
         any = true;
         int g = M.mixGauss[mo + m];
         const float *mu = M.mean + (size_t)g * Dp, *iv = M.ivar + (size_t)g * Dp;
         float a0 = M.gconst[g], a1 = a0;
         for (int k = 0; k < D; k++) {
            float mk = __ldg(mu + k), ik = __ldg(iv + k);
            float d0 = xs[k * GT_FR + lane] - mk, d1 = xs[k * GT_FR + lane + 32] - mk;
            a0 = fmaf(d0 * d0, ik, a0);
            a1 = fmaf(d1 * d1, ik, a1);
         }
         float v0 = -0.5f * a0, v1 = -0.5f * a1;
         if (Mn == 1) { mx0 = v0; mx1 = v1; sm0 = sm1 = 1.f; break; }
         v0 += wt; v1 += wt;
         if (v0 > mx0) { sm0 = sm0 * expf(mx0 - v0) + 1.f; mx0 = v0; } else sm0 += expf(v0 - mx0);
         if (v1 > mx1) { sm1 = sm1 * expf(mx1 - v1) + 1.f; mx1 = v1; } else sm1 += expf(v1 - mx1);
      }
      outT[lane * (GT_SL + 1) + sl] = any ? mx0 + logf(sm0) : (float)LZERO_D;
      outT[(lane + 32) * (GT_SL + 1) + sl] = any ? mx1 + logf(sm1) : (float)LZERO_D;
   }
   __syncthreads();

   float *b = W.b + u.bOff;
   for (int idx = tid; idx < GT_FR * GT_SL; idx += 128) {
      int f = idx >> 5, sl = idx & 31, t = tl.t0 + f, slot = tl.s0 + sl;
      if (t < u.T && slot < u.J) b[(size_t)t * u.J + slot] = outT[f * (GT_SL + 1) + sl];
   }
}

// ------------------------------------------------------------------------------------------
// shared-memory carve-up of the recursion kernels
// ------------------------------------------------------------------------------------------
struct RecSmem {
   double *colA, *colB, *aux;       // two [S] columns; aux[Q] (maxP in beta, mpSelf in alpha)
   double *aux2;                    // [Q] (ex in alpha)
   double *wred;                    // [32]
   int *wlo, *whi;                  // [32] each
   int *sN, *sSoff, *sTr, *sPoff, *sDms;   // [Q] each
};

__host__ __device__ inline size_t rec_smem_bytes(int S, int Q)
{
   return sizeof(double) * ((size_t)2 * S + 2 * Q + 32) + sizeof(int) * ((size_t)64 + 5 * Q);
}

__device__ __forceinline__ RecSmem rec_carve(unsigned char *raw, int S, int Q)
{
   RecSmem r;
   r.colA = (double *)raw; r.colB = r.colA + S; r.aux = r.colB + S; r.aux2 = r.aux + Q;
   r.wred = r.aux2 + Q;
   r.wlo = (int *)(r.wred + 32); r.whi = r.wlo + 32;
   r.sN = r.whi + 32; r.sSoff = r.sN + Q; r.sTr = r.sSoff + Q; r.sPoff = r.sTr + Q; r.sDms = r.sPoff + Q;
   return r;
}

// ------------------------------------------------------------------------------------------
// K2: beta pass with beam pruning and the whole-utterance retry loop
// ------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(256) beta_kernel(DevModel M, Wave W)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) {                             // rejected by the host (CreateInsts checks)
      if (threadIdx.x == 0 && out->status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
      return;
   }
   const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
   const int T = u.T, Q = u.Q, S = u.S, J = u.J;
   RecSmem sm = rec_carve(smraw, S, Q);
   for (int q = tid; q < Q; q += nt) {
      sm.sN[q] = W.mN[u.modOff + q]; sm.sSoff[q] = W.mSoff[u.modOff + q];
      sm.sTr[q] = W.mTrans[u.modOff + q]; sm.sPoff[q] = W.mPoff[u.modOff + q];
      sm.sDms[q] = W.mDms[u.modOff + q];
   }
   const float *A0 = M.transLogA;
   const int *posSlot = W.posSlot + u.posOff;
   const float *bU = W.b + u.bOff;
   double *betaU = W.beta + u.betaOff;
   short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   const int *pre = W.mPre + u.modOff, *suf = W.mSuf + u.modOff;

   double thresh = W.pruneInit, pr = LZERO_D;
   int retries = 0, status = 0;
   __syncthreads();

   for (;;) {
      // ---- SetBeamTaper, HFB.c:1116-1145, closed form: qHi[t] = max{q : sum_{q'<q} dms <= t},
      //      qLo[t] = min{q : sum_{q'>q} dms <= T-1-t}
      for (int t = tid; t < T; t += nt) {
         int lo = 0, hi = Q;                            // first q with pre[q] > t
         while (lo < hi) { int mid = (lo + hi) >> 1; if (pre[mid] <= t) lo = mid + 1; else hi = mid; }
         qHi[t] = (short)(lo - 1);
         int r = T - 1 - t; lo = 0; hi = Q;             // first q with suf[q] <= r (suf non-increasing)
         while (lo < hi) { int mid = (lo + hi) >> 1; if (suf[mid] > r) lo = mid + 1; else hi = mid; }
         qLo[t] = (short)lo;
      }
      __syncthreads();

      double *cur = sm.colA, *prev = sm.colB;
      // ---- t = T-1 (last column), HFB.c:1176-1198
      int lo1 = qLo[T - 1], hi1 = Q - 1, lastq = lo1;
      if (tid == 0) {
         double bn = 0.0, a1N = 0.0; int lN = 0;
         for (int q = Q - 1; q >= lo1; q--) {
            int N = sm.sN[q];
            bn = (q == Q - 1) ? 0.0 : cur[sm.sSoff[q + 1] + lN - 1] + a1N;
            cur[sm.sSoff[q] + N - 1] = bn;
            lN = N; a1N = A0[sm.sTr[q] + N - 1];
         }
         qHi[T - 1] = (short)(Q - 1);
      }
      __syncthreads();
      for (int q = tid; q < Q; q += nt) {
         if (q < lo1) continue;
         const int N = sm.sN[q], so = sm.sSoff[q];
         const float *A = A0 + sm.sTr[q];
         const float *bt = bU + (size_t)(T - 1) * J;
         const int *ps = posSlot + sm.sPoff[q];
         double bn = cur[so + N - 1];
         for (int i = 1; i < N - 1; i++) cur[so + i] = (double)A[i * N + N - 1] + bn;
         double x = LZERO_D;
         for (int j = 1; j < N - 1; j++) {
            double a = A[j], y = cur[so + j];
            if (a > LSMALL_D && y > LSMALL_D) x = ladd<EXACT>(x, a + (double)bt[ps[j - 1]] + y);
         }
         cur[so] = x;
         double *bg = betaU + (size_t)(T - 1) * S + so;
         for (int i = 0; i < N; i++) bg[i] = cur[so + i];
      }
      __syncthreads();
      { double *tmp = cur; cur = prev; prev = tmp; }

      // ---- t = T-2 .. 0, HFB.c:1205-1277
      bool fail = false;
      for (int t = T - 2; t >= 0; t--) {
         const int tapLo = qLo[t], tapHi = qHi[t];
         int startq = hi1;
         int endq = (lo1 == 0) ? 0 : ((tapLo >= lo1) ? tapLo : lo1 - 1);
         while (endq > 0 && sm.sDms[endq - 1] == 0) endq--;
         lastq = endq;
         const float *bt = bU + (size_t)t * J, *bt1 = bt + J;
         double myMax = LZERO_D;
         for (int q = tid; q < Q; q += nt) {
            if (q < endq || q > startq) continue;
            const int N = sm.sN[q], so = sm.sSoff[q];
            const float *A = A0 + sm.sTr[q];
            const int *ps = posSlot + sm.sPoff[q];
            const bool in1 = (q >= lo1 && q <= hi1);
            // exit state (:1225-1227)
            double ex = LZERO_D;
            if (q < Q - 1) {
               if (q + 1 >= lo1 && q + 1 <= hi1) ex = prev[sm.sSoff[q + 1]];
               if (q < startq) {
                  const int N1 = sm.sN[q + 1];
                  const double a1N = A0[sm.sTr[q + 1] + N1 - 1];
                  if (a1N > LSMALL_D) {                 // q+1 is a tee model: its exit value this frame
                     double y = (q + 2 < Q && q + 2 >= lo1 && q + 2 <= hi1) ? prev[sm.sSoff[q + 2]] : LZERO_D;
                     ex = ladd<EXACT>(ex, y + a1N);
                  }
               }
            }
            cur[so + N - 1] = ex;
            double lMax = LZERO_D;
            for (int i = N - 2; i >= 1; i--) {
               double x = (double)A[i * N + N - 1] + ex;
               if (in1)
                  for (int j = 1; j < N - 1; j++) {
                     double a = A[i * N + j], y = prev[so + j];
                     if (a > LSMALL_D && y > LSMALL_D) x = ladd<EXACT>(x, a + (double)bt1[ps[j - 1]] + y);
                  }
               cur[so + i] = x;
               lMax = fmax(lMax, x);
            }
            double x = LZERO_D;
            for (int j = 1; j < N - 1; j++) {
               double a = A[j], y = cur[so + j];
               if (a > LSMALL_D && y > LSMALL_D) x = ladd<EXACT>(x, a + (double)bt[ps[j - 1]] + y);
            }
            cur[so] = x;
            sm.aux[q] = lMax;
            myMax = fmax(myMax, lMax);
            double *bg = betaU + (size_t)t * S + so;
            for (int i = 0; i < N; i++) bg[i] = cur[so + i];
         }
         myMax = warp_max(myMax);
         if (lane == 0) sm.wred[wid] = myMax;
         __syncthreads();
         double gMax = LZERO_D;
         for (int w = 0; w < nw; w++) gMax = fmax(gMax, sm.wred[w]);
         // ---- pruning (:1254-1272)
         int myHi = -1, myLo = 0x7fffffff;
         for (int q = tid; q < Q; q += nt) {
            if (q < endq || q > startq) continue;
            if (!(gMax - sm.aux[q] > thresh)) { myHi = max(myHi, q); myLo = min(myLo, q); }
         }
         myHi = warp_maxi(myHi); myLo = warp_mini(myLo);
         if (lane == 0) { sm.whi[wid] = myHi; sm.wlo[wid] = myLo; }
         __syncthreads();
         int nhi = -1, nlo = 0x7fffffff;
         for (int w = 0; w < nw; w++) { nhi = max(nhi, sm.whi[w]); nlo = min(nlo, sm.wlo[w]); }
         if (nhi < 0) { fail = true; status = HFB_UTT_EBETA; break; }   // HError 7323
         if (nhi > tapHi) nhi = tapHi;                                   // "on taper" (:1259-1263)
         if (nlo > nhi) { fail = true; break; }                          // beam empty -> LZERO (:1268-1270)
         if (tid == 0) { qHi[t] = (short)nhi; qLo[t] = (short)nlo; }
         hi1 = nhi; lo1 = nlo;
         { double *tmp = cur; cur = prev; prev = tmp; }
      }
      if (status != 0) break;
      if (!fail) {
         pr = prev[sm.sSoff[lastq]];                   // utt->pr = bqt[1] (:1280)
         if (pr > LSMALL_D) break;
      }
      // ---- StepBack retry (:1349-1361)
      thresh += W.pruneInc;
      if (thresh > W.pruneLim || W.pruneInc == 0.0) { status = HFB_UTT_SKIPPED; break; }
      retries++;
      __syncthreads();
   }
   if (tid == 0) {
      out->status = status; out->retries = retries; out->pr = (status == 0) ? pr : LZERO_D;
      out->thresh = thresh;
      if (status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
   }
}

// ------------------------------------------------------------------------------------------
// K3: alpha pass, alpha beam, occupancies, transition counts
// ------------------------------------------------------------------------------------------
template <bool EXACT>
__global__ void __launch_bounds__(256) alpha_kernel(DevModel M, Wave W)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) return;
   const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
   const int T = u.T, Q = u.Q, S = u.S, J = u.J, P = u.P;
   RecSmem sm = rec_carve(smraw, S, Q);
   for (int q = tid; q < Q; q += nt) {
      sm.sN[q] = W.mN[u.modOff + q]; sm.sSoff[q] = W.mSoff[u.modOff + q];
      sm.sTr[q] = W.mTrans[u.modOff + q]; sm.sPoff[q] = W.mPoff[u.modOff + q];
      sm.sDms[q] = W.mDms[u.modOff + q];
      sm.aux[q] = LZERO_D; sm.aux2[q] = LZERO_D;
      W.mTmin[u.modOff + q] = 0x7fffffff; W.mTmax[u.modOff + q] = -1;
      atomicAdd(&W.acc[M.L.numEgs + W.mHmm[u.modOff + q]], 1.0);       // HFB.c:1768-1772
   }
   const float *A0 = M.transLogA;
   const int *posSlot = W.posSlot + u.posOff, *posState = W.posState + u.posOff;
   const float *bU = W.b + u.bOff;
   const double *betaU = W.beta + u.betaOff;
   double *occU = W.occ + u.occOff;
   const short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const double pr = out->pr, minF = W.minFrwdP;
   const int uf = W.uFlags;
   const bool doMix = (uf & (HFB_UPMEANS | HFB_UPVARS | HFB_UPMIXES)) != 0;
   const bool doTr = (uf & HFB_UPTRANS) != 0;
   double *cur = sm.colA, *prev = sm.colB;
   double *mpSelf = sm.aux, *exq = sm.aux2;
   __syncthreads();

   int sq = 0, eq = 0;
   for (int t = 0; t < T; t++) {
      const int loT = qLo[t], hiT = qHi[t];
      if (t == 0) {
         // ---- InitAlpha, HFB.c:616-651 (entry chain through leading tee models is serial)
         eq = hiT; sq = 0;
         if (tid == 0) {
            double a1 = 0.0, a1N = 0.0;
            for (int q = 0; q <= eq; q++) {
               a1 = (q == 0) ? 0.0 : a1 + a1N;
               cur[sm.sSoff[q]] = a1;
               a1N = A0[sm.sTr[q] + sm.sN[q] - 1];
            }
         }
         __syncthreads();
         for (int q = tid; q < Q; q += nt) {
            const int N = sm.sN[q], so = sm.sSoff[q];
            if (q > eq) { for (int i = 0; i < N; i++) cur[so + i] = LZERO_D; continue; }
            const float *A = A0 + sm.sTr[q];
            const int *ps = posSlot + sm.sPoff[q];
            const double a1 = cur[so];
            for (int j = 1; j < N - 1; j++) {
               double a = A[j];
               cur[so + j] = (a > LSMALL_D) ? a1 + a + (double)bU[ps[j - 1]] : LZERO_D;
            }
            double x = LZERO_D;
            for (int i = 1; i < N - 1; i++) {
               double a = A[i * N + N - 1];
               if (a > LSMALL_D) x = ladd<EXACT>(x, cur[so + i] + a);
            }
            cur[so + N - 1] = x;
         }
      } else {
         // ---- alpha beam, HFB.c:701-722, from mpSelf/exq of frame t-1
         const int loP = qLo[t - 1], hiP = qHi[t - 1];
         int mySq = 0x7fffffff;
         for (int q = tid; q < Q; q += nt) {
            if (q < loP) continue;
            double mp = fmax((q > 0) ? exq[q - 1] : LZERO_D, mpSelf[q]);
            if (!(pr - mp > minF)) mySq = min(mySq, q);
         }
         mySq = warp_mini(mySq);
         if (lane == 0) sm.wlo[wid] = mySq;
         __syncthreads();
         int nsq = 0x7fffffff;
         for (int w = 0; w < nw; w++) nsq = min(nsq, sm.wlo[w]);
         if (nsq > hiT) { if (tid == 0) out->status = HFB_UTT_EALPHA; return; }   // HError 7390
         if (nsq < loT) nsq = loT;
         const int eq0 = (hiP < Q - 1) ? hiP + 1 : hiP;
         int myEq = -1;
         for (int q = tid; q < Q; q += nt) {
            if (q > eq0) continue;
            double mp = (q > 0) ? exq[q - 1] : LZERO_D;
            if (q > 0 && q - 1 > nsq) {                                   // chain over a preceding tee model
               const int N1 = sm.sN[q - 1];
               if ((double)A0[sm.sTr[q - 1] + N1 - 1] > LSMALL_D && q >= 2) mp = fmax(mp, exq[q - 2]);
            }
            mp = fmax(mp, mpSelf[q]);
            if (!(pr - mp > minF)) myEq = max(myEq, q);
         }
         myEq = warp_maxi(myEq);
         if (lane == 0) sm.whi[wid] = myEq;
         __syncthreads();
         int neq = -1;
         for (int w = 0; w < nw; w++) neq = max(neq, sm.whi[w]);
         if (neq < nsq) { if (tid == 0) out->status = HFB_UTT_EALPHA; return; }
         while (neq < Q - 1 && sm.sDms[neq] == 0) neq++;
         if (neq > hiT) neq = hiT;
         sq = nsq; eq = neq;
         // ---- alpha column, HFB.c:729-771
         for (int q = tid; q < Q; q += nt) {
            const int N = sm.sN[q], so = sm.sSoff[q];
            if (q < sq || q > eq) { for (int i = 0; i < N; i++) cur[so + i] = LZERO_D; continue; }
            const float *A = A0 + sm.sTr[q];
            const int *ps = posSlot + sm.sPoff[q];
            const float *bt = bU + (size_t)t * J;
            double a1 = LZERO_D;
            if (q > 0) {
               const int N1 = sm.sN[q - 1];
               a1 = prev[sm.sSoff[q - 1] + N1 - 1];
               const double a1N = A0[sm.sTr[q - 1] + N1 - 1];
               if (q > sq && a1N > LSMALL_D) {                            // through a tee model this frame
                  double y = (q >= 2) ? prev[sm.sSoff[q - 2] + sm.sN[q - 2] - 1] : LZERO_D;
                  a1 = ladd<EXACT>(a1, y + a1N);
               }
            }
            cur[so] = a1;
            for (int j = 1; j < N - 1; j++) {
               double a = A[j];
               double x = (a > LSMALL_D) ? a + a1 : LZERO_D;
               for (int i = 1; i < N - 1; i++) {
                  double aij = A[i * N + j], y = prev[so + i];
                  if (aij > LSMALL_D && y > LSMALL_D) x = ladd<EXACT>(x, y + aij);
               }
               cur[so + j] = x + (double)bt[ps[j - 1]];
            }
            double x = LZERO_D;
            for (int i = 1; i < N - 1; i++) {
               double a = A[i * N + N - 1], y = cur[so + i];
               if (a > LSMALL_D && y > LSMALL_D) x = ladd<EXACT>(x, y + a);
            }
            cur[so + N - 1] = x;
         }
      }
      if (tid == 0) { sqA[t] = (short)sq; eqA[t] = (short)eq; }

      // ---- accumulation for the models inside the alpha beam (StepForward, HFB.c:1790-1806)
      const float *bt = bU + (size_t)t * J;
      const bool haveT1 = (t + 1 < T);
      const int loT1 = haveT1 ? qLo[t + 1] : 1, hiT1 = haveT1 ? qHi[t + 1] : 0;
      for (int q = tid; q < Q; q += nt) {
         if (q < sq || q > eq) { mpSelf[q] = LZERO_D; exq[q] = LZERO_D; continue; }
         const int N = sm.sN[q], so = sm.sSoff[q];
         const float *A = A0 + sm.sTr[q];
         const int *ps = posSlot + sm.sPoff[q];
         const double *bq = betaU + (size_t)t * S + so;
         const bool hasB1 = haveT1 && q >= loT1 && q <= hiT1;
         const bool hasBq1 = (q < Q - 1) && (q + 1 >= loT) && (q + 1 <= hiT);
         const double bq1 = hasBq1 ? betaU[(size_t)t * S + sm.sSoff[q + 1]] : LZERO_D;
         const double a1N = A[N - 1];
         const int gq = u.modOff + q;
         if (W.mTmin[gq] > t) W.mTmin[gq] = t;
         W.mTmax[gq] = t;
         double mps = LZERO_D;
         for (int i = 0; i < N - 1; i++) mps = fmax(mps, cur[so + i] + bq[i]);
         mpSelf[q] = mps;
         exq[q] = cur[so + N - 1] + bq[N - 1];
         if (doTr) {
            double *tacc = W.acc + W.mTrAcc[gq], *oacc = W.acc + W.mTrOcc[gq];
            // SetOcct (:399-418) feeding ta->occ (:1388-1389)
            for (int i = 0; i < N - 1; i++) {
               double x = cur[so + i] + bq[i];
               if (i == 0 && hasBq1 && a1N > LSMALL_D) x = ladd<EXACT>(x, cur[so] + bq1 + a1N);
               x -= pr;
               if (x > MINEARG_D) { float o = (float)exp(x); if (o != 0.f) atomicAdd(&oacc[i], (double)o); }
            }
            // UpTranParms (:1390-1410)
            for (int j = 1; j < N - 1; j++) {
               double x = cur[so] + (double)A[j] + (double)bt[ps[j - 1]] + bq[j] - pr;
               if (x > MINEARG_D) atomicAdd(&tacc[j], exp(x));
            }
            if (hasB1) {
               const double *bq1t = betaU + (size_t)(t + 1) * S + so;
               const float *bt1 = bt + J;
               for (int i = 1; i < N - 1; i++)
                  for (int j = 1; j < N - 1; j++) {
                     double x = cur[so + i] + (double)A[i * N + j] + (double)bt1[ps[j - 1]] + bq1t[j] - pr;
                     if (x > MINEARG_D) atomicAdd(&tacc[i * N + j], exp(x));
                  }
            }
            for (int i = 1; i < N - 1; i++) {
               double x = cur[so + i] + (double)A[i * N + N - 1] + bq[N - 1] - pr;
               if (x > MINEARG_D) atomicAdd(&tacc[i * N + N - 1], exp(x));
            }
            if (a1N > LSMALL_D && hasBq1) {
               double x = cur[so] + a1N + bq1 - pr;
               if (x > MINEARG_D) atomicAdd(&tacc[N - 1], exp(x));
            }
         }
         if (doMix) {
            const int *pst = posState + sm.sPoff[q];
            double *oc = occU + (size_t)t * P + sm.sPoff[q];
            for (int j = 1; j < N - 1; j++) {
               int s = pst[j - 1];
               int Mn = M.stateMixOff[s + 1] - M.stateMixOff[s];
               double x;
               if (Mn == 1) x = cur[so + j] + bq[j] - pr;                 // :1575-1576
               else {                                                     // initx, :1480-1489
                  x = (double)A[j] + cur[so];
                  if (t > 0)
                     for (int i = 1; i < N - 1; i++) {
                        double a = A[i * N + j];
                        if (a > LSMALL_D) x = ladd<EXACT>(x, prev[so + i] + a);
                     }
                  x += bq[j] - pr;
               }
               oc[j - 1] = x;
            }
         }
      }
      __syncthreads();
      { double *tmp = cur; cur = prev; prev = tmp; }
   }
   if (tid == 0) {
      atomicAdd(&W.acc[M.L.totalT], (double)T);                          // HERest.c:779-780
      atomicAdd(&W.acc[M.L.totalPr], pr);
      atomicAdd(&W.acc[M.L.numOk], 1.0);
   }
}

// ------------------------------------------------------------------------------------------
// K4: per-mixture statistics, one warp per emitting state position
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
stats_kernel(DevModel M, Wave W, const PosRef *__restrict__ pos, int numPos)
{
   const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
   if (wg >= numPos) return;
   const PosRef p = pos[wg];
   if (W.out[p.utt].status != 0) return;
   const UttDesc u = W.utt[p.utt];
   const int gq = u.modOff + p.q;
   const int tmin = W.mTmin[gq], tmax = W.mTmax[gq];
   if (tmin > tmax) return;
   const int D = M.D, Dp = M.Dp, P = u.P;
   const int pp = W.mPoff[gq] + p.j;
   const int s = W.posState[u.posOff + pp];
   const int mo = M.stateMixOff[s], Mn = M.stateMixOff[s + 1] - mo;
   const double *occ = W.occ + u.occOff + pp;
   const short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const float *feat = W.feat + (size_t)u.featOff * D;
   const double minF = W.minFrwdP;
   const int uf = W.uFlags;
   const bool upM = (uf & HFB_UPMEANS) != 0, upV = (uf & HFB_UPVARS) != 0, upW = (uf & HFB_UPMIXES) != 0;
   const int k0 = lane, k1 = lane + 32;            // D <= 64 on this path (checked at create)
   double wsum = 0.0;
   for (int m = 0; m < Mn; m++) {
      const float wt = M.mixLogWt[mo + m];
      if (!(wt > LMINMIX_F)) continue;                                   // HFB.c:1573
      const int g = M.mixGauss[mo + m];
      const float mu0 = (k0 < D) ? M.mean[(size_t)g * Dp + k0] : 0.f, mu1 = (k1 < D) ? M.mean[(size_t)g * Dp + k1] : 0.f;
      const float iv0 = (k0 < D) ? M.ivar[(size_t)g * Dp + k0] : 0.f, iv1 = (k1 < D) ? M.ivar[(size_t)g * Dp + k1] : 0.f;
      const float gc = M.gconst[g];
      double am0 = 0, am1 = 0, av0 = 0, av1 = 0, aocc = 0;
      for (int t = tmin; t <= tmax; t++) {
         if (p.q < sqA[t] || p.q > eqA[t]) continue;
         double x = occ[(size_t)t * P];
         if (x < -1.0e29) continue;                                       // pre-pruned by the alpha kernel
         const float *o = feat + (size_t)t * D;
         const float d0 = (k0 < D) ? o[k0] - mu0 : 0.f, d1 = (k1 < D) ? o[k1] - mu1 : 0.f;
         if (Mn > 1) {
            float part = warp_sumf(fmaf(d0 * d0, iv0, d1 * d1 * iv1));
            float mixp = -0.5f * (gc + part);
            x = x + (double)wt + (double)mixp;                            // :1581-1599
         }
         if (-x < minF) {                                                 // :1606
            const double Lr = exp(x);
            aocc += Lr;
            const double z0 = (double)d0 * Lr, z1 = (double)d1 * Lr;
            am0 += z0; am1 += z1;
            av0 += z0 * (double)d0; av1 += z1 * (double)d1;
         }
      }
      if (aocc > 0.0) {
         if (upM) {
            double *mu = W.acc + M.L.muSum + (size_t)M.meanId[g] * D;
            if (k0 < D) atomicAdd(&mu[k0], am0);
            if (k1 < D) atomicAdd(&mu[k1], am1);
            if (lane == 0) atomicAdd(&W.acc[M.L.muOcc + M.meanId[g]], aocc);
         }
         if (upV) {
            double *va = W.acc + M.L.vaSum + (size_t)M.varId[g] * D;
            if (k0 < D) atomicAdd(&va[k0], av0);
            if (k1 < D) atomicAdd(&va[k1], av1);
            if (lane == 0) atomicAdd(&W.acc[M.L.vaOcc + M.varId[g]], aocc);
         }
         if (upW && lane == 0) atomicAdd(&W.acc[M.L.wtC + mo + m], aocc);
         wsum += aocc;
      }
   }
   if (lane == 0 && wsum > 0.0) atomicAdd(&W.acc[M.L.wtOcc + s], wsum);  // :1736
}
