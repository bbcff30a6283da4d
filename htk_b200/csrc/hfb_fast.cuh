// hfb_fast.cuh -- register-resident recursion kernels for the common case (HMMs with at most
// E+2 states, E = 3 or 6; transcriptions that fit one CTA / one 32-model alpha window).
//
// The generic kernels (hfb_kernels.cuh / hfb_kernels2.cuh) walk N x N transition matrices in
// global memory through shared-memory columns.  Here every thread keeps ITS model's transition
// logs, output-probability slots and state values in registers, the inner loops are unrolled over
// E, and a log-add whose running value is still "log zero" is replaced by an assignment (exactly
// what LAdd returns in that case, HMath.c:1584-1585).  Same arithmetic, ~5x fewer instructions.
//
//   beta_fast_kernel       SetBeamTaper + SetBeta + StepBack retry loop (HFB.c:1116-1366)
//   alpha_fast_kernel      InitAlpha / StepAlpha + alpha beam (HFB.c:616-784): recursion ONLY --
//                          it stores alpha and the beams; every accumulation is done by the
//                          parallel stats3_kernel (hfb_kernels2.cuh)
#pragma once
#include "hfb_kernels.cuh"

// x (+) term in the log domain; `term` must already satisfy the caller's LSMALL guards
template <bool EXACT>
__device__ __forceinline__ double lacc(double x, double term)
{
   if (x < LSMALL_D && term > LSMALL_D) return term;      // LAdd(x, term) returns term exactly here
   return ladd<EXACT>(x, term);
}

template <int E>
struct ModelRegs {
   float aEnt[E];          // entry -> emitting j
   float aInt[E][E];       // emitting i -> emitting j
   float aExit[E];         // emitting i -> exit
   float aTee;             // entry -> exit
   int slot[E];            // output-probability slot of emitting j
   int N, so, po, dms;     // states, offset in a beta/alpha column, offset in an occ row, min duration
};

template <int E>
__device__ __forceinline__ void load_model(ModelRegs<E> &r, const DevModel &M, const Wave &W, const UttDesc &u, int q)
{
   const int gq = u.modOff + q;
   r.N = W.mN[gq]; r.so = W.mSoff[gq]; r.po = W.mPoff[gq]; r.dms = W.mDms[gq];
   const float *A = M.transLogA + W.mTrans[gq];
   const int N = r.N, ne = N - 2;
   const int *ps = W.posSlot + u.posOff + r.po;
#pragma unroll
   for (int j = 0; j < E; j++) {
      r.aEnt[j] = (j < ne) ? A[1 + j] : (float)LZERO_D;
      r.aExit[j] = (j < ne) ? A[(1 + j) * N + N - 1] : (float)LZERO_D;
      r.slot[j] = (j < ne) ? ps[j] : 0;
#pragma unroll
      for (int k = 0; k < E; k++) r.aInt[j][k] = (j < ne && k < ne) ? A[(1 + j) * N + 1 + k] : (float)LZERO_D;
   }
   r.aTee = A[N - 1];
}

__host__ __device__ inline size_t beta_fast_smem_bytes(int Q)
{
   return sizeof(double) * ((size_t)2 * (Q + 2) + 32) + sizeof(int) * 64 + sizeof(float) * (size_t)(Q + 2) + sizeof(int) * (size_t)(Q + 2);
}

// ------------------------------------------------------------------------------------------
// K2 fast: one CTA per utterance, ONE thread per model (Q <= blockDim)
// ------------------------------------------------------------------------------------------
template <bool EXACT, int E>
__global__ void __launch_bounds__(256) beta_fast_kernel(DevModel M, Wave W)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) {
      if (threadIdx.x == 0 && out->status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
      return;
   }
   const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
   const int T = u.T, Q = u.Q, S = u.S, J = u.J;
   double *entA = (double *)smraw, *entB = entA + (Q + 2), *wred = entB + (Q + 2);
   int *wlo = (int *)(wred + 32), *whi = wlo + 32;
   float *sTee = (float *)(whi + 32);                  // entry->exit log prob of every model
   int *sDms = (int *)(sTee + (Q + 2));
   const int q = tid;
   const bool mine = q < Q;
   ModelRegs<E> r;
   if (mine) { load_model<E>(r, M, W, u, q); sTee[q] = r.aTee; sDms[q] = r.dms; }
   else { r.N = 2; r.so = 0; r.po = 0; r.dms = 1; r.aTee = (float)LZERO_D; }
   const float *bU = W.b + u.bOff;
   double *betaU = W.beta + u.betaOff;
   short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   const int *pre = W.mPre + u.modOff, *suf = W.mSuf + u.modOff;
   const int ne = r.N - 2;

   double thresh = W.pruneInit, pr = LZERO_D;
   int retries = 0, status = 0;
   __syncthreads();

   for (;;) {
      // ---- SetBeamTaper (closed form, see beta_kernel)
      for (int t = tid; t < T; t += nt) {
         int lo = 0, hi = Q;
         while (lo < hi) { int mid = (lo + hi) >> 1; if (pre[mid] <= t) lo = mid + 1; else hi = mid; }
         qHi[t] = (short)(lo - 1);
         int rr = T - 1 - t; lo = 0; hi = Q;
         while (lo < hi) { int mid = (lo + hi) >> 1; if (suf[mid] > rr) lo = mid + 1; else hi = mid; }
         qLo[t] = (short)lo;
      }
      __syncthreads();

      double *cur = entA, *prev = entB;                // entry-state beta of every model, frames t / t+1
      double bE[E], bExit = LZERO_D;                   // this model's emitting / exit beta at frame t+1
      float b1[E];                                     // b_j(o_{t+1})
#pragma unroll
      for (int j = 0; j < E; j++) { bE[j] = LZERO_D; b1[j] = 0.f; }

      // ---- t = T-1, HFB.c:1176-1198 (the exit chain through trailing tee models is serial)
      int lo1 = qLo[T - 1], hi1 = Q - 1, lastq = lo1;
      if (tid == 0) {
         double bn = 0.0;
         for (int qq = Q - 1; qq >= lo1; qq--) {
            bn = (qq == Q - 1) ? 0.0 : bn + (double)sTee[qq + 1];
            prev[qq] = bn;                              // borrowed as scratch for the exit values
         }
         qHi[T - 1] = (short)(Q - 1);
      }
      __syncthreads();
      if (mine && q >= lo1) {
         const float *bt = bU + (size_t)(T - 1) * J;
         bExit = prev[q];
         double x = LZERO_D;
#pragma unroll
         for (int j = 0; j < E; j++) {
            if (j < ne) {
               bE[j] = (double)r.aExit[j] + bExit;
               b1[j] = bt[r.slot[j]];
               if (r.aEnt[j] > (float)LSMALL_D && bE[j] > LSMALL_D) x = ladd<EXACT>(x, (double)r.aEnt[j] + (double)b1[j] + bE[j]);
            }
         }
         double *bg = betaU + (size_t)(T - 1) * S + r.so;
         bg[0] = x;
#pragma unroll
         for (int j = 0; j < E; j++) if (j < ne) bg[1 + j] = bE[j];
         bg[r.N - 1] = bExit;
         cur[q] = x;
      }
      __syncthreads();
      { double *tmp = cur; cur = prev; prev = tmp; }

      // ---- t = T-2 .. 0, HFB.c:1205-1277
      bool fail = false;
      for (int t = T - 2; t >= 0; t--) {
         const int tapLo = qLo[t], tapHi = qHi[t];
         int startq = hi1;
         int endq = (lo1 == 0) ? 0 : ((tapLo >= lo1) ? tapLo : lo1 - 1);
         while (endq > 0 && sDms[endq - 1] == 0) endq--;
         lastq = endq;
         const float *bt = bU + (size_t)t * J;
         double lMax = LZERO_D;
         const bool active = mine && q >= endq && q <= startq;
         if (mine && q >= endq - 2 && q <= startq && t >= 2) {
            const float *bt2 = bU + (size_t)(t - 2) * J;
#pragma unroll
            for (int j = 0; j < E; j++) if (j < ne) prefetch_l1(bt2 + r.slot[j]);
         }
         if (active) {
            const bool in1 = (q >= lo1 && q <= hi1);
            float b0[E];
#pragma unroll
            for (int j = 0; j < E; j++) b0[j] = (j < ne) ? bt[r.slot[j]] : 0.f;
            // exit state (:1225-1227)
            double ex = LZERO_D;
            if (q < Q - 1) {
               if (q + 1 >= lo1 && q + 1 <= hi1) ex = prev[q + 1];
               if (q < startq) {
                  const double a1N = sTee[q + 1];
                  if (a1N > LSMALL_D) {
                     double y = (q + 2 < Q && q + 2 >= lo1 && q + 2 <= hi1) ? prev[q + 2] : LZERO_D;
                     ex = ladd<EXACT>(ex, y + a1N);
                  }
               }
            }
            double uj[E];                               // b_j(o_{t+1}) + beta_j(t+1), or log zero
#pragma unroll
            for (int j = 0; j < E; j++) uj[j] = (in1 && j < ne && bE[j] > LSMALL_D) ? (double)b1[j] + bE[j] : LZERO_D;
            double nb[E];
#pragma unroll
            for (int i = E - 1; i >= 0; i--) {
               double x = LZERO_D;
               if (i < ne) {
                  x = (double)r.aExit[i] + ex;
#pragma unroll
                  for (int j = 0; j < E; j++)
                     if (r.aInt[i][j] > (float)LSMALL_D && uj[j] > LSMALL_D) x = lacc<EXACT>(x, (double)r.aInt[i][j] + uj[j]);
                  lMax = fmax(lMax, x);
               }
               nb[i] = x;
            }
            double x = LZERO_D;
#pragma unroll
            for (int j = 0; j < E; j++) {
               bE[j] = nb[j]; b1[j] = b0[j];
               if (j < ne && r.aEnt[j] > (float)LSMALL_D && nb[j] > LSMALL_D)
                  x = lacc<EXACT>(x, (double)r.aEnt[j] + (double)b0[j] + nb[j]);
            }
            bExit = ex;
            double *bg = betaU + (size_t)t * S + r.so;
            bg[0] = x;
#pragma unroll
            for (int j = 0; j < E; j++) if (j < ne) bg[1 + j] = nb[j];
            bg[r.N - 1] = ex;
            cur[q] = x;
         }
         double gMax = warp_max(lMax);
         if (lane == 0) wred[wid] = gMax;
         __syncthreads();
         gMax = LZERO_D;
         for (int w = 0; w < nw; w++) gMax = fmax(gMax, wred[w]);
         // ---- pruning (:1254-1272)
         const bool keep = active && !(gMax - lMax > thresh);
         int myHi = keep ? q : -1, myLo = keep ? q : 0x7fffffff;
         myHi = warp_maxi(myHi); myLo = warp_mini(myLo);
         if (lane == 0) { whi[wid] = myHi; wlo[wid] = myLo; }
         __syncthreads();
         int nhi = -1, nlo = 0x7fffffff;
         for (int w = 0; w < nw; w++) { nhi = max(nhi, whi[w]); nlo = min(nlo, wlo[w]); }
         if (nhi < 0) { fail = true; status = HFB_UTT_EBETA; break; }
         if (nhi > tapHi) nhi = tapHi;
         if (nlo > nhi) { fail = true; break; }
         if (tid == 0) { qHi[t] = (short)nhi; qLo[t] = (short)nlo; }
         hi1 = nhi; lo1 = nlo;
         { double *tmp = cur; cur = prev; prev = tmp; }
      }
      if (status != 0) break;
      if (!fail) {
         pr = prev[lastq];                             // utt->pr = bqt[1] (:1280)
         if (pr > LSMALL_D) break;
      }
      thresh += W.pruneInc;                            // StepBack retry (:1349-1361)
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
