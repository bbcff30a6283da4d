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
   if (EXACT && x < LSMALL_D && term > LSMALL_D) return term;      // LAdd(x, term) returns term exactly here
   return EXACT ? ladd<EXACT>(x, term) : ladd_nz(x, term);         // (the branch-free form does so by itself)
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
         int nhi, nlo;
         if (thresh >= 0.5 * HFB_NOPRUNE) {
            // pruning off: gMax - maxP[q] > thresh is never true (finite log values), the beam is
            // the candidate range; only the entry values need to become visible to the neighbours
            nhi = startq; nlo = endq;
            __syncthreads();
         } else {
            double gMax = warp_max(lMax);
            if (lane == 0) wred[wid] = gMax;
            __syncthreads();
            gMax = LZERO_D;
            for (int w = 0; w < nw; w++) gMax = fmax(gMax, wred[w]);
            // ---- pruning (:1254-1272)
            const bool keep = active && !(gMax - lMax > thresh);
            int myHi = __reduce_max_sync(0xffffffffu, keep ? q : -1);
            int myLo = __reduce_min_sync(0xffffffffu, keep ? q : 0x7fffffff);
            if (lane == 0) { whi[wid] = myHi; wlo[wid] = myLo; }
            __syncthreads();
            nhi = -1; nlo = 0x7fffffff;
            for (int w = 0; w < nw; w++) { nhi = max(nhi, whi[w]); nlo = min(nlo, wlo[w]); }
         }
         if (nhi < 0) { fail = true; if (endq == 0) status = HFB_UTT_EBETA; break; }
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

// ------------------------------------------------------------------------------------------
// K3 fast: alpha recursion + alpha beam, one WARP per utterance, lane L owns the model q with
// q = L (mod 32) inside the sliding 32-model window that starts at the beam's lower end.
// Neighbour values travel by warp shuffles; nothing but alpha, the beams and first/last active
// frame per model is written.  Falls back (out->redo) if a beam ever needs more than 32 models.
// ------------------------------------------------------------------------------------------
template <bool EXACT, int E>
__global__ void __launch_bounds__(32) alpha_fast_kernel(DevModel M, Wave W, int forceRedo)
{
   const UttDesc u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) return;
   if (forceRedo) { if (threadIdx.x == 0) out->redo = 1; return; }
   const unsigned FULL = 0xffffffffu;
   const int lane = threadIdx.x;
   const int T = u.T, Q = u.Q, S = u.S, J = u.J, P = u.P;
   const float *bU = W.b + u.bOff;
   const double *betaU = W.beta + u.betaOff;
   double *occU = W.occ + u.occOff, *aentU = W.aent + u.aentOff;
   const short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const double pr = out->pr, minF = W.minFrwdP;
   int *gTmin = W.mTmin + u.modOff, *gTmax = W.mTmax + u.modOff;
   for (int q = lane; q < Q; q += 32) { gTmin[q] = 0x7fffffff; gTmax[q] = -1; }

   ModelRegs<E> r;
   int myq = lane;
   bool have = myq < Q;
   if (have) load_model<E>(r, M, W, u, myq);
   else { r.N = 2; r.so = 0; r.po = 0; r.dms = 1; r.aTee = (float)LZERO_D; }
   double aEm[E], aEx = LZERO_D, mpS = LZERO_D, exv = LZERO_D;
#pragma unroll
   for (int j = 0; j < E; j++) aEm[j] = LZERO_D;
   int tmin = 0x7fffffff, tmax = -1;
   int sq = 0, eq = 0;

   for (int t = 0; t < T; t++) {
      const int loT = qLo[t], hiT = qHi[t];
      const int ne = r.N - 2;
      // loads that do not depend on the recursion go first
      const bool inWin = have && myq >= sq && myq <= ((t == 0) ? hiT : min(Q - 1, eq + 3));
      float b0[E];
      double bEn = LZERO_D, bEm[E], bX = LZERO_D;
#pragma unroll
      for (int j = 0; j < E; j++) { b0[j] = 0.f; bEm[j] = LZERO_D; }
      if (inWin && myq >= loT && myq <= hiT) {
         const float *bt = bU + (size_t)t * J;
         const double *bq = betaU + (size_t)t * S + r.so;
         bEn = bq[0]; bX = bq[r.N - 1];
#pragma unroll
         for (int j = 0; j < E; j++) if (j < ne) { b0[j] = bt[r.slot[j]]; bEm[j] = bq[1 + j]; }
      }
      if (inWin && t + 2 < T) {                              // first touched two frames from now
         const float *bt2 = bU + (size_t)(t + 2) * J;
         const double *bq2 = betaU + (size_t)(t + 2) * S + r.so;
         prefetch_l1(bq2); prefetch_l1(bq2 + r.N - 1);
#pragma unroll
         for (int j = 0; j < E; j++) if (j < ne) prefetch_l1(bt2 + r.slot[j]);
      }
      double a1 = LZERO_D, nEm[E], nEx = LZERO_D;
#pragma unroll
      for (int j = 0; j < E; j++) nEm[j] = LZERO_D;
      int nsq, neq;
      if (t == 0) {
         // ---- InitAlpha, HFB.c:616-651
         nsq = 0; neq = hiT;
         if (neq + 3 >= 32) { if (lane == 0) out->redo = 1; return; }
         if (lane == 0) a1 = 0.0;
         for (int qq = 1; qq <= neq; qq++) {                  // entry chain through leading tee models
            double v = __shfl_sync(FULL, a1, qq - 1) + (double)__shfl_sync(FULL, r.aTee, qq - 1);
            if (lane == qq) a1 = v;
         }
         if (have && myq <= neq) {
#pragma unroll
            for (int j = 0; j < E; j++)
               if (j < ne) nEm[j] = (r.aEnt[j] > (float)LSMALL_D) ? a1 + (double)r.aEnt[j] + (double)b0[j] : LZERO_D;
            double x = LZERO_D;
#pragma unroll
            for (int i = 0; i < E; i++)
               if (i < ne && r.aExit[i] > (float)LSMALL_D) x = ladd<EXACT>(x, nEm[i] + (double)r.aExit[i]);
            nEx = x;
         } else a1 = LZERO_D;
      } else {
         // ---- alpha beam, HFB.c:701-722
         const int loP = qLo[t - 1], hiP = qHi[t - 1];
         const int q1 = __shfl_sync(FULL, myq, (lane + 31) & 31), q2 = __shfl_sync(FULL, myq, (lane + 30) & 31);
         const double ex1 = __shfl_sync(FULL, exv, (lane + 31) & 31), ex2 = __shfl_sync(FULL, exv, (lane + 30) & 31);
         const double ax1 = __shfl_sync(FULL, aEx, (lane + 31) & 31), ax2 = __shfl_sync(FULL, aEx, (lane + 30) & 31);
         const float tee1 = __shfl_sync(FULL, r.aTee, (lane + 31) & 31);
         const bool ok1 = (q1 == myq - 1), ok2 = (q2 == myq - 2);
         const double e1 = ok1 ? ex1 : LZERO_D, e2 = ok2 ? ex2 : LZERO_D;
         double mp = fmax(e1, mpS);
         int c = (inWin && myq >= loP && !(pr - mp > minF)) ? myq : 0x7fffffff;
         nsq = __reduce_min_sync(FULL, c);
         if (nsq > hiT) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }        // HError 7390
         if (nsq < loT) nsq = loT;
         const int eq0 = (hiP < Q - 1) ? hiP + 1 : hiP;
         double mp2 = e1;
         if (myq >= 2 && myq - 1 > nsq && ok1 && tee1 > (float)LSMALL_D) mp2 = fmax(mp2, e2);
         mp2 = fmax(mp2, mpS);
         c = (inWin && myq <= eq0 && !(pr - mp2 > minF)) ? myq : -1;
         neq = __reduce_max_sync(FULL, c);
         if (neq < nsq) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }
         while (neq < Q - 1) {                                 // while (eq<Q && qDms[eq]==0) eq++
            const int hq = __shfl_sync(FULL, myq, neq & 31), hd = __shfl_sync(FULL, r.dms, neq & 31);
            if (hq != neq) { if (lane == 0) out->redo = 1; return; }
            if (hd == 0) neq++; else break;
         }
         if (neq > hiT) neq = hiT;
         if (neq + 3 - nsq >= 32) { if (lane == 0) out->redo = 1; return; }
         // ---- alpha column, HFB.c:729-771
         if (have && myq >= nsq && myq <= neq) {
            a1 = (myq > 0 && ok1) ? ax1 : LZERO_D;
            if (myq > nsq && ok1 && tee1 > (float)LSMALL_D) {
               double y = (myq >= 2 && ok2) ? ax2 : LZERO_D;
               a1 = ladd<EXACT>(a1, y + (double)tee1);
            }
#pragma unroll
            for (int j = 0; j < E; j++) {
               if (j < ne) {
                  double x = (r.aEnt[j] > (float)LSMALL_D) ? (double)r.aEnt[j] + a1 : LZERO_D;
#pragma unroll
                  for (int i = 0; i < E; i++)
                     if (r.aInt[i][j] > (float)LSMALL_D && aEm[i] > LSMALL_D) x = lacc<EXACT>(x, aEm[i] + (double)r.aInt[i][j]);
                  nEm[j] = x + (double)b0[j];
               }
            }
            double x = LZERO_D;
#pragma unroll
            for (int i = 0; i < E; i++)
               if (i < ne && r.aExit[i] > (float)LSMALL_D && nEm[i] > LSMALL_D) x = lacc<EXACT>(x, nEm[i] + (double)r.aExit[i]);
            nEx = x;
         }
      }
      sq = nsq; eq = neq;
      const bool inBeam = have && myq >= sq && myq <= eq;
      if (lane == 0) { sqA[t] = (short)sq; eqA[t] = (short)eq; }
      aEx = nEx;
      mpS = LZERO_D; exv = LZERO_D;
      if (inBeam) {
         double m = a1 + bEn;
#pragma unroll
         for (int j = 0; j < E; j++) if (j < ne) m = fmax(m, nEm[j] + bEm[j]);
         mpS = m; exv = nEx + bX;
         double *oc = occU + (size_t)t * P + r.po;
#pragma unroll
         for (int j = 0; j < E; j++) if (j < ne) oc[j] = nEm[j];
         aentU[(size_t)t * Q + myq] = a1;
         if (tmin > t) tmin = t;
         tmax = t;
      }
#pragma unroll
      for (int j = 0; j < E; j++) aEm[j] = nEm[j];
      // ---- slide: a lane whose model fell below the beam takes the model 32 further on
      if (have && myq < sq) {
         if (tmax >= 0) { gTmin[myq] = tmin; gTmax[myq] = tmax; }
         myq += 32; have = myq < Q;
         tmin = 0x7fffffff; tmax = -1;
         aEx = LZERO_D; mpS = LZERO_D; exv = LZERO_D;
#pragma unroll
         for (int j = 0; j < E; j++) aEm[j] = LZERO_D;
         if (have) load_model<E>(r, M, W, u, myq);
         else { r.N = 2; r.so = 0; r.po = 0; r.dms = 1; r.aTee = (float)LZERO_D; }
      }
   }
   if (have && tmax >= 0) { gTmin[myq] = tmin; gTmax[myq] = tmax; }
   for (int q = lane; q < Q; q += 32) atomicAdd(&W.acc[M.L.numEgs + W.mHmm[u.modOff + q]], 1.0);   // HFB.c:1768-1772
   if (lane == 0) {
      atomicAdd(&W.acc[M.L.totalT], (double)T);                             // HERest.c:779-780
      atomicAdd(&W.acc[M.L.totalPr], pr);
      atomicAdd(&W.acc[M.L.numOk], 1.0);
   }
}
