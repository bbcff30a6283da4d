// hfb_kernels2.cuh -- the general (any topology, any transcription length) forward pass and the FP32
// statistics kernel.
//
//   alpha_warp_kernel  one WARP per utterance (no block barriers): the lanes cover the window of
//                      models around the alpha beam, [sq(t-1), eq(t-1)+3], which is all StepAlpha
//                      can reach in one frame (HFB.c:701-722: the beam is re-derived from the
//                      previous column); state columns in shared memory.  Redoes the utterances the
//                      register-resident kernels (hfb_fast.cuh, hfb_l2r.cuh) gave up on.
//   stats_tran_chunk   SetOcct + UpTranParms for one 32-frame chunk of a position (shared with stats5)
//   stats3_kernel      one warp per emitting state position, all sums on the FP32 pipe; component
//                      log-densities in the reference's own operation order (IDOutP,
//                      HModel.c:5425-5430).  Used for single-Gaussian sets and D > 39; the
//                      mixture sets go through stats5_kernel (hfb_stats_mma.cuh).
#pragma once
#include "hfb_kernels.cuh"

#define OCC_SKIP (-1.0e30)

__host__ __device__ inline size_t alpha_warp_smem_bytes(int S, int Q)
{
   return sizeof(double) * ((size_t)2 * S + 2 * Q) + sizeof(int) * ((size_t)2 * Q);
}


template <bool EXACT>
__global__ void __launch_bounds__(32) alpha_warp_kernel(DevModel M, Wave W, int onlyRedo)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) return;
   if (onlyRedo && !out->redo) return;
   const int lane = threadIdx.x;
   const int T = u.T, Q = u.Q, S = u.S, J = u.J, P = u.P;
   double *cur = (double *)smraw, *prev = cur + S, *mpSelf = prev + S, *exq = mpSelf + Q;
   int *sTmin = (int *)(exq + Q), *sTmax = sTmin + Q;
   const int *mN = W.mN + u.modOff, *mSoff = W.mSoff + u.modOff, *mTr = W.mTrans + u.modOff;
   const int *mPoff = W.mPoff + u.modOff, *mDms = W.mDms + u.modOff;
   const float *A0 = M.transLogA;
   const int *posSlot = W.posSlot + u.posOff;
   const float *bU = W.b + u.bOff;
   const double *betaU = W.beta + u.betaOff;
   double *occU = W.occ + u.occOff, *aentU = W.aent + u.aentOff;
   const short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const double pr = out->pr, minF = W.minFrwdP;

   for (int q = lane; q < Q; q += 32) {
      mpSelf[q] = LZERO_D; exq[q] = LZERO_D;
      sTmin[q] = 0x7fffffff; sTmax[q] = -1;
   }
   for (int i = lane; i < 2 * S; i += 32) cur[i] = LZERO_D;
   __syncwarp();

   int sq = 0, eq = 0, sqP2 = 0, eqP2 = -1;
   for (int t = 0; t < T; t++) {
      const int loT = qLo[t], hiT = qHi[t];
      if (t == 0) {
         // ---- InitAlpha, HFB.c:616-651
         eq = hiT; sq = 0;
         if (lane == 0) {
            double a1 = 0.0, a1N = 0.0;
            for (int q = 0; q <= eq; q++) {
               a1 = (q == 0) ? 0.0 : a1 + a1N;
               cur[mSoff[q]] = a1;
               a1N = A0[mTr[q] + mN[q] - 1];
            }
         }
         __syncwarp();
         for (int q = lane; q <= eq; q += 32) {
            const int N = mN[q], so = mSoff[q];
            const float *A = A0 + mTr[q];
            const int *ps = posSlot + mPoff[q];
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
         { double *tmp = cur; cur = prev; prev = tmp; }
         // ---- alpha beam, HFB.c:701-722; candidates can only lie in [sq, eq+3] of the previous frame
         const int loP = qLo[t - 1], hiP = qHi[t - 1];
         const int wLo = sq, wHi = min(Q - 1, eq + 3);
         int mySq = 0x7fffffff;
         for (int q = wLo + lane; q <= wHi; q += 32) {
            if (q < loP) continue;
            double mp = fmax((q > 0) ? exq[q - 1] : LZERO_D, mpSelf[q]);
            if (!(pr - mp > minF)) mySq = min(mySq, q);
         }
         int nsq = warp_mini(mySq);
         if (nsq > hiT) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }       // HError 7390
         if (nsq < loT) nsq = loT;
         const int eq0 = (hiP < Q - 1) ? hiP + 1 : hiP;
         int myEq = -1;
         for (int q = wLo + lane; q <= wHi; q += 32) {
            if (q > eq0) continue;
            double mp = (q > 0) ? exq[q - 1] : LZERO_D;
            if (q > 0 && q - 1 > nsq && q >= 2) {
               if ((double)A0[mTr[q - 1] + mN[q - 1] - 1] > LSMALL_D) mp = fmax(mp, exq[q - 2]);
            }
            mp = fmax(mp, mpSelf[q]);
            if (!(pr - mp > minF)) myEq = max(myEq, q);
         }
         int neq = warp_maxi(myEq);
         if (neq < nsq) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }
         while (neq < Q - 1 && mDms[neq] == 0) neq++;
         if (neq > hiT) neq = hiT;
         // ---- alpha column, HFB.c:729-771; also clears what the column of frame t-2 left behind
         const int cLo = sqP2, cHi = min(Q - 1, max(max(eqP2, eq), neq) + 3);
         const float *bt = bU + (size_t)t * J;
         for (int q = cLo + lane; q <= cHi; q += 32) {
            const int N = mN[q], so = mSoff[q];
            if (q < nsq || q > neq) { for (int i = 0; i < N; i++) cur[so + i] = LZERO_D; continue; }
            const float *A = A0 + mTr[q];
            const int *ps = posSlot + mPoff[q];
            double a1 = LZERO_D;
            if (q > 0) {
               const int N1 = mN[q - 1];
               a1 = prev[mSoff[q - 1] + N1 - 1];
               const double a1N = A0[mTr[q - 1] + N1 - 1];
               if (q > nsq && a1N > LSMALL_D) {
                  double y = (q >= 2) ? prev[mSoff[q - 2] + mN[q - 2] - 1] : LZERO_D;
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
         sqP2 = sq; eqP2 = eq;
         sq = nsq; eq = neq;
      }
      __syncwarp();
      if (lane == 0) { sqA[t] = (short)sq; eqA[t] = (short)eq; }
      // pull what frame t+1 will touch for the first time (beta and b of frame t+2) towards L1 now:
      // the T-step chain is latency bound and these addresses do not depend on the recursion
      if (t + 2 < T) {
         const int pHi = min(Q - 1, eq + 3);
         const float *bt2 = bU + (size_t)(t + 2) * J;
         for (int q = sq + lane; q <= pHi; q += 32) {
            const int N = mN[q];
            const double *b2 = betaU + (size_t)(t + 2) * S + mSoff[q];
            prefetch_l1(b2); prefetch_l1(b2 + N - 1);
            const int *ps = posSlot + mPoff[q];
            for (int j = 0; j < N - 2; j++) prefetch_l1(bt2 + ps[j]);
         }
      }

      // ---- what the next frame's beam needs (MaxModelProb, HFB.c:655-682) + alpha for stats3_kernel
      const int aLo = (t == 0) ? 0 : sqP2, aHi = (t == 0) ? eq : min(Q - 1, max(eqP2, eq) + 3);
      for (int q = aLo + lane; q <= aHi; q += 32) {
         if (q < sq || q > eq) { mpSelf[q] = LZERO_D; exq[q] = LZERO_D; continue; }
         const int N = mN[q], so = mSoff[q];
         const double *bq = betaU + (size_t)t * S + so;
         if (sTmin[q] > t) sTmin[q] = t;
         sTmax[q] = t;
         double mps = LZERO_D;
         for (int i = 0; i < N - 1; i++) mps = fmax(mps, cur[so + i] + bq[i]);
         mpSelf[q] = mps;
         exq[q] = cur[so + N - 1] + bq[N - 1];
         double *oc = occU + (size_t)t * P + mPoff[q];
         for (int j = 1; j < N - 1; j++) oc[j - 1] = cur[so + j];
         aentU[(size_t)t * Q + q] = cur[so];
      }
      __syncwarp();
   }
   for (int q = lane; q < Q; q += 32) {
      W.mTmin[u.modOff + q] = sTmin[q]; W.mTmax[u.modOff + q] = sTmax[q];
      atomicAdd(&W.acc[M.L.numEgs + W.mHmm[u.modOff + q]], 1.0);            // HFB.c:1768-1772
   }
   if (lane == 0) {
      atomicAdd(&W.acc[M.L.totalT], (double)T);                             // HERest.c:779-780
      atomicAdd(&W.acc[M.L.totalPr], pr);
      atomicAdd(&W.acc[M.L.numOk], 1.0);
   }
}

// ------------------------------------------------------------------------------------------
// K4: every accumulation of StepForward, in parallel over emitting state positions.
//
// One warp per position (q, j).  For the frames where q is inside the alpha beam (lanes <-> frames
// of a 32-frame chunk) it forms, from the stored alpha / beta / b:
//   * SetOcct + UpTranParms (HFB.c:399-418, :1371-1423): occupancy of state j, transitions
//     entry->j, j->j2 (all j2), j->exit; the warp handling the FIRST emitting state also does the
//     model-level terms (entry occupancy incl. the tee path, entry->exit); partial sums are reduced
//     over the warp so that one atomic per chunk reaches the (hot) transition accumulators;
//   * UpMixParms (HFB.c:1426-1744): x = alpha_j + beta_j - pr for single-mixture states, else
//     initx + log weight + component log-density with initx = alpha_j - b_j + beta_j - pr (equal to
//     HFB.c:1480-1489 because alpha_j(t) = [that sum] + b_j(o_t)), the minimum-occupancy rule
//     -x < minFrwdP, and the centred mean / variance / weight sums.
// Component log-densities use the reference's operation order (IDOutP, HModel.c:5425-5430).
// ------------------------------------------------------------------------------------------
#define ST_WARPS 4
__host__ __device__ inline int stats_ostride(int D) { return D | 1; }
__host__ __device__ inline size_t stats_smem_bytes(int D)
{
   return ST_WARPS * (sizeof(float) * ((size_t)32 * stats_ostride(D) + 32 * 33) + sizeof(double) * 32 + sizeof(int) * 32);
}

__device__ __forceinline__ float warp_sum_nz(float v)
{
   for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
   return v;
}


// Everything stats3/stats4 need to know about "their" emitting state position
struct StatPos {
   const double *alphaJ, *aent, *betaU;
   const float *bU, *A;
   const int *ps;
   const short *qLo, *qHi;
   double *tacc, *oacc;
   double pr;
   int P, S, Q, J, T, N, so, sq1, j, q;
   float aEntJ, aExitJ, aTee;
};

// SetOcct + UpTranParms (HFB.c:399-418, :1371-1423) for frame t = chunk start + lane of one position:
// occupancy of state j, transitions entry->j, j->j2, j->exit and, for the first emitting state, the
// model-level terms; warp-reduced so that one atomic per chunk reaches the (hot) transition accumulators.
__device__ __forceinline__ void stats_tran_chunk(const StatPos &c, int t, bool inb, int lane)
{
   const double *alphaJ = c.alphaJ, *aent = c.aent, *betaU = c.betaU;
   const float *bU = c.bU, *A = c.A;
   const int *ps = c.ps;
   const short *qLo = c.qLo, *qHi = c.qHi;
   double *tacc = c.tacc, *oacc = c.oacc;
   const double pr = c.pr;
   const int P = c.P, S = c.S, Q = c.Q, J = c.J, T = c.T, N = c.N, so = c.so, sq1 = c.sq1, j = c.j, q = c.q;
   const float aEntJ = c.aEntJ, aExitJ = c.aExitJ, aTee = c.aTee;
   {
         // ---- SetOcct / UpTranParms for this state (and, for j == 0, the model-level terms)
         float oJ = 0.f, tEnt = 0.f, tExit = 0.f, oEnt = 0.f, tTee = 0.f;
         const bool hasB1 = inb && (t + 1 < T) && q >= qLo[t + 1] && q <= qHi[t + 1];   // bqt1 != NULL
         if (inb) {
            const double aj = alphaJ[(size_t)t * P], ae = aent[(size_t)t * Q];
            const double *bq = betaU + (size_t)t * S + so;
            const float bjt = bU[(size_t)t * J + ps[j]];
            double x = aj + bq[1 + j] - pr;
            if (x > -87.0) oJ = expf((float)x);                              // occt[j] -> ta->occ
            x = ae + (double)aEntJ + (double)bjt + bq[1 + j] - pr;           // entry -> j
            if (x > -87.0) tEnt = expf((float)x);
            x = aj + (double)aExitJ + bq[N - 1] - pr;                        // j -> exit
            if (x > -87.0) tExit = expf((float)x);
            if (j == 0) {
               const bool hasBq1 = (q < Q - 1) && (q + 1 >= qLo[t]) && (q + 1 <= qHi[t]);   // bq1t != NULL
               const double bnx = hasBq1 ? betaU[(size_t)t * S + sq1] : LZERO_D;
               x = ae + bq[0];
               if (hasBq1 && aTee > (float)LSMALL_D) x = ladd<false>(x, ae + bnx + (double)aTee);
               x -= pr;
               if (x > -87.0) oEnt = expf((float)x);                         // occt[1]
               if (hasBq1 && aTee > (float)LSMALL_D) {
                  x = ae + (double)aTee + bnx - pr;
                  if (x > -87.0) tTee = expf((float)x);
               }
            }
         }
         if (__ballot_sync(0xffffffffu, inb)) {
            oJ = warp_sum_nz(oJ); tEnt = warp_sum_nz(tEnt); tExit = warp_sum_nz(tExit);
            if (lane == 0) {
               if (oJ != 0.f) atomicAdd(&oacc[1 + j], (double)oJ);
               if (tEnt != 0.f) atomicAdd(&tacc[1 + j], (double)tEnt);
               if (tExit != 0.f) atomicAdd(&tacc[(1 + j) * N + N - 1], (double)tExit);
            }
            for (int j2 = 0; j2 < N - 2; j2++) {                             // j -> j2 (internal)
               const float a = A[(1 + j) * N + 1 + j2];
               if (!(a > (float)LSMALL_D)) continue;
               float v = 0.f;
               if (hasB1) {
                  const double x = alphaJ[(size_t)t * P] + (double)a + (double)bU[(size_t)(t + 1) * J + ps[j2]] +
                                   betaU[(size_t)(t + 1) * S + so + 1 + j2] - pr;
                  if (x > -87.0) v = expf((float)x);
               }
               v = warp_sum_nz(v);
               if (lane == 0 && v != 0.f) atomicAdd(&tacc[(1 + j) * N + 1 + j2], (double)v);
            }
            if (j == 0) {
               oEnt = warp_sum_nz(oEnt); tTee = warp_sum_nz(tTee);
               if (lane == 0) {
                  if (oEnt != 0.f) atomicAdd(&oacc[0], (double)oEnt);
                  if (tTee != 0.f) atomicAdd(&tacc[N - 1], (double)tTee);
               }
            }
         }
         }
}

__global__ void __launch_bounds__(32 * ST_WARPS)
stats3_kernel(DevModel M, Wave W)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const int wInB = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int wg = blockIdx.x * ST_WARPS + wInB;
   if (wg >= W.totalPos) return;
   const int ui = upper_index(W.posPre, W.numUtt, wg);
   if (W.out[ui].status != 0) return;
   const UttDesc u = W.utt[ui];
   const int lp = wg - W.posPre[ui];
   const int q = upper_index(W.mPoff + u.modOff, u.Q, lp);
   const int gq = u.modOff + q;
   const int j = lp - W.mPoff[gq];                     // emitting index 0..N-3
   const int tmin = W.mTmin[gq], tmax = W.mTmax[gq];
   if (tmin > tmax) return;
   const int D = M.D, Dp = M.Dp, P = u.P, S = u.S, Q = u.Q, J = u.J, T = u.T, ostr = stats_ostride(D);
   const size_t perWarp = sizeof(float) * ((size_t)32 * ostr + 32 * 33) + sizeof(double) * 32 + sizeof(int) * 32;
   unsigned char *mine = smraw + perWarp * wInB;
   double *x0s = (double *)mine;                       // [32] initx / log occupancy per chunk frame
   float *os = (float *)(x0s + 32);                    // [32][ostr] observation rows
   float *lrs = os + 32 * ostr;                        // [32 mixtures][33] occupancies Lr
   int *ts = (int *)(lrs + 32 * 33);                   // [32] frame numbers

   const int N = W.mN[gq], so = W.mSoff[gq];
   const int s = W.posState[u.posOff + lp];
   const int mo = M.stateMixOff[s], Mn = M.stateMixOff[s + 1] - mo;
   const float *A = M.transLogA + W.mTrans[gq];
   const int *ps = W.posSlot + u.posOff + W.mPoff[gq];
   const double *alphaJ = W.occ + u.occOff + lp, *aent = W.aent + u.aentOff + q;
   const double *betaU = W.beta + u.betaOff;
   const float *bU = W.b + u.bOff;
   const short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   const float *feat = W.feat + (size_t)u.featOff * D;
   // HERest -r (HFB.c:1603-1611, :1731): posteriors from `feat`, mean / variance sums from the second data stream
   const float *feat2 = W.feat2 ? W.feat2 + (size_t)u.featOff * D : nullptr;
   const double pr = W.out[ui].pr, minF = W.minFrwdP;
   const int uf = W.uFlags;
   const bool upM = (uf & HFB_UPMEANS) != 0, upV = (uf & HFB_UPVARS) != 0, upW = (uf & HFB_UPMIXES) != 0;
   const bool doMix = upM || upV || upW, doTr = (uf & HFB_UPTRANS) != 0;
   const float aEntJ = A[1 + j], aExitJ = A[(1 + j) * N + N - 1], aTee = A[N - 1];
   const int sq1 = (q < Q - 1) ? W.mSoff[gq + 1] : 0;  // column offset of the next model
   double *tacc = W.acc + W.mTrAcc[gq], *oacc = W.acc + W.mTrOcc[gq];
   const int k0 = lane, k1 = lane + 32;                // D <= 64 (checked at create)
   double wsum = 0.0;
   StatPos sp;
   sp.alphaJ = alphaJ; sp.aent = aent; sp.betaU = betaU; sp.bU = bU; sp.A = A; sp.ps = ps; sp.qLo = qLo; sp.qHi = qHi;
   sp.tacc = tacc; sp.oacc = oacc; sp.pr = pr; sp.P = P; sp.S = S; sp.Q = Q; sp.J = J; sp.T = T; sp.N = N; sp.so = so;
   sp.sq1 = sq1; sp.j = j; sp.q = q; sp.aEntJ = aEntJ; sp.aExitJ = aExitJ; sp.aTee = aTee;

   for (int t0 = tmin; t0 <= tmax; t0 += 32) {
      const int t = t0 + lane;
      const bool inb = (t <= tmax) && q >= sqA[t] && q <= eqA[t];
      double x0 = OCC_SKIP;
      if (inb) {
         const double aj = alphaJ[(size_t)t * P];
         const double *bq = betaU + (size_t)t * S + so;
         const float bjt = bU[(size_t)t * J + ps[j]];
         const double lg = aj + bq[1 + j] - pr;                              // log occupancy of state j
         if (!(lg < -(minF + 0.25))) x0 = (Mn == 1) ? lg : lg - (double)bjt;  // :1575-1576 / initx :1480-1489
      }
      if (doTr) stats_tran_chunk(sp, t, inb, lane);
      if (!doMix) continue;
      // ---- frames of this chunk that can contribute to the mixture statistics
      const bool valid = inb && x0 > -1.0e29;
      const unsigned mask = __ballot_sync(0xffffffffu, valid);
      const int nT = __popc(mask);
      if (nT == 0) continue;
      if (valid) { int idx = __popc(mask & ((1u << lane) - 1)); ts[idx] = t; x0s[idx] = x0; }
      __syncwarp();
      for (int e = lane; e < nT * D; e += 32) {               // all loads of the chunk in flight at once
         const int ti = e / D, k = e - ti * D;
         os[ti * ostr + k] = feat[(size_t)ts[ti] * D + k];
      }
      __syncwarp();
      for (int mb = 0; mb < Mn; mb += 32) {
         const int Mc = min(32, Mn - mb);
         // ---- phase 1: lanes <-> (mixture, frame) pairs
         unsigned act = 0;                             // mixtures with any occupancy in this chunk
         for (int pi = lane; pi < ((Mc * nT + 31) & ~31); pi += 32) {
            float Lr = 0.f;
            const int mi = pi / nT, ti = pi - mi * nT;
            if (mi < Mc) {
               const float wt = M.mixLogWt[mo + mb + mi];
               if (wt > LMINMIX_F) {                                         // HFB.c:1573
                  double x = x0s[ti];
                  if (Mn > 1) {
                     const int g = M.mixGauss[mo + mb + mi];
                     const float *mu = M.mean + (size_t)g * Dp, *iv = M.ivar + (size_t)g * Dp;
                     const float *o = os + ti * ostr;
                     float sum = M.gconst[g];
                     int k = 0;
                     for (; k + 4 <= D; k += 4) {                            // rows are 16-byte aligned (Dp % 4 == 0)
                        const float4 m4 = *reinterpret_cast<const float4 *>(mu + k);
                        const float4 v4 = *reinterpret_cast<const float4 *>(iv + k);
                        float d = __fsub_rn(o[k], m4.x);     sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), v4.x));
                        d = __fsub_rn(o[k + 1], m4.y);       sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), v4.y));
                        d = __fsub_rn(o[k + 2], m4.z);       sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), v4.z));
                        d = __fsub_rn(o[k + 3], m4.w);       sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), v4.w));
                     }
                     for (; k < D; k++) {
                        const float d = __fsub_rn(o[k], mu[k]);
                        sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), iv[k]));
                     }
                     const float mixp = -0.5f * sum;
                     x = (x + (double)wt) + (double)mixp;                    // :1581-1599
                  }
                  if (-x < minF) Lr = expf((float)x);                        // :1606, :1612
               }
               lrs[mi * 33 + ti] = Lr;
            }
            act |= __reduce_or_sync(0xffffffffu, (Lr > 0.f) ? (1u << mi) : 0u);
         }
         __syncwarp();
         // ---- phase 2: lanes <-> feature dimensions, centred sums over the chunk's frames
         for (unsigned b = act; b; b &= b - 1) {
            const int mi = __ffs(b) - 1;
            const int g = M.mixGauss[mo + mb + mi];
            const float mu0 = (k0 < D) ? M.mean[(size_t)g * Dp + k0] : 0.f, mu1 = (k1 < D) ? M.mean[(size_t)g * Dp + k1] : 0.f;
            float am0 = 0.f, am1 = 0.f, av0 = 0.f, av1 = 0.f, aocc = 0.f;
            for (int ti = 0; ti < nT; ti++) {
               const float Lr = lrs[mi * 33 + ti];
               float o0, o1;
               if (feat2) {
                  const float *r2 = feat2 + (size_t)ts[ti] * D;
                  o0 = (k0 < D) ? r2[k0] : 0.f; o1 = (k1 < D) ? r2[k1] : 0.f;
               } else { o0 = os[ti * ostr + k0]; o1 = os[ti * ostr + k1]; }
               const float d0 = (k0 < D) ? o0 - mu0 : 0.f, d1 = (k1 < D) ? o1 - mu1 : 0.f;
               const float z0 = d0 * Lr, z1 = d1 * Lr;                       // zmeanlr, :1675
               aocc += Lr; am0 += z0; am1 += z1;
               av0 = fmaf(z0, d0, av0); av1 = fmaf(z1, d1, av1);
            }
            if (upM) {
               double *mu = W.acc + M.L.muSum + (size_t)M.meanId[g] * D;
               if (k0 < D) atomicAdd(&mu[k0], (double)am0);
               if (k1 < D) atomicAdd(&mu[k1], (double)am1);
               if (lane == 0) atomicAdd(&W.acc[M.L.muOcc + M.meanId[g]], (double)aocc);
            }
            if (upV) {
               double *va = W.acc + M.L.vaSum + (size_t)M.varId[g] * D;
               if (k0 < D) atomicAdd(&va[k0], (double)av0);
               if (k1 < D) atomicAdd(&va[k1], (double)av1);
               if (lane == 0) atomicAdd(&W.acc[M.L.vaOcc + M.varId[g]], (double)aocc);
            }
            if (upW && lane == 0) atomicAdd(&W.acc[M.L.wtC + mo + mb + mi], (double)aocc);
            wsum += (double)aocc;
         }
         __syncwarp();
      }
   }
   if (lane == 0 && wsum > 0.0) atomicAdd(&W.acc[M.L.wtOcc + s], wsum);     // :1736
}

// ------------------------------------------------------------------------------------------
// Two-model re-estimation (UseAlignHMMSet, HFB.c:296-333; HERest ALIGNMODELMMF): the wave was aligned with set A
// (output probabilities, beams, alpha, beta, pr -- every kernel above ran on it); the statistics belong to set U.
// One warp per emitting position (utterance, label q, state j), like stats3_kernel.  Per frame inside the alpha beam
// (HFB.c:1518-1547, :1573-1578):
//    comp_prob[m] = wght_m + log N(o_t; U's component m)            float, all M components of U's state
//    norm         = LAdd over m of comp_prob[m]                      rounded to float after every LAdd
//    x            = comp_prob[m] + alpha_j(t) + beta_j(t) - pr - norm   (M = 1: alpha_j + beta_j - pr)
//    kept if wght_m > LMINMIX and -x < minFrwdP; Lr = exp(x) goes into U's accumulators, sums centred on U's means.
// compLevel (HFB: ALIGNCOMPLEVEL = T, :1521-1530): comp_prob / norm are those of A's state at the position instead (same
// number of components, checked when the batch is submitted); what is kept and where it goes stays U's.
// With a second data stream (HERest -r, W.feat2): the alignment saw the first stream; comp_prob is evaluated on the SECOND
// one (:1533-1541) -- on the first one with compLevel (:1534-1535) -- and the sums are the second stream's (:1603-1611).
// Transition statistics do not exist in this mode (HFB.c:313-316); numEgs counts U's physical HMMs (:1768-1772).
// Not a bench path: FP32 CUDA cores, the reference's own evaluation order.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t stats_two_smem_bytes(int D)
{
   return ST_WARPS * (sizeof(double) * 64 + sizeof(float) * ((size_t)32 * stats_ostride(D) + 32 * 33 + 32) + sizeof(int) * 32);
}

__global__ void __launch_bounds__(32 * ST_WARPS)
stats_two_kernel(DevModel A, DevModel U, Wave W, const int *__restrict__ labUp, int compLevel)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const int wInB = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int wg = blockIdx.x * ST_WARPS + wInB;
   if (wg >= W.totalPos) return;
   const int ui = upper_index(W.posPre, W.numUtt, wg);
   if (W.out[ui].status != 0) return;
   const UttDesc u = W.utt[ui];
   const int lp = wg - W.posPre[ui];
   if (lp == 0)                                                               // up_hmm->hook, HFB.c:1768-1772
      for (int q2 = lane; q2 < u.Q; q2 += 32) atomicAdd(&W.acc[U.L.numEgs + labUp[u.labOff + q2]], 1.0);
   const int q = upper_index(W.mPoff + u.modOff, u.Q, lp);
   const int gq = u.modOff + q;
   const int j = lp - W.mPoff[gq];                     // emitting index 0..N-3
   const int tmin = W.mTmin[gq], tmax = W.mTmax[gq];
   const int uf = W.uFlags;
   const bool upM = (uf & HFB_UPMEANS) != 0, upV = (uf & HFB_UPVARS) != 0, upW = (uf & HFB_UPMIXES) != 0;
   if (tmin > tmax || !(upM || upV || upW)) return;
   const int D = U.D, Dp = U.Dp, P = u.P, S = u.S, ostr = stats_ostride(D);
   const size_t perWarp = sizeof(double) * 64 + sizeof(float) * ((size_t)32 * ostr + 32 * 33 + 32) + sizeof(int) * 32;
   unsigned char *mine = smraw + perWarp * wInB;
   double *ajs = (double *)mine, *bjs = ajs + 32;      // [32] alpha_j(t), beta_j(t) of the chunk's frames
   float *os = (float *)(bjs + 32);                    // [32][ostr] observation rows
   float *lrs = os + 32 * ostr;                        // [32 components][33] comp_prob, then Lr
   float *nrm = lrs + 32 * 33;                         // [32] norm per frame
   int *ts = (int *)(nrm + 32);                        // [32] frame numbers

   const int so = W.mSoff[gq];
   const int pU = labUp[u.labOff + q];
   const int s = U.hmmState[U.hmmStateOff[pU] + j];    // the update set's tied state at this position
   const int mo = U.stateMixOff[s], Mn = U.stateMixOff[s + 1] - mo;
   const double *alphaJ = W.occ + u.occOff + lp;
   const double *betaU = W.beta + u.betaOff;
   const short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const float *feat1 = W.feat + (size_t)u.featOff * D;
   const float *feat2 = W.feat2 ? W.feat2 + (size_t)u.featOff * D : nullptr;
   const float *feat = (feat2 && !compLevel) ? feat2 : feat1;      // rows comp_prob is evaluated on (kept in shared memory)
   const float *featS = feat2 ? feat2 : feat1;                     // rows the sums are formed from
   const bool sameRows = featS == feat;
   const double pr = W.out[ui].pr, minF = W.minFrwdP;
   const int k0 = lane, k1 = lane + 32;                // D <= 64 (checked at create)
   double wsum = 0.0;
   // the set whose components give comp_prob / norm: U's state, or (ALIGNCOMPLEVEL) A's state at this position
   const DevModel &C = compLevel ? A : U;
   const int cmo = compLevel ? A.stateMixOff[W.posState[u.posOff + lp]] : mo;
   const int CDp = C.Dp;

   auto comp_prob = [&](int m, const float *o) -> float {          // wght + MOutP, HFB.c:1544-1545 (IDOutP: HModel.c:5420-5431)
      const int g = C.mixGauss[cmo + m];
      const float *mu = C.mean + (size_t)g * CDp, *iv = C.ivar + (size_t)g * CDp;
      float sum = C.gconst[g];
      for (int k = 0; k < D; k++) {
         const float d = __fsub_rn(o[k], mu[k]);
         sum = __fadd_rn(sum, __fmul_rn(__fmul_rn(d, d), iv[k]));
      }
      return __fadd_rn(C.mixLogWt[cmo + m], -0.5f * sum);
   };

   for (int t0 = tmin; t0 <= tmax; t0 += 32) {
      const int t = t0 + lane;
      const bool inb = (t <= tmax) && q >= sqA[t] && q <= eqA[t];
      double aj = 0.0, bj = 0.0;
      bool valid = false;
      if (inb) {
         aj = alphaJ[(size_t)t * P];
         bj = betaU[(size_t)t * S + so + 1 + j];
         valid = !((aj + bj) - pr < -(minF + 0.25));                         // posteriors are <= 1: nothing can pass :1606
      }
      const unsigned mask = __ballot_sync(0xffffffffu, valid);
      const int nT = __popc(mask);
      if (nT == 0) continue;
      if (valid) { int idx = __popc(mask & ((1u << lane) - 1)); ts[idx] = t; ajs[idx] = aj; bjs[idx] = bj; }
      __syncwarp();
      for (int e = lane; e < nT * D; e += 32) {
         const int ti = e / D, k = e - ti * D;
         os[ti * ostr + k] = feat[(size_t)ts[ti] * D + k];
      }
      __syncwarp();
      // ---- norm of every frame: components in order, float after every LAdd
      float norm = (float)LZERO_D;
      if (Mn > 1)
         for (int mb = 0; mb < Mn; mb += 32) {
            const int Mc = min(32, Mn - mb);
            for (int pi = lane; pi < Mc * nT; pi += 32) {
               const int mi = pi / nT, ti = pi - mi * nT;
               lrs[mi * 33 + ti] = comp_prob(mb + mi, os + ti * ostr);
            }
            __syncwarp();
            if (lane < nT)
               for (int mi = 0; mi < Mc; mi++) norm = (float)ladd<true>((double)norm, (double)lrs[mi * 33 + lane]);
            __syncwarp();
         }
      nrm[lane] = norm;
      __syncwarp();
      for (int mb = 0; mb < Mn; mb += 32) {
         const int Mc = min(32, Mn - mb);
         // ---- phase 1: lanes <-> (component, frame) pairs
         unsigned act = 0;
         for (int pi = lane; pi < ((Mc * nT + 31) & ~31); pi += 32) {
            float Lr = 0.f;
            const int mi = pi / nT, ti = pi - mi * nT;
            if (mi < Mc) {
               const float wt = U.mixLogWt[mo + mb + mi];
               if (wt > LMINMIX_F) {                                         // HFB.c:1573
                  double x;
                  if (Mn == 1) x = (ajs[ti] + bjs[ti]) - pr;                 // :1575-1576
                  else {
                     const float cp = (Mn <= 32) ? lrs[mi * 33 + ti] : comp_prob(mb + mi, os + ti * ostr);
                     x = ((((double)cp + ajs[ti]) + bjs[ti]) - pr) - (double)nrm[ti];   // :1578
                  }
                  if (-x < minF) Lr = (float)exp(x);                         // :1606, :1612
               }
            }
            if (mi < Mc) lrs[mi * 33 + ti] = Lr;                             // the lane that read this comp_prob overwrites it
            act |= __reduce_or_sync(0xffffffffu, (Lr > 0.f) ? (1u << mi) : 0u);
         }
         __syncwarp();
         // ---- phase 2: lanes <-> feature dimensions, sums centred on the UPDATE set's means
         for (unsigned b = act; b; b &= b - 1) {
            const int mi = __ffs(b) - 1;
            const int g = U.mixGauss[mo + mb + mi];
            const float mu0 = (k0 < D) ? U.mean[(size_t)g * Dp + k0] : 0.f, mu1 = (k1 < D) ? U.mean[(size_t)g * Dp + k1] : 0.f;
            float am0 = 0.f, am1 = 0.f, av0 = 0.f, av1 = 0.f, aocc = 0.f;
            for (int ti = 0; ti < nT; ti++) {
               const float Lr = lrs[mi * 33 + ti];
               float o0, o1;
               if (sameRows) { o0 = os[ti * ostr + k0]; o1 = os[ti * ostr + k1]; }
               else {
                  const float *r2 = featS + (size_t)ts[ti] * D;
                  o0 = (k0 < D) ? r2[k0] : 0.f; o1 = (k1 < D) ? r2[k1] : 0.f;
               }
               const float d0 = (k0 < D) ? o0 - mu0 : 0.f, d1 = (k1 < D) ? o1 - mu1 : 0.f;
               const float z0 = d0 * Lr, z1 = d1 * Lr;                       // zmeanlr, :1675
               aocc += Lr; am0 += z0; am1 += z1;
               av0 = fmaf(z0, d0, av0); av1 = fmaf(z1, d1, av1);
            }
            if (upM) {
               double *mu = W.acc + U.L.muSum + (size_t)U.meanId[g] * D;
               if (k0 < D) atomicAdd(&mu[k0], (double)am0);
               if (k1 < D) atomicAdd(&mu[k1], (double)am1);
               if (lane == 0) atomicAdd(&W.acc[U.L.muOcc + U.meanId[g]], (double)aocc);
            }
            if (upV) {
               double *va = W.acc + U.L.vaSum + (size_t)U.varId[g] * D;
               if (k0 < D) atomicAdd(&va[k0], (double)av0);
               if (k1 < D) atomicAdd(&va[k1], (double)av1);
               if (lane == 0) atomicAdd(&W.acc[U.L.vaOcc + U.varId[g]], (double)aocc);
            }
            if (upW && lane == 0) atomicAdd(&W.acc[U.L.wtC + mo + mb + mi], (double)aocc);
            wsum += (double)aocc;
         }
         __syncwarp();
      }
   }
   if (lane == 0 && wsum > 0.0) atomicAdd(&W.acc[U.L.wtOcc + s], wsum);     // :1736
}
