// hfb_l2r.cuh -- recursion kernels specialised for HTK's standard topology: every HMM of the set
// has 5 states (3 emitting) and the only transitions are entry->2, i->i, i->i+1 and 4->exit
// (what MakeProtoHMMSet / HInit produce and what all BASELINE configs use; checked once per
// model set in hfbgpu_create, see model_is_l2r()).  No tee models, no skips, so
//
//   * SetBeta   (HFB.c:1205-1277) is three log-adds per (model, frame):
//        beta_4(t) = (a_4N + beta_N(t))          (+) (a_44 + u_4)        u_j = b_j(o_t+1) + beta_j(t+1)
//        beta_3(t) = (a_33 + u_3)                (+) (a_34 + u_4)
//        beta_2(t) = (a_22 + u_2)                (+) (a_23 + u_3)
//        beta_1(t) =  a_12 + b_2(o_t) + beta_2(t)
//   * StepAlpha (HFB.c:729-771) likewise:
//        alpha_2(t) = [(a_12 + alpha_1(t)) (+) (alpha_2(t-1) + a_22)] + b_2(o_t)   etc.
//
// The transition logs live in registers as doubles (no per-frame conversions), the structural
// "log zero" tests of the generic code disappear at compile time, and the guarded
// `if (y > LSMALL) x = LAdd(x, y)` of the reference becomes an unconditional branch-free log-add:
// with one operand at log zero the other is returned exactly, with both the result stays below
// LSMALL, i.e. log zero for every later test.  Same numbers as beta_fast/alpha_fast_kernel for
// every value above LSMALL; ~2.5x fewer instructions per frame.
#pragma once
#include "hfb_fast.cuh"

struct L2RRegs {
   double aE, a00, a01, a11, a12, a22, a2x;    // entry->0, i->i, i->i+1, 2->exit (emitting states 0..2)
   int s0, s1, s2;                             // output-probability slots of the emitting states
};

__device__ __forceinline__ void load_l2r(L2RRegs &r, const DevModel &M, const Wave &W, const UttDesc &u, int q)
{
   const float *A = M.transLogA + W.mTrans[u.modOff + q];      // 5 x 5, row-major
   r.aE = (double)A[1];
   r.a00 = (double)A[6];  r.a01 = (double)A[7];
   r.a11 = (double)A[12]; r.a12 = (double)A[13];
   r.a22 = (double)A[18]; r.a2x = (double)A[19];
   const int *ps = W.posSlot + u.posOff + 3 * q;
   r.s0 = ps[0]; r.s1 = ps[1]; r.s2 = ps[2];
}

__device__ __forceinline__ double dmax(double a, double b) { return (a > b) ? a : b; }

// ------------------------------------------------------------------------------------------
// K2 (standard topology): one CTA per utterance, one thread per model
// ------------------------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT, MAXT == 128 ? 7 : 1024 / MAXT) beta_l2r_kernel(DevModel M, Wave W, int onlyRedo)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc &u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   // onlyRedo: second launch after beta_l2r_slide_kernel, for the utterances whose beam outgrew its window
   if (onlyRedo ? (out->status != HFB_UTT_BETAWIDE) : (out->status != 0)) {
      if (!onlyRedo && threadIdx.x == 0 && out->status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
      return;
   }
   __syncthreads();                                     // everybody has read the status before thread 0 rewrites it
   const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
   const int T = u.T, Q = u.Q, J = u.J;
   const size_t S = (size_t)5 * Q;
   double *entA = (double *)smraw, *entB = entA + (Q + 2), *wred = entB + (Q + 2);
   int *wlo = (int *)(wred + 32), *whi = wlo + 32;
   const int q = tid;
   const bool mine = q < Q;
   L2RRegs r;
   if (mine) load_l2r(r, M, W, u, q);
   else { r.aE = r.a00 = r.a01 = r.a11 = r.a12 = r.a22 = r.a2x = LZERO_D; r.s0 = r.s1 = r.s2 = 0; }
   const float *bU = W.b + u.bOff;
   double *betaQ = W.beta + u.betaOff + 5 * q;
   short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;

   double thresh = W.pruneInit, pr = LZERO_D;
   int retries = 0, status = 0;

   for (;;) {
      // ---- SetBeamTaper (HFB.c:1116-1145): every model has minimum duration 3 here, so the taper is
      //      qHi(t) = min(Q-1, t / 3), qLo(t) = max(0, Q-1 - (T-1-t) / 3) -- evaluated in the frame loop
      //      with running quotients, no table and no loads
      const bool noPrune = thresh >= 0.5 * HFB_NOPRUNE;

      double *cur = entA, *prev = entB;                // entry-state beta of every model, frames t / t+1
      double u0 = LZERO_D, u1 = LZERO_D, u2 = LZERO_D; // b_j(o_{t+1}) + beta_j(t+1), log zero outside the beam

      // ---- t = T-1, HFB.c:1176-1198
      int lo1 = Q - 1, hi1 = Q - 1, lastq = lo1;
      if (tid == 0) { qHi[T - 1] = (short)(Q - 1); qLo[T - 1] = (short)(Q - 1); }
      if (mine && q >= lo1) {
         const float *bt = bU + (size_t)(T - 1) * J;
         const double bExit = (q == Q - 1) ? 0.0 : LZERO_D;
         const double n0 = LZERO_D + bExit, n1 = LZERO_D + bExit, n2 = r.a2x + bExit;
         u0 = (double)bt[r.s0] + n0; u1 = (double)bt[r.s1] + n1; u2 = (double)bt[r.s2] + n2;
         const double x = (n0 > LSMALL_D) ? r.aE + u0 : LZERO_D;
         double *bg = betaQ + (size_t)(T - 1) * S;
         bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = bExit;
         cur[q] = x;
      }
      // output probabilities travel two frames ahead of their use, in registers
      float bA0 = 0.f, bA1 = 0.f, bA2 = 0.f, bB0 = 0.f, bB1 = 0.f, bB2 = 0.f;
      if (mine) {
         if (T >= 2) { const float *b2 = bU + (size_t)(T - 2) * J; bA0 = b2[r.s0]; bA1 = b2[r.s1]; bA2 = b2[r.s2]; }
         if (T >= 3) { const float *b2 = bU + (size_t)(T - 3) * J; bB0 = b2[r.s0]; bB1 = b2[r.s1]; bB2 = b2[r.s2]; }
      }
      __syncthreads();
      { double *tmp = cur; cur = prev; prev = tmp; }

      // ---- t = T-2 .. 0, HFB.c:1205-1277
      bool fail = false;
      double *bg = betaQ + (size_t)(T - 1) * S;        // running pointers: this model's beta column at t,
      const float *bp = bU + (size_t)(T - 3) * J;      // the output-probability row of frame t-2
      int hiC = (T - 1) / 3, hiR = (T - 1) % 3, loC = 0, loR = 0;      // t / 3 and (T-1-t) / 3 with remainders, at t = T-1
      for (int t = T - 2; t >= 0; t--) {
         bg -= S; bp -= J;
         if (hiR == 0) { hiR = 2; hiC--; } else hiR--;
         if (loR == 2) { loR = 0; loC++; } else loR++;
         const int tapLo = max(0, Q - 1 - loC), tapHi = min(Q - 1, hiC);
         const int startq = hi1;
         const int endq = (lo1 == 0) ? 0 : ((tapLo >= lo1) ? tapLo : lo1 - 1);
         lastq = endq;
         const bool active = mine && q >= endq && q <= startq;
         const float c0 = bA0, c1 = bA1, c2 = bA2;
         bA0 = bB0; bA1 = bB1; bA2 = bB2;
         if (t >= 2 && mine && q >= endq - 2 && q <= startq) { bB0 = bp[r.s0]; bB1 = bp[r.s1]; bB2 = bp[r.s2]; }
         double lMax = LZERO_D, un0 = LZERO_D, un1 = LZERO_D, un2 = LZERO_D;
         if (active) {
            const double ex = (q + 1 >= lo1 && q + 1 <= hi1) ? prev[q + 1] : LZERO_D;      // :1225
            const double n2 = ladd_nz_b(r.a2x + ex, r.a22 + u2);                             // :1228-1236
            const double n1 = ladd_nz_b(r.a11 + u1, r.a12 + u2);
            const double n0 = ladd_nz_b(r.a00 + u0, r.a01 + u1);
            un0 = (double)c0 + n0; un1 = (double)c1 + n1; un2 = (double)c2 + n2;
            const double x = r.aE + un0;                                                   // :1242-1250
            bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = ex;
            cur[q] = x;
            lMax = dmax(dmax(n0, n1), n2);
         }
         int nhi, nlo;
         if (noPrune) {
            // gMax - maxP[q] > thresh is never true for finite log values: the beam is the candidate
            // range and only the entry values have to become visible to the neighbours
            nhi = startq; nlo = endq;
            __syncthreads();
         } else {
            double gMax = warp_max(lMax);
            if (lane == 0) wred[wid] = gMax;
            __syncthreads();
            gMax = LZERO_D;
            for (int w = 0; w < nw; w++) gMax = dmax(gMax, wred[w]);
            // ---- pruning (:1254-1272)
            const bool keep = active && !(gMax - lMax > thresh);
            const int myHi = __reduce_max_sync(0xffffffffu, keep ? q : -1);
            const int myLo = __reduce_min_sync(0xffffffffu, keep ? q : 0x7fffffff);
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
         const bool inNew = active && q >= nlo && q <= nhi;
         u0 = inNew ? un0 : LZERO_D; u1 = inNew ? un1 : LZERO_D; u2 = inNew ? un2 : LZERO_D;
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
// K2, transcriptions of up to 128 labels: ONE WARP per utterance, four models per lane
// ------------------------------------------------------------------------------------------
// The one-thread-per-model kernel above spends its time waiting, not issuing: a barrier per frame (22 % of the stall
// samples), and per thread one short dependent chain with nothing to overlap it (issue slots 40 % used).  Here lane L
// owns the models q = L + 32 k, k = 0..3: their twelve log-adds per frame are independent (instruction-level
// parallelism instead of warps), the neighbour's entry beta comes by shuffle (model q + 1 sits in lane L + 1, for
// lane 31 in lane 0 one row up), the beam reductions are shuffles / redux -- no shared memory, no barrier -- and the
// uniform beam-taper arithmetic is done once per four models.  Same arithmetic, same order, same results as
// beta_l2r_kernel (tests: HFBGPU_NO_BETA_WARP).  (The 8-byte stores of the five betas of a model touch 32 sectors per
// instruction and cost 0.26 of the 1.19 ms; staging them through shared memory for 256-byte contiguous stores cost more
// than it saved: 1.69 ms.)
#define BW_NM 4
// -DBW_RING_MINB=9 compiles the ring-window variant for 9 CTAs per SM = 165 registers without spills (left alone the
// compiler takes 224); part of the "small waves" experiment in launch_wave (hfbgpu.cu), 2.5 % slower on config #5: off.
#ifndef BW_RING_MINB
#define BW_RING_MINB 1
#endif
// RING: transcriptions of any length under a beam (config #5: 667 labels, beam ~40 models).  The 128 (lane, slot) pairs
// form a ring: pair (L, k) holds the model q = L + 32 k (mod 128) nearest below the beam's upper end and takes the one
// 128 below when its model has left the beam for good (the beta beam only ever moves towards the start of the
// transcription).  No shared memory, no barrier -- the 256-thread sliding block kernel (beta_l2r_slide_kernel) needs two
// barriers per frame.  A beam wider than 124 models flags the utterance HFB_UTT_BETAWIDE and beta_l2r_kernel<1024> redoes it.
// (245 registers = 8 warps per SM.  Capped at 168 registers -- __launch_bounds__(32, 12), 300 bytes of spills -- for 12 warps
// per SM the pass is slower at every wave size: 1.44 against 1.13 ms per 1 024 000 frames at 1184 utterances, 1.57 against
// 1.38 at 1776; the spills sit on the dependent chain.)
template <bool ALUCVT, bool RING>
__global__ void __launch_bounds__(32, RING ? BW_RING_MINB : 1) beta_l2r_warp_kernel(DevModel M, Wave W)
{
   const UttDesc &u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   const int lane = threadIdx.x;
   if (out->status != 0) {
      if (lane == 0 && out->status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
      return;
   }
   __syncwarp();
   const int T = u.T, Q = u.Q, J = u.J;
   const size_t S = (size_t)5 * Q;
   L2RRegs r[BW_NM];
   bool mine[BW_NM];
   int qk[BW_NM];                                      // model held by slot k (RING: changes as the window slides)
   auto place = [&]() {                                // window at the top of the transcription
#pragma unroll
      for (int k = 0; k < BW_NM; k++) {
         const int res = lane + 32 * k;
         qk[k] = RING ? ((res <= Q - 1) ? res + 128 * ((Q - 1 - res) / 128) : -1) : res;
         mine[k] = qk[k] >= 0 && qk[k] < Q;
         if (mine[k]) load_l2r(r[k], M, W, u, qk[k]);
         else { r[k].aE = r[k].a00 = r[k].a01 = r[k].a11 = r[k].a12 = r[k].a22 = r[k].a2x = LZERO_D; r[k].s0 = r[k].s1 = r[k].s2 = 0; }
      }
   };
   if (!RING) place();
   const float *bU = W.b + u.bOff;
   double *betaL = W.beta + u.betaOff;
   short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;

   double thresh = W.pruneInit, pr = LZERO_D;
   int retries = 0, status = 0;

   for (;;) {
      const bool noPrune = thresh >= 0.5 * HFB_NOPRUNE;
      if (RING) place();                               // every retry starts again at the top
      double u0[BW_NM], u1[BW_NM], u2[BW_NM], xPrev[BW_NM];   // b_j(o_{t+1}) + beta_j(t+1); entry beta at t+1
      float bA0[BW_NM], bA1[BW_NM], bA2[BW_NM], bB0[BW_NM], bB1[BW_NM], bB2[BW_NM];
      // ---- t = T-1, HFB.c:1176-1198
      int lo1 = Q - 1, hi1 = Q - 1, lastq = lo1;
      if (lane == 0) { qHi[T - 1] = (short)(Q - 1); qLo[T - 1] = (short)(Q - 1); }
#pragma unroll
      for (int k = 0; k < BW_NM; k++) {
         const int q = qk[k];
         u0[k] = u1[k] = u2[k] = xPrev[k] = LZERO_D;
         bA0[k] = bA1[k] = bA2[k] = bB0[k] = bB1[k] = bB2[k] = 0.f;
         if (mine[k] && q >= lo1) {
            const float *bt = bU + (size_t)(T - 1) * J;
            const double bExit = (q == Q - 1) ? 0.0 : LZERO_D;
            const double n0 = LZERO_D + bExit, n1 = LZERO_D + bExit, n2 = r[k].a2x + bExit;
            u0[k] = (double)bt[r[k].s0] + n0; u1[k] = (double)bt[r[k].s1] + n1; u2[k] = (double)bt[r[k].s2] + n2;
            const double x = (n0 > LSMALL_D) ? r[k].aE + u0[k] : LZERO_D;
            double *bg = betaL + (size_t)(T - 1) * S + 5 * q;
            bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = bExit;
            xPrev[k] = x;
         }
         // output probabilities travel two frames ahead of their use, in registers
         if (mine[k]) {
            if (T >= 2) { const float *b2 = bU + (size_t)(T - 2) * J; bA0[k] = b2[r[k].s0]; bA1[k] = b2[r[k].s1]; bA2[k] = b2[r[k].s2]; }
            if (T >= 3) { const float *b2 = bU + (size_t)(T - 3) * J; bB0[k] = b2[r[k].s0]; bB1[k] = b2[r[k].s1]; bB2[k] = b2[r[k].s2]; }
         }
      }

      // ---- t = T-2 .. 0, HFB.c:1205-1277
      bool fail = false;
      double *bgT = betaL + (size_t)(T - 1) * S;       // beta column block of frame t
      const float *bp = bU + (size_t)(T - 3) * J;      // the output-probability row of frame t-2
      int hiC = (T - 1) / 3, hiR = (T - 1) % 3, loC = 0, loR = 0;      // closed-form taper, see beta_l2r_kernel
      for (int t = T - 2; t >= 0; t--) {
         bgT -= S; bp -= J;
         if (hiR == 0) { hiR = 2; hiC--; } else hiR--;
         if (loR == 2) { loR = 0; loC++; } else loR++;
         const int tapLo = max(0, Q - 1 - loC), tapHi = min(Q - 1, hiC);
         const int startq = hi1;
         const int endq = (lo1 == 0) ? 0 : ((tapLo >= lo1) ? tapLo : lo1 - 1);
         lastq = endq;
         double un0[BW_NM], un1[BW_NM], un2[BW_NM], xNew[BW_NM], lMax[BW_NM];
         bool active[BW_NM];
#pragma unroll
         for (int k = 0; k < BW_NM; k++) {
            const int q = qk[k];
            // entry beta of model q + 1 at t + 1: lane + 1 same slot, lane 31 -> lane 0 one slot up (RING: the slots wrap)
            const double y = (lane == 0) ? (RING ? xPrev[(k + 1) % BW_NM] : ((k + 1 < BW_NM) ? xPrev[k + 1] : LZERO_D)) : xPrev[k];
            const double exN = __shfl_sync(0xffffffffu, y, (lane + 1) & 31);
            active[k] = mine[k] && q >= endq && q <= startq;
            const float c0 = bA0[k], c1 = bA1[k], c2 = bA2[k];
            bA0[k] = bB0[k]; bA1[k] = bB1[k]; bA2[k] = bB2[k];
            if (t >= 2 && mine[k] && q >= endq - 2 && q <= startq) { bB0[k] = bp[r[k].s0]; bB1[k] = bp[r[k].s1]; bB2[k] = bp[r[k].s2]; }
            lMax[k] = LZERO_D; un0[k] = un1[k] = un2[k] = xNew[k] = LZERO_D;
            if (active[k]) {
               const double ex = (q + 1 >= lo1 && q + 1 <= hi1) ? exN : LZERO_D;            // :1225
               const double n2 = (ALUCVT ? ladd_nz : ladd_nz_b)(r[k].a2x + ex, r[k].a22 + u2[k]);               // :1228-1236
               const double n1 = (ALUCVT ? ladd_nz : ladd_nz_b)(r[k].a11 + u1[k], r[k].a12 + u2[k]);
               const double n0 = (ALUCVT ? ladd_nz : ladd_nz_b)(r[k].a00 + u0[k], r[k].a01 + u1[k]);
               un0[k] = (double)c0 + n0; un1[k] = (double)c1 + n1; un2[k] = (double)c2 + n2;
               const double x = r[k].aE + un0[k];                                          // :1242-1250
               double *bg = bgT + 5 * q;
               bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = ex;
               xNew[k] = x;
               lMax[k] = dmax(dmax(n0, n1), n2);
            }
         }
         int nhi, nlo;
         if (noPrune) { nhi = startq; nlo = endq; }
         else {
            double gMax = dmax(dmax(lMax[0], lMax[1]), dmax(lMax[2], lMax[3]));
            gMax = warp_max(gMax);
            // ---- pruning (:1254-1272)
            int myHi = -1, myLo = 0x7fffffff;
#pragma unroll
            for (int k = 0; k < BW_NM; k++) {
               const bool keep = active[k] && !(gMax - lMax[k] > thresh);
               if (keep) { myHi = max(myHi, qk[k]); myLo = min(myLo, qk[k]); }
            }
            nhi = __reduce_max_sync(0xffffffffu, myHi);
            nlo = __reduce_min_sync(0xffffffffu, myLo);
         }
         if (nhi < 0) { fail = true; if (endq == 0) status = HFB_UTT_EBETA; break; }
         if (nhi > tapHi) nhi = tapHi;
         if (nlo > nhi) { fail = true; break; }
         if (lane == 0) { qHi[t] = (short)nhi; qLo[t] = (short)nlo; }
         hi1 = nhi; lo1 = nlo;
#pragma unroll
         for (int k = 0; k < BW_NM; k++) {
            const int q = qk[k];
            const bool inNew = active[k] && q >= nlo && q <= nhi;
            u0[k] = inNew ? un0[k] : LZERO_D; u1[k] = inNew ? un1[k] : LZERO_D; u2[k] = inNew ? un2[k] : LZERO_D;
            xPrev[k] = xNew[k];
         }
         if (RING) {
            // everything the next frame can touch is [lo1 - 3, hi1] (beam, one model of growth, two of look-ahead)
            if (hi1 - lo1 + 4 > 32 * BW_NM) { fail = true; status = HFB_UTT_BETAWIDE; break; }
#pragma unroll
            for (int k = 0; k < BW_NM; k++)
               if (mine[k] && qk[k] > hi1) {           // my model has left the beam for good: take the one 128 below
                  do qk[k] -= 32 * BW_NM; while (qk[k] > hi1);
                  mine[k] = qk[k] >= 0;
                  u0[k] = u1[k] = u2[k] = xPrev[k] = LZERO_D;
                  bA0[k] = bA1[k] = bA2[k] = bB0[k] = bB1[k] = bB2[k] = 0.f;
                  if (mine[k]) {
                     load_l2r(r[k], M, W, u, qk[k]);
                     if (t >= 1) { const float *b1 = bp + J; bA0[k] = b1[r[k].s0]; bA1[k] = b1[r[k].s1]; bA2[k] = b1[r[k].s2]; }
                     if (t >= 2) { bB0[k] = bp[r[k].s0]; bB1[k] = bp[r[k].s1]; bB2[k] = bp[r[k].s2]; }
                  }
               }
         }
      }
      if (status != 0) break;
      if (!fail) {
         // utt->pr = bqt[1] (:1280): entry beta of model lastq at the last frame processed
         const int kq = (lastq >> 5) % BW_NM;
         const double v = (kq == 0) ? xPrev[0] : (kq == 1) ? xPrev[1] : (kq == 2) ? xPrev[2] : xPrev[3];
         pr = __shfl_sync(0xffffffffu, v, lastq & 31);
         if (pr > LSMALL_D) break;
      }
      thresh += W.pruneInc;                            // StepBack retry (:1349-1361)
      if (thresh > W.pruneLim || W.pruneInc == 0.0) { status = HFB_UTT_SKIPPED; break; }
      retries++;
   }
   if (lane == 0) {
      out->status = status; out->retries = retries; out->pr = (status == 0) ? pr : LZERO_D;
      out->thresh = thresh;
      if (status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
   }
}

// Long transcriptions with beam pruning (config #5: 667 labels, beam ~40 models): the same recursion with a SLIDING
// window of blockDim models instead of one thread per label -- thread i owns the model q = i (mod blockDim) nearest
// below the beam's upper end and moves blockDim models down when its model leaves the beam for good (the beta beam
// only ever moves towards the start of the transcription).  Eight warps per barrier instead of 32.  If a beam ever
// needs more than blockDim - 3 models the utterance is flagged HFB_UTT_BETAWIDE and beta_l2r_kernel<1024> redoes it.
__global__ void __launch_bounds__(256, 4) beta_l2r_slide_kernel(DevModel M, Wave W)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   const UttDesc &u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) {
      if (threadIdx.x == 0 && out->status == HFB_UTT_SKIPPED) atomicAdd(&W.acc[M.L.numSkipped], 1.0);
      return;
   }
   const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, wid = tid >> 5, nw = nt >> 5;
   const int T = u.T, Q = u.Q, J = u.J;
   const size_t S = (size_t)5 * Q;
   double *entA = (double *)smraw, *entB = entA + (Q + 2), *wred = entB + (Q + 2);
   int *wlo = (int *)(wred + 32), *whi = wlo + 32;
   int q = -1;
   bool mine = false;
   L2RRegs r;
   const float *bU = W.b + u.bOff;
   double *betaU = W.beta + u.betaOff;
   short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;

   double thresh = W.pruneInit, pr = LZERO_D;
   int retries = 0, status = 0;

   for (;;) {
      // ---- SetBeamTaper (HFB.c:1116-1145): every model has minimum duration 3 here, so the taper is
      //      qHi(t) = min(Q-1, t / 3), qLo(t) = max(0, Q-1 - (T-1-t) / 3) -- evaluated in the frame loop
      //      with running quotients, no table and no loads
      const bool noPrune = thresh >= 0.5 * HFB_NOPRUNE;
      // window at the top of the transcription: the largest q <= Q-1 with q = tid (mod nt)
      q = (tid <= Q - 1) ? tid + nt * ((Q - 1 - tid) / nt) : -1;
      mine = q >= 0;
      if (mine) load_l2r(r, M, W, u, q);
      else { r.aE = r.a00 = r.a01 = r.a11 = r.a12 = r.a22 = r.a2x = LZERO_D; r.s0 = r.s1 = r.s2 = 0; }

      double *cur = entA, *prev = entB;                // entry-state beta of every model, frames t / t+1
      double u0 = LZERO_D, u1 = LZERO_D, u2 = LZERO_D; // b_j(o_{t+1}) + beta_j(t+1), log zero outside the beam

      // ---- t = T-1, HFB.c:1176-1198
      int lo1 = Q - 1, hi1 = Q - 1, lastq = lo1;
      if (tid == 0) { qHi[T - 1] = (short)(Q - 1); qLo[T - 1] = (short)(Q - 1); }
      if (mine && q >= lo1) {
         const float *bt = bU + (size_t)(T - 1) * J;
         const double bExit = (q == Q - 1) ? 0.0 : LZERO_D;
         const double n0 = LZERO_D + bExit, n1 = LZERO_D + bExit, n2 = r.a2x + bExit;
         u0 = (double)bt[r.s0] + n0; u1 = (double)bt[r.s1] + n1; u2 = (double)bt[r.s2] + n2;
         const double x = (n0 > LSMALL_D) ? r.aE + u0 : LZERO_D;
         double *bg = betaU + (size_t)(T - 1) * S + 5 * q;
         bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = bExit;
         cur[q] = x;
      }
      // output probabilities travel two frames ahead of their use, in registers
      float bA0 = 0.f, bA1 = 0.f, bA2 = 0.f, bB0 = 0.f, bB1 = 0.f, bB2 = 0.f;
      if (mine) {
         if (T >= 2) { const float *b2 = bU + (size_t)(T - 2) * J; bA0 = b2[r.s0]; bA1 = b2[r.s1]; bA2 = b2[r.s2]; }
         if (T >= 3) { const float *b2 = bU + (size_t)(T - 3) * J; bB0 = b2[r.s0]; bB1 = b2[r.s1]; bB2 = b2[r.s2]; }
      }
      __syncthreads();
      { double *tmp = cur; cur = prev; prev = tmp; }

      // ---- t = T-2 .. 0, HFB.c:1205-1277
      bool fail = false;
      double *bgRow = betaU + (size_t)(T - 1) * S;     // running pointers: the beta column block of frame t,
      const float *bp = bU + (size_t)(T - 3) * J;      // the output-probability row of frame t-2
      int hiC = (T - 1) / 3, hiR = (T - 1) % 3, loC = 0, loR = 0;      // t / 3 and (T-1-t) / 3 with remainders, at t = T-1
      for (int t = T - 2; t >= 0; t--) {
         bgRow -= S; bp -= J;
         if (hiR == 0) { hiR = 2; hiC--; } else hiR--;
         if (loR == 2) { loR = 0; loC++; } else loR++;
         const int tapLo = max(0, Q - 1 - loC), tapHi = min(Q - 1, hiC);
         const int startq = hi1;
         const int endq = (lo1 == 0) ? 0 : ((tapLo >= lo1) ? tapLo : lo1 - 1);
         lastq = endq;
         const bool active = mine && q >= endq && q <= startq;
         const float c0 = bA0, c1 = bA1, c2 = bA2;
         bA0 = bB0; bA1 = bB1; bA2 = bB2;
         if (t >= 2 && mine && q >= endq - 2 && q <= startq) { bB0 = bp[r.s0]; bB1 = bp[r.s1]; bB2 = bp[r.s2]; }
         double lMax = LZERO_D, un0 = LZERO_D, un1 = LZERO_D, un2 = LZERO_D;
         if (active) {
            const double ex = (q + 1 >= lo1 && q + 1 <= hi1) ? prev[q + 1] : LZERO_D;      // :1225
            const double n2 = ladd_nz_b(r.a2x + ex, r.a22 + u2);                             // :1228-1236
            const double n1 = ladd_nz_b(r.a11 + u1, r.a12 + u2);
            const double n0 = ladd_nz_b(r.a00 + u0, r.a01 + u1);
            un0 = (double)c0 + n0; un1 = (double)c1 + n1; un2 = (double)c2 + n2;
            const double x = r.aE + un0;                                                   // :1242-1250
            double *bg = bgRow + 5 * q;
            bg[0] = x; bg[1] = n0; bg[2] = n1; bg[3] = n2; bg[4] = ex;
            cur[q] = x;
            lMax = dmax(dmax(n0, n1), n2);
         }
         int nhi, nlo;
         if (noPrune) {
            // gMax - maxP[q] > thresh is never true for finite log values: the beam is the candidate
            // range and only the entry values have to become visible to the neighbours
            nhi = startq; nlo = endq;
            __syncthreads();
         } else {
            double gMax = warp_max(lMax);
            if (lane == 0) wred[wid] = gMax;
            __syncthreads();
            gMax = LZERO_D;
            for (int w = 0; w < nw; w++) gMax = dmax(gMax, wred[w]);
            // ---- pruning (:1254-1272)
            const bool keep = active && !(gMax - lMax > thresh);
            const int myHi = __reduce_max_sync(0xffffffffu, keep ? q : -1);
            const int myLo = __reduce_min_sync(0xffffffffu, keep ? q : 0x7fffffff);
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
         const bool inNew = active && q >= nlo && q <= nhi;
         u0 = inNew ? un0 : LZERO_D; u1 = inNew ? un1 : LZERO_D; u2 = inNew ? un2 : LZERO_D;
         { double *tmp = cur; cur = prev; prev = tmp; }
         // ---- the window: everything the next frame can touch is [lo1 - 3, hi1] (beam, one model of growth, two of
         //      output-probability look-ahead)
         if (hi1 - lo1 + 4 > nt) { fail = true; status = HFB_UTT_BETAWIDE; break; }
         if (mine && q > hi1) {                            // my model has left the beam for good: take the one nt below
            do q -= nt; while (q > hi1);
            mine = q >= 0;
            u0 = u1 = u2 = LZERO_D;
            if (mine) {
               load_l2r(r, M, W, u, q);
               if (t >= 1) { const float *b1 = bp + J; bA0 = b1[r.s0]; bA1 = b1[r.s1]; bA2 = b1[r.s2]; }
               if (t >= 2) { bB0 = bp[r.s0]; bB1 = bp[r.s1]; bB2 = bp[r.s2]; }
            }
         }
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
// K3 (standard topology): one WARP per utterance, lane L owns the model q = L (mod 32) of the
// sliding window that starts at the alpha beam's lower end.  Without tee models StepAlpha can
// reach [sq(t-1), eq(t-1)+1] only.  Writes alpha, the beams and first/last active frame per
// model; gives up (out->redo) if a beam ever needs more than 32 models.
// (Tried: two utterances per warp in lock step, for the instruction-level parallelism that pays in
// beta_l2r_warp_kernel.  1.21 ms instead of 0.78: the per-utterance branches (window, beam, slide) keep the two chains
// in separate basic blocks, so they run one after the other; it would need fully predicated code.)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) alpha_l2r_kernel(DevModel M, Wave W, int forceRedo)
{
   const UttDesc &u = W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) return;
   if (forceRedo) { if (threadIdx.x == 0) out->redo = 1; return; }
   const unsigned FULL = 0xffffffffu;
   const int lane = threadIdx.x;
   const int T = u.T, Q = u.Q, J = u.J;
   const size_t S = (size_t)5 * Q, P = (size_t)3 * Q;
   const float *bU = W.b + u.bOff;
   const double *betaU = W.beta + u.betaOff;
   double *occU = W.occ + u.occOff, *aentU = W.aent + u.aentOff;
   const short *qLo = W.qLo + u.frameBase, *qHi = W.qHi + u.frameBase;
   short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   const double pr = out->pr, minF = W.minFrwdP;
   int *gTmin = W.mTmin + u.modOff, *gTmax = W.mTmax + u.modOff;
   for (int q = lane; q < Q; q += 32) { gTmin[q] = 0x7fffffff; gTmax[q] = -1; }

   L2RRegs r;
   int myq = lane;
   bool have = myq < Q;
   if (have) load_l2r(r, M, W, u, myq);
   else { r.aE = r.a00 = r.a01 = r.a11 = r.a12 = r.a22 = r.a2x = LZERO_D; r.s0 = r.s1 = r.s2 = 0; }
   double e0 = LZERO_D, e1 = LZERO_D, e2 = LZERO_D;     // alpha of the emitting states at t-1
   double aEx = LZERO_D, mpS = LZERO_D, exv = LZERO_D;  // alpha_N(t-1); max_i alpha_i+beta_i; alpha_N+beta_N
   int tmin = 0x7fffffff, tmax = -1;
   int sq = 0, eq = 0;

   // running pointers: frame t of the beta column block, the output-probability row, the alpha rows
   const double *bqT = betaU;
   const float *btT = bU;
   double *ocT = occU, *aeT = aentU;
   int loP = 0, hiP = 0;                                   // beta beam of the previous frame
   int loT = qLo[0], hiT = qHi[0];
   for (int t = 0; t < T; t++) {
      // loads that do not depend on the recursion go first
      const bool inWin = have && myq >= sq && myq <= ((t == 0) ? hiT : min(Q - 1, eq + 1));
      float b0 = 0.f, b1 = 0.f, b2 = 0.f;
      double bEn = LZERO_D, bE0 = LZERO_D, bE1 = LZERO_D, bE2 = LZERO_D, bX = LZERO_D;
      if (inWin && myq >= loT && myq <= hiT) {
         const double *bq = bqT + 5 * myq;
         bEn = bq[0]; bE0 = bq[1]; bE1 = bq[2]; bE2 = bq[3]; bX = bq[4];
         b0 = btT[r.s0]; b1 = btT[r.s1]; b2 = btT[r.s2];
      }
      if (inWin && t + 2 < T) {                              // first touched two frames from now
         const float *bt2 = btT + 2 * (size_t)J;
         const double *bq2 = bqT + 2 * S + 5 * myq;
         prefetch_l1(bq2); prefetch_l1(bq2 + 4);
         prefetch_l1(bt2 + r.s0); prefetch_l1(bt2 + r.s1); prefetch_l1(bt2 + r.s2);
      }
      int loN = 0, hiN = 0;                                  // next frame's beta beam (read early, used next iteration)
      if (t + 1 < T) { loN = qLo[t + 1]; hiN = qHi[t + 1]; }
      double a1 = LZERO_D, n0 = LZERO_D, n1 = LZERO_D, n2 = LZERO_D, nEx = LZERO_D;
      int nsq, neq;
      if (t == 0) {
         // ---- InitAlpha, HFB.c:616-651 (no tee models: only the first model can start)
         nsq = 0; neq = hiT;
         if (neq + 1 >= 32) { if (lane == 0) out->redo = 1; return; }
         if (have && myq <= neq) {
            a1 = (myq == 0) ? 0.0 : LZERO_D;
            n0 = a1 + r.aE + (double)b0;
         }
      } else {
         // ---- alpha beam, HFB.c:701-722
         const int q1 = __shfl_sync(FULL, myq, (lane + 31) & 31);
         const double ex1 = __shfl_sync(FULL, exv, (lane + 31) & 31);
         const double ax1 = __shfl_sync(FULL, aEx, (lane + 31) & 31);
         const bool ok1 = (q1 == myq - 1);
         const double mp = dmax(ok1 ? ex1 : LZERO_D, mpS);                  // MaxModelProb(q, t-1), :655-682
         const bool keep = inWin && !(pr - mp > minF);
         // the alpha column (HFB.c:729-771) does not depend on the beam decisions: its log-adds are issued first and
         // run under the latency of the two warp reductions; the beam only decides whether the values are kept
         const double a1s = (myq > 0 && ok1) ? ax1 : LZERO_D;
         const double s0 = ladd_nz(r.aE + a1s, e0 + r.a00) + f2d_alu(b0);
         const double s1 = ladd_nz(e0 + r.a01, e1 + r.a11) + f2d_alu(b1);
         const double s2 = ladd_nz(e1 + r.a12, e2 + r.a22) + f2d_alu(b2);
         const int eq0 = (hiP < Q - 1) ? hiP + 1 : hiP;
         nsq = __reduce_min_sync(FULL, (keep && myq >= loP) ? myq : 0x7fffffff);
         neq = __reduce_max_sync(FULL, (keep && myq <= eq0) ? myq : -1);
         if (nsq > hiT) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }        // HError 7390
         if (nsq < loT) nsq = loT;
         if (neq < nsq) { if (lane == 0) out->status = HFB_UTT_EALPHA; return; }
         if (neq > hiT) neq = hiT;
         if (neq + 1 - nsq >= 32) { if (lane == 0) out->redo = 1; return; }
         if (have && myq >= nsq && myq <= neq) {
            a1 = a1s; n0 = s0; n1 = s1; n2 = s2;
            nEx = (n2 > LSMALL_D) ? n2 + r.a2x : LZERO_D;
         }
      }
      sq = nsq; eq = neq;
      const bool inBeam = have && myq >= sq && myq <= eq;
      if (lane == 0) { sqA[t] = (short)sq; eqA[t] = (short)eq; }
      aEx = nEx; e0 = n0; e1 = n1; e2 = n2;
      mpS = LZERO_D; exv = LZERO_D;
      if (inBeam) {
         mpS = dmax(dmax(a1 + bEn, n0 + bE0), dmax(n1 + bE1, n2 + bE2));
         exv = nEx + bX;
         double *oc = ocT + 3 * myq;
         oc[0] = n0; oc[1] = n1; oc[2] = n2;
         aeT[myq] = a1;
         if (tmin > t) tmin = t;
         tmax = t;
      }
      // ---- slide: a lane whose model fell below the beam takes the model 32 further on
      if (have && myq < sq) {
         if (tmax >= 0) { gTmin[myq] = tmin; gTmax[myq] = tmax; }
         myq += 32; have = myq < Q;
         tmin = 0x7fffffff; tmax = -1;
         aEx = LZERO_D; mpS = LZERO_D; exv = LZERO_D; e0 = e1 = e2 = LZERO_D;
         if (have) load_l2r(r, M, W, u, myq);
         else { r.aE = r.a00 = r.a01 = r.a11 = r.a12 = r.a22 = r.a2x = LZERO_D; r.s0 = r.s1 = r.s2 = 0; }
      }
      bqT += S; btT += J; ocT += P; aeT += Q;
      loP = loT; hiP = hiT; loT = loN; hiT = hiN;
   }
   if (have && tmax >= 0) { gTmin[myq] = tmin; gTmax[myq] = tmax; }
   for (int q = lane; q < Q; q += 32) atomicAdd(&W.acc[M.L.numEgs + W.mHmm[u.modOff + q]], 1.0);   // HFB.c:1768-1772
   if (lane == 0) {
      atomicAdd(&W.acc[M.L.totalT], (double)T);                             // HERest.c:779-780
      atomicAdd(&W.acc[M.L.totalPr], pr);
      atomicAdd(&W.acc[M.L.numOk], 1.0);
   }
}
