// hfb_kernels.cuh -- shared device helpers (log-add, warp reductions) and the general kernels of the E-step.
//
//   ladd / ladd_nz    LAdd (HTKLib/HMath.c:1576-1590) with the correction term on the special-function unit
//   gmm_fp32_kernel   log b_j(o_t) for every (frame, distinct tied state) of an utterance:
//                     diagonal Gaussians (IDOutP, HTKLib/HModel.c:5420-5431) + log-sum-exp
//                     over mixtures (ShStrP, HTKLib/HFB.c:898-988) on the FP32 pipe: single-Gaussian
//                     sets and D > 46; mixture sets use the tcgen05 kernels of gmm_tc.cuh.
//   prep_kernel       CreateInsts (HFB.c:508-574): per-utterance tables, built on the device
//   beta_kernel       SetBeamTaper + SetBeta + the StepBack retry loop (HFB.c:1116-1145, :1149-1296,
//                     :1321-1366) for any topology (N <= 16) and any transcription length: one CTA per
//                     utterance, one thread per model, two beta columns in shared memory, two block
//                     barriers per frame.  The register-resident versions are in hfb_fast.cuh (N <= 8)
//                     and hfb_l2r.cuh (HTK's standard topology).
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
//
// The FP32 correction is branch-free and uses the two special-function-unit approximations
// directly: e = ex2.approx(d log2 e) (relative error <= 2^-22), then lg2.approx(1 + e) (absolute
// error <= 2^-22 on [1, 2]) -- together <= 3.5e-7 absolute on the correction, which is bounded by
// log 2.  With d < minLogExp the correction is dropped, so LAdd(log zero, y) == y exactly as in
// the reference (HMath.c:1584-1585).  ~18 instructions instead of ~50 for log1pf(expf(d)).
__device__ __forceinline__ float ex2_approx(float x)
{
   float r;
   asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}
__device__ __forceinline__ float lg2_approx(float x)
{
   float r;
   asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
   return r;
}

// Conversions between FP64 and FP32 execute on the XU pipe, next to ex2 / lg2 and at a quarter of their rate, and
// that pipe is what bounds the recursion kernels (ncu on beta_l2r_kernel: sm__inst_executed_pipe_xu 150 % of its
// sustained peak with 4 conversions + 2 MUFU per log-add).  The two conversions of a log-add are therefore done with
// integer operations on the ALU pipe: both are exact replacements of cvt.rn for the values that occur here.
//   f2d_alu:      float -> double for zero and normal floats (log values, the bounded correction; no denormal, inf
//                 or NaN reaches it): re-bias the exponent, shift the mantissa.
//   neg_abs_d2f:  -|v| as a float, round to nearest even; |v| below 2^-126 gives -0 (the ex2 that follows flushes
//                 denormals anyway), |v| is far below 2^128 (log zero is -1e10).
__device__ __forceinline__ double f2d_alu(float f)
{
   const unsigned b = __float_as_uint(f), mag = b & 0x7fffffffu;
   const unsigned hi = (b & 0x80000000u) | (mag ? (mag >> 3) + 0x38000000u : 0u);
   return __hiloint2double((int)hi, (int)(b << 29));
}
__device__ __forceinline__ float neg_abs_d2f(double v)
{
   const unsigned hi = (unsigned)__double2hiint(v) & 0x7fffffffu, lo = (unsigned)__double2loint(v);
   unsigned r = ((hi - 0x38000000u) << 3) | (lo >> 29);
   const unsigned rem = lo & 0x1fffffffu;
   r += (rem > 0x10000000u || (rem == 0x10000000u && (r & 1u))) ? 1u : 0u;
   r = (hi < 0x38100000u) ? 0u : r;
   return __uint_as_float(r | 0x80000000u);
}

template <bool EXACT>
__device__ __forceinline__ double ladd(double x, double y)
{
   if (EXACT) {
      if (x < y) { double t = x; x = y; y = t; }
      double d = y - x;
      if (d < MINLOGEXP) return (x < LSMALL_D) ? LZERO_D : x;
      return x + log(1.0 + exp(d));
   }
   const double hi = (x > y) ? x : y;
   const float d = neg_abs_d2f(x - y);
   float c = lg2_approx(1.0f + ex2_approx(d * 1.4426950408889634f)) * 0.6931471805599453f;
   c = (d < (float)MINLOGEXP) ? 0.f : c;
   return (hi < LSMALL_D) ? LZERO_D : hi + f2d_alu(c);
}

// The same for callers that have already checked one operand against LSMALL (every guarded
// `if (term > LSMALL) x = LAdd(x, term)` of HFB.c): the result cannot be log zero.
__device__ __forceinline__ double ladd_nz(double x, double y)
{
   const double hi = (x > y) ? x : y;
   const float d = neg_abs_d2f(x - y);
   float c = lg2_approx(1.0f + ex2_approx(d * 1.4426950408889634f)) * 0.6931471805599453f;
   c = (d < (float)MINLOGEXP) ? 0.f : c;
   return hi + f2d_alu(c);
}

// beta kernels (64 registers, one thread per model): the integer conversions cost more issue slots and registers
// there than the XU conversions they save (measured: 1.81 -> 1.90 ms with both on the ALU, 1.95 ms with only the
// float -> double one), so they keep cvt
__device__ __forceinline__ double ladd_nz_b(double x, double y)
{
   const double hi = (x > y) ? x : y;
   const float d = -fabsf((float)(x - y));
   float c = lg2_approx(1.0f + ex2_approx(d * 1.4426950408889634f)) * 0.6931471805599453f;
   c = (d < (float)MINLOGEXP) ? 0.f : c;
   return hi + (double)c;
}

__device__ __forceinline__ void prefetch_l1(const void *p)
{
   asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
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

struct GmmTile { int utt, t0, s0; };

// largest i in [0, n) with pre[i] <= v (pre non-decreasing, pre[0] <= v)
__device__ __forceinline__ int upper_index(const int *__restrict__ pre, int n, int v)
{
   int lo = 0, hi = n;
   while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (pre[mid] <= v) lo = mid; else hi = mid; }
   return lo;
}

__global__ void __launch_bounds__(128)
gmm_fp32_kernel(DevModel M, Wave W)
{
   extern __shared__ float gsm[];
   float *xs = gsm;                              // [D][GT_FR]
   float *outT = gsm + M.D * GT_FR;              // [GT_FR][GT_SL+1]
   GmmTile tl;
   tl.utt = upper_index(W.tilePre, W.numUtt, (int)blockIdx.x);
   const UttDesc u = W.utt[tl.utt];
   {
      const int lt = (int)blockIdx.x - W.tilePre[tl.utt], nSl = (u.P + GT_SL - 1) / GT_SL;
      tl.t0 = (lt / nSl) * GT_FR; tl.s0 = (lt % nSl) * GT_SL;
   }
   if (W.out[tl.utt].status != 0 || tl.s0 >= u.Jt) return;
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
      if (slot >= u.Jt) break;
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
      if (t < u.T && slot < u.Jt) b[(size_t)t * u.J + slot] = outT[f * (GT_SL + 1) + sl];
   }
}

// ------------------------------------------------------------------------------------------
// K0: per-utterance tables, CreateInsts (HFB.c:508-574) on the device.  One CTA per utterance.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) prep_kernel(DevModel M, Wave W)
{
   __shared__ int sStatus;
   UttDesc *u = &W.utt[blockIdx.x];
   UttOut *out = &W.out[blockIdx.x];
   if (out->status != 0) return;                        // rejected by the host (bad label index ...)
   const int tid = threadIdx.x, nt = blockDim.x;
   const int Q = u->Q, T = u->T, P = u->P;
   const int *lab = W.lab + u->labOff;
   int *mN = W.mN + u->modOff, *mTrans = W.mTrans + u->modOff, *mSoff = W.mSoff + u->modOff;
   int *mPoff = W.mPoff + u->modOff, *mDms = W.mDms + u->modOff, *mPre = W.mPre + u->modOff;
   int *mSuf = W.mSuf + u->modOff, *mHmm = W.mHmm + u->modOff;
   long long *mTrAcc = W.mTrAcc + u->modOff, *mTrOcc = W.mTrOcc + u->modOff;
   int *posState = W.posState + u->posOff, *posSlot = W.posSlot + u->posOff, *slotState = W.slotState + u->slotOff;
   for (int q = tid; q < Q; q += nt) {
      const int p = lab[q], tr = M.hmmTrans[p];
      mN[q] = M.hmmN[p]; mTrans[q] = M.transOffF[tr]; mDms[q] = M.transMinDur[tr]; mHmm[q] = p;
      mTrAcc[q] = M.tranAccOff[tr]; mTrOcc[q] = M.tranOccOff[tr];
   }
   __syncthreads();
   if (tid == 0) {
      int S = 0, Pp = 0, qt = 0, bad = 0, prevD = 1;
      for (int q = 0; q < Q; q++) {
         const int N = mN[q], d = mDms[q];
         mSoff[q] = S; mPoff[q] = Pp; mPre[q] = qt;
         S += N; Pp += N - 2; qt += d;
         if (q > 0 && d == 0 && prevD == 0) bad = HFB_UTT_ETEE;                // HFB.c:557
         prevD = d;
      }
      int acc = 0;
      for (int q = Q - 1; q >= 0; q--) { mSuf[q] = acc; acc += mDms[q]; }
      if (mDms[0] == 0 || mDms[Q - 1] == 0) bad = HFB_UTT_ETEE;                // HFB.c:564
      if (!bad && qt > T) bad = HFB_UTT_SKIPPED;                               // HFB.c:1339-1343
      sStatus = bad;
   }
   __syncthreads();
   if (sStatus != 0) { if (tid == 0) { out->status = sStatus; out->J = 0; } return; }
   for (int q = tid; q < Q; q += nt) {
      const int p = lab[q], n = mN[q] - 2, so = M.hmmStateOff[p], po = mPoff[q];
      for (int j = 0; j < n; j++) { posState[po + j] = M.hmmState[so + j]; W.posQ[u->posOff + po + j] = q; }
   }
   __syncthreads();
   // the reference evaluates each tied state once per frame (HFB.c:910-912): find the first
   // position using the same state, then number the distinct states ("slots")
   __shared__ int sJ;
   if (W.globalSlots > 0) {
      // small single-Gaussian sets: slot = tied state (every utterance uses most of them anyway), so that the
      // tensor-core kernel reads its B operand as contiguous rows
      for (int pp = tid; pp < P; pp += nt) posSlot[pp] = posState[pp];
      for (int j = tid; j < W.globalSlots; j += nt) slotState[j] = j;
      if (tid == 0) { sJ = W.globalSlots; u->Jt = sJ; u->J = (sJ + 3) & ~3; out->J = sJ; }
   } else {
      for (int pp = tid; pp < P; pp += nt) {
         const int s = posState[pp];
         int f = pp;
         for (int p2 = 0; p2 < pp; p2++) if (posState[p2] == s) { f = p2; break; }
         posSlot[pp] = f;
      }
      __syncthreads();
      if (tid == 0) {
         int J = 0;
         for (int pp = 0; pp < P; pp++) {
            const int f = posSlot[pp];
            if (f == pp) { slotState[J] = posState[pp]; posSlot[pp] = J++; }
            else posSlot[pp] = posSlot[f];
         }
         sJ = J; u->Jt = J; u->J = (J + 3) & ~3; out->J = J;
      }
   }
   __syncthreads();
   // ---- frames in which each slot can be needed (see Wave::slotFirst)
   const int J = sJ;
   int *slotFirst = W.slotFirst + u->slotOff, *slotLast = W.slotLast + u->slotOff;
   for (int j = tid; j < J; j += nt) { slotFirst[j] = W.noTaperSkip ? 0 : 0x7fffffff; slotLast[j] = W.noTaperSkip ? T - 1 : -1; }
   __syncthreads();
   if (!W.noTaperSkip)
      for (int pp = tid; pp < P; pp += nt) {
         const int q = W.posQ[u->posOff + pp], q2 = min(q + 2, Q - 1);
         const int tf = max(0, mPre[q] - 3), tl = min(T - 1, T - 1 - mSuf[q2] + 2);
         atomicMin(&slotFirst[posSlot[pp]], tf);
         atomicMax(&slotLast[posSlot[pp]], tl);
      }
   __syncthreads();
   long long mine = 0;
   for (int j = tid; j < J; j += nt) mine += max(0, slotLast[j] - slotFirst[j] + 1);
   if (W.spt > 0) {
      int2 *tileIv = W.tileIv + u->slotOff;
      const int nTiles = (J + W.spt - 1) / W.spt;
      for (int n = tid; n < nTiles; n += nt) {
         int f = 0x7fffffff, l = -1;
         for (int j = n * W.spt; j < min(J, (n + 1) * W.spt); j++) { f = min(f, slotFirst[j]); l = max(l, slotLast[j]); }
         tileIv[n] = make_int2(f, l);
      }
   }
   __shared__ long long sPairs[4];
   for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
   if ((tid & 31) == 0) sPairs[tid >> 5] = mine;
   __syncthreads();
   if (tid == 0) { long long a = 0; for (int w2 = 0; w2 < (nt >> 5); w2++) a += sPairs[w2]; out->pairs = a; }
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
__global__ void __launch_bounds__(1024) beta_kernel(DevModel M, Wave W)
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
            if (q < endq - 2 || q > startq) continue;
            if (t >= 2) {                                  // b of frame t-2 is first touched two steps from now
               const int *ps2 = posSlot + sm.sPoff[q];
               const float *bt2 = bU + (size_t)(t - 2) * J;
               for (int j = 0; j < sm.sN[q] - 2; j++) prefetch_l1(bt2 + ps2[j]);
            }
            if (q < endq) continue;
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
         if (nhi < 0) { fail = true; if (endq == 0) status = HFB_UTT_EBETA; break; }   // HError 7323
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

