// hfb_mstep.cuh -- the M-step of HERest on the device: new transition matrices, mixture
// weights, means, variances and gConsts from the resident FP64 accumulators, so that further EM passes need
// neither the accumulator download nor the scatter into HTK's structures.
//
// Restates MLUpdateModels (HTKTools/HERest.c:1262-1321) for PLAINHS / SHAREDHS diagonal-covariance sets:
//   * a physical HMM with fewer than minEgs examples is "copied" (:1286-1298); a shared structure (tied state,
//     ~t matrix, mean / variance vector) is updated iff some HMM that uses it is updated -- the reference gets
//     this from clearing the accumulator hook after the first use; here mstep_enable_kernel marks them;
//   * UpdateTrans (:795-816):   a_ij = tran_ij / occ_i for i = 1..N-1, j = 2..N, rows with occ_i = 0 kept;
//   * UpdateWeights (:897-971): w_m = c_m / occ clamped to 1, zero below MINMIX, then FloorMixes (:819-840);
//   * UpdateVars (:1045-1122):  var_k = va_k / occ - (mu_k / occ_mean)^2, the second term dropped when the
//     variance vector is shared or its mean has no occupancy, floored per dimension; only for components whose
//     NEW weight exceeds MINMIX;
//   * UpdateMeans (:974-1012):  mean_k += mu_k / occ (the accumulators are centred on the old mean);
//   * FixGConsts (HModel.c:5688, :5641): gConst = D log 2 pi + sum log var_k for the updated components.
// The accumulators are FP64 here and float in the reference: agreement to float rounding.
#pragma once
#include "hfb_kernels.cuh"

#define HFB_MINMIX_F 1.0e-5f
#define HFB_MINLARG_F 2.45e-308

struct MStepDev {
   const double *acc;
   int minEgs, uFlags, maxM;
   float mixWeightFloor;
   const float *vFloor;          // [D]
   // per-structure flags (zeroed before the pass)
   int *transOn, *stateOn, *meanOn, *varOn;
   const int *varUse;            // number of Gaussians that use the variance vector
   const int *gaussFirstOfVar;   // 1 if this Gaussian is the first user of its variance vector (counts floors once)
   int *counters;                // [0] floored variance elements, [1] mixes with floored variances, [2] models copied,
                                 // [3] structures with zero occupancy
   // outputs
   float *mean, *var, *gConst, *mixWeight, *transP;
};

// one thread per physical HMM: which shared structures does an updated HMM touch
__global__ void mstep_enable_kernel(DevModel M, MStepDev S)
{
   const int p = blockIdx.x * blockDim.x + threadIdx.x;
   if (p >= M.P) return;
   const double n = S.acc[M.L.numEgs + p];
   if (!(n >= (double)S.minEgs && n > 0.0)) { atomicAdd(&S.counters[2], 1); return; }
   S.transOn[M.hmmTrans[p]] = 1;
   for (int i = M.hmmStateOff[p]; i < M.hmmStateOff[p + 1]; i++) S.stateOn[M.hmmState[i]] = 1;
}

// one thread per (transition matrix, row)
__global__ void mstep_trans_kernel(DevModel M, MStepDev S, const int *__restrict__ transN)
{
   const int tr = blockIdx.x, i = threadIdx.x, N = transN[tr];
   if (i >= N) return;
   const float *A = M.transLogA + M.transOffF[tr];
   float *out = S.transP + M.transOffF[tr];
   const bool upd = (S.uFlags & HFB_UPTRANS) && S.transOn[tr] && i < N - 1;
   const double occ = S.acc[M.tranOccOff[tr] + i];
   for (int j = 0; j < N; j++) {
      float v = (A[i * N + j] > (float)LSMALL_D) ? expf(A[i * N + j]) : 0.f;
      if (upd && occ > 0.0 && j >= 1) {
         const float x = (float)(S.acc[M.tranAccOff[tr] + i * N + j] / occ);
         v = (x > (float)HFB_MINLARG_F) ? x : 0.f;
      }
      out[i * N + j] = v;
   }
   if (upd && !(occ > 0.0)) atomicAdd(&S.counters[3], 1);
}

// one thread per tied state: weights, and which mean / variance vectors its live components touch
__global__ void mstep_state_kernel(DevModel M, MStepDev S)
{
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if (s >= M.J) return;
   const int mo = M.stateMixOff[s], Mn = M.stateMixOff[s + 1] - mo;
   const bool on = S.stateOn[s] != 0;
   const bool updW = on && (S.uFlags & HFB_UPMIXES) && S.maxM > 1;
   const double occ = S.acc[M.L.wtOcc + s];
   for (int m = 0; m < Mn; m++) {
      const float lw = M.mixLogWt[mo + m];
      S.mixWeight[mo + m] = (lw > (float)LSMALL_D) ? expf(lw) : 0.f;
   }
   if (updW) {
      if (occ > 0.0) {
         for (int m = 0; m < Mn; m++) {
            float x = (float)(S.acc[M.L.wtC + mo + m] / occ);
            if (x > 1.f) x = 1.f;
            S.mixWeight[mo + m] = (x > HFB_MINMIX_F) ? x : 0.f;
         }
         const float floor = S.mixWeightFloor;
         if (floor > 0.f) {                                   // FloorMixes
            float sum = 0.f, fsum = 0.f;
            for (int m = 0; m < Mn; m++) {
               if (S.mixWeight[mo + m] > floor) sum += S.mixWeight[mo + m];
               else { fsum += floor; S.mixWeight[mo + m] = floor; }
            }
            if (fsum > 0.f && sum > 0.f) {
               const float scale = (1.f - fsum) / sum;
               for (int m = 0; m < Mn; m++) if (S.mixWeight[mo + m] > floor) S.mixWeight[mo + m] *= scale;
            }
         }
      } else atomicAdd(&S.counters[3], 1);
   }
   if (on)
      for (int m = 0; m < Mn; m++)
         if (S.mixWeight[mo + m] > HFB_MINMIX_F) {
            const int g = M.mixGauss[mo + m];
            S.meanOn[M.meanId[g]] = 1;
            S.varOn[M.varId[g]] = 1;
         }
}

// one thread per Gaussian
__global__ void mstep_gauss_kernel(DevModel M, MStepDev S)
{
   const int g = blockIdx.x * blockDim.x + threadIdx.x;
   if (g >= M.G) return;
   const int D = M.D, Dp = M.Dp, mId = M.meanId[g], vId = M.varId[g];
   const bool updM = (S.uFlags & HFB_UPMEANS) && S.meanOn[mId];
   const bool updV = (S.uFlags & HFB_UPVARS) && S.varOn[vId];
   const double mocc = S.acc[M.L.muOcc + mId], vocc = S.acc[M.L.vaOcc + vId];
   const bool haveMu = (S.uFlags & HFB_UPMEANS) && mocc > 0.0;          // ma != NULL && ma->occ > 0
   const bool shared = S.varUse[vId] > 1 || !haveMu;
   const bool first = S.gaussFirstOfVar[g] != 0;
   bool floored = false;
   float gc = (float)(D * log(6.283185307179586));
   for (int k = 0; k < D; k++) {
      const float oldMean = M.mean[(size_t)g * Dp + k], oldVar = 1.0f / M.ivar[(size_t)g * Dp + k];
      float mean = oldMean, var = oldVar;
      if (updM && mocc > 0.0) mean = oldMean + (float)(S.acc[M.L.muSum + (size_t)mId * D + k] / mocc);
      if (updV && vocc > 0.0) {
         const float md = shared ? 0.f : (float)(S.acc[M.L.muSum + (size_t)mId * D + k] / mocc);
         float x = (float)(S.acc[M.L.vaSum + (size_t)vId * D + k] / vocc) - md * md;
         if (x < S.vFloor[k]) { x = S.vFloor[k]; floored = true; if (first) atomicAdd(&S.counters[0], 1); }
         var = x;
      }
      S.mean[(size_t)g * D + k] = mean;
      S.var[(size_t)g * D + k] = var;
      gc += (var <= 0.f) ? (float)LZERO_D : logf(var);
   }
   if (floored && first) atomicAdd(&S.counters[1], 1);
   if ((updM && !(mocc > 0.0)) || (updV && !(vocc > 0.0))) { if (first) atomicAdd(&S.counters[3], 1); }
   // FixGConsts touches the live components of updated models; the others keep the stored value
   S.gConst[g] = ((S.uFlags & (HFB_UPMEANS | HFB_UPVARS)) && S.varOn[vId]) ? gc : M.gconst[g];
}
