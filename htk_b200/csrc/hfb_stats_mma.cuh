// hfb_stats_mma.cuh -- K4 with the occupancy-weighted sums on the tensor cores.
//
// UpMixParms (HTKLib/HFB.c:1426-1744) accumulates, per mixture component m of state j,
//     occ += Lr,   mu[k] += Lr (o_t[k] - mean_m[k]),   var[k] += Lr (o_t[k] - mean_m[k])^2
// over the frames where the state is inside the alpha beam and the component passes the
// minimum-occupancy rule.  stats3_kernel does these sums on the FP32 pipe, ~14 warp instructions
// per (component, frame).  Here they are the matrix product north_star asks for,
//     [ S0 | S1 | S2 ](m, :) = Lr[m, t] x [ 1 | o_t - c_j | (o_t - c_j)^2 ][t, :],
// evaluated with mma.sync.m16n8k8 (16 components x 8 frames x 8 dims), 3xTF32 (hi*hi + hi*lo + lo*hi,
// FP32 accumulate, ~2^-21 relative).  The occupancy matrix of an utterance is ~2 % dense (alpha beam,
// HFB.c:701-722), so the product runs only over the frames that survive -- a dense tcgen05 GEMM over
// [all states x all frames] would do ~50x the work.
//
// STATE-MAJOR ORDER.  A position (one emitting state of one label) is inside the alpha beam for ~17
// frames, but it updates 16 x (2 D + 3) accumulators: per-position flushing made the atomics and their
// address arithmetic a third of the kernel.  So the positions of the whole wave are first bucketed by
// tied state (statpos_* kernels: histogram, scan, scatter -- a counting sort), a warp takes a slice of
// S5_CAP consecutive entries of the sorted list, keeps the accumulator fragments in registers across
// positions and flushes only when the state changes: ~20x fewer flushes, and the Gaussian parameters
// of a state are read by neighbouring warps at the same time.
//
// The reference centres on each component's own mean.  A matrix product needs ONE centre per row
// block, so the sums are taken about the state's centre c_j (mean of its component means, uploaded
// once) and moved to the component mean at flush time in FP64:
//     mu  = S1 - d S0,    var = S2 - 2 d S1 + d^2 S0,    d = mean_m - c_j .
// |d| is the spread of the components inside one state (~sigma), so the cancellation costs about
// one bit -- unlike un-centred sums, which would lose mean^2/sigma^2.
//
// The component posteriors themselves (phase 1) keep stats3's code: reference operation order on the
// FP32 pipe (IDOutP, HModel.c:5425-5430), lanes <-> (component, frame) pairs.
#pragma once
#include "hfb_kernels2.cuh"

#define S4_WARPS 4
#define S4_LSTR 36                    // row stride of the Lr tile: conflict-free A-fragment loads
#define S4_CSTR 40                    // row stride of the state-centre table

// TF32 split on the integer pipe (cvt.rna.tf32 is a quarter-rate conversion): hi = x rounded to the 19
// bits the tensor core reads (add half an ulp, mask), lo = x - hi exactly (signed, <= 12 significant
// bits, of which the tensor core keeps 11): |x - hi - lo_read| <= 2^-23 |x| and unbiased, unlike plain
// truncation, whose always-towards-zero error would bias every occupancy sum.
__device__ __forceinline__ uint32_t s4_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ uint32_t s4_lo(float x) { return __float_as_uint(x - __uint_as_float(s4_hi(x))); }
__device__ __forceinline__ void s4_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
   asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}


#define S5_CAP 16                     // sorted positions per warp: minimum; the host raises it to 32 / 64 for large waves

// ---- counting sort of the wave's positions by tied state ------------------------------------------
// cnt[J] must be zero on entry.  Only positions whose model was ever inside the alpha beam are listed.
__global__ void __launch_bounds__(128) statpos_count_kernel(Wave W, int *__restrict__ cnt)
{
   const UttDesc &u = W.utt[blockIdx.x];
   if (W.out[blockIdx.x].status != 0) return;
   const int *posState = W.posState + u.posOff, *posQ = W.posQ + u.posOff;
   const int *tmin = W.mTmin + u.modOff, *tmax = W.mTmax + u.modOff;
   for (int pp = threadIdx.x; pp < u.P; pp += blockDim.x) {
      const int q = posQ[pp];
      if (tmin[q] <= tmax[q]) atomicAdd(&cnt[posState[pp]], 1);
   }
}
// off[s] = exclusive prefix sum of cnt, off[J] = total; fill[] zeroed.  One CTA.
__global__ void __launch_bounds__(1024) statpos_scan_kernel(const int *__restrict__ cnt, int *__restrict__ off,
                                                            int *__restrict__ fill, int J)
{
   __shared__ int part[1024];
   const int tid = threadIdx.x, per = (J + 1023) / 1024, b = tid * per, e = min(J, b + per);
   int sum = 0;
   for (int i = b; i < e; i++) sum += cnt[i];
   part[tid] = sum;
   __syncthreads();
   for (int o = 1; o < 1024; o <<= 1) {
      int v = (tid >= o) ? part[tid - o] : 0;
      __syncthreads();
      part[tid] += v;
      __syncthreads();
   }
   int run = part[tid] - sum;
   for (int i = b; i < e; i++) { off[i] = run; run += cnt[i]; fill[i] = 0; }
   if (tid == 1023) off[J] = part[1023];
}
// Everything stats5 needs about one listed position, gathered here (in parallel over positions) so that
// the statistics warp reads ONE flat record instead of chasing ~15 dependent table look-ups per position.
struct __align__(16) PosRec {
   long long alphaOff, aentOff, betaOff, bOff, frameBase, featOff, trAcc, trOcc;   // element offsets into the wave arrays
   long long vOff;                  // first entry of this position in the valid-frame list (stats_pre_kernel), -1: none
   double pr;
   int s, q, j, tmin, tmax, N, so, sq1, P, S, Q, J, T, transOff, utt, pad;
   int ps[HFB_MAXN];                // output-probability slots of the model's emitting states
};

__global__ void __launch_bounds__(128) statpos_scatter_kernel(Wave W, const int *__restrict__ off, int *__restrict__ fill,
                                                              PosRec *__restrict__ list, unsigned long long *vCursor,
                                                              long long vCap, int *__restrict__ posIdx, int *__restrict__ overflow)
{
   const UttDesc &u = W.utt[blockIdx.x];
   if (W.out[blockIdx.x].status != 0) return;
   const int *posState = W.posState + u.posOff, *posQ = W.posQ + u.posOff;
   const int *tmin = W.mTmin + u.modOff, *tmax = W.mTmax + u.modOff;
   const int lane = threadIdx.x & 31;
   for (int base = 0; base < u.P; base += blockDim.x) {
      const int pp = base + threadIdx.x;
      const bool in = pp < u.P;
      const int q = in ? posQ[pp] : 0, gq = u.modOff + q;
      const bool on = in && tmin[q] <= tmax[q];
      // room for one list entry per frame of the model's alpha-beam span: one atomic per warp on the cursor
      // (one per position serialised 300 k atomics on a single address: 0.2 ms)
      const long long span = on ? (long long)tmax[q] - tmin[q] + 1 : 0;
      long long incl = span;
      for (int o = 1; o < 32; o <<= 1) { const long long v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
      const long long tot = __shfl_sync(0xffffffffu, incl, 31);
      long long wbase = 0;
      if (lane == 0 && tot > 0 && vCap > 0) wbase = (long long)atomicAdd(vCursor, (unsigned long long)tot);
      wbase = __shfl_sync(0xffffffffu, wbase, 0);
      if (on) {
         const int s = posState[pp];
         PosRec r;
         r.s = s; r.q = q; r.j = pp - W.mPoff[gq]; r.tmin = tmin[q]; r.tmax = tmax[q];
         r.N = W.mN[gq]; r.so = W.mSoff[gq]; r.sq1 = (q < u.Q - 1) ? W.mSoff[gq + 1] : 0;
         r.P = u.P; r.S = u.S; r.Q = u.Q; r.J = u.J; r.T = u.T; r.transOff = W.mTrans[gq]; r.utt = (int)blockIdx.x; r.pad = 0;
         r.alphaOff = u.occOff + pp; r.aentOff = u.aentOff + q; r.betaOff = u.betaOff; r.bOff = u.bOff;
         r.frameBase = u.frameBase; r.featOff = u.featOff; r.trAcc = W.mTrAcc[gq]; r.trOcc = W.mTrOcc[gq];
         r.pr = W.out[blockIdx.x].pr;
         const long long o = wbase + incl - span;
         r.vOff = (vCap > 0 && o + span <= vCap) ? o : -1;    // positions that do not fit keep the inline front
         if (r.vOff < 0) *overflow = 1;                       // ... of stats5_kernel: the tcgen05 kernel then leaves the wave to it
         const int *ps = W.posSlot + u.posOff + W.mPoff[gq];
#pragma unroll
         for (int i = 0; i < HFB_MAXN; i++) r.ps[i] = (i < r.N - 2) ? ps[i] : 0;
         const int at = off[s] + atomicAdd(&fill[s], 1);
         list[at] = r;
         posIdx[u.posOff + pp] = at;
      } else if (in)
         posIdx[u.posOff + pp] = -1;
   }
}

// ---- packed FP32 pairs (FADD2 / FMUL2 / FFMA2 on sm_100): two frames of one component per instruction in phase 1
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t f2_pack(float lo, float hi) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(f32x2_t v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2_t f2_sub(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_mul(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t f2_fma(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
// one dimension of two frames: sum += ((o - mu)^2) * ivar, the product and the sum fused (one rounding fewer than
// IDOutP, HModel.c:5425-5430: ~1e-7 relative on log N, three orders below the parity bound)
__device__ __forceinline__ f32x2_t f2_step(f32x2_t sum, f32x2_t o, float mu, float iv)
{
   const f32x2_t d = f2_sub(o, f2_pack(mu, mu));
   return f2_fma(f2_mul(d, d), f2_pack(iv, iv), sum);
}

// ---- front half of the statistics as its own kernel ------------------------------------------------------------
// Alpha-beam test, state occupancy, SetOcct + UpTranParms, and per position the list of frames that can contribute
// to the mixture statistics with their initx.  Inside the state-sorted stats5_kernel this part was 43 % of the stall
// samples and 2.7 GB of DRAM reads per step: a position touches ~16 sectors per frame (alpha, entry alpha, two beta
// columns, b at t and t+1, beams) and the three states of a model -- which share nearly all of them -- sit in
// different corners of the sorted list.  Here one warp owns a MODEL of an utterance (lanes <-> frames of a 32-frame
// window, loop over its emitting states), so those sectors are fetched once, at 64 registers and 32 warps/SM, and
// stats5_kernel gets dense chunks of 32 VALID frames instead of 32-frame windows of the span.
struct __align__(16) ValidFrame { double x0; int t; int pad; };

#define SPRE_WARPS 4
__global__ void __launch_bounds__(32 * SPRE_WARPS, 8) stats_pre_kernel(DevModel M, Wave W, const PosRec *__restrict__ list,
                                                                       const int *__restrict__ posIdx,
                                                                       ValidFrame *__restrict__ vbuf, int *__restrict__ vcnt)
{
   const UttDesc &u = W.utt[blockIdx.x];
   if (W.out[blockIdx.x].status != 0) return;
   const int lane = threadIdx.x & 31;
   const int uf = W.uFlags;
   const bool doMix = (uf & (HFB_UPMEANS | HFB_UPVARS | HFB_UPMIXES)) != 0, doTr = (uf & HFB_UPTRANS) != 0;
   const double minF = W.minFrwdP;
   const short *sqA = W.sq + u.frameBase, *eqA = W.eq + u.frameBase;
   StatPos sp;
   sp.betaU = W.beta + u.betaOff; sp.bU = W.b + u.bOff; sp.qLo = W.qLo + u.frameBase; sp.qHi = W.qHi + u.frameBase;
   sp.pr = W.out[blockIdx.x].pr; sp.P = u.P; sp.S = u.S; sp.Q = u.Q; sp.J = u.J; sp.T = u.T;
   const double pr = sp.pr;
   for (int q = blockIdx.y * SPRE_WARPS + (threadIdx.x >> 5); q < u.Q; q += SPRE_WARPS * gridDim.y) {
      const int gq = u.modOff + q, tmin = W.mTmin[gq], tmax = W.mTmax[gq];
      if (tmin > tmax) continue;
      const int N = W.mN[gq], nj = N - 2, pp0 = W.mPoff[gq];
      const float *A = M.transLogA + W.mTrans[gq];
      sp.aent = W.aent + u.aentOff + q; sp.A = A; sp.ps = W.posSlot + u.posOff + pp0;
      sp.tacc = W.acc + W.mTrAcc[gq]; sp.oacc = W.acc + W.mTrOcc[gq];
      sp.N = N; sp.so = W.mSoff[gq]; sp.sq1 = (q < u.Q - 1) ? W.mSoff[gq + 1] : 0; sp.q = q; sp.aTee = A[N - 1];
      // lane j keeps what belongs to emitting state j: list index, list offset, single-Gaussian flag, entries so far
      int myIdx = -1, myOne = 0, myCnt = 0;
      long long myOff = -1;
      if (lane < nj) {
         myIdx = posIdx[u.posOff + pp0 + lane];
         if (myIdx >= 0) myOff = list[myIdx].vOff;
         const int st = W.posState[u.posOff + pp0 + lane];
         myOne = (M.stateMixOff[st + 1] - M.stateMixOff[st]) == 1;
      }
      for (int t0 = tmin; t0 <= tmax; t0 += 32) {
         const int t = t0 + lane;
         const bool inb = (t <= tmax) && q >= sqA[t] && q <= eqA[t];
         if (!__ballot_sync(0xffffffffu, inb)) continue;
         for (int j = 0; j < nj; j++) {
            const long long vOff = __shfl_sync(0xffffffffu, myOff, j);
            if (vOff < 0) continue;                    // no room in the list: stats5_kernel does this position inline
            const int one = __shfl_sync(0xffffffffu, myOne, j), n = __shfl_sync(0xffffffffu, myCnt, j);
            sp.alphaJ = W.occ + u.occOff + pp0 + j; sp.j = j;
            sp.aEntJ = A[1 + j]; sp.aExitJ = A[(1 + j) * N + N - 1];
            if (doTr) stats_tran_chunk(sp, t, inb, lane);
            if (!doMix) continue;
            double x0 = OCC_SKIP;
            if (inb) {
               const double aj = sp.alphaJ[(size_t)t * u.P];
               const double *bq = sp.betaU + (size_t)t * u.S + sp.so;
               const float bjt = sp.bU[(size_t)t * u.J + sp.ps[j]];
               const double lg = aj + bq[1 + j] - pr;                              // log occupancy of state j
               if (!(lg < -(minF + 0.25))) x0 = one ? lg : lg - (double)bjt;       // :1575-1576 / initx :1480-1489
            }
            const bool valid = inb && x0 > -1.0e29;
            const unsigned mask = __ballot_sync(0xffffffffu, valid);
            if (valid) {
               ValidFrame v; v.x0 = x0; v.t = t; v.pad = 0;
               vbuf[vOff + n + __popc(mask & ((1u << lane) - 1))] = v;
            }
            if (lane == j) myCnt += __popc(mask);
         }
      }
      if (lane < nj && myOff >= 0) vcnt[myIdx] = myCnt;
   }
}

#define S5_GSTR 44                   // row stride (floats) of the staged Gaussian parameters: 16-byte aligned rows
// observation tile: frames interleaved in pairs, os[(row >> 1) * S5_OPS + 2 * col + (row & 1)], so that phase 1 reads
// {o_a[k], o_b[k], o_a[k+1], o_b[k+1]} of a frame pair with one 128-bit load; 84 = 2 * 40 columns + 4 (pair rows of a
// quarter warp on distinct banks)
#define S5_OPS 84
#define S5_OS(row, col) ((((row) >> 1) * S5_OPS) + 2 * (col) + ((row) & 1))
// Shared memory per warp: x0s[32] doubles, ts[32] ints, the observation tile (16 frame pairs x S5_OPS; columns: D
// dimensions, a ones column, zero padding), the Lr tile [16][S4_LSTR] (both reused as flush staging
// [16][16 NT]) and the state's Gaussians (means, inverse variances, gConst, log weights).
template <int NT>
__host__ __device__ inline size_t stats5_warp_bytes(int D)
{
   const size_t work = sizeof(float) * ((size_t)16 * S5_OPS + 16 * S4_LSTR);
   const size_t flush = sizeof(float) * 16 * 16 * NT;
   const size_t gauss = sizeof(float) * (2 * 16 * S5_GSTR + 32);
   return sizeof(double) * 32 + sizeof(int) * 32 + ((((work > flush) ? work : flush) + 15) & ~(size_t)15) + gauss;
}

// NT = 8-column tiles of the right-hand side: columns 0..D-1 the dimensions, column D the ones column (D + 1 <= 8 NT)
template <int NT>
__global__ void __launch_bounds__(32 * S4_WARPS, 3)
stats5_kernel(DevModel M, Wave W, const float *__restrict__ centre, const PosRec *__restrict__ list,
              const int *__restrict__ listEnd, const ValidFrame *__restrict__ vbuf, const int *__restrict__ vcnt, int cap,
              const int *__restrict__ onlyIfOverflow)
{
   extern __shared__ __align__(16) unsigned char smraw[];
   if (onlyIfOverflow != nullptr && *onlyIfOverflow == 0) return;   // stats_tc_kernel has done the wave
   const int wInB = threadIdx.x >> 5, lane = threadIdx.x & 31;
   const int nSorted = *listEnd;
   const int i0 = (blockIdx.x * S4_WARPS + wInB) * cap, i1 = min(nSorted, i0 + cap);
   if (i0 >= i1) return;
   const int D = M.D, Dp = M.Dp;
   unsigned char *mine = smraw + stats5_warp_bytes<NT>(D) * wInB;
   double *x0s = (double *)mine;                       // [32] initx / log occupancy per chunk frame
   int *ts = (int *)(x0s + 32);                        // [32] frame numbers
   float *os = (float *)(ts + 32);                     // [32][ostr] observations | 1 | 0...
   float *lrs = os + 16 * S5_OPS;                        // [16 components][S4_LSTR] occupancies Lr
   float *fb = os;                                     // flush staging [16][16 NT]
   constexpr int FSTR = 16 * NT;
   float *gmu = (float *)(mine + stats5_warp_bytes<NT>(D) - sizeof(float) * (2 * 16 * S5_GSTR + 32));
   float *giv = gmu + 16 * S5_GSTR, *ggc = giv + 16 * S5_GSTR, *gwt = ggc + 16;   // the current state's 16 components
   const double minF = W.minFrwdP;
   const int uf = W.uFlags;
   const bool upM = (uf & HFB_UPMEANS) != 0, upV = (uf & HFB_UPVARS) != 0, upW = (uf & HFB_UPMIXES) != 0;
   const bool doMix = upM || upV || upW, doTr = (uf & HFB_UPTRANS) != 0;
   // fragment coordinates (PTX ISA, mma.m16n8k8 .tf32): A rows g / g+8, cols c / c+4; B rows c / c+4, col g
   const int fg = lane >> 2, fc = lane & 3;

   for (int mb = 0; mb < M.maxM; mb += 16) {           // component tiles of 16 (one pass for M <= 16)
      float acc1[NT][4], acc2[NT][4];                  // sum Lr (o - c), sum Lr (o - c)^2; column D of acc1 = sum Lr
      float cenB[NT];                                  // state centre at "my" right-hand-side column of each tile
      int curS = -1, mo = 0, Mn = 0, Mc = 0;
      bool any = false;

      // moves the accumulated fragments of state curS to the FP64 accumulators (HFB.c:1665-1678, :1724-1736)
      auto flush = [&]() {
         __syncwarp();
#pragma unroll
         for (int nt = 0; nt < NT; nt++) {
            const int col = nt * 8 + 2 * fc;
            fb[fg * FSTR + col] = acc1[nt][0];            fb[fg * FSTR + col + 1] = acc1[nt][1];
            fb[(fg + 8) * FSTR + col] = acc1[nt][2];      fb[(fg + 8) * FSTR + col + 1] = acc1[nt][3];
            fb[fg * FSTR + 8 * NT + col] = acc2[nt][0];   fb[fg * FSTR + 8 * NT + col + 1] = acc2[nt][1];
            fb[(fg + 8) * FSTR + 8 * NT + col] = acc2[nt][2]; fb[(fg + 8) * FSTR + 8 * NT + col + 1] = acc2[nt][3];
         }
         __syncwarp();
         const float *cen = centre + (size_t)curS * S4_CSTR;
         double wsum = 0.0;
         for (int mi = 0; mi < Mc; mi++) {
            const float *row = fb + mi * FSTR;
            const double S0 = (double)row[D];
            if (!(S0 > 0.0)) continue;                 // component never passed the minimum-occupancy rule
            const int g = M.mixGauss[mo + mb + mi];
            const int mId = M.meanId[g], vId = M.varId[g];
#pragma unroll
            for (int h = 0; h < 2; h++) {
               const int k = lane + 32 * h;
               if (k < D) {
                  const double d = (double)M.mean[(size_t)g * Dp + k] - (double)cen[k];
                  const double s1 = (double)row[k], s2 = (double)row[8 * NT + k];
                  if (upM) atomicAdd(&W.acc[M.L.muSum + (size_t)mId * D + k], s1 - d * S0);
                  if (upV) atomicAdd(&W.acc[M.L.vaSum + (size_t)vId * D + k], s2 - 2.0 * d * s1 + d * d * S0);
               }
            }
            if (lane == 0) {
               if (upM) atomicAdd(&W.acc[M.L.muOcc + mId], S0);
               if (upV) atomicAdd(&W.acc[M.L.vaOcc + vId], S0);
               if (upW) atomicAdd(&W.acc[M.L.wtC + mo + mb + mi], S0);
            }
            wsum += S0;
         }
         if (lane == 0 && wsum > 0.0) atomicAdd(&W.acc[M.L.wtOcc + curS], wsum);
         __syncwarp();
      };

      // software prefetch pipeline along the sorted list (long-scoreboard stalls were the largest share): record
      // fields of position it + 3 -> frame numbers of it + 2 -> observation rows of it + 1 into L2
      int tB = -1, cntA = 0;
      long long fB = 0, offA = -1, featA = 0;
      for (int it = i0; it < i1; it++) {
         const PosRec &R = list[it];
         if (it + 1 < i1) prefetch_l1(&list[it + 1]);
         if (tB >= 0) {
            const float *row = W.feat + (size_t)(fB + tB) * D;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(row + 32));
         }
         tB = (offA >= 0 && lane < cntA) ? vbuf[offA + lane].t : -1;
         fB = featA;
         offA = -1;
         if (it + 3 < i1) { offA = list[it + 3].vOff; cntA = min(32, vcnt[it + 3]); featA = list[it + 3].featOff; }
         const int s = R.s;
         if (s != curS) {
            if (any) flush();
            curS = s; any = false;
            mo = M.stateMixOff[s]; Mn = M.stateMixOff[s + 1] - mo; Mc = min(16, Mn - mb);
            const float *cen = centre + (size_t)s * S4_CSTR;
#pragma unroll
            for (int nt = 0; nt < NT; nt++) {
               const int dim = nt * 8 + fg;
               cenB[nt] = (dim < D) ? cen[dim] : 0.f;             // 0 for the ones column and the padding
#pragma unroll
               for (int i = 0; i < 4; i++) { acc1[nt][i] = 0.f; acc2[nt][i] = 0.f; }
            }
            // stage the tile's Gaussians in shared memory: every position of this state reads them from there
            __syncwarp();
            if (Mn > 1)
               for (int e = lane; e < Mc * (Dp >> 2); e += 32) {
                  const int mi = e / (Dp >> 2), k4 = e - mi * (Dp >> 2);
                  const int g = M.mixGauss[mo + mb + mi];
                  *reinterpret_cast<float4 *>(gmu + mi * S5_GSTR + 4 * k4) = *reinterpret_cast<const float4 *>(M.mean + (size_t)g * Dp + 4 * k4);
                  *reinterpret_cast<float4 *>(giv + mi * S5_GSTR + 4 * k4) = *reinterpret_cast<const float4 *>(M.ivar + (size_t)g * Dp + 4 * k4);
               }
            if (lane < 16) {
               const bool on = lane < Mc;
               gwt[lane] = on ? M.mixLogWt[mo + mb + lane] : LZERO_D;
               ggc[lane] = on ? M.gconst[M.mixGauss[mo + mb + lane]] : 0.f;
            }
            __syncwarp();
         }
         if (Mc <= 0) continue;                        // state has no components in this tile
         const int q = R.q, j = R.j, tmin = R.tmin, tmax = R.tmax;
         const int P = R.P, S = R.S, J = R.J, so = R.so;
         const int *ps = R.ps;
         const long long frameBase = R.frameBase;
         const short *sqA = W.sq + frameBase, *eqA = W.eq + frameBase;
         const float *feat = W.feat + (size_t)R.featOff * D;
         StatPos sp;
         const float *A = M.transLogA + R.transOff;
         sp.alphaJ = W.occ + R.alphaOff; sp.aent = W.aent + R.aentOff; sp.betaU = W.beta + R.betaOff;
         sp.bU = W.b + R.bOff; sp.A = A; sp.ps = ps; sp.qLo = W.qLo + frameBase; sp.qHi = W.qHi + frameBase;
         sp.tacc = W.acc + R.trAcc; sp.oacc = W.acc + R.trOcc; sp.pr = R.pr;
         sp.P = P; sp.S = S; sp.Q = R.Q; sp.J = J; sp.T = R.T; sp.N = R.N; sp.so = so;
         sp.sq1 = R.sq1; sp.j = j; sp.q = q;
         sp.aEntJ = A[1 + j]; sp.aExitJ = A[(1 + j) * R.N + R.N - 1]; sp.aTee = A[R.N - 1];
         const double pr = sp.pr;

         // two fronts: the dense list of valid frames written by stats_pre_kernel (chunks of 32 valid frames), or -- for
         // positions that did not fit into the list -- the inline version over 32-frame windows of the span
         const long long vOff = R.vOff;
         const int nV = (vOff >= 0) ? vcnt[it] : 0;
         for (int ch = 0;; ch++) {
            int nT;
            if (vOff >= 0) {
               if (ch * 32 >= nV) break;
               nT = min(32, nV - ch * 32);
               __syncwarp();
               if (lane < nT) { const ValidFrame v = vbuf[vOff + ch * 32 + lane]; ts[lane] = v.t; x0s[lane] = v.x0; }
            } else {
               const int t0 = tmin + 32 * ch;
               if (t0 > tmax) break;
               const int t = t0 + lane;
               const bool inb = (t <= tmax) && q >= sqA[t] && q <= eqA[t];
               double x0 = OCC_SKIP;
               if (inb) {
                  const double aj = sp.alphaJ[(size_t)t * P];
                  const double *bq = sp.betaU + (size_t)t * S + so;
                  const float bjt = sp.bU[(size_t)t * J + ps[j]];
                  const double lg = aj + bq[1 + j] - pr;                              // log occupancy of state j
                  if (!(lg < -(minF + 0.25))) x0 = (Mn == 1) ? lg : lg - (double)bjt;  // :1575-1576 / initx :1480-1489
               }
               if (doTr && mb == 0) stats_tran_chunk(sp, t, inb, lane);
               if (!doMix) continue;
               // ---- frames of this chunk that can contribute to the mixture statistics
               const bool valid = inb && x0 > -1.0e29;
               const unsigned mask = __ballot_sync(0xffffffffu, valid);
               nT = __popc(mask);
               if (nT == 0) continue;
               __syncwarp();
               if (valid) { int idx = __popc(mask & ((1u << lane) - 1)); ts[idx] = t; x0s[idx] = x0; }
            }
            const int nT8 = (nT + 7) & ~7;
            for (int e = lane; e < 16 * S4_LSTR; e += 32) lrs[e] = 0.f;
            __syncwarp();
            // (4-byte cp.async straight into the tile, all rows in flight, no register staging: measured 4 % SLOWER)
            // rows nT..nT8-1 are padding: their occupancies stay zero, so they may hold any finite row (frame ts[0])
            if (lane >= nT && lane < nT8) ts[lane] = ts[0];
            __syncwarp();
            const bool col0 = lane < D, col1 = lane + 32 < D;
            const float pad0 = (lane == D) ? 1.f : 0.f, pad1 = (lane + 32 == D) ? 1.f : 0.f;   // the ones column: sum Lr in phase 2
            for (int tb = 0; tb < nT8; tb += 8) {                  // observation rows, 8 at a time (16 loads in flight per
               float v0[8], v1[8];                                 // lane)
#pragma unroll
               for (int r = 0; r < 8; r++) {
                  const float *src = feat + (size_t)ts[tb + r] * D;
                  v0[r] = col0 ? src[lane] : pad0;
                  v1[r] = col1 ? src[lane + 32] : pad1;
               }
#pragma unroll
               for (int r = 0; r < 8; r++) {
                  float *dst = os + S5_OS(tb + r, 0);
                  if (lane < 8 * NT) dst[2 * lane] = v0[r];
                  if (lane + 32 < 8 * NT) dst[2 * (lane + 32)] = v1[r];
               }
            }
            __syncwarp();
            // ---- phase 1: component posteriors on the FP32 pipe, lanes <-> (component, frame) pairs -> Lr tile.
            //      (Tried twice as a second mma.sync 3xTF32 product about the state centre.  Accumulating in the
            //      fragment: the tensor core's truncating FP32 accumulation biased log N by ~5e-5, past the parity
            //      bound.  With fresh fragments and FP32-pipe running sums: parity fine (wtC 1.4e-5 vs 1.1e-5) and
            //      3.7x fewer instructions for this phase, but the kernel got 9-14 % SLOWER on B200 -- the legacy
            //      HMMA path cannot keep up; the tcgen05 version needs frames gathered per state, see DESIGN.md.)
            unsigned anyLr = 0;
            // lane <-> (component, four frames): the frames share the component's parameter loads and go through the
            // packed FP32 pipe as two independent pairs; rows past nT hold finite data whose results are dropped
            const int nTq = (nT + 3) >> 2;
            const float rnT = 1.0f / (float)nTq;                 // (pi + 0.5) / nTq is never within 1/16 of an integer
            for (int pi = lane; pi < ((Mc * nTq + 31) & ~31); pi += 32) {
               float Lr[4] = {0.f, 0.f, 0.f, 0.f};
               const int mi = (int)(((float)pi + 0.5f) * rnT), tq = pi - mi * nTq, ti = 4 * tq;
               if (mi < Mc) {
                  const float wt = gwt[mi];
                  if (wt > LMINMIX_F) {                                         // HFB.c:1573
                     double x[4];
#pragma unroll
                     for (int j = 0; j < 4; j++) x[j] = (ti + j < nT) ? x0s[ti + j] : 0.0;
                     if (Mn > 1) {
                        const float *mu = gmu + mi * S5_GSTR, *iv = giv + mi * S5_GSTR;
                        const float *o = os + 2 * tq * S5_OPS;
                        f32x2_t sA = f2_pack(ggc[mi], ggc[mi]), sB = sA;
                        // up to Dp = D rounded up to 4: the padding of the parameter rows is zero (inverse variance 0: the
                        // ones column and the zero columns of the tile contribute exactly nothing)
                        for (int k = 0; k < Dp; k += 4) {
                           const float4 m4 = *reinterpret_cast<const float4 *>(mu + k);
                           const float4 v4 = *reinterpret_cast<const float4 *>(iv + k);
                           const ulonglong2 oa = *reinterpret_cast<const ulonglong2 *>(o + 2 * k);
                           const ulonglong2 ob = *reinterpret_cast<const ulonglong2 *>(o + 2 * k + 4);
                           const ulonglong2 oc = *reinterpret_cast<const ulonglong2 *>(o + S5_OPS + 2 * k);
                           const ulonglong2 od = *reinterpret_cast<const ulonglong2 *>(o + S5_OPS + 2 * k + 4);
                           sA = f2_step(sA, oa.x, m4.x, v4.x); sB = f2_step(sB, oc.x, m4.x, v4.x);
                           sA = f2_step(sA, oa.y, m4.y, v4.y); sB = f2_step(sB, oc.y, m4.y, v4.y);
                           sA = f2_step(sA, ob.x, m4.z, v4.z); sB = f2_step(sB, od.x, m4.z, v4.z);
                           sA = f2_step(sA, ob.y, m4.w, v4.w); sB = f2_step(sB, od.y, m4.w, v4.w);
                        }
                        float sv[4];
                        f2_unpack(sA, sv[0], sv[1]); f2_unpack(sB, sv[2], sv[3]);
#pragma unroll
                        for (int j = 0; j < 4; j++) x[j] = (x[j] + (double)wt) + (double)(-0.5f * sv[j]);   // :1581-1599
                     }
#pragma unroll
                     for (int j = 0; j < 4; j++)
                        if (ti + j < nT && -x[j] < minF) Lr[j] = expf((float)x[j]);   // :1606, :1612
                  }
#pragma unroll
                  for (int j = 0; j < 4; j++)
                     if (ti + j < nT) lrs[mi * S4_LSTR + ti + j] = Lr[j];
               }
               anyLr |= __ballot_sync(0xffffffffu, Lr[0] > 0.f || Lr[1] > 0.f || Lr[2] > 0.f || Lr[3] > 0.f);
            }
            __syncwarp();
            if (!anyLr) continue;
            any = true;
            // ---- phase 2: [16 components x nT frames] x [nT frames x (dims | 1)] on the tensor cores
            for (int ks = 0; ks < nT8; ks += 8) {
               uint32_t ah[4], al[4];
               {
                  const float a0 = lrs[fg * S4_LSTR + ks + fc], a1 = lrs[(fg + 8) * S4_LSTR + ks + fc];
                  const float a2 = lrs[fg * S4_LSTR + ks + fc + 4], a3 = lrs[(fg + 8) * S4_LSTR + ks + fc + 4];
                  ah[0] = s4_hi(a0); ah[1] = s4_hi(a1); ah[2] = s4_hi(a2); ah[3] = s4_hi(a3);
                  al[0] = s4_lo(a0); al[1] = s4_lo(a1); al[2] = s4_lo(a2); al[3] = s4_lo(a3);
               }
               const float *o0 = os + S5_OS(ks + fc, 0), *o1 = o0 + 2 * S5_OPS;
#pragma unroll
               for (int nt = 0; nt < NT; nt++) {
                  const float v0 = o0[2 * (nt * 8 + fg)] - cenB[nt], v1 = o1[2 * (nt * 8 + fg)] - cenB[nt];   // column D of the tile holds 1
                  // the tensor core forms the 8-frame partial products in a fresh fragment; the running sums are
                  // kept on the FP32 pipe (round to nearest): its own accumulation truncates, which over the
                  // ~100 MMAs of a flush interval would bias every occupancy sum by several 1e-6
                  uint32_t h0 = s4_hi(v0), h1 = s4_hi(v1), l0 = s4_lo(v0), l1 = s4_lo(v1);
                  float p[4] = {0.f, 0.f, 0.f, 0.f};
                  s4_mma(p, ah, l0, l1); s4_mma(p, al, h0, h1); s4_mma(p, ah, h0, h1);
                  acc1[nt][0] += p[0]; acc1[nt][1] += p[1]; acc1[nt][2] += p[2]; acc1[nt][3] += p[3];
                  const float w0 = v0 * v0, w1 = v1 * v1;
                  h0 = s4_hi(w0); h1 = s4_hi(w1); l0 = s4_lo(w0); l1 = s4_lo(w1);
                  p[0] = p[1] = p[2] = p[3] = 0.f;
                  s4_mma(p, ah, l0, l1); s4_mma(p, al, h0, h1); s4_mma(p, ah, h0, h1);
                  acc2[nt][0] += p[0]; acc2[nt][1] += p[1]; acc2[nt][2] += p[2]; acc2[nt][3] += p[3];
               }
            }
         }
      }
      if (any) flush();
   }
}
