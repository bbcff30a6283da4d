/* hfb_oracle.c -- CPU restatement of HTK 3.4.1's embedded Baum-Welch E-step.
 *
 * TEST INFRASTRUCTURE ONLY (see hfb_oracle.h).  Plain C, sequential, same float /
 * double choices as the reference so that it agrees with the reference's HERest to
 * float rounding.  Every routine cites the reference lines it restates; nothing
 * here is shared with the CUDA product.
 *
 * Conventions: model numbers q = 1..Q and state numbers i = 1..N are 1-based as in
 * HTK; arrays are allocated one element larger and slot 0 is unused.  Where the
 * reference stores NULL for a beta vector outside the beam this code tests the
 * final beam limits qLo[t]..qHi[t] instead (equivalent: every vector outside the
 * final beam is set to NULL at HFB.c:1254-1267).
 */
#include "hfb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <pthread.h>

#define LZERO   HFB_LZERO
#define LSMALL  HFB_LSMALL
#define MINEARG HFB_MINEARG

/* ------------------------------------------------------------------ log add */

/* HMath.c:1576-1590 with minLogExp = -log(-LZERO) (HMath.c:1680) */
static double ladd(double x, double y)
{
   static const double minLogExp = -23.025850929940457; /* -log(1e10) */
   double d;
   if (x < y) { d = x; x = y; y = d; }
   d = y - x;
   if (d < minLogExp) return (x < LSMALL) ? LZERO : x;
   return x + log(1.0 + exp(d));
}

/* ------------------------------------------------------------------ layout */

static void layout_of(const hfb_model *m, hfb_acc_layout *L)
{
   int64_t o = 0, nn = 0, n = 0, i;
   for (i = 0; i < m->numTrans; i++) { nn += (int64_t)m->transN[i] * m->transN[i]; n += m->transN[i]; }
   L->tran = o;    o += nn;
   L->tranOcc = o; o += n;
   L->wtC = o;     o += m->stateMixOff[m->numStates];
   L->wtOcc = o;   o += m->numStates;
   L->muSum = o;   o += (int64_t)m->numMeanAcc * m->vecSize;
   L->muOcc = o;   o += m->numMeanAcc;
   L->vaSum = o;   o += (int64_t)m->numVarAcc * m->vecSize;
   L->vaOcc = o;   o += m->numVarAcc;
   L->numEgs = o;  o += m->numHmm;
   L->totalT = o++; L->totalPr = o++; L->numOk = o++; L->numSkipped = o++;
   L->count = o; L->tranOccStride = 0;
}

/* ------------------------------------------------------------------ min durations */

/* FindStateOrder, HFB.c:91-102: post-order numbering of a depth-first walk over
 * predecessors, started from the exit state.                                      */
static void state_order(const float *A, int N, int *ord, int s, int *cnt)
{
   int p;
   ord[s] = 0;
   for (p = 1; p < N; p++)
      if (A[(p - 1) * N + (s - 1)] > LSMALL && p != s && ord[p] < 0)
         state_order(A, N, ord, p, cnt);
   ord[s] = ++(*cnt);
}

/* SetMinDurs, HFB.c:106-155 */
static int min_dur_of(const float *A, int N)
{
   int *md = (int *)malloc(sizeof(int) * (N + 1) * 2), *so = md + N + 1;
   int i, j, k, cnt = 0, d, r;
   for (i = 1; i <= N; i++) so[i] = md[i] = -1;
   state_order(A, N, md, N, &cnt);
   for (i = 1; i <= N; i++) if (md[i] > 0) so[md[i]] = i;
   for (i = 1; i <= N; i++) md[i] = N;
   md[1] = 0;
   for (k = 1; k <= cnt; k++) {
      i = so[k];
      if (i < 1 || i > N) continue;
      for (j = 1; j < N; j++)
         if (A[(j - 1) * N + (i - 1)] > LSMALL) {
            d = md[j] + ((i == N) ? 0 : 1);
            if (d < md[i]) md[i] = d;
         }
   }
   if (md[N] < 0 || md[N] >= N)        /* HFB.c:144-149 "discontinuity": under-estimate */
      r = (A[0 * N + (N - 1)] > LSMALL) ? 0 : 1;
   else
      r = md[N];
   free(md);
   return r;
}

int hfbo_min_durs(const hfb_model *m, int32_t *out)
{
   int i;
   for (i = 0; i < m->numTrans; i++)
      out[i] = min_dur_of(m->transLogA + m->transOff[i], m->transN[i]);
   return HFB_OK;
}

/* ------------------------------------------------------------------ context */

typedef struct {
   const hfb_model *m;       /* the set that aligns (al_hset); the only set unless two-model re-estimation is on */
   const hfb_model *up;      /* the set whose statistics are collected (up_hset, HFB.c:253, :331)               */
   int twoModels;            /* UseAlignHMMSet, HFB.c:296-333                                                     */
   int compLevelMismatch;    /* ALIGNCOMPLEVEL with different component counts (fatal HError 999 in the reference)         */
   hfb_options opt;
   hfb_acc_layout L;
   int accDouble;
   void *acc;
   int maxM;
   int32_t *minDur;          /* per transition matrix */
   int64_t *tranAccOff;      /* per transition matrix: offset of tran[][] in acc */
   int64_t *tranOccOff;
   /* two-entry output-probability cache per tied state (WtAcc.time/prob, HFB.c:910-912) */
   int *stamp;               /* [2][J] */
   float *probs;             /* [2][sumM + J]: per state vector [0..M] */
   int64_t *probOff;         /* [J] */
   int epoch;
   float *occOut;            /* optional [T][P] occupancies of the current utterance */
   double totalPr;           /* LogDouble totalPr, HERest.c:131 (cast to float only in the dump) */
} Ctx;

static void acc_add(Ctx *c, int64_t idx, double v)
{
   /* "f += d" on a float accumulator is (float)((double)f + d) in C */
   if (c->accDouble) ((double *)c->acc)[idx] += v;
   else { float *a = (float *)c->acc; a[idx] = (float)((double)a[idx] + v); }
}

static int ctx_init(Ctx *c, const hfb_model *m, const hfb_options *opt, void *acc, int accDouble)
{
   int j, i;
   int64_t o;
   memset(c, 0, sizeof(*c));
   c->up = m; c->opt = *opt; c->acc = acc; c->accDouble = accDouble;
   layout_of(m, &c->L);                                   /* accumulators belong to the update set */
   if (opt->alignModel) {
      m = opt->alignModel; c->twoModels = 1;
      c->opt.uFlags &= ~HFB_UPTRANS;                      /* HFB.c:313-316 */
   }
   c->m = m;
   c->maxM = 0;
   for (j = 0; j < m->numStates; j++) {
      int M = m->stateMixOff[j + 1] - m->stateMixOff[j];
      if (M > c->maxM) c->maxM = M;
   }
   c->minDur = (int32_t *)malloc(sizeof(int32_t) * (m->numTrans + 1));
   hfbo_min_durs(m, c->minDur);
   c->tranAccOff = (int64_t *)malloc(sizeof(int64_t) * (m->numTrans + 1) * 2);
   c->tranOccOff = c->tranAccOff + m->numTrans + 1;
   {
      int64_t a = c->L.tran, b = c->L.tranOcc;
      for (i = 0; i < m->numTrans; i++) {
         c->tranAccOff[i] = a; c->tranOccOff[i] = b;
         a += (int64_t)m->transN[i] * m->transN[i]; b += m->transN[i];
      }
   }
   c->stamp = (int *)malloc(sizeof(int) * 2 * (m->numStates + 1));
   c->probOff = (int64_t *)malloc(sizeof(int64_t) * (m->numStates + 1));
   for (o = 0, j = 0; j < m->numStates; j++) {
      c->probOff[j] = o; o += (m->stateMixOff[j + 1] - m->stateMixOff[j]) + 1;
   }
   c->probs = (float *)malloc(sizeof(float) * 2 * (o + 1));
   c->probOff[m->numStates] = o;
   for (j = 0; j < 2 * m->numStates; j++) c->stamp[j] = -1;
   c->epoch = 0;
   return HFB_OK;
}

static void ctx_free(Ctx *c)
{
   free(c->minDur); free(c->tranAccOff); free(c->stamp); free(c->probOff); free(c->probs);
}

/* ------------------------------------------------------------------ output probs */

/* IDOutP, HModel.c:5420-5431: float, sequential, starting from gConst */
static float gauss_logp(const hfb_model *m, int g, const float *x)
{
   const float *mu = m->mean + (size_t)g * m->vecSize, *iv = m->ivar + (size_t)g * m->vecSize;
   float sum = m->gConst[g], d;
   int k;
   for (k = 0; k < m->vecSize; k++) { d = x[k] - mu[k]; sum += d * d * iv[k]; }
   return (float)(-0.5 * sum);
}

/* ShStrP, HFB.c:898-988 (plain branches: M==1 at :917-928, M>1 at :949-960).
 * v[0] = state log prob, v[1..M] = component log densities (LZERO if skipped).    */
static void state_logp_vec(const hfb_model *m, int s, const float *x, float *v)
{
   int o = m->stateMixOff[s], M = m->stateMixOff[s + 1] - o, k;
   float acc, mixp, wt;
   if (M == 1) { v[0] = gauss_logp(m, m->mixGauss[o], x); v[1] = v[0]; return; }
   acc = (float)LZERO;
   for (k = 1; k <= M; k++) v[k] = (float)LZERO;          /* NewOtprobVec, HFB.c:883-895 */
   for (k = 0; k < M; k++) {
      wt = m->mixLogWt[o + k];
      if (wt > HFB_LMINMIX) {
         mixp = gauss_logp(m, m->mixGauss[o + k], x);
         acc = (float)ladd((double)acc, (double)(float)(wt + mixp));  /* x is LogFloat, :907 */
         v[k + 1] = mixp;
      }
   }
   v[0] = acc;
}

/* cached access; `t` doubles as the time stamp, slot = t & 1 */
static const float *outp(Ctx *c, int s, const float *feat, int t)
{
   int slot = t & 1, J = c->m->numStates;
   float *v = c->probs + (size_t)slot * (c->probOff[J] + 1) + c->probOff[s];
   if (c->stamp[slot * J + s] != t) {
      state_logp_vec(c->m, s, feat + (size_t)(t - 1) * c->m->vecSize, v);
      c->stamp[slot * J + s] = t;
   }
   return v;
}

static void outp_reset(Ctx *c)
{
   int j;
   for (j = 0; j < 2 * c->m->numStates; j++) c->stamp[j] = -1;
}

int hfbo_state_loglik(const hfb_model *m, const float *feat, int32_t T,
                      const int32_t *states, int32_t n, float *out, float *mixOut)
{
   int t, i, k, maxM = 0;
   int64_t mo, sumM = 0;
   float *v;
   for (i = 0; i < n; i++) {
      int M = m->stateMixOff[states[i] + 1] - m->stateMixOff[states[i]];
      if (M > maxM) maxM = M;
      sumM += M;
   }
   v = (float *)malloc(sizeof(float) * (maxM + 2));
   for (t = 0; t < T; t++) {
      mo = 0;
      for (i = 0; i < n; i++) {
         int M = m->stateMixOff[states[i] + 1] - m->stateMixOff[states[i]];
         state_logp_vec(m, states[i], feat + (size_t)t * m->vecSize, v);
         out[(size_t)t * n + i] = v[0];
         if (mixOut) for (k = 0; k < M; k++) mixOut[(size_t)t * sumM + mo + k] = v[k + 1];
         mo += M;
      }
   }
   free(v);
   return HFB_OK;
}

/* ------------------------------------------------------------------ one utterance */

typedef struct {
   int T, Q;
   const float *feat;
   const int32_t *lab;     /* al_qList: physical HMMs of the aligning set */
   const int32_t *labUp;   /* up_qList (HFB.c:521-531): the same array unless two-model re-estimation is on */
   int *N;            /* [Q+2] states per model */
   const float **A;   /* [Q+2] transition logs (row-major N*N, 0-based) */
   int *dms;          /* [Q+2] minimum durations (qDms) */
   int *off;          /* [Q+2] offset of model q's states in a column */
   int S;             /* states per column */
   int *poff;         /* [Q+2] offset of model q's emitting states in an occupancy row */
   int P;
   short *qLo, *qHi;  /* [T+2] */
   double *beta;      /* [T+1][S] */
   double *alpha, *alpha1;   /* [S] each */
   double *alphaBase;
   double *maxP;      /* [Q+2] */
} Utt;

#define TR(u, q, i, j)   ((u)->A[q][((i) - 1) * (u)->N[q] + ((j) - 1)])
#define BETA(u, t, q, i) ((u)->beta[(size_t)(t) * (u)->S + (u)->off[q] + (i) - 1])
#define AL(u, q, i)      ((u)->alpha[(u)->off[q] + (i) - 1])
#define AL1(u, q, i)     ((u)->alpha1[(u)->off[q] + (i) - 1])

static int in_beam(const Utt *u, int t, int q)
{
   return q >= 1 && q <= u->Q && q >= u->qLo[t] && q <= u->qHi[t];
}

/* log b_j(o_t) of emitting state j (2..N-1) of model q */
static float bprob(Ctx *c, const Utt *u, int q, int j, int t)
{
   int p = u->lab[q - 1];
   int s = c->m->hmmState[c->m->hmmStateOff[p] + (j - 2)];
   return outp(c, s, u->feat, t)[0];
}

/* SetBeamTaper, HFB.c:1116-1145 */
static void beam_taper(Utt *u)
{
   int q, dq, i, t, Q = u->Q, T = u->T;
   q = 1; dq = u->dms[q]; i = 0;
   for (t = 1; t <= T; t++) {
      while (i == dq) { i = 0; if (q < Q) { q++; dq = u->dms[q]; } else dq = -1; }
      u->qHi[t] = (short)q; i++;
   }
   q = Q; dq = u->dms[q]; i = 0;
   for (t = T; t >= 1; t--) {
      while (i == dq) { i = 0; if (q > 1) { q--; dq = u->dms[q]; } else dq = -1; }
      u->qLo[t] = (short)q; i++;
   }
}

/* SetBeta, HFB.c:1149-1296.  Returns pr or LZERO on failure; *err set on HError 7323 */
static double beta_pass(Ctx *c, Utt *u, double thresh, int *err)
{
   int Q = u->Q, T = u->T, t, q, i, j, Nq, lNq = 0, startq, endq, lastq = Q;
   double x, y, a, a1N = 0.0, gMax, lMax, pr;

   /* t = T, HFB.c:1176-1198 */
   u->qHi[T] = (short)Q; endq = u->qLo[T];
   for (q = Q; q >= endq; q--) {
      Nq = u->N[q];
      BETA(u, T, q, Nq) = (q == Q) ? 0.0 : BETA(u, T, q + 1, lNq) + a1N;
      for (i = 2; i < Nq; i++) BETA(u, T, q, i) = TR(u, q, i, Nq) + BETA(u, T, q, Nq);
      x = LZERO;
      for (j = 2; j < Nq; j++) {
         a = TR(u, q, 1, j); y = BETA(u, T, q, j);
         if (a > LSMALL && y > LSMALL) x = ladd(x, a + bprob(c, u, q, j, T) + y);
      }
      BETA(u, T, q, 1) = x;
      lNq = Nq; a1N = TR(u, q, 1, Nq); lastq = q;
   }

   /* t = T-1 .. 1, HFB.c:1205-1277 */
   for (t = T - 1; t >= 1; t--) {
      int lo1 = u->qLo[t + 1], hi1 = u->qHi[t + 1];
      gMax = LZERO;
      startq = hi1;
      endq = (lo1 == 1) ? 1 : ((u->qLo[t] >= lo1) ? u->qLo[t] : lo1 - 1);
      while (endq > 1 && u->dms[endq - 1] == 0) endq--;
      for (q = startq; q >= endq; q--) {
         int inner = (q >= lo1 && q <= hi1);
         lMax = LZERO; Nq = u->N[q];
         /* exit state: entry of the next model one frame later, plus the path
          * through a following tee model in the same frame (:1225-1227)         */
         x = (q < Q && q + 1 >= lo1 && q + 1 <= hi1) ? BETA(u, t + 1, q + 1, 1) : LZERO;
         if (q < startq && a1N > LSMALL) x = ladd(x, BETA(u, t, q + 1, lNq) + a1N);
         BETA(u, t, q, Nq) = x;
         for (i = Nq - 1; i > 1; i--) {
            x = TR(u, q, i, Nq) + BETA(u, t, q, Nq);
            if (inner)
               for (j = 2; j < Nq; j++) {
                  a = TR(u, q, i, j); y = BETA(u, t + 1, q, j);
                  if (a > LSMALL && y > LSMALL) x = ladd(x, a + bprob(c, u, q, j, t + 1) + y);
               }
            BETA(u, t, q, i) = x;
            if (x > lMax) lMax = x;
            if (x > gMax) gMax = x;
         }
         x = LZERO;
         for (j = 2; j < Nq; j++) {
            a = TR(u, q, 1, j); y = BETA(u, t, q, j);
            if (a > LSMALL && y > LSMALL) x = ladd(x, a + bprob(c, u, q, j, t) + y);
         }
         BETA(u, t, q, 1) = x;
         u->maxP[q] = lMax;
         lNq = Nq; a1N = TR(u, q, 1, Nq); lastq = q;
      }
      /* beam pruning, HFB.c:1254-1272 */
      while (gMax - u->maxP[startq] > thresh) { if (--startq < 1) { *err = HFB_UTT_EBETA; return LZERO; } }
      while (u->qHi[t] < startq)              { if (--startq < 1) { *err = HFB_UTT_EBETA; return LZERO; } }
      u->qHi[t] = (short)startq;
      while (gMax - u->maxP[endq] > thresh) { if (++endq > startq) return LZERO; }
      u->qLo[t] = (short)endq;
   }
   pr = BETA(u, 1, lastq, 1);   /* utt->pr = bqt[1] of the last vector computed (:1280) */
   if (pr <= LSMALL) return LZERO;
   return pr;
}

/* MaxModelProb, HFB.c:655-682 (alpha = column t, passed in as `al`) */
static double max_model_prob(const Utt *u, const double *al, int q, int t, int minq)
{
   double maxP, x;
   int qx, i, Nq;
   if (q == 1) maxP = LZERO;
   else {
      int Nq1 = u->N[q - 1];
      maxP = in_beam(u, t, q - 1) ? al[u->off[q - 1] + Nq1 - 1] + BETA(u, t, q - 1, Nq1) : LZERO;
      for (qx = q - 1; qx > minq && TR(u, qx, 1, u->N[qx]) > LSMALL; qx--) {
         int qx1 = qx - 1, N1 = u->N[qx1];
         x = in_beam(u, t, qx1) ? al[u->off[qx1] + N1 - 1] + BETA(u, t, qx1, N1) : LZERO;
         if (x > maxP) maxP = x;
      }
   }
   Nq = u->N[q];
   if (in_beam(u, t, q))
      for (i = 1; i < Nq; i++) {
         x = al[u->off[q] + i - 1] + BETA(u, t, q, i);
         if (x > maxP) maxP = x;
      }
   return maxP;
}

static void zero_alpha(Utt *u, int qlo, int qhi)
{
   int q, i;
   for (q = qlo; q <= qhi; q++) for (i = 1; i <= u->N[q]; i++) AL(u, q, i) = LZERO;
}

/* InitAlpha, HFB.c:616-651 */
static void init_alpha(Ctx *c, Utt *u, int *start, int *end)
{
   int q, i, j, Nq, eq = u->qHi[1];
   double x, a, a1N = 0.0;
   for (q = 1; q <= eq; q++) {
      Nq = u->N[q];
      AL(u, q, 1) = (q == 1) ? 0.0 : AL(u, q - 1, 1) + a1N;
      for (j = 2; j < Nq; j++) {
         a = TR(u, q, 1, j);
         AL(u, q, j) = (a > LSMALL) ? AL(u, q, 1) + a + bprob(c, u, q, j, 1) : LZERO;
      }
      x = LZERO;
      for (i = 2; i < Nq; i++) { a = TR(u, q, i, Nq); if (a > LSMALL) x = ladd(x, AL(u, q, i) + a); }
      AL(u, q, Nq) = x;
      a1N = TR(u, q, 1, Nq);
   }
   zero_alpha(u, eq + 1, u->Q);
   *start = 1; *end = eq;
}

/* StepAlpha, HFB.c:686-784.  Returns 0 or HFB_UTT_EALPHA. */
static int step_alpha(Ctx *c, Utt *u, int t, int *start, int *end, double pr)
{
   int sq, eq, i, j, q, Nq, Q = u->Q;
   double x = 0.0, y, a, a1N = 0.0, *tmp;
   double minF = c->opt.minFrwdP;

   sq = u->qLo[t - 1];
   while (pr - max_model_prob(u, u->alpha, sq, t - 1, sq) > minF) {
      ++sq;
      if (sq > u->qHi[t]) return HFB_UTT_EALPHA;
   }
   if (sq < u->qLo[t]) sq = u->qLo[t];
   eq = u->qHi[t - 1] < Q ? u->qHi[t - 1] + 1 : u->qHi[t - 1];
   while (pr - max_model_prob(u, u->alpha, eq, t - 1, sq) > minF) {
      --eq;
      if (eq < sq) return HFB_UTT_EALPHA;
   }
   while (eq < Q && u->dms[eq] == 0) eq++;
   if (eq > u->qHi[t]) eq = u->qHi[t];

   tmp = u->alpha1; u->alpha1 = u->alpha; u->alpha = tmp;
   if (sq > 1) zero_alpha(u, 1, sq - 1);
   for (q = sq; q <= eq; q++) {
      Nq = u->N[q];
      if (q == 1) AL(u, q, 1) = LZERO;
      else {
         AL(u, q, 1) = AL1(u, q - 1, u->N[q - 1]);
         if (q > sq && a1N > LSMALL) AL(u, q, 1) = ladd(AL(u, q, 1), AL(u, q - 1, 1) + a1N);
      }
      for (j = 2; j < Nq; j++) {
         a = TR(u, q, 1, j);
         x = (a > LSMALL) ? a + AL(u, q, 1) : LZERO;
         for (i = 2; i < Nq; i++) {
            a = TR(u, q, i, j); y = AL1(u, q, i);
            if (a > LSMALL && y > LSMALL) x = ladd(x, y + a);
         }
         AL(u, q, j) = x + bprob(c, u, q, j, t);
      }
      x = LZERO;
      for (i = 2; i < Nq; i++) {
         a = TR(u, q, i, Nq); y = AL(u, q, i);
         if (a > LSMALL && y > LSMALL) x = ladd(x, y + a);
      }
      AL(u, q, Nq) = x; a1N = TR(u, q, 1, Nq);
   }
   if (eq < Q) zero_alpha(u, eq + 1, Q);
   *start = sq; *end = eq;
   return 0;
}

/* SetOcct (HFB.c:399-418), UpMixParms (:1426-1744, plain single-stream branches)
 * and UpTranParms (:1371-1423) for model q at time t.                            */
static void accumulate_model(Ctx *c, Utt *u, int t, int q, double pr)
{
   const hfb_model *m = c->up;                           /* UpMixParms / UpTranParms get up_hmm, HFB.c:1802-1805 */
   int N = u->N[q], i, j, k, mx, D = m->vecSize, p = u->labUp[q - 1];
   int hasB = in_beam(u, t, q);                          /* always true inside the alpha beam */
   int hasB1 = (t < u->T) && in_beam(u, t + 1, q);       /* bqt1 != NULL */
   int hasBq1 = (q < u->Q) && in_beam(u, t, q + 1);      /* bq1t != NULL */
   int uf = c->opt.uFlags;
   int trId = m->hmmTrans[p];
   float occt[64];
   double x;
   const float *o = u->feat + (size_t)(t - 1) * D;
   (void)hasB;

   /* SetOcct */
   for (i = 1; i <= N; i++) {
      x = AL(u, q, i) + BETA(u, t, q, i);
      if (i == 1 && hasBq1 && TR(u, q, 1, N) > LSMALL)
         x = ladd(x, AL(u, q, 1) + BETA(u, t, q + 1, 1) + TR(u, q, 1, N));
      x -= pr;
      occt[i] = (x > MINEARG) ? (float)exp(x) : 0.0f;
      if (c->occOut && i > 1 && i < N)
         c->occOut[(size_t)(t - 1) * u->P + u->poff[q] + (i - 2)] = occt[i];
   }

   if (uf & (HFB_UPMEANS | HFB_UPVARS | HFB_UPMIXES)) {
      for (j = 2; j < N; j++) {
         int s = m->hmmState[m->hmmStateOff[p] + (j - 2)];
         int mo = m->stateMixOff[s], M = m->stateMixOff[s + 1] - mo;
         const float *ov = c->twoModels ? NULL : outp(c, s, u->feat, t);
         double initx = LZERO, steSumLr = 0.0, Lr;
         float a, norm = 0.0f, comp_prob[256];
         if (c->maxM > 1) {                                       /* :1480-1489 */
            initx = TR(u, q, 1, j) + AL(u, q, 1);
            if (t > 1)
               for (i = 2; i < N; i++) {
                  a = TR(u, q, i, j);
                  if (a > LSMALL) initx = ladd(initx, AL1(u, q, i) + a);
               }
            initx += BETA(u, t, q, j) - pr;
         }
         if (c->twoModels) {                                      /* component probs of the update hmm, :1518-1547 */
            /* ALIGNCOMPLEVEL (:1521-1530): ... of the ALIGNMENT hmm's state j instead (same number of components) */
            const hfb_model *cm = m;
            int cmo = mo;
            if (c->opt.flags & HFB_OPT_ALIGN_COMP_LEVEL) {
               int sa = c->m->hmmState[c->m->hmmStateOff[u->lab[q - 1]] + (j - 2)];
               cm = c->m; cmo = cm->stateMixOff[sa];
               if (cm->stateMixOff[sa + 1] - cmo != M) { c->compLevelMismatch = 1; continue; }   /* HError 999, :1524 */
            }
            if (M > 255) M = 255;
            norm = (float)LZERO;
            for (mx = 1; mx <= M; mx++) {
               comp_prob[mx] = cm->mixLogWt[cmo + mx - 1] + gauss_logp(cm, cm->mixGauss[cmo + mx - 1], o);
               norm = (float)ladd((double)norm, (double)comp_prob[mx]);
            }
         }
         for (mx = 1; mx <= M; mx++) {
            float wght = m->mixLogWt[mo + mx - 1];
            int g = m->mixGauss[mo + mx - 1];
            if (!(wght > HFB_LMINMIX)) continue;                  /* :1573 */
            if (M == 1) x = AL(u, q, j) + BETA(u, t, q, j) - pr; /* :1575-1576 */
            else if (c->twoModels) x = comp_prob[mx] + AL(u, q, j) + BETA(u, t, q, j) - pr - norm;   /* :1577-1578 */
            else { x = initx + wght; x += ov[mx]; }              /* :1581-1599 */
            if (-x < c->opt.minFrwdP) {                           /* :1606 */
               const float *mean = m->mean + (size_t)g * D;
               Lr = exp(x);
               steSumLr += Lr;
               if ((uf & HFB_UPMEANS) && (uf & HFB_UPVARS)) {     /* :1665-1678 */
                  int64_t mu0 = c->L.muSum + (int64_t)m->meanId[g] * D;
                  int64_t va0 = c->L.vaSum + (int64_t)m->varId[g] * D;
                  acc_add(c, c->L.muOcc + m->meanId[g], Lr);
                  acc_add(c, c->L.vaOcc + m->varId[g], Lr);
                  for (k = 0; k < D; k++) {
                     float zmean = o[k] - mean[k];
                     float zmeanlr = (float)(zmean * Lr);
                     acc_add(c, mu0 + k, zmeanlr);
                     acc_add(c, va0 + k, (float)(zmean * zmeanlr));
                  }
               } else if (uf & HFB_UPMEANS) {                     /* :1693-1699 */
                  int64_t mu0 = c->L.muSum + (int64_t)m->meanId[g] * D;
                  acc_add(c, c->L.muOcc + m->meanId[g], Lr);
                  for (k = 0; k < D; k++) acc_add(c, mu0 + k, (float)(o[k] - mean[k]) * Lr);
               } else if (uf & HFB_UPVARS) {                      /* :1700-1709 */
                  int64_t va0 = c->L.vaSum + (int64_t)m->varId[g] * D;
                  acc_add(c, c->L.vaOcc + m->varId[g], Lr);
                  for (k = 0; k < D; k++) {
                     float zmean = o[k] - mean[k];
                     acc_add(c, va0 + k, (float)(zmean * zmean) * Lr);
                  }
               }
               if (uf & HFB_UPMIXES) acc_add(c, c->L.wtC + mo + mx - 1, Lr);   /* :1724-1725 */
            }
         }
         acc_add(c, c->L.wtOcc + s, steSumLr);                    /* :1736 */
      }
   }

   if (uf & HFB_UPTRANS) {                                        /* UpTranParms */
      int64_t t0 = c->tranAccOff[trId], o0 = c->tranOccOff[trId];
      for (i = 1; i < N; i++) acc_add(c, o0 + i - 1, occt[i]);
      for (i = 1; i < N; i++)
         for (j = 2; j <= N; j++) {
            double ai = TR(u, q, i, j);
            if (i == 1 && j < N) {
               x = AL(u, q, 1) + ai + bprob(c, u, q, j, t) + BETA(u, t, q, j) - pr;
               if (x > MINEARG) acc_add(c, t0 + (i - 1) * N + (j - 1), exp(x));
            } else if (i > 1 && j < N && hasB1) {
               x = AL(u, q, i) + ai + bprob(c, u, q, j, t + 1) + BETA(u, t + 1, q, j) - pr;
               if (x > MINEARG) acc_add(c, t0 + (i - 1) * N + (j - 1), exp(x));
            } else if (i > 1 && j == N) {
               x = AL(u, q, i) + ai + BETA(u, t, q, N) - pr;
               if (x > MINEARG) acc_add(c, t0 + (i - 1) * N + (N - 1), exp(x));
            }
            if (i == 1 && j == N && ai > LSMALL && hasBq1) {
               x = AL(u, q, 1) + ai + BETA(u, t, q + 1, 1) - pr;
               if (x > MINEARG) acc_add(c, t0 + (i - 1) * N + (N - 1), exp(x));
            }
         }
   }
}

/* FBFile = StepBack (HFB.c:1321-1366) + StepForward (:1752-1810) */
static void fb_utt(Ctx *c, const float *feat, int T, const int32_t *lab, const int32_t *labUp, int Q,
                   hfb_utt_result *r, int16_t *bLo, int16_t *bHi, int16_t *aLo, int16_t *aHi)
{
   const hfb_model *m = c->m;
   Utt u;
   int q, t, qt = 0, S = 0, P = 0, err = 0, start, end, maxN = 0;
   double thresh, lbeta = LZERO;

   memset(&u, 0, sizeof(u));
   r->status = HFB_UTT_OK; r->retries = 0; r->pr = LZERO; r->pruneThresh = c->opt.pruneInit;
   u.T = T; u.Q = Q; u.feat = feat; u.lab = lab; u.labUp = labUp ? labUp : lab;
   u.N = (int *)calloc((size_t)(Q + 2) * 4, sizeof(int));
   u.dms = u.N + (Q + 2); u.off = u.dms + (Q + 2); u.poff = u.off + (Q + 2);
   u.A = (const float **)calloc(Q + 2, sizeof(float *));
   /* CreateInsts, HFB.c:508-574 */
   for (q = 1; q <= Q; q++) {
      int p = lab[q - 1], tr = m->hmmTrans[p];
      u.N[q] = m->hmmNumStates[p];
      u.A[q] = m->transLogA + m->transOff[tr];
      u.dms[q] = c->minDur[tr];
      u.off[q] = S; S += u.N[q];
      u.poff[q] = P; P += u.N[q] - 2;
      if (u.N[q] > maxN) maxN = u.N[q];
      qt += u.dms[q];
      if (q > 1 && u.dms[q] == 0 && u.dms[q - 1] == 0) err = HFB_UTT_ETEE;
      if (c->twoModels && c->up->hmmNumStates[u.labUp[q - 1]] != u.N[q]) err = HFB_EINVAL;   /* HError 999, :549-551 */
      else if (c->twoModels && (c->opt.flags & HFB_OPT_ALIGN_COMP_LEVEL)) {                  /* HError 999, :1523-1524 */
         int j, pu = u.labUp[q - 1];
         for (j = 0; j < u.N[q] - 2; j++) {
            int sa = m->hmmState[m->hmmStateOff[p] + j], su = c->up->hmmState[c->up->hmmStateOff[pu] + j];
            if (m->stateMixOff[sa + 1] - m->stateMixOff[sa] != c->up->stateMixOff[su + 1] - c->up->stateMixOff[su]) err = HFB_EINVAL;
         }
      }
   }
   u.S = S; u.P = P;
   if (Q < 1 || u.dms[1] == 0 || u.dms[Q] == 0) err = HFB_UTT_ETEE;
   if (err || maxN > 62) { r->status = err ? err : HFB_UTT_ETEE; goto done; }

   u.qLo = (short *)calloc((size_t)(T + 2) * 2, sizeof(short)); u.qHi = u.qLo + (T + 2);
   u.maxP = (double *)calloc(Q + 2, sizeof(double));
   u.beta = (double *)malloc(sizeof(double) * (size_t)(T + 1) * S);
   u.alphaBase = u.alpha = (double *)malloc(sizeof(double) * (size_t)S * 2); u.alpha1 = u.alpha + S;

   /* StepBack retry loop */
   thresh = c->opt.pruneInit;
   for (;;) {
      if (qt > T) { r->status = HFB_UTT_SKIPPED; goto done; }          /* :1339-1343 */
      outp_reset(c);                                                  /* ResetHMMWtAccs, :559-562 */
      beam_taper(&u);
      r->pruneThresh = thresh;
      lbeta = beta_pass(c, &u, thresh, &err);
      if (err) { r->status = err; goto done; }
      if (lbeta > LSMALL) break;
      thresh += c->opt.pruneInc;
      if (thresh > c->opt.pruneLim || c->opt.pruneInc == 0.0) { r->status = HFB_UTT_SKIPPED; goto done; }
      r->retries++;
   }
   r->pr = lbeta;

   /* StepForward */
   init_alpha(c, &u, &start, &end);
   for (q = 1; q <= Q; q++) acc_add(c, c->L.numEgs + u.labUp[q - 1], 1.0);  /* up_hmm->hook, :1768-1772 */
   for (t = 1; t <= T; t++) {
      if (t > 1) {
         err = step_alpha(c, &u, t, &start, &end, lbeta);
         if (err) { r->status = err; goto done; }   /* reference: fatal HError 7390 */
      }
      if (aLo) { aLo[t - 1] = (int16_t)start; aHi[t - 1] = (int16_t)end; }
      for (q = start; q <= end; q++) accumulate_model(c, &u, t, q, lbeta);
   }
   acc_add(c, c->L.totalT, (double)T);                                  /* HERest.c:779-780 */
   c->totalPr += lbeta;
   acc_add(c, c->L.numOk, 1.0);

done:
   if (r->status == HFB_UTT_SKIPPED) acc_add(c, c->L.numSkipped, 1.0);
   if (bLo && u.qLo) for (t = 1; t <= T; t++) { bLo[t - 1] = u.qLo[t]; bHi[t - 1] = u.qHi[t]; }
   free(u.N); free((void *)u.A); free(u.qLo); free(u.maxP); free(u.beta); free(u.alphaBase);
}

/* utterance-parallel driver for the multi-core CPU baseline (not a reference feature:
 * the reference parallelises with N processes + a file merge, HERest.c:366-367,514-521) */
typedef struct {
   const hfb_model *m; const hfb_options *opt; const hfb_batch *b;
   hfb_utt_result *res; const hfb_beams *beams;
   int next; int64_t count; pthread_mutex_t mu;
} Work;
typedef struct { Work *w; double *acc; } WorkArg;

static void *worker(void *p)
{
   WorkArg *a = (WorkArg *)p;
   Work *w = a->w;
   const hfb_batch *b = w->b;
   const hfb_beams *beams = w->beams;
   int D = w->m->vecSize, u;
   Ctx c;
   a->acc = (double *)calloc(w->count, sizeof(double));
   ctx_init(&c, w->m, w->opt, a->acc, 1);
   for (;;) {
      pthread_mutex_lock(&w->mu); u = w->next++; pthread_mutex_unlock(&w->mu);
      if (u >= b->numUtt) break;
      {
         int64_t f0 = b->frameOff[u];
         int T = (int)(b->frameOff[u + 1] - f0), Q = b->labOff[u + 1] - b->labOff[u];
         fb_utt(&c, b->feat + (size_t)f0 * D, T, (b->labAlign ? b->labAlign : b->lab) + b->labOff[u],
                b->lab + b->labOff[u], Q, &w->res[u],
                beams && beams->qLo ? beams->qLo + f0 : NULL, beams && beams->qHi ? beams->qHi + f0 : NULL,
                beams && beams->sq ? beams->sq + f0 : NULL, beams && beams->eq ? beams->eq + f0 : NULL);
      }
   }
   a->acc[c.L.totalPr] += c.totalPr;
   ctx_free(&c);
   return NULL;
}

/* ------------------------------------------------------------------ public */

int hfbo_accumulate(const hfb_model *m, const hfb_options *opt, const hfb_batch *b,
                    hfb_utt_result *res, const hfb_beams *beams,
                    void *acc, int accDouble, int threads)
{
   int u, D = m->vecSize;
   if (threads <= 1) {
      Ctx c;
      ctx_init(&c, m, opt, acc, accDouble);
      for (u = 0; u < b->numUtt; u++) {
         int64_t f0 = b->frameOff[u];
         int T = (int)(b->frameOff[u + 1] - f0), Q = b->labOff[u + 1] - b->labOff[u];
         fb_utt(&c, b->feat + (size_t)f0 * D, T, (b->labAlign ? b->labAlign : b->lab) + b->labOff[u],
                b->lab + b->labOff[u], Q, &res[u],
                beams && beams->qLo ? beams->qLo + f0 : NULL, beams && beams->qHi ? beams->qHi + f0 : NULL,
                beams && beams->sq ? beams->sq + f0 : NULL, beams && beams->eq ? beams->eq + f0 : NULL);
      }
      if (accDouble) ((double *)acc)[c.L.totalPr] += c.totalPr;
      else ((float *)acc)[c.L.totalPr] = (float)((double)((float *)acc)[c.L.totalPr] + c.totalPr);
      ctx_free(&c);
      return HFB_OK;
   }
   if (!accDouble) return HFB_EINVAL;
   {
      hfb_acc_layout L;
      int nt = threads > 256 ? 256 : threads, k;
      Work w;
      pthread_t *th = (pthread_t *)calloc(nt, sizeof(pthread_t));
      WorkArg *wa = (WorkArg *)calloc(nt, sizeof(WorkArg));
      layout_of(m, &L);
      w.m = m; w.opt = opt; w.b = b; w.res = res; w.beams = beams; w.next = 0; w.count = L.count;
      pthread_mutex_init(&w.mu, NULL);
      for (k = 0; k < nt; k++) { wa[k].w = &w; wa[k].acc = NULL; pthread_create(&th[k], NULL, worker, &wa[k]); }
      for (k = 0; k < nt; k++) {                 /* merge in thread order */
         int64_t i;
         pthread_join(th[k], NULL);
         if (!wa[k].acc) continue;
         for (i = 0; i < L.count; i++) ((double *)acc)[i] += wa[k].acc[i];
         free(wa[k].acc);
      }
      pthread_mutex_destroy(&w.mu);
      free(th); free(wa);
      (void)u; (void)D;
      return HFB_OK;
   }
}

int hfbo_utt_occupancy(const hfb_model *m, const hfb_options *opt,
                       const float *feat, int32_t T, const int32_t *lab, int32_t Q,
                       float *occ, hfb_utt_result *res)
{
   Ctx c;
   hfb_acc_layout L;
   double *acc;
   int q, P = 0;
   layout_of(m, &L);
   acc = (double *)calloc(L.count, sizeof(double));
   ctx_init(&c, m, opt, acc, 1);
   for (q = 0; q < Q; q++) P += m->hmmNumStates[lab[q]] - 2;
   memset(occ, 0, sizeof(float) * (size_t)T * P);
   c.occOut = occ;
   fb_utt(&c, feat, T, lab, NULL, Q, res, NULL, NULL, NULL, NULL);
   ctx_free(&c);
   free(acc);
   return HFB_OK;
}
