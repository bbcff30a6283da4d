/* hfbgpu_bridge.c -- see hfbgpu_bridge.h.  Compiled against the reference's headers:
 *   gcc -I$REF/HTKLib -I include -c bridge/hfbgpu_bridge.c
 * Pointer map between HTK memory and the flat hfb_model: SURVEY.md 8(b).
 */
#include "HShell.h"
#include "HMem.h"
#include "HMath.h"
#include "HSigP.h"
#include "HAudio.h"
#include "HWave.h"
#include "HVQ.h"
#include "HParm.h"
#include "HLabel.h"
#include "HModel.h"
#include "HTrain.h"
#include "HUtil.h"
#include "HAdapt.h"
#include "HFB.h"

#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <unistd.h>
#include <fcntl.h>
#include <pthread.h>
#include <time.h>

#include "hfbgpu.h"
#include "hfbgpu_bridge.h"

/* ------------------------------------------------------------------ pointer -> index map */
typedef struct { const void **key; int *val; int cap, n; } PMap;

static void pm_init(PMap *m, int cap)
{
   int c = 64;
   while (c < cap * 2) c *= 2;
   m->cap = c; m->n = 0;
   m->key = (const void **)calloc(c, sizeof(void *));
   m->val = (int *)calloc(c, sizeof(int));
}
static int pm_slot(const PMap *m, const void *p)
{
   size_t h = ((size_t)p >> 4) * 2654435761u;
   int i = (int)(h & (size_t)(m->cap - 1));
   while (m->key[i] && m->key[i] != p) i = (i + 1) & (m->cap - 1);
   return i;
}
static int pm_get(const PMap *m, const void *p) { int i = pm_slot(m, p); return m->key[i] ? m->val[i] : -1; }
static int pm_add(PMap *m, const void *p)          /* returns index, new or existing */
{
   int i = pm_slot(m, p);
   if (!m->key[i]) { m->key[i] = p; m->val[i] = m->n++; }
   return m->val[i];
}

/* ------------------------------------------------------------------ state */
/* One HMM set flattened for the library, with the HTK objects behind every flat index */
typedef struct {
   hfb_model m;
   HLink *hmm; StreamElem **ste; MixPDF **mp; SVector *meanV, *varV; SMatrix *trans;
   PMap pmHmm, pmSte, pmMp, pmMean, pmVar, pmTr;
   int nHmm, nSte, nMp, nMean, nVar, nTr;
} FlatSet;

static struct {
   HMMSet *hset;                         /* the set whose accumulators are filled (up_hset) */
   HMMSet *alHset;                       /* 2-model re-estimation: the set that aligns (fbInfo->al_hset), else NULL */
   hfbgpu_ctx *ctx;
   FlatSet u, a;                         /* update set; alignment set (2-model re-estimation only) */
   hfb_acc_layout L;
   int D, batchUtts;
   long batchFrames;
   UPDSet uFlags;
   PMap pmLab; int *labPhys, *labPhysAl; int labCap;   /* label id (LabId) -> physical HMM index in each set, filled on first use */
   /* totals */
   long nOk, nSkipped;
   int twoData;                          /* HERest -r */
   int trace;                            /* HERest's own trace flags (-T): bit 0 = the per-utterance line */
   /* raw feature-file reader (see "fast loader" below) */
   int fastState;                        /* 0 = first file not seen yet, 1 = validated, -1 = off */
   int fastSwap;                         /* payload is byte-swapped on this host */
   int fastComp, fastCrc;                /* the validated files are `_C` compressed (16-bit integers + A / B) / carry a `_K` check sum */
   int fastStatic, cols;                 /* the files hold `cols` static coefficients; the library forms the target kind's
                                            differentials / normalisation on the device (hfbgpu_set_qualifiers) */
   hfb_qualifiers qual; int qualActive;  /* ... with this description; what the context is currently set to */
   short fastKind, fastSize;             /* header fields every fast-loaded file must carry */
   int fastFd; long fastT;               /* file opened by HFBGPU_FastLoad, consumed by HFBGPU_Queue */
   long nFast, nSlow;
   /* host-side profile of the file loop (printed under -T 1): seconds spent in each part of the bridge */
   double tInit, tLoop0, tLast, sQueue, sFastLoad, sReaderWait, sSubmit, sComplete, sOutside;
} B;

static double now_s(void)
{
   struct timespec ts;
   clock_gettime(CLOCK_MONOTONIC, &ts);
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/* Two pending batches: while the library works on one (hfbgpu_submit is asynchronous), HERest's own
   file loop (LoadLabs / LoadData, HERest.c:753-787) fills the other.  Features live in pinned host
   memory (hfbgpu_host_alloc) so that the upload is an asynchronous DMA. */
typedef struct {
   float *feat; long featCap, nFrames;
   float *feat2; long feat2Cap;          /* single-pass retraining (-r): the second parameterisation, HFB.c:445 */
   int64_t *frameOff; int32_t *labOff, *lab, *labAl; int nUtt, labCap, nLab;
   char **names;
   hfb_utt_result *res;
   int comp;                             /* this batch holds the 16-bit integers of `_C` files in `feat` (hfb_compressed) */
   int stat;                             /* this batch's rows hold B.cols static coefficients (device qualifiers) */
   float *scaleA, *scaleB; long abCap;   /* their A / B vectors, [nUtt][D] */
   int inflight;
   int64_t ticket;                       /* of the hfbgpu_submit that took this batch */
   int jobs;                             /* reader jobs not finished yet (guarded by R.mu) */
} Pending;
static Pending P[2];
static int cur = 0;

static void *xrealloc(void *p, size_t n)
{
   void *q = realloc(p, n);
   if (!q) HError(7399, "hfbgpu bridge: out of host memory");
   return q;
}

/* ------------------------------------------------------------------ fast loader
   HTK parameter files are a 12-byte header + nSamples x sampSize bytes of big-endian floats (HWave.c:1399-1433,
   HParm.c OpenParmFile).  When the file kind already IS the target kind -- the first utterance is loaded by the
   reference's own LoadData / ReadAsTable (HFB.c:1837-1879, HParm.c:4616) and the raw payload must reproduce those
   observations bit for bit -- later files skip HParm altogether: the main thread reads the header, reserves the rows in
   the pinned batch buffer and a pool of reader threads preads + byte-swaps the payload straight into it while HERest's
   loop goes on resolving labels.  `_C` compressed files (HASCOMPX; 12-byte header, vectors A and B, rows of 16-bit
   integers, HParm.c:3680-3699) take the same route when the first one decodes -- v = ((float)s + B) / A, HParm.c:3492-3493
   -- to exactly what HParm delivered: the readers then bring the INTEGERS into the pinned buffer, check the `_K` check
   sum like the reference (UpdateCRCC, HParm.c:3357-3380; HError 6350), and the library decodes them on the device
   (hfbgpu_submit_compressed): half the bytes over PCIe.  Files that do not carry the validated header (other kind,
   other width) and everything under -r fall back to LoadData.  HFBGPU_READERS = threads (0 = off). */
typedef struct { int fd; size_t bytes; float *dst; int swap; Pending *owner;
                 int comp, crc, cols; long rows; float *A, *Bv; } Job;
static struct {
   pthread_t *th; int nTh;
   pthread_mutex_t mu; pthread_cond_t work, done;
   Job *q; int cap, head, tail, n; int stop; int failed;
} R;

static void swap_floats(float *p, size_t n)
{
   unsigned int *u = (unsigned int *)p;
   size_t i;
   for (i = 0; i < n; i++) u[i] = __builtin_bswap32(u[i]);
}

static int read_payload(int fd, float *dst, size_t bytes, int swap)
{
   size_t got = 0;
   while (got < bytes) {
      ssize_t r = pread(fd, (char *)dst + got, bytes - got, (off_t)(12 + got));
      if (r <= 0) return -1;
      got += (size_t)r;
   }
   if (swap) swap_floats(dst, bytes / 4);
   return 0;
}

/* Uncompressed file with a `_K` check sum (HCopy's default, SAVEWITHCRC = T): the payload as read_payload delivers it plus
   the reference's check -- the sum runs over the 16-bit words in FILE order, i.e. the most significant half of every
   (host-order) float first (UpdateCRCC, HParm.c:3357-3380); 0 = ok, -1 short read, -2 mismatch (HParm.c:4515) */
static int read_payload_crc(int fd, float *dst, size_t bytes, int swap)
{
   const unsigned int *u = (const unsigned int *)dst;
   unsigned int crc = 0;
   unsigned char tail[2];
   size_t i;
   if (read_payload(fd, dst, bytes, swap) != 0) return -1;
   if (!swap) {                                             /* NATURALREADORDER on a little-endian host: file order = low half first */
      for (i = 0; i < bytes / 4; i++) { crc = (crc * 65536 + (u[i] & 0xffff)) % 36897; crc = (crc * 65536 + (u[i] >> 16)) % 36897; }
   } else
      for (i = 0; i < bytes / 4; i++) { crc = (crc * 65536 + (u[i] >> 16)) % 36897; crc = (crc * 65536 + (u[i] & 0xffff)) % 36897; }
   if (pread(fd, tail, 2, (off_t)(12 + bytes)) != 2) return -1;
   if (crc != (unsigned int)(swap ? ((tail[0] << 8) | tail[1]) : ((tail[1] << 8) | tail[0]))) return -2;
   return 0;
}

/* `_C` file: A, B into the batch's scale rows, the integers into the pinned rows, all in host byte order; returns 0, -1
   (short read) or -2 (check sum of a `_K` file does not match, HParm.c:4515) */
static int read_compressed(const Job *j)
{
   const size_t ab = (size_t)j->cols * sizeof(float), body = (size_t)j->rows * j->cols * sizeof(short);
   short *sp = (short *)j->dst;
   size_t got = 0, i;
   unsigned int crc = 0;
   unsigned char tail[2];
   if (pread(j->fd, j->A, ab, 12) != (ssize_t)ab || pread(j->fd, j->Bv, ab, (off_t)(12 + ab)) != (ssize_t)ab) return -1;
   while (got < body) {
      ssize_t r = pread(j->fd, (char *)sp + got, body - got, (off_t)(12 + 2 * ab + got));
      if (r <= 0) return -1;
      got += (size_t)r;
   }
   if (j->swap) {
      unsigned short *u = (unsigned short *)sp;
      swap_floats(j->A, (size_t)j->cols); swap_floats(j->Bv, (size_t)j->cols);
      for (i = 0; i < (size_t)j->rows * j->cols; i++) u[i] = __builtin_bswap16(u[i]);
   }
   if (j->crc) {
      /* the sum runs over the 16-bit words in FILE order: most significant half of every float first */
      const unsigned int *a = (const unsigned int *)j->A, *b = (const unsigned int *)j->Bv;
      const unsigned short *u = (const unsigned short *)sp;
      const int hs = j->swap ? 16 : 0, ls = j->swap ? 0 : 16;   /* which half of a host-order float comes first in the file */
      for (i = 0; i < (size_t)j->cols; i++) { crc = (crc * 65536 + ((a[i] >> hs) & 0xffff)) % 36897; crc = (crc * 65536 + ((a[i] >> ls) & 0xffff)) % 36897; }
      for (i = 0; i < (size_t)j->cols; i++) { crc = (crc * 65536 + ((b[i] >> hs) & 0xffff)) % 36897; crc = (crc * 65536 + ((b[i] >> ls) & 0xffff)) % 36897; }
      for (i = 0; i < (size_t)j->rows * j->cols; i++) crc = (crc * 65536 + u[i]) % 36897;
      if (pread(j->fd, tail, 2, (off_t)(12 + 2 * ab + body)) != 2) return -1;
      if (crc != (unsigned int)(j->swap ? ((tail[0] << 8) | tail[1]) : ((tail[1] << 8) | tail[0]))) return -2;
   }
   return 0;
}

static void *reader_main(void *arg)
{
   (void)arg;
   for (;;) {
      Job j;
      pthread_mutex_lock(&R.mu);
      while (R.n == 0 && !R.stop) pthread_cond_wait(&R.work, &R.mu);
      if (R.n == 0 && R.stop) { pthread_mutex_unlock(&R.mu); return NULL; }
      j = R.q[R.head]; R.head = (R.head + 1) % R.cap; R.n--;
      pthread_mutex_unlock(&R.mu);
      {
         int bad = j.comp ? read_compressed(&j) : j.crc ? read_payload_crc(j.fd, j.dst, j.bytes, j.swap) : read_payload(j.fd, j.dst, j.bytes, j.swap);
         close(j.fd);
         pthread_mutex_lock(&R.mu);
         if (bad == -2) R.failed = 2; else if (bad && !R.failed) R.failed = 1;
         j.owner->jobs--;
         pthread_cond_broadcast(&R.done);
         pthread_mutex_unlock(&R.mu);
      }
   }
}

static void reader_start(int n)
{
   int i;
   memset(&R, 0, sizeof(R));
   if (n <= 0) return;
   pthread_mutex_init(&R.mu, NULL); pthread_cond_init(&R.work, NULL); pthread_cond_init(&R.done, NULL);
   R.cap = 8192; R.q = (Job *)calloc(R.cap, sizeof(Job));
   R.th = (pthread_t *)calloc(n, sizeof(pthread_t));
   for (i = 0; i < n; i++) if (pthread_create(&R.th[R.nTh], NULL, reader_main, NULL) == 0) R.nTh++;
}

static void reader_push(Job j)
{
   pthread_mutex_lock(&R.mu);
   while (R.n == R.cap) pthread_cond_wait(&R.done, &R.mu);
   R.q[R.tail] = j; R.tail = (R.tail + 1) % R.cap; R.n++;
   j.owner->jobs++;
   pthread_cond_signal(&R.work);
   pthread_mutex_unlock(&R.mu);
}

static void reader_wait(Pending *p)                      /* every payload of this batch is in the pinned buffer */
{
   if (R.nTh == 0) return;
   pthread_mutex_lock(&R.mu);
   while (p->jobs > 0) pthread_cond_wait(&R.done, &R.mu);
   pthread_mutex_unlock(&R.mu);
   if (R.failed == 2) HError(6350, "CloseBuffer: Crc error (hfbgpu bridge, fast loader)");      /* HParm.c:4515 */
   if (R.failed) HError(7350, "hfbgpu bridge: short read in a parameter file (fast loader)");
}

static void reader_stop(void)
{
   int i;
   if (R.nTh == 0) return;
   pthread_mutex_lock(&R.mu); R.stop = 1; pthread_cond_broadcast(&R.work); pthread_mutex_unlock(&R.mu);
   for (i = 0; i < R.nTh; i++) pthread_join(R.th[i], NULL);
   free(R.th); free(R.q);
   memset(&R, 0, sizeof(R));
}

/* ------------------------------------------------------------------ flatten the HMMSet */
static void Flatten(HMMSet *hset, FlatSet *F)
{
   HMMScanState hss;
   int p, j, m, k, D = hset->vecSize, sumM = 0, sumE = 0, sumNN = 0;
   int32_t *stateMixOff, *mixGauss, *hmmN, *hmmStateOff, *hmmState, *hmmTrans, *transN, *transOff, *meanId, *varId;
   float *mixLogWt, *mean, *ivar, *gConst, *transLogA;

   if (hset->swidth[0] != 1) HError(7399, "hfbgpu bridge: only single-stream sets are accelerated");
   if (hset->hsKind != PLAINHS && hset->hsKind != SHAREDHS)
      HError(7399, "hfbgpu bridge: only PLAINHS/SHAREDHS sets are accelerated");
   pm_init(&F->pmHmm, hset->numPhyHMM); pm_init(&F->pmSte, hset->numStates + 16);
   pm_init(&F->pmMp, hset->numMix + 16); pm_init(&F->pmMean, hset->numMix + 16);
   pm_init(&F->pmVar, hset->numMix + 16); pm_init(&F->pmTr, hset->numPhyHMM);
   F->hmm = (HLink *)calloc(hset->numPhyHMM + 1, sizeof(HLink));

   /* pass 1: number physical HMMs (HMMScan order = dump order), states, pdfs, vectors, matrices */
   NewHMMScan(hset, &hss);
   do {
      HLink hmm = hss.hmm;
      p = pm_add(&F->pmHmm, hmm); F->hmm[p] = hmm;
      pm_add(&F->pmTr, hmm->transP);
      for (j = 2; j < hmm->numStates; j++) {
         StreamElem *ste = hmm->svec[j].info->pdf + 1;
         int M = ste->nMix < 0 ? -ste->nMix : ste->nMix;
         if (pm_get(&F->pmSte, ste) < 0) {
            pm_add(&F->pmSte, ste);
            for (m = 1; m <= M; m++) {
               MixPDF *mp = ste->spdf.cpdf[m].mpdf;
               if (mp->ckind != INVDIAGC && mp->ckind != DIAGC)
                  HError(7399, "hfbgpu bridge: only diagonal covariances are accelerated");
               pm_add(&F->pmMp, mp); pm_add(&F->pmMean, mp->mean); pm_add(&F->pmVar, mp->cov.var);
            }
            sumM += M;
         }
         sumE++;
      }
   } while (GoNextHMM(&hss));
   EndHMMScan(&hss);
   F->nHmm = F->pmHmm.n; F->nSte = F->pmSte.n; F->nMp = F->pmMp.n; F->nMean = F->pmMean.n; F->nVar = F->pmVar.n; F->nTr = F->pmTr.n;

   F->ste = (StreamElem **)calloc(F->nSte + 1, sizeof(void *)); F->mp = (MixPDF **)calloc(F->nMp + 1, sizeof(void *));
   F->meanV = (SVector *)calloc(F->nMean + 1, sizeof(SVector)); F->varV = (SVector *)calloc(F->nVar + 1, sizeof(SVector));
   F->trans = (SMatrix *)calloc(F->nTr + 1, sizeof(SMatrix));
   stateMixOff = (int32_t *)calloc(F->nSte + 1, sizeof(int32_t));
   mixGauss = (int32_t *)calloc(sumM + 1, sizeof(int32_t)); mixLogWt = (float *)calloc(sumM + 1, sizeof(float));
   mean = (float *)calloc((size_t)F->nMp * D + 1, sizeof(float)); ivar = (float *)calloc((size_t)F->nMp * D + 1, sizeof(float));
   gConst = (float *)calloc(F->nMp + 1, sizeof(float));
   meanId = (int32_t *)calloc(F->nMp + 1, sizeof(int32_t)); varId = (int32_t *)calloc(F->nMp + 1, sizeof(int32_t));
   hmmN = (int32_t *)calloc(F->nHmm + 1, sizeof(int32_t)); hmmStateOff = (int32_t *)calloc(F->nHmm + 2, sizeof(int32_t));
   hmmState = (int32_t *)calloc(sumE + 1, sizeof(int32_t)); hmmTrans = (int32_t *)calloc(F->nHmm + 1, sizeof(int32_t));
   transN = (int32_t *)calloc(F->nTr + 1, sizeof(int32_t)); transOff = (int32_t *)calloc(F->nTr + 2, sizeof(int32_t));

   /* pass 2: fill (state / pdf numbering follows first-use order of pass 1, so offsets are
      assigned by walking the states in index order) */
   {
      int *mOfState = (int *)calloc(F->nSte + 1, sizeof(int));
      for (p = 0; p < F->nHmm; p++) {
         HLink hmm = F->hmm[p];
         for (j = 2; j < hmm->numStates; j++) {
            StreamElem *ste = hmm->svec[j].info->pdf + 1;
            int s = pm_get(&F->pmSte, ste);
            F->ste[s] = ste; mOfState[s] = ste->nMix < 0 ? -ste->nMix : ste->nMix;
         }
      }
      for (j = 0; j < F->nSte; j++) stateMixOff[j + 1] = stateMixOff[j] + mOfState[j];
      free(mOfState);
   }
   for (j = 0; j < F->nSte; j++) {
      StreamElem *ste = F->ste[j];
      int M = stateMixOff[j + 1] - stateMixOff[j];
      for (m = 1; m <= M; m++) {
         MixPDF *mp = ste->spdf.cpdf[m].mpdf;
         int g = pm_get(&F->pmMp, mp), o = stateMixOff[j] + m - 1;
         mixGauss[o] = g;
         mixLogWt[o] = MixLogWeight(hset, ste->spdf.cpdf[m].weight);     /* log weight (ConvLogWt done) */
         if (!F->mp[g]) {
            F->mp[g] = mp;
            meanId[g] = pm_get(&F->pmMean, mp->mean); varId[g] = pm_get(&F->pmVar, mp->cov.var);
            F->meanV[meanId[g]] = mp->mean; F->varV[varId[g]] = mp->cov.var;
            gConst[g] = mp->gConst;                                      /* as stored, never recomputed */
            for (k = 1; k <= D; k++) {
               mean[(size_t)g * D + k - 1] = mp->mean[k];
               ivar[(size_t)g * D + k - 1] = (mp->ckind == INVDIAGC) ? mp->cov.var[k] : 1.0f / mp->cov.var[k];
            }
         }
      }
   }
   for (p = 0; p < F->nHmm; p++) {
      HLink hmm = F->hmm[p];
      int t = pm_get(&F->pmTr, hmm->transP);
      hmmN[p] = hmm->numStates; hmmTrans[p] = t;
      hmmStateOff[p + 1] = hmmStateOff[p] + hmm->numStates - 2;
      for (j = 2; j < hmm->numStates; j++) hmmState[hmmStateOff[p] + j - 2] = pm_get(&F->pmSte, hmm->svec[j].info->pdf + 1);
      if (!F->trans[t]) { F->trans[t] = hmm->transP; transN[t] = hmm->numStates; }
   }
   for (j = 0; j < F->nTr; j++) { transOff[j + 1] = transOff[j] + transN[j] * transN[j]; }
   sumNN = transOff[F->nTr];
   transLogA = (float *)calloc(sumNN + 1, sizeof(float));
   for (j = 0; j < F->nTr; j++) {
      int N = transN[j], a, b2;
      for (a = 1; a <= N; a++) for (b2 = 1; b2 <= N; b2++) transLogA[transOff[j] + (a - 1) * N + b2 - 1] = F->trans[j][a][b2];
   }
   memset(&F->m, 0, sizeof(F->m));
   F->m.vecSize = D; F->m.numGauss = F->nMp; F->m.mean = mean; F->m.ivar = ivar; F->m.gConst = gConst;
   F->m.meanId = meanId; F->m.varId = varId; F->m.numMeanAcc = F->nMean; F->m.numVarAcc = F->nVar;
   F->m.numStates = F->nSte; F->m.stateMixOff = stateMixOff; F->m.mixGauss = mixGauss; F->m.mixLogWt = mixLogWt;
   F->m.numHmm = F->nHmm; F->m.hmmNumStates = hmmN; F->m.hmmStateOff = hmmStateOff; F->m.hmmState = hmmState;
   F->m.hmmTrans = hmmTrans; F->m.numTrans = F->nTr; F->m.transN = transN; F->m.transOff = transOff; F->m.transLogA = transLogA;
}

static void FreeFlat(FlatSet *F)
{
   PMap *pm[6]; int k;
   pm[0] = &F->pmHmm; pm[1] = &F->pmSte; pm[2] = &F->pmMp; pm[3] = &F->pmMean; pm[4] = &F->pmVar; pm[5] = &F->pmTr;
   for (k = 0; k < 6; k++) { free((void *)pm[k]->key); free(pm[k]->val); }
   free(F->hmm); free(F->ste); free(F->mp); free(F->meanV); free(F->varV); free(F->trans);
   free((void *)F->m.mean); free((void *)F->m.ivar); free((void *)F->m.gConst); free((void *)F->m.meanId); free((void *)F->m.varId);
   free((void *)F->m.stateMixOff); free((void *)F->m.mixGauss); free((void *)F->m.mixLogWt); free((void *)F->m.hmmNumStates);
   free((void *)F->m.hmmStateOff); free((void *)F->m.hmmState); free((void *)F->m.hmmTrans); free((void *)F->m.transN);
   free((void *)F->m.transOff); free((void *)F->m.transLogA);
   memset(F, 0, sizeof(*F));
}

/* ------------------------------------------------------------------ public */
void HFBGPU_Init(HMMSet *hset, FBInfo *fbInfo, LogDouble pruneInit, LogDouble pruneInc,
                 LogDouble pruneLim, float minFrwdP, UPDSet uFlags, int herestTrace)
{
   hfb_options opt;
   ConfParam *cParm[MAXGLOBS];
   int nParm, rc;
   double d;
   char *env;

   memset(&B, 0, sizeof(B));
   B.hset = hset; B.uFlags = uFlags; B.trace = herestTrace; B.fastFd = -1;
   B.tInit = now_s();
   B.alHset = fbInfo->twoModels ? fbInfo->al_hset : NULL;
   if (hset->xf != NULL || (uFlags & (UPXFORM | UPSEMIT | UPMAP)))
      HError(7399, "hfbgpu bridge: transforms / MAP updates are not accelerated");
   /* the reference applies these inside Setotprob / UpMixParms (ApplyCompFXForm + Jacobian); the library does not,
      and silently training on untransformed features would be a wrong model */
   if (hset->semiTied != NULL || hset->projSize > 0)
      HError(7399, "hfbgpu bridge: semi-tied / projected (HLDA) sets are not accelerated");
   if (fbInfo->inXForm != NULL || fbInfo->al_inXForm != NULL || fbInfo->paXForm != NULL)
      HError(7399, "hfbgpu bridge: input / parent transforms (-a, -J, -E) are not accelerated");
   if (!hset->logWt) HError(7399, "hfbgpu bridge: expected log weights (ConvLogWt)");
   Flatten(hset, &B.u);
   B.D = B.cols = hset->vecSize;
   pm_init(&B.pmLab, hset->numLogHMM + (B.alHset ? B.alHset->numLogHMM : 0) + 1024);
   hfbgpu_default_options(&opt);
   if (B.alHset != NULL) {
      /* 2-model re-estimation (ALIGNMODELMMF ..., HERest.c:647-684; UseAlignHMMSet, HFB.c:296-333): the library aligns
         with this set and collects the statistics of `hset` */
      Boolean bv;
      if (B.alHset->xf != NULL || B.alHset->semiTied != NULL || B.alHset->projSize > 0)
         HError(7399, "hfbgpu bridge: transforms on the alignment set are not accelerated");
      if (!B.alHset->logWt) HError(7399, "hfbgpu bridge: expected log weights in the alignment set (ConvLogWt)");
      nParm = GetConfig("HFB", TRUE, cParm, MAXGLOBS);
      if (nParm > 0 && GetConfBool(cParm, nParm, "ALIGNCOMPLEVEL", &bv) && bv)
         opt.flags |= HFB_OPT_ALIGN_COMP_LEVEL;          /* HFB.c:231, :1521-1530: posteriors from the alignment set's components */
      Flatten(B.alHset, &B.a);
      opt.alignModel = &B.a.m;
   }
   /* same precedence as InitFB + InitialiseForBack (HFB.c:221-233, :270-276): config file first,
      command line overrides */
   nParm = GetConfig("HFB", TRUE, cParm, MAXGLOBS);
   if (nParm > 0) {
      if (GetConfFlt(cParm, nParm, "PRUNEINIT", &d)) opt.pruneInit = d;
      if (GetConfFlt(cParm, nParm, "PRUNEINC", &d)) opt.pruneInc = d;
      if (GetConfFlt(cParm, nParm, "PRUNELIM", &d)) opt.pruneLim = d;
      if (GetConfFlt(cParm, nParm, "MINFORPROB", &d)) opt.minFrwdP = (float)d;
   }
   if (pruneInit < NOPRUNE) { opt.pruneInit = pruneInit; opt.pruneInc = pruneInc; opt.pruneLim = pruneLim; }
   if (minFrwdP < NOPRUNE) opt.minFrwdP = minFrwdP;
   opt.uFlags = 0;
   if (uFlags & UPMEANS) opt.uFlags |= HFB_UPMEANS;
   if (uFlags & UPVARS) opt.uFlags |= HFB_UPVARS;
   if (uFlags & UPTRANS) opt.uFlags |= HFB_UPTRANS;
   if (uFlags & UPMIXES) opt.uFlags |= HFB_UPMIXES;
   env = getenv("HFBGPU_DEVICE"); opt.device = env ? atoi(env) : 0;
   /* 1184 = 8 utterances per SM: the recursion kernels run one warp per utterance and 8 of them fill an SM's register
      file (bench.py DEFAULT_UTTS has the sweep); larger batches only enlarge the wave workspace (8.4 MB per
      1000-frame utterance), whose first-use cudaMalloc a short run pays in full */
   env = getenv("HFBGPU_BATCH_UTTS"); B.batchUtts = env ? atoi(env) : 1184;
   /* two batches are in flight at most (P[0], P[1]): two wave slots, not the library's four, so that only two
      workspaces are ever allocated */
   setenv("HFBGPU_STREAMS", "2", 0);
   env = getenv("HFBGPU_BATCH_FRAMES"); B.batchFrames = env ? atol(env) : 4000000L;
   {
      long nc = sysconf(_SC_NPROCESSORS_ONLN);
      env = getenv("HFBGPU_READERS");
      reader_start(env ? atoi(env) : (int)(nc > 8 ? 8 : (nc > 1 ? nc - 1 : 1)));
   }
   /* HFBGPU_DEVICES = "0,1,2,3" or "all": one HERest process drives every listed GPU (hfbgpu_create_multi) -- what
      the reference does with N `-p k` processes and a `-p 0` merge */
   env = getenv("HFBGPU_DEVICES");
   if (env != NULL && *env != '\0') {
      int32_t devs[16];
      int nd = 0;
      if (strcmp(env, "all") == 0) { int n = hfbgpu_device_count(); for (nd = 0; nd < n && nd < 16; nd++) devs[nd] = nd; }
      else { char *e = env; while (*e && nd < 16) { devs[nd++] = (int32_t)strtol(e, &e, 10); while (*e == ',' || *e == ' ') e++; } }
      if (nd < 1) HError(7399, "hfbgpu bridge: HFBGPU_DEVICES lists no device");
      rc = hfbgpu_create_multi(&B.ctx, &B.u.m, &opt, devs, nd);
   } else
      rc = hfbgpu_create(&B.ctx, &B.u.m, &opt);
   if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_create failed: %s (%s)", hfbgpu_strerror(rc), hfbgpu_last_error());
   hfbgpu_acc_layout(&B.u.m, &B.L);
   {  /* a device group cuts every batch into one range per GPU: keep the per-GPU share at the default */
      const int nd = hfbgpu_num_devices(B.ctx);
      if (nd > 1 && getenv("HFBGPU_BATCH_UTTS") == NULL) B.batchUtts *= nd;
      if (nd > 1 && getenv("HFBGPU_BATCH_FRAMES") == NULL) B.batchFrames *= nd;
   }
   printf("hfbgpu: %d physical HMMs, %d tied states, %d Gaussians, %d transition matrices on %d GPU(s)\n",
          B.u.nHmm, B.u.nSte, B.u.nMp, B.u.nTr, hfbgpu_num_devices(B.ctx));
   fflush(stdout);
   {  /* the two pinned batch buffers, once (cudaMallocHost of ~0.6 GB takes a few tenths of a second) */
      int i2;
      const long ncap = B.batchFrames + 40000;
      for (i2 = 0; i2 < 2; i2++) {
         P[i2].feat = (float *)hfbgpu_host_alloc(sizeof(float) * (size_t)ncap * B.D);
         P[i2].featCap = P[i2].feat ? ncap : 0;
      }
   }
   B.tLoop0 = B.tLast = now_s();
}

/* Completes ONE batch (and any older one) and reports its per-utterance outcomes in submission order; a younger batch
   keeps running on the GPU meanwhile (hfbgpu_wait_ticket). */
static void Complete(Pending *p)
{
   int u, rc;
   double t0;
   if (!p->inflight) return;
   t0 = now_s();
   rc = hfbgpu_wait_ticket(B.ctx, p->ticket);
   B.sComplete += now_s() - t0;
   if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_wait_ticket failed: %s (%s)", hfbgpu_strerror(rc), hfbgpu_last_error());
   for (u = 0; u < p->nUtt; u++) {
      const hfb_utt_result *r = &p->res[u];
      if (r->status == HFB_UTT_OK) {
         B.nOk++;
         if (B.trace & 1) {                                            /* HFB.c:1286-1293, under HERest -T 1 */
            printf(" Utterance prob per frame = %e\n", r->pr / (double)(p->frameOff[u + 1] - p->frameOff[u]));
            fflush(stdout);
         }
      } else if (r->status == HFB_UTT_SKIPPED) {                       /* HFB.c:1342, :1354 */
         HError(-7324, "StepBack: File %s - bad data or over pruning\n", p->names[u]);
         B.nSkipped++;
      } else if (r->status == HFB_UTT_ETEE)
         HError(7332, "CreateInsts: Cannot have Tee models at start or end of transcription / successive Tee models (%s)", p->names[u]);
      else
         HError(r->status, "hfbgpu: forward-backward failed for %s (%s)", p->names[u], hfbgpu_strerror(r->status));
      free(p->names[u]);
   }
   free(p->res); p->res = NULL;
   p->inflight = 0; p->nUtt = 0; p->nFrames = 0; p->nLab = 0;
}

static void Drain(void)
{
   Complete(&P[cur]);                                       /* P[cur] was submitted before P[cur ^ 1] */
   Complete(&P[cur ^ 1]);
}

/* Hands the current batch to the library (asynchronously) and switches to the other buffer. */
static void Flush(void)
{
   Pending *p = &P[cur];
   hfb_batch b;
   int rc;
   if (p->nUtt == 0) return;
   { double t0 = now_s(); reader_wait(p); B.sReaderWait += now_s() - t0; }
   p->frameOff[p->nUtt] = p->nFrames; p->labOff[p->nUtt] = p->nLab;
   b.numUtt = p->nUtt; b.frameOff = p->frameOff; b.feat = p->feat; b.labOff = p->labOff; b.lab = p->lab;
   b.labAlign = (B.alHset != NULL) ? p->labAl : NULL;
   p->res = (hfb_utt_result *)calloc(p->nUtt, sizeof(hfb_utt_result));
   if (p->stat != B.qualActive) {                           /* static-coefficient batches and full-width ones never mix */
      rc = hfbgpu_set_qualifiers(B.ctx, p->stat ? &B.qual : NULL);
      if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_set_qualifiers failed: %s (%s)", hfbgpu_strerror(rc), hfbgpu_last_error());
      B.qualActive = p->stat;
   }
   if (B.twoData) {
      /* -r: alignment on the first file of each pair, mean / variance sums from the second (HFB.c:1603-1611);
         the library's entry for it is blocking, so the batch is completed and reported right away */
      rc = hfbgpu_accumulate_retrain(B.ctx, &b, p->feat2, p->res, NULL, 0);
      if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_accumulate_retrain failed: %s (%s)", hfbgpu_strerror(rc), hfbgpu_last_error());
      p->inflight = 1; p->ticket = hfbgpu_last_ticket(B.ctx);
      cur ^= 1;
      Drain();
      return;
   }
   if (p->comp) {
      hfb_compressed cf;
      double t0 = now_s();
      cf.feat = (const int16_t *)p->feat; cf.scaleA = p->scaleA; cf.scaleB = p->scaleB;
      b.feat = NULL;
      rc = hfbgpu_submit_compressed(B.ctx, &b, &cf, p->res, NULL);
      B.sSubmit += now_s() - t0;
   } else { double t0 = now_s(); rc = hfbgpu_submit(B.ctx, &b, p->res, NULL, 0); B.sSubmit += now_s() - t0; }
   if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_submit failed: %s (%s)", hfbgpu_strerror(rc), hfbgpu_last_error());
   p->inflight = 1; p->ticket = hfbgpu_last_ticket(B.ctx);
   cur ^= 1;
   Complete(&P[cur]);                                       /* the buffer about to be refilled must be done; the batch
                                                               just submitted keeps the GPU busy meanwhile */
}

/* The rows of an open parameter file as floats [T][cols], `_C` rows decoded as HParm does (HParm.c:3492-3493); 0 = ok */
static int read_rows(int fd, int sw, unsigned short kd, int T, int cols, float *out)
{
   Job j;
   float *A;
   short *sp;
   int ok, e;
   if (!(kd & HASCOMPX))
      return (kd & HASCRCC) ? read_payload_crc(fd, out, (size_t)T * cols * sizeof(float), sw) : read_payload(fd, out, (size_t)T * cols * sizeof(float), sw);
   A = (float *)malloc(sizeof(float) * 2 * (size_t)cols);
   sp = (short *)malloc(sizeof(short) * (size_t)T * cols + 2);
   memset(&j, 0, sizeof(j));
   j.fd = fd; j.dst = (float *)sp; j.swap = sw; j.comp = 1; j.crc = (kd & HASCRCC) ? 1 : 0; j.cols = cols; j.rows = T; j.A = A; j.Bv = A + cols;
   ok = A && sp && read_compressed(&j) == 0;
   for (e = 0; ok && e < T * cols; e++) out[e] = ((float)sp[e] + A[cols + e % cols]) / A[e % cols];
   free(A); free(sp);
   return ok ? 0 : -1;
}

/* From the files' kind to the set's kind with qualifiers the library can add on the device (HParm.c:1618-1722 AddQualifiers:
   _D _A _T from DELTAWINDOW / ACCWINDOW / THIRDWINDOW / SIMPLEDIFFS, _Z, _N); FALSE = HParm has to do it. */
static Boolean MakeQualifiers(unsigned short fileKind, unsigned short tgtKind, int cols, int D, hfb_qualifiers *q)
{
   const unsigned short added = (unsigned short)(HASDELTA | HASACCS | HASTHIRD | HASZEROM | HASNULLE);
   const unsigned short fk = (unsigned short)(fileKind & ~(HASCOMPX | HASCRCC)), tk = (unsigned short)(tgtKind & ~(HASCOMPX | HASCRCC));
   ConfParam *cParm[MAXGLOBS];
   int nParm, iv, orders;
   Boolean bv;
   if ((fk & BASEMASK) != (tk & BASEMASK) || (fk & added) || (fk & ~tk) || ((tk & ~fk) & ~added) || (tk & HASVQ)) return FALSE;
   if ((tk & HASNULLE) && (!(tk & HASDELTA) || !(tk & (HASENERGY | HASZEROC)) || ((tk & HASENERGY) && (tk & HASZEROC)))) return FALSE;
   memset(q, 0, sizeof(*q));
   q->numStatic = cols;
   q->delWin = q->accWin = q->thirdWin = 2;                 /* HParm.c:838-871 defaults */
   nParm = GetConfig("HPARM", TRUE, cParm, MAXGLOBS);
   if (nParm > 0) {
      if (GetConfInt(cParm, nParm, "DELTAWINDOW", &iv)) q->delWin = iv;
      if (GetConfInt(cParm, nParm, "ACCWINDOW", &iv)) q->accWin = iv;
      if (GetConfInt(cParm, nParm, "THIRDWINDOW", &iv)) q->thirdWin = iv;
      if (GetConfBool(cParm, nParm, "SIMPLEDIFFS", &bv)) q->simpleDiffs = bv ? 1 : 0;
   }
   if (!(tk & HASDELTA)) q->delWin = 0;
   if (!(tk & HASACCS)) q->accWin = 0;
   if (!(tk & HASTHIRD)) q->thirdWin = 0;
   q->suppressEnergy = (tk & HASNULLE) ? 1 : 0;
   if (tk & HASZEROM) {                                      /* cepstra and c0, not the energy (HParm.c:1709-1712) */
      q->zeroMeanCols = cols - ((tk & HASENERGY) ? 1 : 0);
      if (q->suppressEnergy && q->zeroMeanCols > cols - 1) q->zeroMeanCols = cols - 1;
   }
   orders = 1 + (q->delWin > 0) + (q->accWin > 0) + (q->thirdWin > 0);
   return (cols * orders - q->suppressEnergy == D) ? TRUE : FALSE;
}

/* First utterance (loaded by the reference's own LoadData): does the raw payload of the file reproduce, bit for bit,
   the observations HParm delivered -- as it is (files of the target kind, plain or `_C` compressed), or after the
   library's own qualifier expansion (files holding the static coefficients of the target kind)?  Only then may later
   files with the same header skip HParm.  `p`, `want`: the batch and the rows HParm's observations were copied to. */
static void FastValidate(UttInfo *utt, char *datafn, Pending *p, float *want, int T)
{
   unsigned char h[12];
   int fd, sw, D = B.D;
   float *tmp, *raw;
   B.fastState = -1;
   if (R.nTh == 0 || utt->twoDataFiles) return;
   fd = open(datafn, O_RDONLY);
   if (fd < 0) return;
   if (pread(fd, h, 12, 0) != 12) { close(fd); return; }
   tmp = (float *)malloc((size_t)T * D * sizeof(float));
   raw = (float *)malloc((size_t)T * D * sizeof(float));
   for (sw = 1; sw >= 0 && tmp && raw; sw--) {              /* HTK files are big-endian unless NATURALREADORDER */
      unsigned int ns = *(unsigned int *)h; unsigned short ss = *(unsigned short *)(h + 8), kd = *(unsigned short *)(h + 10);
      const int64_t one[2] = {0, T};
      hfb_qualifiers q;
      int comp, esz, cols, useQ = 0, ok;
      if (sw) { ns = __builtin_bswap32(ns); ss = __builtin_bswap16(ss); kd = __builtin_bswap16(kd); }
      if (kd & HASVQ) continue;                              /* VQ files stay with HParm */
      comp = (kd & HASCOMPX) ? 1 : 0;
      esz = comp ? (int)sizeof(short) : (int)sizeof(float);
      if (ss == 0 || ss % esz != 0) continue;
      cols = ss / esz;
      if ((long)ns != (long)T + (comp ? 4 : 0)) continue;    /* compressed: nSamples counts the A / B vectors as 4 rows */
      if (cols > D) continue;
      if (cols < D) {
         if (p->nUtt != 0 || !MakeQualifiers(kd, (unsigned short)B.hset->pkind, cols, D, &q)) continue;
         useQ = 1;
      }
      if (read_rows(fd, sw, kd, T, cols, raw) != 0) continue;
      if (!useQ) ok = memcmp(raw, want, (size_t)T * D * sizeof(float)) == 0;
      else {
         ok = hfbgpu_set_qualifiers(B.ctx, &q) == HFB_OK && hfbgpu_expand_features(B.ctx, raw, one, 1, tmp) == HFB_OK &&
              memcmp(tmp, want, (size_t)T * D * sizeof(float)) == 0;
         if (!ok) hfbgpu_set_qualifiers(B.ctx, NULL);
      }
      if (!ok) continue;
      B.fastState = 1; B.fastSwap = sw; B.fastKind = (short)kd; B.fastSize = (short)ss;
      B.fastComp = comp; B.fastCrc = (kd & HASCRCC) ? 1 : 0;
      if (useQ) {
         /* from here on every fast-loaded batch holds `cols` columns; so does this one: the rows HParm expanded are
            replaced by the file's own static coefficients (this is the first utterance of its batch) */
         B.fastStatic = 1; B.cols = cols; B.qual = q; B.qualActive = 1;
         memcpy(want, raw, (size_t)T * cols * sizeof(float));
         p->stat = 1;
      }
      break;
   }
   free(tmp); free(raw);
   close(fd);
   if (B.trace & 1) {
      printf("hfbgpu: fast loader %s%s\n", B.fastState != 1 ? "off (files need HParm's conversions)" :
             B.fastComp ? "on (compressed files: integers to the device, decoded there; first file verified against HParm)"
                        : "on (payload verified against HParm on the first file)",
             B.fastState != 1 ? "" : B.fastStatic ? "; files hold static coefficients, the target kind's qualifiers are formed on the device"
                                                  : "; file kind = target kind");
      fflush(stdout);
   }
}

/* Called instead of LoadData (HFB.c:1837-1879) once the fast loader is validated: reads the 12-byte header, checks
   it against the validated one and leaves the open file for HFBGPU_Queue.  FALSE = the caller runs LoadData. */
Boolean HFBGPU_FastLoad(UttInfo *utt, char *datafn, char *datafn2)
{
   unsigned char h[12];
   unsigned int ns; unsigned short ss, kd;
   int fd;
   double t0;
   if (B.ctx == NULL || B.fastState != 1 || utt->twoDataFiles || datafn2 != NULL) return FALSE;
   t0 = now_s();
   B.sOutside += t0 - B.tLast;                              /* HERest's own code since the bridge was last left (LoadLabs ...) */
   fd = open(datafn, O_RDONLY);
   if (fd < 0) return FALSE;                                /* let LoadData report it */
   if (pread(fd, h, 12, 0) != 12) { close(fd); return FALSE; }
   ns = *(unsigned int *)h; ss = *(unsigned short *)(h + 8); kd = *(unsigned short *)(h + 10);
   if (B.fastSwap) { ns = __builtin_bswap32(ns); ss = __builtin_bswap16(ss); kd = __builtin_bswap16(kd); }
   if ((short)kd != B.fastKind || (short)ss != B.fastSize || ns == 0 || ns > 100000000u) { close(fd); return FALSE; }
   if (B.fastComp) { if (ns <= 4) { close(fd); return FALSE; } ns -= 4; }   /* A and B count as 4 rows (HParm.c:3643-3644, :3686) */
   utt->T = (int)ns;
   B.fastFd = fd; B.fastT = (long)ns;
   B.tLast = now_s(); B.sFastLoad += B.tLast - t0;
   return TRUE;
}

/* Replaces FBFile (HFB.c:1923): the utterance is only buffered here.  Always returns FALSE so
   that the caller's own totalT/totalPr update (HERest.c:777-786) is skipped; HFBGPU_Finish sets
   the totals from the device accumulators instead. */
Boolean HFBGPU_Queue(FBInfo *fbInfo, UttInfo *utt, char *datafn)
{
   LLink lab;
   Pending *p = &P[cur];
   int q, t, k, T = utt->T, Q = utt->Q, D = B.D;
   double tq0 = now_s();
   const int comp = (B.fastFd >= 0 && B.fastComp) ? 1 : 0, stat = (B.fastFd >= 0 && B.fastStatic) ? 1 : 0;
   const int W = stat ? B.cols : D;                         /* columns of this utterance's rows in the batch buffer */
   B.sOutside += tq0 - B.tLast;
   B.twoData = utt->twoDataFiles ? 1 : 0;
   /* a batch is all floats or all integers, all static coefficients or all full width */
   if (p->nUtt > 0 && (p->comp != comp || p->stat != stat)) { Flush(); p = &P[cur]; }
   p->comp = comp; p->stat = stat;
   if (!p->frameOff) {
      p->frameOff = (int64_t *)xrealloc(NULL, sizeof(int64_t) * (B.batchUtts + 2));
      p->labOff = (int32_t *)xrealloc(NULL, sizeof(int32_t) * (B.batchUtts + 2));
      p->names = (char **)xrealloc(NULL, sizeof(char *) * (B.batchUtts + 2));
   }
   if (p->nFrames + T > p->featCap) {                       /* grow the pinned feature buffer */
      long ncap = (p->nFrames + T) * 2 + 1024;
      if (ncap < B.batchFrames + T + 1024) ncap = B.batchFrames + T + 1024;   /* one allocation per buffer in the normal case */
      reader_wait(p);                                       /* nobody writes into the old buffer any more */
      float *nf = (float *)hfbgpu_host_alloc(sizeof(float) * (size_t)ncap * D);
      if (!nf) HError(7399, "hfbgpu bridge: out of pinned host memory");
      if (p->feat) { memcpy(nf, p->feat, sizeof(float) * (size_t)p->nFrames * D); hfbgpu_host_free(p->feat); }
      p->feat = nf; p->featCap = ncap;
   }
   if (B.twoData && p->nFrames + T > p->feat2Cap) {
      long ncap = (p->nFrames + T) * 2 + 1024;
      float *nf = (float *)hfbgpu_host_alloc(sizeof(float) * (size_t)ncap * D);
      if (!nf) HError(7399, "hfbgpu bridge: out of pinned host memory");
      if (p->feat2) { memcpy(nf, p->feat2, sizeof(float) * (size_t)p->nFrames * D); hfbgpu_host_free(p->feat2); }
      p->feat2 = nf; p->feat2Cap = ncap;
   }
   if (p->nLab + Q > p->labCap) {
      p->labCap = (p->nLab + Q) * 2 + 1024;
      p->lab = (int32_t *)xrealloc(p->lab, sizeof(int32_t) * (size_t)p->labCap);
      if (B.alHset != NULL) p->labAl = (int32_t *)xrealloc(p->labAl, sizeof(int32_t) * (size_t)p->labCap);
   }
   p->frameOff[p->nUtt] = p->nFrames; p->labOff[p->nUtt] = p->nLab;
   /* labels -> physical HMM indices (CreateInsts, HFB.c:538-552; with two sets: al_qList and up_qList) */
   for (lab = utt->tr->head->head->succ, q = 0; lab->succ != NULL; lab = lab->succ, q++) {
      int li = pm_get(&B.pmLab, lab->labid), ph, pa = -1;
      if (li < 0) {                                         /* first time this label is seen: the reference's lookup */
         MLink ml;
         if (B.alHset != NULL) {
            HLink ah;
            ml = FindMacroName(B.alHset, 'l', lab->labid);
            if (ml == NULL) HError(7321, "CreateInsts: Unknown label %s", lab->labid->name);
            ah = (HLink)ml->structure;
            pa = pm_get(&B.a.pmHmm, ah);
            if (pa < 0) HError(7321, "hfbgpu bridge: label %s maps to an unknown physical HMM of the alignment set", lab->labid->name);
            ml = FindMacroName(B.hset, 'l', lab->labid);
            if (ml == NULL) HError(2321, "CreateInsts: Unknown update label %s", lab->labid->name);
            if (ah->numStates != ((HLink)ml->structure)->numStates)
               HError(999, "Num states differ in align and update models (%d %d)", ah->numStates, ((HLink)ml->structure)->numStates);
         } else {
            ml = FindMacroName(B.hset, 'l', lab->labid);
            if (ml == NULL) HError(7321, "CreateInsts: Unknown label %s", lab->labid->name);
         }
         ph = pm_get(&B.u.pmHmm, ml->structure);
         if (ph < 0) HError(7321, "hfbgpu bridge: label %s maps to an unknown physical HMM", lab->labid->name);
         if (B.pmLab.n * 2 + 2 < B.pmLab.cap) {
            li = pm_add(&B.pmLab, lab->labid);
            if (li >= B.labCap) {
               B.labCap = li * 2 + 1024;
               B.labPhys = (int *)xrealloc(B.labPhys, sizeof(int) * (size_t)B.labCap);
               B.labPhysAl = (int *)xrealloc(B.labPhysAl, sizeof(int) * (size_t)B.labCap);
            }
            B.labPhys[li] = ph; B.labPhysAl[li] = pa;
         }
      } else { ph = B.labPhys[li]; pa = B.labPhysAl[li]; }
      p->lab[p->nLab + q] = ph;
      if (B.alHset != NULL) p->labAl[p->nLab + q] = pa;
   }
   if (comp && p->nUtt + 1 > p->abCap) {
      long ncap = (p->nUtt + 1) * 2 + B.batchUtts + 16;
      reader_wait(p);                                       /* readers write into the rows */
      p->scaleA = (float *)xrealloc(p->scaleA, sizeof(float) * (size_t)ncap * D);
      p->scaleB = (float *)xrealloc(p->scaleB, sizeof(float) * (size_t)ncap * D);
      p->abCap = ncap;
   }
   if (B.fastFd >= 0) {
      /* HFBGPU_FastLoad opened the file: a reader thread brings the payload into the pinned rows */
      Job j;
      memset(&j, 0, sizeof(j));
      j.fd = B.fastFd; j.bytes = (size_t)T * W * sizeof(float); j.dst = p->feat + (size_t)p->nFrames * W;
      j.crc = B.fastCrc;
      if (comp) {                                           /* the same pinned buffer, viewed as 16-bit integers */
         j.dst = (float *)((short *)p->feat + (size_t)p->nFrames * W);
         j.comp = 1; j.crc = B.fastCrc; j.cols = W; j.rows = T;
         j.A = p->scaleA + (size_t)p->nUtt * W; j.Bv = p->scaleB + (size_t)p->nUtt * W;
      }
      j.swap = B.fastSwap; j.owner = p;
      B.fastFd = -1;
      reader_push(j);
      B.nFast++;
   } else {
      /* observations exactly as the reference reads them, frame by frame (HFB.c:1009, :1778) */
      for (t = 0; t < T; t++) {
         ReadAsTable(utt->pbuf, t, &utt->ot);
         for (k = 1; k <= D; k++) p->feat[(size_t)(p->nFrames + t) * D + k - 1] = utt->ot.fv[1][k];
         if (B.twoData) {                                   /* HFB.c:445 */
            ReadAsTable(utt->pbuf2, t, &utt->ot2);
            for (k = 1; k <= D; k++) p->feat2[(size_t)(p->nFrames + t) * D + k - 1] = utt->ot2.fv[1][k];
         }
      }
      B.nSlow++;
      if (B.fastState == 0) FastValidate(utt, datafn, p, p->feat + (size_t)p->nFrames * D, T);
   }
   p->names[p->nUtt] = strdup(datafn);
   p->nFrames += T; p->nLab += Q; p->nUtt++;
   B.sQueue += now_s() - tq0;
   if (p->nUtt >= B.batchUtts || p->nFrames >= B.batchFrames) Flush();
   B.tLast = now_s();
   return FALSE;
}

/* Scatter the device accumulators into HTK's (float) accumulators. */
void HFBGPU_Finish(int *totalT, LogDouble *totalPr)
{
   double *acc;
   int p, s, g, t, i, j, k, D = B.D, rc;
   long long o;
   double tf0 = now_s(), tLoop;
   B.sOutside += tf0 - B.tLast;
   Flush();
   Drain();
   tLoop = now_s();
   acc = (double *)calloc((size_t)B.L.count, sizeof(double));
   rc = hfbgpu_get_accs(B.ctx, acc);
   if (rc != HFB_OK) HError(7399, "hfbgpu bridge: hfbgpu_get_accs failed: %s", hfbgpu_strerror(rc));
   for (p = 0; p < B.u.nHmm; p++) {
      long n = (long)B.u.hmm[p]->hook + (long)(acc[B.L.numEgs + p] + 0.5);
      B.u.hmm[p]->hook = (void *)n;                                       /* HFB.c:1768-1772 */
   }
   for (t = 0, o = 0; t < B.u.nTr; t++) {
      TrAcc *ta = (TrAcc *)GetHook(B.u.trans[t]);
      int N = B.u.m.transN[t];
      if (ta != NULL && (B.uFlags & UPTRANS))
         for (i = 1; i <= N; i++)
            for (j = 1; j <= N; j++) ta->tran[i][j] += (float)acc[B.L.tran + o + (i - 1) * N + j - 1];
      o += (long long)N * N;
   }
   for (t = 0, o = 0; t < B.u.nTr; t++) {
      TrAcc *ta = (TrAcc *)GetHook(B.u.trans[t]);
      int N = B.u.m.transN[t];
      if (ta != NULL && (B.uFlags & UPTRANS)) for (i = 1; i <= N; i++) ta->occ[i] += (float)acc[B.L.tranOcc + o + i - 1];
      o += N;
   }
   for (s = 0; s < B.u.nSte; s++) {
      WtAcc *wa = (WtAcc *)B.u.ste[s]->hook;
      int M = B.u.m.stateMixOff[s + 1] - B.u.m.stateMixOff[s];
      if (wa == NULL) continue;
      for (k = 1; k <= M; k++) wa->c[k] += (float)acc[B.L.wtC + B.u.m.stateMixOff[s] + k - 1];
      wa->occ += (float)acc[B.L.wtOcc + s];
   }
   for (g = 0; g < B.u.nMean; g++) {
      MuAcc *ma = (MuAcc *)GetHook(B.u.meanV[g]);
      if (ma == NULL) continue;
      for (k = 1; k <= D; k++) ma->mu[k] += (float)acc[B.L.muSum + (long long)g * D + k - 1];
      ma->occ += (float)acc[B.L.muOcc + g];
   }
   for (g = 0; g < B.u.nVar; g++) {
      VaAcc *va = (VaAcc *)GetHook(B.u.varV[g]);
      if (va == NULL) continue;
      for (k = 1; k <= D; k++) va->cov.var[k] += (float)acc[B.L.vaSum + (long long)g * D + k - 1];
      va->occ += (float)acc[B.L.vaOcc + g];
   }
   *totalT += (int)(acc[B.L.totalT] + 0.5);                             /* HERest.c:779-780 */
   *totalPr += acc[B.L.totalPr];
   free(acc);
   reader_stop();
   if (B.trace & 1) {
      printf("hfbgpu: %ld utterances through the fast loader, %ld through HParm\n", B.nFast, B.nSlow);
      printf("hfbgpu: host profile (s): set-up (flatten, CUDA context, model upload, pinned buffers) %.3f; file loop %.3f = HERest's own code (LoadLabs, LoadData ...) %.3f + bridge queue %.3f + header reads %.3f"
             " + wait for readers %.3f + submit %.3f + wait for GPU %.3f; final flush %.3f; download + scatter %.3f\n",
             B.tLoop0 - B.tInit, tf0 - B.tLoop0, B.sOutside, B.sQueue, B.sFastLoad, B.sReaderWait, B.sSubmit, B.sComplete, tLoop - tf0, now_s() - tLoop);
      fflush(stdout);
   }
   for (i = 0; i < 2; i++) {
      hfbgpu_host_free(P[i].feat); hfbgpu_host_free(P[i].feat2);
      free(P[i].frameOff); free(P[i].labOff); free(P[i].lab); free(P[i].labAl); free(P[i].names); free(P[i].res);
      free(P[i].scaleA); free(P[i].scaleB);
      memset(&P[i], 0, sizeof(P[i]));
   }
   hfbgpu_destroy(B.ctx);
   B.ctx = NULL;
   free((void *)B.pmLab.key); free(B.pmLab.val);
   free(B.labPhys); free(B.labPhysAl);
   FreeFlat(&B.u);
   if (B.alHset != NULL) FreeFlat(&B.a);
   memset(&B, 0, sizeof(B)); B.fastFd = -1;
}
