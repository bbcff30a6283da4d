/* hfbgpu.h -- C ABI of libhfbgpu: the B200-native Baum-Welch E-step behind HERest.
 *
 * This is the drop-in boundary.  HTK has no plugin/FFI layer; the
 * seam this library replaces is the three-call API of HTKLib/HFB.h:
 *
 *   InitialiseForBack(fbInfo, heap, hset, uset, pruneInit, pruneInc, pruneLim, minFrwdP)
 *                                             HTKLib/HFB.h:117-119  ->  hfbgpu_create()
 *   FBFile(fbInfo, utt, datafn)               HTKLib/HFB.h:143      ->  hfbgpu_accumulate()
 *   side effects into TrAcc/WtAcc/MuAcc/VaAcc HTKLib/HTrain.h:211-232,
 *   hmm->hook (numEgs, HFB.c:1768-1772), utt->pr, totalT/totalPr
 *   (HTKTools/HERest.c:779-780)                                     ->  hfbgpu_get_accs()
 *
 * Plain C structs, plain pointers and sizes; the caller owns every host buffer,
 * the library owns device memory.  No call ever exits the process: errors come
 * back as HFB_E* codes carrying the reference's own HError numbers where one
 * exists.  There is NO CPU fallback: without a CUDA device every compute entry
 * point returns HFB_ENODEVICE.
 *
 * All model arrays are 0-based.  State indices inside an HMM follow HTK: state 1
 * is the non-emitting entry, state N the non-emitting exit, 2..N-1 emit.
 */
#ifndef HFBGPU_H_
#define HFBGPU_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HFBGPU_ABI_VERSION 3

/* log-arithmetic constants, HTKLib/HMath.h:42-45 and HTKLib/HModel.h:52-53 */
#define HFB_LZERO   (-1.0E10)
#define HFB_LSMALL  (-0.5E10)
#define HFB_MINEARG (-708.3)
#define HFB_MINLARG 2.45E-308
#define HFB_MINMIX  1.0E-5
#define HFB_LMINMIX (-11.5129254649702)
#define HFB_NOPRUNE 1.0E20          /* HTKLib/HFB.h:31 */

/* Limits of one utterance (the per-frame beams are int16 like the reference's `short *qHi, *qLo`,
 * HFB.h:88-89).  A batch holding a longer utterance is rejected as a whole with HFB_EUNSUPPORTED
 * BEFORE anything is accumulated; the caller hands such an utterance to the reference's FBFile. */
#define HFB_MAX_FRAMES 32767
#define HFB_MAX_LABELS 32766

/* hfb_options.flags.  ALIGN_COMP_LEVEL: two-model re-estimation with the configuration variable HFB: ALIGNCOMPLEVEL = T
 * (HFB.c:71, :231, :1521-1530) -- the component posteriors comp_prob[m] / norm are those of the ALIGNMENT set's state
 * (same number of components required, HError 999) instead of the update set's. */
enum { HFB_OPT_ALIGN_COMP_LEVEL = 1 };

/* update flags, HTKLib/HTrain.h:44 (UPMEANS|UPVARS|UPTRANS|UPMIXES) */
enum { HFB_UPMEANS = 1, HFB_UPVARS = 2, HFB_UPTRANS = 4, HFB_UPMIXES = 8 };

/* return codes; positive values in the 73xx range are the reference's HError numbers */
enum {
   HFB_OK          = 0,
   HFB_EINVAL      = -1,   /* malformed argument */
   HFB_ENODEVICE   = -2,   /* no CUDA device / extension missing: never falls back to CPU */
   HFB_ECUDA       = -3,   /* a CUDA call failed, see hfbgpu_last_error() */
   HFB_ENOMEM      = -4,
   HFB_EUNSUPPORTED= -5,   /* model uses a feature outside the path (S>1, full covariance...) */
   HFB_ETEE        = 7332, /* CreateInsts: tee models first/last/successive, HFB.c:557-565 */
   HFB_EALPHAPRUNE = 7390, /* StepAlpha: alpha prune failed, HFB.c:706,718 */
   HFB_EBETAPRUNE  = 7323  /* SetBeta: beta prune failed, HFB.c:1257 */
};

/* per-utterance status */
enum {
   HFB_UTT_OK      = 0,
   HFB_UTT_SKIPPED = 7324, /* warning -7324 + FBFile returns FALSE, HFB.c:1342,1354 */
   HFB_UTT_EALPHA  = 7390,
   HFB_UTT_EBETA   = 7323,
   HFB_UTT_ETEE    = 7332
};

/* ---- flat, read-only model: what HERest holds after ConvDiagC + ConvLogWt --------
 * (HTKTools/HERest.c:592-594,643-645; pointer map in SURVEY.md 8b).               */
typedef struct hfb_model {
   int32_t vecSize;              /* D = hset->vecSize (one stream)                   */

   int32_t numGauss;             /* G distinct MixPDFs                               */
   const float   *mean;          /* [G][D]  mp->mean                                 */
   const float   *ivar;          /* [G][D]  mp->cov.var as INVERSE variances         */
   const float   *gConst;        /* [G]     mp->gConst AS STORED (never recomputed)  */
   const int32_t *meanId;        /* [G] -> mean accumulator (MuAcc) id, ~u sharing   */
   const int32_t *varId;         /* [G] -> variance accumulator (VaAcc) id, ~v       */
   int32_t numMeanAcc;
   int32_t numVarAcc;

   int32_t numStates;            /* J distinct StreamElems (tied states)             */
   const int32_t *stateMixOff;   /* [J+1] offsets into mixGauss/mixLogWt             */
   const int32_t *mixGauss;      /* [sum M] Gaussian index of component m            */
   const float   *mixLogWt;      /* [sum M] log weight; <= HFB_LMINMIX means skipped */

   int32_t numHmm;               /* P physical HMMs, in HMMScan (= dump) order       */
   const int32_t *hmmNumStates;  /* [P] N_p including entry and exit                 */
   const int32_t *hmmStateOff;   /* [P+1] offsets into hmmState                      */
   const int32_t *hmmState;      /* [sum (N_p-2)] tied-state index of states 2..N-1  */
   const int32_t *hmmTrans;      /* [P] transition-matrix id (~t sharing)            */

   int32_t numTrans;
   const int32_t *transN;        /* [numTrans] N of each matrix                      */
   const int32_t *transOff;      /* [numTrans+1] offsets into transLogA (N*N each)   */
   const float   *transLogA;     /* row-major [N][N] float logs, HFB_LZERO if absent */
} hfb_model;

typedef struct hfb_options {
   double  pruneInit;            /* HFB_NOPRUNE = pruning off (HFB.c:83)             */
   double  pruneInc;
   double  pruneLim;
   float   minFrwdP;             /* default 10.0 (HFB.c:83)                          */
   int32_t uFlags;               /* HFB_UP* mask                                     */
   int32_t device;               /* CUDA device ordinal                              */
   int32_t gmmKernel;            /* 0 = auto, 1 = FP32 CUDA-core, 2 = tcgen05 (3xFP16) */
   int32_t flags;                /* HFB_OPT_* bits                                   */
   size_t  workspaceBytes;       /* 0 = default; cap for per-wave beta/outprob pool  */
   /* Two-model re-estimation (UseAlignHMMSet, HFB.c:296-333; HERest ALIGNMODELMMF / ALIGNHMMLIST,
    * HERest.c:163-181, :647-684).  NULL = one set does both jobs.  Otherwise this set ALIGNS -- output
    * probabilities, beams, alpha, beta, utt->pr come from it -- and the model handed to hfbgpu_create is the UPDATE
    * set: component posteriors (HFB.c:1518-1547, :1577-1578), centred sums, numEgs and the accumulator layout are
    * its own.  Both sets must have the same vecSize, and the two HMMs a label resolves to the same number of states
    * (HFB.c:549-551).  HFB_UPTRANS is dropped from uFlags as the reference does (HRError 7392, HFB.c:313-316): the
    * transition accumulators stay zero.  Only read during hfbgpu_create / hfbgpu_create_multi.                     */
   const struct hfb_model *alignModel;
} hfb_options;

/* ---- one batch of loaded utterances ------------------------------------------- */
typedef struct hfb_batch {
   int32_t numUtt;
   const int64_t *frameOff;      /* [numUtt+1] frame offsets into feat               */
   const float   *feat;          /* [totalT][D] row-major, what ReadAsTable yields   */
   const int32_t *labOff;        /* [numUtt+1] offsets into lab                      */
   const int32_t *lab;           /* physical-HMM index per label (utt->tr resolved)  */
   const int32_t *labAlign;      /* two-model re-estimation only (hfb_options.alignModel): the same labels resolved
                                    in the ALIGNMENT set (al_qList, HFB.c:540-542); `lab` is then up_qList (:545-547).
                                    NULL otherwise                                     */
} hfb_batch;

typedef struct hfb_utt_result {
   int32_t status;               /* HFB_UTT_*                                        */
   int32_t retries;              /* beta passes repeated (HFB.c:1349-1361)           */
   double  pr;                   /* utt->pr, total log likelihood                    */
   double  pruneThresh;          /* threshold finally used                           */
} hfb_utt_result;

/* optional per-frame beams of the last batch (debug / parity): 1-based model numbers
 * like the reference's "Beta Beam lo->hi" / "Alpha Beam sq->eq" trace lines.       */
typedef struct hfb_beams {
   int16_t *qLo, *qHi;           /* [totalT] beta beam                               */
   int16_t *sq,  *eq;            /* [totalT] alpha beam                              */
} hfb_beams;

/* ---- layout of the flat FP64 accumulator buffer ------------------------------- */
typedef struct hfb_acc_layout {
   int64_t tran;      /* [sum N*N]  TrAcc.tran, row-major per matrix               */
   int64_t tranOcc;   /* [sum N]    TrAcc.occ                                      */
   int64_t wtC;       /* [sum M]    WtAcc.c                                        */
   int64_t wtOcc;     /* [J]        WtAcc.occ                                      */
   int64_t muSum;     /* [numMeanAcc][D]  MuAcc.mu, CENTRED on the current mean    */
   int64_t muOcc;     /* [numMeanAcc]                                              */
   int64_t vaSum;     /* [numVarAcc][D]   VaAcc.cov.var, centred                   */
   int64_t vaOcc;     /* [numVarAcc]                                               */
   int64_t numEgs;    /* [P]        hmm->hook counters (as doubles, exact < 2^53)  */
   int64_t totalT;    /* [1]        frames of successful utterances                */
   int64_t totalPr;   /* [1]        sum of utt->pr                                 */
   int64_t numOk;     /* [1]        utterances accumulated                         */
   int64_t numSkipped;/* [1]        utterances that returned FALSE                 */
   int64_t count;     /* total doubles                                             */
   int64_t tranOccStride; /* reserved                                              */
} hfb_acc_layout;

typedef struct hfbgpu_ctx hfbgpu_ctx;

/* Host-only helpers (no device needed). */
int  hfbgpu_abi_version(void);
int  hfbgpu_acc_layout(const hfb_model *m, hfb_acc_layout *out);
void hfbgpu_default_options(hfb_options *opt);
const char *hfbgpu_strerror(int code);
const char *hfbgpu_last_error(void);
int  hfbgpu_device_count(void);

/* Replaces InitialiseForBack (HFB.h:117): uploads the model, computes minimum
 * durations (SetMinDurs, HFB.c:106-155), allocates zeroed accumulators.          */
int hfbgpu_create(hfbgpu_ctx **ctx, const hfb_model *m, const hfb_options *opt);
int hfbgpu_destroy(hfbgpu_ctx *ctx);

/* The same on a LIST of devices (SURVEY.md 8b "device list", 8e): one context, one host thread, N GPUs.  Replaces the
 * reference's only parallelism -- N `HERest -p k` processes + the `-p 0` merge of their dumps (HERest.c:514-521,
 * HTrain.c:1626-1687).  Every batch passed to hfbgpu_accumulate / hfbgpu_submit / hfbgpu_accumulate_retrain (HOST
 * features) is cut into numDevices contiguous utterance ranges of equal sum(T * Q), one per GPU; hfbgpu_get_accs,
 * hfbgpu_mstep and hfbgpu_reduce_accs first sum the per-GPU FP64 accumulators into the first device's buffer with a
 * kernel that reads the peers' buffers over NVLink (peer access), so the caller sees ONE set of accumulators.
 * opt->device is ignored; hfbgpu_set_stream and device-resident features are single-device only.                    */
int hfbgpu_create_multi(hfbgpu_ctx **ctx, const hfb_model *m, const hfb_options *opt,
                        const int32_t *devices, int32_t numDevices);
int hfbgpu_num_devices(hfbgpu_ctx *ctx);
int hfbgpu_reduce_accs(hfbgpu_ctx *ctx);

/* Run on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) instead
 * of the library's own, so the caller's events bracket the kernels.  NULL restores it. */
int hfbgpu_set_stream(hfbgpu_ctx *ctx, void *cudaStream);

/* ZeroAccs (HTrain.c:1072). */
int hfbgpu_zero_accs(hfbgpu_ctx *ctx);

/* Replaces the FBFile loop (HERest.c:502-534 / HFB.c:1923): forward-backward and
 * accumulation for every utterance of the batch.  res[numUtt]; beams may be NULL.
 * Host buffers in, copies are part of the call.                                   */
int hfbgpu_accumulate(hfbgpu_ctx *ctx, const hfb_batch *batch,
                      hfb_utt_result *res, const hfb_beams *beams);

/* Same, but batch->feat is a DEVICE pointer already resident in HBM (the other
 * arrays stay on the host).  The caller must have finished writing the features
 * (synchronise its own stream) before the call; the call returns after the results
 * have been copied back.                                                            */
int hfbgpu_accumulate_device(hfbgpu_ctx *ctx, const hfb_batch *batch,
                             hfb_utt_result *res, const hfb_beams *beams);

/* Asynchronous form: hfbgpu_submit() enqueues the batch on one of the library's streams (four
 * slots with private workspaces) and returns; consecutive submits run on different streams, so
 * the host->device copy, table building and kernel tails of one batch overlap the kernels of
 * the others.  The caller keeps `batch` arrays, `res` and `beams` alive until hfbgpu_wait()
 * has returned; hfbgpu_get_accs / _zero_accs wait implicitly.
 * This is what the HERest bridge uses to load the next batch from disk while the GPU works. */
int hfbgpu_submit(hfbgpu_ctx *ctx, const hfb_batch *batch, hfb_utt_result *res,
                  const hfb_beams *beams, int featOnDevice);
int hfbgpu_wait(hfbgpu_ctx *ctx);
/* Per-batch completion: every hfbgpu_submit gets a ticket (1, 2, ...; hfbgpu_last_ticket returns the one of the most
 * recent call).  hfbgpu_wait_ticket completes the batches submitted up to and including `ticket` -- their res[] are
 * valid afterwards -- and leaves younger batches running, so a caller with two buffers can refill the older one while
 * the GPU works on the newer (what the HERest bridge does).                                                      */
int64_t hfbgpu_last_ticket(hfbgpu_ctx *ctx);
int hfbgpu_wait_ticket(hfbgpu_ctx *ctx, int64_t ticket);

/* Single-pass retraining, HERest -r (HERest.c:369, :505-511; HFB.c:1603-1611, :1731): every utterance comes as
 * two parameterisations of the same frames.  Alignment (output probabilities, alpha / beta, beams, state and
 * component occupancies, transition and weight counts) uses batch->feat exactly as hfbgpu_accumulate does; the mean
 * and variance sums take their observations from feat2 = [totalT][vecSize], same frame offsets, centred on the
 * current means as the reference does.  feat2 lives where feat lives (featOnDevice); with qualifiers set, feat is
 * static-only and feat2 is still full width.                                                                      */
int hfbgpu_accumulate_retrain(hfbgpu_ctx *ctx, const hfb_batch *batch, const float *feat2, hfb_utt_result *res,
                              const hfb_beams *beams, int featOnDevice);

/* ---- parameter-kind qualifiers on the device ---------------------------
 * Replaces, for whole utterances, HParm.c:1618-1722 AddQualifiers -> :1552-1599 AddDiffs ->
 * HSigP.c:827-857 Regress and HSigP.c:803-823 FZeroMean: what HERest's loader does when the files
 * hold static coefficients (say MFCC_0) and the configuration asks for TARGETKIND = MFCC_0_D_A[_Z].
 * After hfbgpu_set_qualifiers(ctx, &q) every feature matrix passed to hfbgpu_accumulate /
 * _device / hfbgpu_submit has numStatic columns; the library forms the differentials (bit-identical
 * to the reference: same FP32 operations in the same order) and the cepstral mean normalisation
 * on the device before the output probabilities.  numStatic * (1 + orders) must equal vecSize.
 * hfbgpu_set_qualifiers(ctx, NULL) switches back to full-width matrices.
 * Not expressible (use HParm): _N, _V, global mean / variance files, input transforms, V1COMPAT. */
typedef struct hfb_qualifiers {
   int32_t numStatic;      /* static columns: cepstra, then c0 (_0) and / or energy (_E)  (FindSpans, HParm.c:1430) */
   int32_t delWin;         /* DELTAWINDOW (HParm.c:838), 0 = no _D                              */
   int32_t accWin;         /* ACCWINDOW   (:839),        0 = no _A                              */
   int32_t thirdWin;       /* THIRDWINDOW (:871),        0 = no _T                              */
   int32_t simpleDiffs;    /* SIMPLEDIFFS (:840)                                                */
   int32_t zeroMeanCols;   /* _Z: leading columns to zero-mean per utterance = cepstra (+1 with _0), HParm.c:1709-1712; 0 = no _Z */
   int32_t suppressEnergy; /* _N: the absolute energy / c0 -- the LAST static column -- enters the differentials but is left
                              out of the observation (HParm.c:2882 skipE, :4655-4656): vecSize = numStatic * orders - 1 */
} hfb_qualifiers;
int hfbgpu_set_qualifiers(hfbgpu_ctx *ctx, const hfb_qualifiers *q);
/* The expansion alone (what HCopy with that TARGETKIND writes): src = [frames][numStatic] host
 * floats, frameOff[numUtt + 1] as in hfb_batch, dst = [frames][vecSize] host floats.          */
int hfbgpu_expand_features(hfbgpu_ctx *ctx, const float *src, const int64_t *frameOff, int32_t numUtt, float *dst);

/* ---- HTK compressed parameter files (`_C`, HASCOMPX) on the device ---------------------------------------
 * A compressed file holds, after the 12-byte header, two float vectors A and B (one value per column; they count as
 * 4 of the header's nSamples) and then nSamples - 4 rows of 16-bit integers (+ a 16-bit check sum with `_K`).  The
 * reference's loader turns every integer back into a float as  v[j] = ((float)s[j] + B[j]) / A[j]
 * (HTKLib/HParm.c:3489-3494; A and B are read at :3683-3694; CalcCompress, :4892-4960, wrote them).  These calls take
 * the integers as they are in the file (host byte order) -- half the bytes of the float table ReadAsTable yields cross
 * the PCIe link -- and perform exactly those two FP32 operations on the device: the observations are bit-identical
 * to the reference's.  cols = numStatic with qualifiers set (the files hold static coefficients), else vecSize.
 * batch->feat is ignored (may be NULL); everything else is as for hfbgpu_submit / hfbgpu_accumulate (host buffers,
 * tickets, device groups).  Not combined with device-resident features or -r (two data files).                  */
typedef struct hfb_compressed {
   const int16_t *feat;    /* [totalT][cols] row-major, indexed by batch->frameOff like hfb_batch.feat          */
   const float   *scaleA;  /* [numUtt][cols] vector A of each utterance's file                                   */
   const float   *scaleB;  /* [numUtt][cols] vector B                                                            */
} hfb_compressed;
int hfbgpu_submit_compressed(hfbgpu_ctx *ctx, const hfb_batch *batch, const hfb_compressed *cf,
                             hfb_utt_result *res, const hfb_beams *beams);
int hfbgpu_accumulate_compressed(hfbgpu_ctx *ctx, const hfb_batch *batch, const hfb_compressed *cf,
                                 hfb_utt_result *res, const hfb_beams *beams);
/* The decompression alone (what HList / HCopy print for such a file): dst = [frames][cols] host floats. */
int hfbgpu_decompress_features(hfbgpu_ctx *ctx, const hfb_compressed *cf, const int64_t *frameOff, int32_t numUtt,
                               int32_t cols, float *dst);

/* Pinned (page-locked) host memory for the feature matrix of a batch: uploads from it are
 * asynchronous DMAs that overlap the kernels of the previous batch.  Replaces nothing in the
 * reference (its observations live in HTK's ParmBuf, HParm.c); the bridge copies frames out of
 * ReadAsTable into such a buffer.  Returns NULL on failure.                          */
void *hfbgpu_host_alloc(size_t bytes);
void  hfbgpu_host_free(void *p);

/* Flat FP64 accumulators: device pointer (for an NCCL all-reduce by the caller)
 * and host download.  Layout: hfbgpu_acc_layout().                                */
double *hfbgpu_acc_device_ptr(hfbgpu_ctx *ctx);
int64_t hfbgpu_acc_count(hfbgpu_ctx *ctx);
int hfbgpu_get_accs(hfbgpu_ctx *ctx, double *hostOut);
int hfbgpu_set_accs(hfbgpu_ctx *ctx, const double *hostIn);

/* MOutP/OutP alone (HModel.c:5484, HFB.c:898-988): log b_j(o_t) for `n` tied
 * states over T host frames -> out[T][n].  mixOut (optional) gets the per-mixture
 * log densities laid out [T][sum M of the listed states].                         */
int hfbgpu_state_loglik(hfbgpu_ctx *ctx, const float *feat, int32_t T,
                        const int32_t *states, int32_t n, float *out, float *mixOut);

/* Minimum durations per transition matrix as computed at create (TrAcc.minDur). */
int hfbgpu_get_min_durs(hfbgpu_ctx *ctx, int32_t *out);

/* ---- M-step on the device ----------------------------------------------------
 * Replaces MLUpdateModels (HTKTools/HERest.c:1262-1321: UpdateTrans :795, UpdateWeights :897 with
 * FloorMixes :819, UpdateVars :1045, UpdateMeans :974, FixGConsts HModel.c:5688) for the sets the
 * library accelerates.  Reads the resident accumulators (after the caller's all-reduce) and writes the
 * re-estimated parameters in the flat model's own order; parameters of structures that are not updated
 * (too few examples, zero occupancy, update flag off) are returned unchanged.                       */
typedef struct hfb_mstep_options {
   int32_t minEgs;             /* HERest -m (default 3): models with fewer examples are copied       */
   float   mixWeightFloor;     /* HERest -w f, already multiplied by MINMIX (default 0 = off)          */
   const float *varFloor;      /* [vecSize] variance floor per dimension (HERest -v / ~v varFloor1)    */
} hfb_mstep_options;

typedef struct hfb_mstep_result {
   float *mean;                /* [numGauss][vecSize]                                                  */
   float *var;                 /* [numGauss][vecSize] variances (not inverse)                          */
   float *gConst;              /* [numGauss]  D log 2 pi + sum log var                                 */
   float *mixWeight;           /* [stateMixOff[numStates]] linear weights                              */
   float *transP;              /* all matrices, row-major N*N each, linear probabilities               */
   int32_t nFloorVar, nFloorVarMix;   /* floored variance elements / components (HERest.c:789-790)    */
   int32_t nCopied;            /* physical HMMs with fewer than minEgs examples (warning -2331)        */
   int32_t nNoOcc;             /* structures of updated models without occupancy (warnings -2326/-2330) */
} hfb_mstep_result;

int hfbgpu_mstep(hfbgpu_ctx *ctx, const hfb_mstep_options *opt, hfb_mstep_result *out);

/* Counters for bench.py: kernels launched / device time since the last reset. */
typedef struct hfb_stats {
   int64_t launches;
   int64_t launchesGmm, launchesBeta, launchesAlpha, launchesStats, launchesMisc;
   double  msGmm, msBeta, msAlpha, msStats;   /* CUDA-event time on the library stream */
   int64_t betaCells, alphaCells, gmmPairs;   /* algorithmic units processed             */
   int64_t h2dBytes, d2hBytes;
   int64_t launchesL2R;                       /* of launchesBeta/Alpha: standard-topology kernels */
   double  msExpand;                          /* part of msGmm: feature expansion before the tensor-core kernel */
} hfb_stats;
int hfbgpu_get_stats(hfbgpu_ctx *ctx, hfb_stats *out);
int hfbgpu_reset_stats(hfbgpu_ctx *ctx);
int hfbgpu_set_timing(hfbgpu_ctx *ctx, int enable);

#ifdef __cplusplus
}
#endif
#endif /* HFBGPU_H_ */
