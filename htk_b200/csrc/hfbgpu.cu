// hfbgpu.cu -- host side of libhfbgpu: the C ABI declared in include/hfbgpu.h.
//
// Replaces the reference's InitialiseForBack / FBFile seam (HTKLib/HFB.h:117-119, :143).
// The batch is cut into waves that fit the device workspace (four wave slots, each with its own
// stream and workspace); per wave the host only sizes things and uploads labels + descriptors in one
// copy, then launches
//   K0 prep (tables, CreateInsts HFB.c:508-574) -> K1 gmm -> K2 beta -> K3 alpha -> K4 stats
// on the slot's stream.  Accumulators stay resident in HBM as one flat FP64 buffer
// (layout: hfbgpu_acc_layout) until hfbgpu_get_accs() or the caller's all-reduce.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <unordered_map>
#include <vector>

#include "hfb_common.h"
#include "hfb_kernels.cuh"
#include "hfb_kernels2.cuh"
#include "hfb_fast.cuh"
#include "hfb_l2r.cuh"
#include "hfb_stats_mma.cuh"
#include "hfb_mstep.cuh"
#include "hfb_feat.cuh"
#include "gmm_tc.cuh"
#include "gmm_tc3.cuh"
#include "gmm_tc4.cuh"
#include "hfb_stats_tc.cuh"

static thread_local std::string g_lastError;

#define CK(call)                                                                                   \
   do {                                                                                            \
      cudaError_t e_ = (call);                                                                     \
      if (e_ != cudaSuccess) {                                                                     \
         char buf_[512];                                                                           \
         snprintf(buf_, sizeof buf_, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,          \
                  cudaGetErrorString(e_));                                                         \
         g_lastError = buf_;                                                                       \
         return HFB_ECUDA;                                                                         \
      }                                                                                            \
   } while (0)

namespace {

template <class T>
struct DevBuf {
   T *p = nullptr;
   size_t cap = 0;
   int reserve(size_t n)
   {
      if (n <= cap) return HFB_OK;
      if (p) cudaFree(p);
      p = nullptr; cap = 0;
      size_t want = n + n / 4 + 16;
      if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) {
         cudaGetLastError();
         want = n;
         if (cudaMalloc(&p, want * sizeof(T)) != cudaSuccess) { cudaGetLastError(); return HFB_ENOMEM; }
      }
      cap = want;
      return HFB_OK;
   }
   void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostModel {
   int D, G, J, P, numTrans, maxM, maxN;
   bool l2r;                         // every HMM has HTK's standard 5-state left-to-right topology (hfb_l2r.cuh)
   std::vector<int> stateMixOff, hmmN, hmmStateOff, hmmState, hmmTrans, transN, transOff, minDur;
   std::vector<int> meanId, varId;   // per Gaussian (M-step)
   int numMeanAcc = 0, numVarAcc = 0;
   std::vector<long long> tranAccOff, tranOccOff;
   std::vector<float> transLogA;
};

}  // namespace

namespace {
struct WaveTables {
   std::vector<UttDesc> utt;
   std::vector<UttOut> out;
   std::vector<int> posPre, tilePre;
   std::vector<int2> tcItems;        // (utterance, first frame) blocks of TC_BM frames for the tcgen05 kernel
   std::vector<int2> tcItems2;       // the same in blocks of 2*TC_BM frames for the CTA-pair kernel
   std::vector<int2> tcItems4;       // ... and of 4*TC_BM frames (two blocks per CTA, 3xFP16 path)
   std::vector<int> uttIndex;        // index in the caller's batch
   long long bFloats = 0, betaDoubles = 0, occDoubles = 0, aentDoubles = 0;
   long long totalQ = 0, totalP = 0, totalSl = 0, tiles = 0;   // totalSl: slots allocated (>= positions; global-slot sets: tied states)
   int maxQ = 0, maxS = 0, maxN = 0, maxT = 0;
   int lab0 = 0;                     // first label of the wave in the caller's label array
   void clear()
   {
      utt.clear(); out.clear(); posPre.clear(); tilePre.clear(); tcItems.clear(); tcItems2.clear(); tcItems4.clear(); uttIndex.clear();
      bFloats = betaDoubles = occDoubles = aentDoubles = 0; totalQ = totalP = totalSl = tiles = 0; maxQ = maxS = maxN = maxT = 0; lab0 = 0;
   }
};

}  // namespace
typedef WaveTables *WaveTablesPtr;

struct hfbgpu_ctx {
   int device = 0;
   hfb_options opt;
   HostModel hm;
   DevModel dm;
   hfb_acc_layout L;
   cudaStream_t stream = nullptr;    // stream of model uploads / accumulator copies (caller's if set)
   cudaStream_t ownStream = nullptr;
   cudaStream_t gmmStream = nullptr;    // high priority: the tensor-core kernels of all waves, back to back
   cudaEvent_t evRef = nullptr;         // HFBGPU_TRACE_KERNELS: time origin of the per-wave timeline on stderr
   bool trace = false;
   int waveUtts = 592;                  // utterances per wave (see launch_wave: recursions co-reside with the next GMM)
   bool timing = false;
   FeatQual qual = {};                  // hfbgpu_set_qualifiers: feat matrices hold static coefficients only
   hfb_stats stats;
   // model on device
   DevBuf<float> dMean, dIvar, dGconst, dMixLogWt, dTransLogA;
   DevBuf<float> dCentre;             // [J][S4_CSTR] centre of each tied state's component means (stats4_kernel)
   DevBuf<int> dMeanId, dVarId, dStateMixOff, dMixGauss;
   GmmTcModel tc;                    // expanded / split operands for the tcgen05 path
   GmmTc3Model tc3;                  // operands of gmm_tc3_kernel (fused expansion, taper skipping, single-Gaussian sets)
   bool useV3 = false;               // K1 = gmm_tc3_kernel
   // Two-model re-estimation (hfb_options.alignModel): THIS context holds the alignment set and runs every kernel up to
   // alpha on it; `upd` is a complete context of the update set that owns the accumulators (its layout, plus a tail of
   // P_align doubles where the alignment kernels' numEgs increments land and are ignored), the M-step and the model
   // tables stats_two_kernel reads.
   hfbgpu_ctx *upd = nullptr;
   // accumulators
   DevBuf<double> dAcc;
   DevBuf<int> dHmmN, dHmmStateOff, dHmmState, dHmmTrans, dTransOffF, dTransMinDur;
   DevBuf<long long> dTranAccOff, dTranOccOff;
   // Waves of one call run on NSLOT streams with private workspaces: the latency-bound recursion
   // kernels of one wave overlap the throughput-bound GMM / statistics kernels of the other.
   struct Slot {
      cudaStream_t stream = nullptr;
      cudaStream_t recStream = nullptr;   // HFBGPU_REC_STREAM experiment
      cudaEvent_t ev[6] = {};
      cudaEvent_t evIn = nullptr, evGmm = nullptr;   // inputs uploaded / output probabilities ready
      cudaEvent_t evX = nullptr;        // timing mode: feature expansion done, tensor-core kernel starts
      bool hasX = false;
      DevBuf<float> dFeat;              // only for host-feature calls and for expanded features
      DevBuf<float> dFeatSrc;           // host static coefficients before the qualifier expansion
      DevBuf<float> dFeat2;             // host second data stream (single-pass retraining)
      DevBuf<short> dFeatC;             // host 16-bit integers of `_C` compressed files before feat_decompress_kernel
      DevBuf<float> dB;
      DevBuf<double> dBeta, dOcc, dAent;
      DevBuf<short> dBeams;             // 4 * frames
      DevBuf<unsigned char> dTables, dScratch;
      DevBuf<int> dStateIdx;            // 3 x (J + 2): count / offset / fill arrays of the by-state position sort, then
                                        // the 8-byte cursor of the valid-frame list and one count per position
      DevBuf<ValidFrame> dValid;        // valid-frame list of stats_pre_kernel
      GmmTcWork tcw;
      std::vector<unsigned char> blob;
      unsigned char *hTables = nullptr; // pinned staging
      size_t hTablesCap = 0;
      UttOut *hOut = nullptr;           // pinned
      size_t hOutCap = 0;
      short *hBeams = nullptr;          // pinned
      size_t hBeamsCap = 0;
      // in-flight wave
      bool busy = false;
      int64_t ticket = 0;               // hfbgpu_submit call this wave belongs to
      hfb_utt_result *res = nullptr;    // where the in-flight wave's results go
      hfb_beams beams = {};
      WaveTablesPtr w = nullptr;
      long long waveFrame0 = 0, waveFrames = 0;
      bool wantBeams = false;
   };
   static const int NSLOT = 4;
   Slot slot[NSLOT];
   int numSlots = NSLOT;
   unsigned nextSlot = 0;
   int64_t submitSeq = 0;               // tickets of hfbgpu_submit (hfbgpu_wait_ticket)
   // ---- device group (hfbgpu_create_multi): this context owns no device memory, it drives one child per GPU
   std::vector<hfbgpu_ctx *> kids;
   std::vector<std::vector<int64_t>> kidTickets;   // [group ticket % 64][kid]
   bool reduced = true;                 // children 1.. hold nothing that is not already in child 0
   size_t workspaceBytes = 0;
   int smCount = 148;
   int maxSmemOptin = 0;
};

// ------------------------------------------------------------------------------------------
// host-only helpers
// ------------------------------------------------------------------------------------------
extern "C" int hfbgpu_abi_version(void) { return HFBGPU_ABI_VERSION; }

extern "C" int hfbgpu_acc_layout(const hfb_model *m, hfb_acc_layout *L)
{
   if (!m || !L) return HFB_EINVAL;
   long long o = 0, nn = 0, n = 0;
   for (int i = 0; i < m->numTrans; i++) { nn += (long long)m->transN[i] * m->transN[i]; n += m->transN[i]; }
   L->tran = o;    o += nn;
   L->tranOcc = o; o += n;
   L->wtC = o;     o += m->stateMixOff[m->numStates];
   L->wtOcc = o;   o += m->numStates;
   L->muSum = o;   o += (long long)m->numMeanAcc * m->vecSize;
   L->muOcc = o;   o += m->numMeanAcc;
   L->vaSum = o;   o += (long long)m->numVarAcc * m->vecSize;
   L->vaOcc = o;   o += m->numVarAcc;
   L->numEgs = o;  o += m->numHmm;
   L->totalT = o++; L->totalPr = o++; L->numOk = o++; L->numSkipped = o++;
   L->count = o; L->tranOccStride = 0;
   return HFB_OK;
}

extern "C" void hfbgpu_default_options(hfb_options *o)
{
   if (!o) return;
   memset(o, 0, sizeof(*o));
   o->pruneInit = HFB_NOPRUNE; o->pruneInc = 0.0; o->pruneLim = HFB_NOPRUNE;   // HFB.c:83
   o->minFrwdP = 10.0f;
   o->uFlags = HFB_UPMEANS | HFB_UPVARS | HFB_UPTRANS | HFB_UPMIXES;
   o->device = 0; o->gmmKernel = 0; o->workspaceBytes = 0;
}

extern "C" const char *hfbgpu_strerror(int code)
{
   switch (code) {
   case HFB_OK: return "ok";
   case HFB_EINVAL: return "invalid argument";
   case HFB_ENODEVICE: return "no CUDA device (libhfbgpu has no CPU fallback)";
   case HFB_ECUDA: return "CUDA error";
   case HFB_ENOMEM: return "out of device memory";
   case HFB_EUNSUPPORTED: return "model feature outside the accelerated path";
   case HFB_ETEE: return "CreateInsts: tee model first, last or twice in a row (HError 7332)";
   case HFB_EALPHAPRUNE: return "StepAlpha: alpha prune failed (HError 7390)";
   case HFB_EBETAPRUNE: return "SetBeta: beta prune failed (HError 7323)";
   case HFB_UTT_SKIPPED: return "StepBack: bad data or over pruning (HError -7324)";
   default: return "unknown";
   }
}

extern "C" const char *hfbgpu_last_error(void) { return g_lastError.c_str(); }

extern "C" int hfbgpu_device_count(void)
{
   int n = 0;
   if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
   return n;
}

// FindStateOrder + SetMinDurs, HTKLib/HFB.c:91-155.
static void order_states(const float *A, int N, std::vector<int> &ord, int s, int &cnt)
{
   ord[s] = 0;
   for (int p = 0; p < N - 1; p++)
      if (A[p * N + s] > HFB_LSMALL && p != s && ord[p] < 0) order_states(A, N, ord, p, cnt);
   ord[s] = ++cnt;
}

static int min_duration(const float *A, int N)
{
   std::vector<int> ord(N, -1), byOrd(N + 1, -1), md(N, N);
   int cnt = 0;
   order_states(A, N, ord, N - 1, cnt);
   for (int i = 0; i < N; i++) if (ord[i] > 0) byOrd[ord[i]] = i;
   md[0] = 0;
   for (int k = 1; k <= cnt; k++) {
      int i = byOrd[k];
      if (i < 0) continue;
      for (int j = 0; j < N - 1; j++)
         if (A[j * N + i] > HFB_LSMALL) {
            int d = md[j] + ((i == N - 1) ? 0 : 1);
            if (d < md[i]) md[i] = d;
         }
   }
   if (md[N - 1] < 0 || md[N - 1] >= N) return (A[N - 1] > HFB_LSMALL) ? 0 : 1;   // HFB.c:144-149
   return md[N - 1];
}

// Standard topology test for the specialised recursion kernels (hfb_l2r.cuh): N = 5 and every
// transition other than entry->2, i->i, i->i+1 (i = 2..4) is log zero.
static bool matrix_is_l2r(const float *A, int N)
{
   if (N != 5) return false;
   for (int i = 0; i < N; i++)
      for (int j = 0; j < N; j++) {
         const bool structural = (i == 0 && j == 1) || (i >= 1 && i <= 3 && (j == i || j == i + 1));
         if (!structural && A[i * N + j] > HFB_LSMALL) return false;
      }
   return true;
}

// ------------------------------------------------------------------------------------------
// create / destroy
// ------------------------------------------------------------------------------------------
template <class T>
static int upload(DevBuf<T> &b, const T *src, size_t n, cudaStream_t st)
{
   int rc = b.reserve(n ? n : 1);
   if (rc) return rc;
   if (n) CK(cudaMemcpyAsync(b.p, src, n * sizeof(T), cudaMemcpyHostToDevice, st));
   return HFB_OK;
}

static int create_one(hfbgpu_ctx **out, const hfb_model *m, const hfb_options *opt);

static double *accp(hfbgpu_ctx *c) { return c->upd ? c->upd->dAcc.p : c->dAcc.p; }

extern "C" int hfbgpu_create(hfbgpu_ctx **out, const hfb_model *m, const hfb_options *opt)
{
   if (!out || !m || !opt) return HFB_EINVAL;
   *out = nullptr;
   if (!opt->alignModel) return create_one(out, m, opt);
   // ---- two-model re-estimation: UseAlignHMMSet, HFB.c:296-333
   const hfb_model *al = opt->alignModel;
   if (al->vecSize != m->vecSize) { g_lastError = "alignment and update sets differ in vector size (HError 7392)"; return HFB_EINVAL; }
   hfb_options o = *opt;
   o.alignModel = nullptr;
   hfbgpu_ctx *u = nullptr, *a = nullptr;
   hfb_options ou = o;
   ou.gmmKernel = 1;                                    // the update set evaluates nothing on the tensor cores
   int rc = create_one(&u, m, &ou);
   if (rc) return rc;
   o.uFlags &= ~HFB_UPTRANS;                            // "Don't update transitions on a 2-model alignment", HFB.c:313-316
   if ((rc = create_one(&a, al, &o))) { hfbgpu_destroy(u); return rc; }
   // the accumulators are the update set's; the alignment kernels index totalT / totalPr / numOk / numSkipped through
   // their DevModel's layout, and numEgs by alignment HMM: that one goes to a tail nobody reads
   a->dAcc.release();
   a->L = u->L;
   a->dm.L = u->L;
   a->dm.L.numEgs = u->L.count;
   if ((rc = u->dAcc.reserve((size_t)u->L.count + (size_t)al->numHmm + 1))) { hfbgpu_destroy(u); hfbgpu_destroy(a); return rc; }
   if (cudaMemset(u->dAcc.p, 0, ((size_t)u->L.count + (size_t)al->numHmm + 1) * sizeof(double)) != cudaSuccess) {
      cudaGetLastError(); hfbgpu_destroy(u); hfbgpu_destroy(a); return HFB_ECUDA;
   }
   a->upd = u;
   cudaFuncSetAttribute(stats_two_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, a->maxSmemOptin);
   *out = a;
   return HFB_OK;
}

static int create_one(hfbgpu_ctx **out, const hfb_model *m, const hfb_options *opt)
{
   if (m->vecSize < 1 || m->numGauss < 1 || m->numStates < 1 || m->numHmm < 1 || m->numTrans < 1) return HFB_EINVAL;
   if (m->vecSize > 64) { g_lastError = "vecSize > 64 is outside the accelerated path"; return HFB_EUNSUPPORTED; }
   int ndev = hfbgpu_device_count();
   if (ndev <= 0) { g_lastError = "no CUDA device visible"; return HFB_ENODEVICE; }
   if (opt->device < 0 || opt->device >= ndev) return HFB_EINVAL;
   CK(cudaSetDevice(opt->device));

   hfbgpu_ctx *c = new hfbgpu_ctx();
   c->device = opt->device;
   c->opt = *opt;
   memset(&c->stats, 0, sizeof(c->stats));
   hfbgpu_acc_layout(m, &c->L);
   cudaDeviceProp prop;
   CK(cudaGetDeviceProperties(&prop, c->device));
   c->smCount = prop.multiProcessorCount;
   c->maxSmemOptin = (int)prop.sharedMemPerBlockOptin;
   CK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
   c->stream = c->ownStream;
   int prLo = 0, prHi = 0;
   CK(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
   CK(cudaStreamCreateWithPriority(&c->gmmStream, cudaStreamNonBlocking, prHi));
   for (auto &sl : c->slot) {
      CK(cudaStreamCreateWithFlags(&sl.stream, cudaStreamNonBlocking));
      CK(cudaStreamCreateWithPriority(&sl.recStream, cudaStreamNonBlocking, prHi));
      for (auto &e : sl.ev) CK(cudaEventCreate(&e));
      CK(cudaEventCreateWithFlags(&sl.evIn, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&sl.evGmm, cudaEventDisableTiming));
      CK(cudaEventCreate(&sl.evX));
   }
   c->trace = getenv("HFBGPU_TRACE_KERNELS") != nullptr;
   if (c->trace) { CK(cudaEventCreate(&c->evRef)); CK(cudaEventRecord(c->evRef, c->stream)); }
   c->waveUtts = 16384;
   if (const char *wu = getenv("HFBGPU_WAVE_UTTS")) c->waveUtts = std::max(1, atoi(wu));
   if (const char *ns = getenv("HFBGPU_STREAMS")) c->numSlots = std::max(1, std::min((int)hfbgpu_ctx::NSLOT, atoi(ns)));

   HostModel &h = c->hm;
   h.D = m->vecSize; h.G = m->numGauss; h.J = m->numStates; h.P = m->numHmm; h.numTrans = m->numTrans;
   h.stateMixOff.assign(m->stateMixOff, m->stateMixOff + h.J + 1);
   h.hmmN.assign(m->hmmNumStates, m->hmmNumStates + h.P);
   h.hmmStateOff.assign(m->hmmStateOff, m->hmmStateOff + h.P + 1);
   h.hmmState.assign(m->hmmState, m->hmmState + m->hmmStateOff[h.P]);
   h.hmmTrans.assign(m->hmmTrans, m->hmmTrans + h.P);
   h.transN.assign(m->transN, m->transN + h.numTrans);
   h.transOff.assign(m->transOff, m->transOff + h.numTrans + 1);
   h.transLogA.assign(m->transLogA, m->transLogA + m->transOff[h.numTrans]);
   h.meanId.assign(m->meanId, m->meanId + h.G); h.varId.assign(m->varId, m->varId + h.G);
   h.numMeanAcc = m->numMeanAcc; h.numVarAcc = m->numVarAcc;
   h.maxM = 0; h.maxN = 0; h.l2r = true;
   for (int j = 0; j < h.J; j++) h.maxM = std::max(h.maxM, h.stateMixOff[j + 1] - h.stateMixOff[j]);
   h.minDur.resize(h.numTrans);
   h.tranAccOff.resize(h.numTrans); h.tranOccOff.resize(h.numTrans);
   long long a = c->L.tran, b = c->L.tranOcc;
   for (int i = 0; i < h.numTrans; i++) {
      int N = h.transN[i];
      h.maxN = std::max(h.maxN, N);
      h.minDur[i] = min_duration(h.transLogA.data() + h.transOff[i], N);       // SetMinDurs
      if (!matrix_is_l2r(h.transLogA.data() + h.transOff[i], N) || h.minDur[i] != 3) h.l2r = false;
      h.tranAccOff[i] = a; h.tranOccOff[i] = b;
      a += (long long)N * N; b += N;
   }
   for (int p = 0; p < h.P; p++)
      if (h.hmmN[p] != h.transN[h.hmmTrans[p]] || h.hmmStateOff[p + 1] - h.hmmStateOff[p] != h.hmmN[p] - 2) {
         delete c; g_lastError = "hmmNumStates inconsistent with transN / hmmStateOff"; return HFB_EINVAL;
      }
   if (h.maxN > HFB_MAXN) { delete c; g_lastError = "HMM with more than 16 states"; return HFB_EUNSUPPORTED; }

   // model upload (means / inverse variances padded to a multiple of 4 floats per row)
   const int D = h.D, Dp = (D + 3) & ~3;
   std::vector<float> mean((size_t)h.G * Dp, 0.f), ivar((size_t)h.G * Dp, 0.f);
   for (int g = 0; g < h.G; g++)
      for (int k = 0; k < D; k++) {
         mean[(size_t)g * Dp + k] = m->mean[(size_t)g * D + k];
         ivar[(size_t)g * Dp + k] = m->ivar[(size_t)g * D + k];
      }
   int rc;
   const int sumM = h.stateMixOff[h.J];
   if ((rc = upload(c->dMean, mean.data(), mean.size(), c->stream)) ||
       (rc = upload(c->dIvar, ivar.data(), ivar.size(), c->stream)) ||
       (rc = upload(c->dGconst, m->gConst, (size_t)h.G, c->stream)) ||
       (rc = upload(c->dMeanId, m->meanId, (size_t)h.G, c->stream)) ||
       (rc = upload(c->dVarId, m->varId, (size_t)h.G, c->stream)) ||
       (rc = upload(c->dStateMixOff, m->stateMixOff, (size_t)h.J + 1, c->stream)) ||
       (rc = upload(c->dMixGauss, m->mixGauss, (size_t)sumM, c->stream)) ||
       (rc = upload(c->dMixLogWt, m->mixLogWt, (size_t)sumM, c->stream)) ||
       (rc = upload(c->dTransLogA, h.transLogA.data(), h.transLogA.size(), c->stream)) ||
       (rc = upload(c->dHmmN, h.hmmN.data(), h.hmmN.size(), c->stream)) ||
       (rc = upload(c->dHmmStateOff, h.hmmStateOff.data(), h.hmmStateOff.size(), c->stream)) ||
       (rc = upload(c->dHmmState, h.hmmState.data(), h.hmmState.size(), c->stream)) ||
       (rc = upload(c->dHmmTrans, h.hmmTrans.data(), h.hmmTrans.size(), c->stream)) ||
       (rc = upload(c->dTransOffF, h.transOff.data(), (size_t)h.numTrans, c->stream)) ||
       (rc = upload(c->dTransMinDur, h.minDur.data(), h.minDur.size(), c->stream)) ||
       (rc = upload(c->dTranAccOff, h.tranAccOff.data(), h.tranAccOff.size(), c->stream)) ||
       (rc = upload(c->dTranOccOff, h.tranOccOff.data(), h.tranOccOff.size(), c->stream))) {
      hfbgpu_destroy(c); return rc;
   }
   {  // state centres for the tensor-core statistics kernel (hfb_stats_mma.cuh)
      std::vector<float> cen((size_t)h.J * S4_CSTR, 0.f);
      if (D <= S4_CSTR)
         for (int s2 = 0; s2 < h.J; s2++) {
            const int o = h.stateMixOff[s2], n = h.stateMixOff[s2 + 1] - o;
            for (int k = 0; k < D; k++) {
               double a = 0.0;
               for (int i = 0; i < n; i++) a += m->mean[(size_t)m->mixGauss[o + i] * D + k];
               cen[(size_t)s2 * S4_CSTR + k] = (float)(a / std::max(n, 1));
            }
         }
      if ((rc = upload(c->dCentre, cen.data(), cen.size(), c->stream))) { hfbgpu_destroy(c); return rc; }
   }
   CK(cudaStreamSynchronize(c->stream));
   DevModel &d = c->dm;
   d.D = D; d.Dp = Dp; d.G = h.G; d.J = h.J; d.P = h.P; d.numTrans = h.numTrans; d.maxM = h.maxM;
   d.mean = c->dMean.p; d.ivar = c->dIvar.p; d.gconst = c->dGconst.p;
   d.meanId = c->dMeanId.p; d.varId = c->dVarId.p;
   d.stateMixOff = c->dStateMixOff.p; d.mixGauss = c->dMixGauss.p; d.mixLogWt = c->dMixLogWt.p;
   d.transLogA = c->dTransLogA.p;
   d.hmmN = c->dHmmN.p; d.hmmStateOff = c->dHmmStateOff.p; d.hmmState = c->dHmmState.p; d.hmmTrans = c->dHmmTrans.p;
   d.transOffF = c->dTransOffF.p; d.transMinDur = c->dTransMinDur.p;
   d.tranAccOff = c->dTranAccOff.p; d.tranOccOff = c->dTranOccOff.p;
   d.L = c->L;

   if ((rc = c->dAcc.reserve((size_t)c->L.count))) { hfbgpu_destroy(c); return rc; }
   CK(cudaMemsetAsync(c->dAcc.p, 0, (size_t)c->L.count * sizeof(double), c->stream));

   // tensor-core operands (expanded quadratic form, 3xTF32 split)
   if ((rc = gmm_tc_prepare(c->tc, m, c->stream))) { hfbgpu_destroy(c); return rc; }
   {
      void *fn = c->tc.encodeFn;
      cudaDriverEntryPointQueryResult qr;
      if (!fn && (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || !fn)) { cudaGetLastError(); fn = nullptr; }
      if (!getenv("HFBGPU_GMM_V2") && !getenv("HFBGPU_TC_TF32") && !getenv("HFBGPU_NO_PAIR") && opt->gmmKernel != 1 &&
          (rc = gmm_tc3_prepare(c->tc3, m, c->stream, fn))) { hfbgpu_destroy(c); return rc; }
      c->useV3 = c->tc3.ready && c->smCount >= 2;
   }
   // HFBGPU_L2_PERSIST=1 (experiment): the Gaussian operand of K1 (41 MB for config #3, 66 MB for #5) is read once per
   // 512-frame work item by every CTA pair and should live in the 126 MB L2; the output probabilities K1 writes and the
   // arrays the recursion kernels of other waves stream through it compete for the same lines.  The window asks the L2 to
   // keep the operand (persisting) on every stream the library launches on.
   if (c->useV3 && getenv("HFBGPU_L2_PERSIST") && atoi(getenv("HFBGPU_L2_PERSIST")) != 0 && c->tc3.bBytes > 0) {
      cudaDeviceProp pr2;
      cudaGetDeviceProperties(&pr2, c->device);
      const size_t want = std::min((size_t)pr2.persistingL2CacheMaxSize, c->tc3.bBytes);
      cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
      cudaStreamAttrValue av;
      memset(&av, 0, sizeof(av));
      av.accessPolicyWindow.base_ptr = c->tc3.dBhi;
      av.accessPolicyWindow.num_bytes = std::min((size_t)pr2.accessPolicyMaxWindowSize, c->tc3.bBytes);
      av.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)av.accessPolicyWindow.num_bytes);
      av.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
      av.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
      cudaStreamSetAttribute(c->gmmStream, cudaStreamAttributeAccessPolicyWindow, &av);
      for (auto &sl : c->slot) cudaStreamSetAttribute(sl.stream, cudaStreamAttributeAccessPolicyWindow, &av);
      fprintf(stderr, "[hfbgpu] L2 window: %zu MB of the %zu MB operand persisting (device maximum %zu MB)\n", want >> 20,
              c->tc3.bBytes >> 20, (size_t)pr2.persistingL2CacheMaxSize >> 20);
      cudaGetLastError();
   }
   if (opt->gmmKernel == 2 && !gmm_tc_available(c->tc) && !c->useV3) {
      hfbgpu_destroy(c); g_lastError = "tcgen05 GMM kernel requested but not available for this model";
      return HFB_EUNSUPPORTED;
   }

   size_t freeB = 0, totalB = 0;
   CK(cudaMemGetInfo(&freeB, &totalB));
   // default: 70 % of what is free now (the per-wave buffers grow on demand up to this; models, accumulators and the
   // caller's own feature buffers live in the rest)
   size_t ws = opt->workspaceBytes ? opt->workspaceBytes : freeB / 10 * 7;
   if (ws > freeB / 10 * 8) ws = freeB / 10 * 8;
   c->workspaceBytes = ws;

   int maxOpt = c->maxSmemOptin;
   cudaFuncSetAttribute(beta_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(beta_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(alpha_warp_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(alpha_warp_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(stats3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(beta_l2r_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(stats5_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   cudaFuncSetAttribute(stats5_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, maxOpt);
   // HFBGPU_CARVEOUT=1 (experiment, "small waves" in launch_wave): the recursion kernels ask for the same shared-memory /
   // L1 split as the tensor-core kernel (maximum shared memory), because CTAs of kernels whose splits differ cannot share an SM
   if (getenv("HFBGPU_CARVEOUT") && atoi(getenv("HFBGPU_CARVEOUT")) != 0) {
      const int co = cudaSharedmemCarveoutMaxShared;
      cudaFuncSetAttribute(beta_l2r_warp_kernel<false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaFuncSetAttribute(beta_l2r_warp_kernel<false, false>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaFuncSetAttribute(beta_l2r_kernel<1024>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaFuncSetAttribute(alpha_l2r_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaFuncSetAttribute(alpha_warp_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaFuncSetAttribute(prep_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co);
      cudaGetLastError();
   }
   stats_tc_set_attributes();
   gmm_tc4_set_attributes();
   CK(cudaStreamSynchronize(c->stream));
   *out = c;
   return HFB_OK;
}

extern "C" int hfbgpu_destroy(hfbgpu_ctx *c)
{
   if (!c) return HFB_EINVAL;
   if (!c->kids.empty()) {
      for (auto *k : c->kids) hfbgpu_destroy(k);
      delete c;
      return HFB_OK;
   }
   cudaSetDevice(c->device);
   if (c->stream) cudaStreamSynchronize(c->stream);
   if (c->upd) { hfbgpu_destroy(c->upd); c->upd = nullptr; }
   c->dCentre.release();
   c->dMean.release(); c->dIvar.release(); c->dGconst.release(); c->dMixLogWt.release(); c->dTransLogA.release();
   c->dMeanId.release(); c->dVarId.release(); c->dStateMixOff.release(); c->dMixGauss.release();
   gmm_tc_release(c->tc);
   gmm_tc3_release(c->tc3);
   c->dAcc.release();
   for (auto &sl : c->slot) {
      if (sl.stream) cudaStreamSynchronize(sl.stream);
      sl.dFeat.release(); sl.dB.release(); sl.dBeta.release(); sl.dOcc.release(); sl.dAent.release(); sl.dBeams.release();
      sl.dTables.release(); sl.dScratch.release(); sl.dStateIdx.release(); sl.dValid.release(); sl.dFeatSrc.release(); sl.dFeat2.release(); sl.tcw.release();
      if (sl.hTables) cudaFreeHost(sl.hTables);
      if (sl.hOut) cudaFreeHost(sl.hOut);
      if (sl.hBeams) cudaFreeHost(sl.hBeams);
      for (auto &e : sl.ev) if (e) cudaEventDestroy(e);
      if (sl.evIn) cudaEventDestroy(sl.evIn);
      if (sl.evGmm) cudaEventDestroy(sl.evGmm);
      if (sl.evX) cudaEventDestroy(sl.evX);
      if (sl.stream) cudaStreamDestroy(sl.stream);
      if (sl.recStream) cudaStreamDestroy(sl.recStream);
      delete sl.w;
   }
   c->dHmmN.release(); c->dHmmStateOff.release(); c->dHmmState.release(); c->dHmmTrans.release();
   c->dTransOffF.release(); c->dTransMinDur.release(); c->dTranAccOff.release(); c->dTranOccOff.release();
   if (c->gmmStream) { cudaStreamSynchronize(c->gmmStream); cudaStreamDestroy(c->gmmStream); }
   if (c->ownStream) cudaStreamDestroy(c->ownStream);
   delete c;
   return HFB_OK;
}

extern "C" int hfbgpu_set_stream(hfbgpu_ctx *c, void *st)
{
   if (!c) return HFB_EINVAL;
   if (!c->kids.empty()) { g_lastError = "a device group runs on the library's own streams"; return HFB_EUNSUPPORTED; }
   CK(cudaSetDevice(c->device));
   CK(cudaStreamSynchronize(c->stream));
   c->stream = st ? (cudaStream_t)st : c->ownStream;
   return HFB_OK;
}

static int wait_impl(hfbgpu_ctx *c);
static bool is_group(const hfbgpu_ctx *c);
static int group_reduce(hfbgpu_ctx *g);

extern "C" int hfbgpu_zero_accs(hfbgpu_ctx *c)
{
   if (!c) return HFB_EINVAL;
   if (is_group(c)) {
      int rcAll = HFB_OK;
      for (auto *k : c->kids) { int rc = hfbgpu_zero_accs(k); if (rc && !rcAll) rcAll = rc; }
      c->reduced = true;
      return rcAll;
   }
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }
   CK(cudaMemsetAsync(accp(c), 0, ((size_t)c->L.count + (c->upd ? (size_t)c->hm.P + 1 : 0)) * sizeof(double), c->stream));
   CK(cudaStreamSynchronize(c->stream));
   return HFB_OK;
}

extern "C" double *hfbgpu_acc_device_ptr(hfbgpu_ctx *c) { return !c ? nullptr : (is_group(c) ? accp(c->kids[0]) : accp(c)); }
extern "C" int64_t hfbgpu_acc_count(hfbgpu_ctx *c) { return c ? c->L.count : 0; }

static int wait_impl(hfbgpu_ctx *c);

extern "C" int hfbgpu_get_accs(hfbgpu_ctx *c, double *hostOut)
{
   if (!c || !hostOut) return HFB_EINVAL;
   if (is_group(c)) { int rc = group_reduce(c); return rc ? rc : hfbgpu_get_accs(c->kids[0], hostOut); }
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }
   CK(cudaStreamSynchronize(c->stream));
   CK(cudaMemcpy(hostOut, accp(c), (size_t)c->L.count * sizeof(double), cudaMemcpyDeviceToHost));
   c->stats.d2hBytes += c->L.count * (int64_t)sizeof(double);
   return HFB_OK;
}

extern "C" int hfbgpu_set_accs(hfbgpu_ctx *c, const double *hostIn)
{
   if (!c || !hostIn) return HFB_EINVAL;
   if (is_group(c)) {
      int rc = hfbgpu_zero_accs(c);
      return rc ? rc : hfbgpu_set_accs(c->kids[0], hostIn);
   }
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }        // never under an in-flight wave
   CK(cudaStreamSynchronize(c->stream));
   CK(cudaMemcpy(accp(c), hostIn, (size_t)c->L.count * sizeof(double), cudaMemcpyHostToDevice));
   return HFB_OK;
}

extern "C" int hfbgpu_get_min_durs(hfbgpu_ctx *c, int32_t *out)
{
   if (!c || !out) return HFB_EINVAL;
   if (is_group(c)) return hfbgpu_get_min_durs(c->kids[0], out);
   for (int i = 0; i < c->hm.numTrans; i++) out[i] = c->hm.minDur[i];
   return HFB_OK;
}

extern "C" int hfbgpu_get_stats(hfbgpu_ctx *c, hfb_stats *o)
{
   if (!c || !o) return HFB_EINVAL;
   if (!is_group(c)) { *o = c->stats; return HFB_OK; }
   memset(o, 0, sizeof(*o));                           // counters: sums over the devices; times: the slowest device
   for (auto *k : c->kids) {
      const hfb_stats &s = k->stats;
      o->launches += s.launches; o->launchesGmm += s.launchesGmm; o->launchesBeta += s.launchesBeta; o->launchesAlpha += s.launchesAlpha;
      o->launchesStats += s.launchesStats; o->launchesMisc += s.launchesMisc; o->launchesL2R += s.launchesL2R;
      o->betaCells += s.betaCells; o->alphaCells += s.alphaCells; o->gmmPairs += s.gmmPairs; o->h2dBytes += s.h2dBytes; o->d2hBytes += s.d2hBytes;
      o->msGmm = std::max(o->msGmm, s.msGmm); o->msBeta = std::max(o->msBeta, s.msBeta); o->msAlpha = std::max(o->msAlpha, s.msAlpha);
      o->msStats = std::max(o->msStats, s.msStats); o->msExpand = std::max(o->msExpand, s.msExpand);
   }
   return HFB_OK;
}
extern "C" int hfbgpu_reset_stats(hfbgpu_ctx *c)
{
   if (!c) return HFB_EINVAL;
   for (auto *k : c->kids) hfbgpu_reset_stats(k);
   memset(&c->stats, 0, sizeof(c->stats));
   return HFB_OK;
}
extern "C" int hfbgpu_set_timing(hfbgpu_ctx *c, int on)
{
   if (!c) return HFB_EINVAL;
   for (auto *k : c->kids) hfbgpu_set_timing(k, on);
   c->timing = on != 0;
   return HFB_OK;
}

// ------------------------------------------------------------------------------------------
// wave construction: the host only sizes things; the tables are built by prep_kernel
// ------------------------------------------------------------------------------------------
namespace {

// Sizes one utterance (sum of N over its labels) and appends its descriptor.
// Returns the workspace bytes it needs.
size_t add_utterance(const HostModel &h, WaveTables &w, int uidx, int T, const int32_t *lab, int Q, int labOff,
                     long long featOff, int globalSlots)
{
   UttDesc u;
   memset(&u, 0, sizeof(u));
   UttOut o;
   memset(&o, 0, sizeof(o));
   o.pr = HFB_LZERO;
   const int uLocal = (int)w.utt.size();
   int S = 0;
   bool bad = (Q < 1 || T < 1);
   for (int q = 0; q < Q && !bad; q++) {
      const int p = lab[q];
      if (p < 0 || p >= h.P) bad = true; else { S += h.hmmN[p]; w.maxN = std::max(w.maxN, h.hmmN[p]); }
   }
   if (bad) { o.status = HFB_UTT_ETEE; Q = 0; S = 0; }
   const int Pp = S - 2 * Q;
   const int slots = bad ? 0 : std::max(Pp, globalSlots);                     // output-probability slots to allocate
   u.T = T; u.Q = Q; u.S = S; u.P = Pp; u.J = (slots + 3) & ~3; u.Jt = slots;  // J, Jt: bounds, prep_kernel sets them
   w.maxT = std::max(w.maxT, T);
   u.labOff = labOff; u.modOff = (int)w.totalQ; u.slotOff = (int)w.totalSl; u.posOff = (int)w.totalP;
   u.featOff = featOff; u.frameBase = featOff;
   u.bOff = w.bFloats; u.betaOff = w.betaDoubles; u.occOff = w.occDoubles; u.aentOff = w.aentDoubles;
   w.posPre.push_back((int)w.totalP);
   w.tilePre.push_back((int)w.tiles);
   size_t bytes = 0;
   if (!bad) {
      w.bFloats += (long long)T * ((slots + 3) & ~3); w.betaDoubles += (long long)T * S; w.occDoubles += (long long)T * Pp;
      w.aentDoubles += (long long)T * Q;
      bytes = (size_t)T * ((size_t)Pp * 12 + (size_t)S * 8 + (size_t)Q * 8);
      w.totalQ += Q; w.totalP += Pp; w.totalSl += slots;
      w.tiles += (long long)((T + GT_FR - 1) / GT_FR) * ((Pp + GT_SL - 1) / GT_SL);
      for (int t0 = 0; t0 < T; t0 += TC_BM) w.tcItems.push_back(make_int2(uLocal, t0));
      for (int t0 = 0; t0 < T; t0 += 2 * TC_BM) w.tcItems2.push_back(make_int2(uLocal, t0));
      for (int t0 = 0; t0 < T; t0 += 4 * TC_BM) w.tcItems4.push_back(make_int2(uLocal, t0));
      w.maxQ = std::max(w.maxQ, Q); w.maxS = std::max(w.maxS, S);
   }
   w.utt.push_back(u); w.out.push_back(o); w.uttIndex.push_back(uidx);
   return bytes;
}

template <class T>
size_t blob_put(std::vector<unsigned char> &blob, const T *v, size_t n)
{
   size_t off = (blob.size() + 255) & ~(size_t)255;
   blob.resize(off + n * sizeof(T) + 8);
   if (n) memcpy(blob.data() + off, v, n * sizeof(T));
   return off;
}
template <class T>
size_t blob_put(std::vector<unsigned char> &blob, const std::vector<T> &v) { return blob_put(blob, v.data(), v.size()); }

// device scratch written by prep_kernel / alpha kernel
struct ScratchLayout {
   size_t mN, mTrans, mSoff, mPoff, mDms, mPre, mSuf, mHmm, mTmin, mTmax, mTrAcc, mTrOcc, slotState, posSlot, posState, posQ, posList, bytes;
   size_t slotFirst, slotLast, tileIv;
   ScratchLayout(long long totalQ, long long totalP, long long totalSl)
   {
      size_t o = 0;
      auto take = [&](size_t n, size_t el) { size_t r = o; o += ((n * el + 255) & ~(size_t)255) + 256; return r; };
      const size_t q = (size_t)totalQ + 1, pp = (size_t)totalP + 1, ns = (size_t)totalSl + 1;
      mTrAcc = take(q, 8); mTrOcc = take(q, 8);
      mN = take(q, 4); mTrans = take(q, 4); mSoff = take(q, 4); mPoff = take(q, 4); mDms = take(q, 4);
      mPre = take(q, 4); mSuf = take(q, 4); mHmm = take(q, 4); mTmin = take(q, 4); mTmax = take(q, 4);
      slotState = take(ns, 4); slotFirst = take(ns, 4); slotLast = take(ns, 4); tileIv = take(ns, 8);
      posSlot = take(pp, 4); posState = take(pp, 4); posQ = take(pp, 4); posList = take(pp, sizeof(PosRec));
      bytes = o;
   }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// one wave: launch (asynchronous) and finish (synchronise + hand results to the caller)
// ------------------------------------------------------------------------------------------
static int launch_wave_impl(hfbgpu_ctx *c, hfbgpu_ctx::Slot &S, const int32_t *labBase, const int32_t *labUpBase, const float *feat, const float *feat2,
                            bool featOnDevice, long long waveFrame0, long long waveFrames, bool wantBeams,
                            const hfb_compressed *cf = nullptr, int cfUtt0 = 0)
{
   WaveTables &w = *S.w;
   const int nU = (int)w.utt.size();
   cudaStream_t st = S.stream;
   int rc;
   const int D = c->hm.D;
   S.waveFrame0 = waveFrame0; S.waveFrames = waveFrames; S.wantBeams = wantBeams;
   S.busy = true;
   // ---- features (with qualifiers: the caller's matrix holds numStatic columns, expanded below into S.dFeat)
   const FeatQual &fq = c->qual;
   const int Dsrc = fq.enabled ? fq.numStatic : D;
   const float *dFeat, *dFeatSrc = nullptr;
   float *dStage = nullptr;                           // where feat_decompress_kernel writes (compressed input only)
   if (featOnDevice) dFeat = feat + (size_t)waveFrame0 * Dsrc;
   else if (cf) {
      // `_C` files: the 16-bit integers travel (half the bytes), the floats are formed on the device below
      DevBuf<float> &stage = fq.enabled ? S.dFeatSrc : S.dFeat;
      if ((rc = stage.reserve((size_t)waveFrames * Dsrc + 4)) || (rc = S.dFeatC.reserve((size_t)waveFrames * Dsrc + 8))) return rc;
      CK(cudaMemcpyAsync(S.dFeatC.p, cf->feat + (size_t)waveFrame0 * Dsrc, (size_t)waveFrames * Dsrc * sizeof(int16_t),
                         cudaMemcpyHostToDevice, st));
      c->stats.h2dBytes += (int64_t)waveFrames * Dsrc * sizeof(int16_t);
      dFeat = dStage = stage.p;
   } else {
      DevBuf<float> &stage = fq.enabled ? S.dFeatSrc : S.dFeat;
      if ((rc = stage.reserve((size_t)waveFrames * Dsrc + 4))) return rc;
      CK(cudaMemcpyAsync(stage.p, feat + (size_t)waveFrame0 * Dsrc, (size_t)waveFrames * Dsrc * sizeof(float),
                         cudaMemcpyHostToDevice, st));
      c->stats.h2dBytes += (int64_t)waveFrames * Dsrc * sizeof(float);
      dFeat = stage.p;
   }
   if (fq.enabled) {
      if ((rc = S.dFeat.reserve((size_t)waveFrames * D + 4))) return rc;
      dFeatSrc = dFeat; dFeat = S.dFeat.p;
   }
   const float *dFeat2 = nullptr;                     // second data stream: always full width
   if (feat2) {
      if (featOnDevice) dFeat2 = feat2 + (size_t)waveFrame0 * D;
      else {
         if ((rc = S.dFeat2.reserve((size_t)waveFrames * D + 4))) return rc;
         CK(cudaMemcpyAsync(S.dFeat2.p, feat2 + (size_t)waveFrame0 * D, (size_t)waveFrames * D * sizeof(float),
                            cudaMemcpyHostToDevice, st));
         c->stats.h2dBytes += (int64_t)waveFrames * D * sizeof(float);
         dFeat2 = S.dFeat2.p;
      }
   }
   // ---- pack + upload the (small) host tables
   w.posPre.push_back((int)w.totalP);
   w.tilePre.push_back((int)w.tiles);
   std::vector<unsigned char> &blob = S.blob;
   blob.clear();
   size_t oUtt = blob_put(blob, w.utt), oOut = blob_put(blob, w.out), oPp = blob_put(blob, w.posPre),
          oTp = blob_put(blob, w.tilePre), oIt = blob_put(blob, w.tcItems), oIt2 = blob_put(blob, w.tcItems2), oIt4 = blob_put(blob, w.tcItems4);
   const size_t nLab = (size_t)w.utt.back().labOff + (size_t)w.utt.back().Q;
   size_t oLab = blob_put(blob, labBase, nLab);
   const size_t oLabUp = labUpBase ? blob_put(blob, labUpBase, nLab) : 0;     // two-model re-estimation: up_qList
   const size_t oCA = cf ? blob_put(blob, cf->scaleA + (size_t)cfUtt0 * Dsrc, (size_t)nU * Dsrc) : 0;   // vectors A, B of the wave's files
   const size_t oCB = cf ? blob_put(blob, cf->scaleB + (size_t)cfUtt0 * Dsrc, (size_t)nU * Dsrc) : 0;
   if (blob.size() > S.hTablesCap) {
      if (S.hTables) cudaFreeHost(S.hTables);
      S.hTablesCap = blob.size() * 2;
      CK(cudaMallocHost(&S.hTables, S.hTablesCap));
   }
   memcpy(S.hTables, blob.data(), blob.size());
   const ScratchLayout sl(w.totalQ, w.totalP, w.totalSl);
   if ((rc = S.dTables.reserve(blob.size())) || (rc = S.dScratch.reserve(sl.bytes))) return rc;
   CK(cudaMemcpyAsync(S.dTables.p, S.hTables, blob.size(), cudaMemcpyHostToDevice, st));
   c->stats.h2dBytes += (int64_t)blob.size();
   if ((rc = S.dB.reserve((size_t)w.bFloats + 1)) || (rc = S.dBeta.reserve((size_t)w.betaDoubles + 1)) ||
       (rc = S.dOcc.reserve((size_t)w.occDoubles + 1)) || (rc = S.dAent.reserve((size_t)w.aentDoubles + 1)) ||
       (rc = S.dBeams.reserve((size_t)waveFrames * 4 + 4)))
      return rc;

   unsigned char *base = S.dTables.p, *sc = S.dScratch.p;
   Wave W;
   memset(&W, 0, sizeof(W));
   W.utt = (UttDesc *)(base + oUtt); W.out = (UttOut *)(base + oOut); W.numUtt = nU;
   W.lab = (const int *)(base + oLab); W.posPre = (const int *)(base + oPp); W.tilePre = (const int *)(base + oTp);
   W.totalPos = (int)w.totalP;
   W.mN = (int *)(sc + sl.mN); W.mTrans = (int *)(sc + sl.mTrans); W.mSoff = (int *)(sc + sl.mSoff);
   W.mPoff = (int *)(sc + sl.mPoff); W.mDms = (int *)(sc + sl.mDms); W.mPre = (int *)(sc + sl.mPre);
   W.mSuf = (int *)(sc + sl.mSuf); W.mHmm = (int *)(sc + sl.mHmm);
   W.mTrAcc = (long long *)(sc + sl.mTrAcc); W.mTrOcc = (long long *)(sc + sl.mTrOcc);
   W.mTmin = (int *)(sc + sl.mTmin); W.mTmax = (int *)(sc + sl.mTmax);
   W.slotState = (int *)(sc + sl.slotState); W.slotFirst = (int *)(sc + sl.slotFirst); W.slotLast = (int *)(sc + sl.slotLast);
   W.tileIv = (int2 *)(sc + sl.tileIv);
   W.posSlot = (int *)(sc + sl.posSlot); W.posState = (int *)(sc + sl.posState); W.posQ = (int *)(sc + sl.posQ);
   W.feat = dFeat; W.feat2 = dFeat2;
   if (cf) {
      feat_decompress_launch(W.utt, nU, w.maxT, S.dFeatC.p, (const float *)(base + oCA), (const float *)(base + oCB), dStage, Dsrc, st);
      c->stats.launches++; c->stats.launchesMisc++;
   }
   if (fq.enabled) {
      int nl = 0;
      feat_expand_launch(fq, W.utt, nU, w.maxT, dFeatSrc, S.dFeat.p, D, st, &nl);
      c->stats.launches += nl; c->stats.launchesMisc += nl;
   }
   W.b = S.dB.p; W.beta = S.dBeta.p; W.occ = S.dOcc.p; W.aent = S.dAent.p;
   W.qLo = S.dBeams.p; W.qHi = W.qLo + waveFrames; W.sq = W.qHi + waveFrames; W.eq = W.sq + waveFrames;
   W.acc = accp(c);
   W.pruneInit = c->opt.pruneInit; W.pruneInc = c->opt.pruneInc; W.pruneLim = c->opt.pruneLim;
   W.minFrwdP = (double)c->opt.minFrwdP; W.uFlags = c->opt.uFlags;

   {
      const int mp = c->useV3 ? c->tc3.MP : (gmm_tc_available(c->tc) ? c->tc.MP : 0);
      W.spt = (c->opt.gmmKernel != 1 && mp > 0) ? TC_BN / mp : 0;
      W.globalSlots = (c->useV3 && c->opt.gmmKernel != 1) ? c->tc3.globalSlots : 0;
      W.noTaperSkip = (getenv("HFBGPU_NO_TAPER_SKIP") || !(c->useV3 && c->opt.gmmKernel != 1)) ? 1 : 0;
   }
   const bool tm = c->timing;
   // Experiment kept behind HFBGPU_GMM_STREAM=1 (+ HFBGPU_WAVE_UTTS=592): the table-building and tensor-core
   // kernels of ALL waves go through one high-priority stream, back to back, so that the latency-bound
   // beta/alpha CTAs of wave w co-reside with gmm(w+1).  Measured on B200 (HFBGPU_TRACE_KERNELS timeline):
   // the kernels do overlap, but every one of them slows down by about the overlap gained (alpha, one warp
   // per utterance, 0.9 -> 4-6 ms; gmm 2.2 -> 3.6 ms) -- 110-112 M frames/s either way -- so the default
   // keeps each wave on its own stream and lets the hardware overlap only copies and kernel tails.
   // Small waves (long utterances, config #5: 96 per wave, four waves in flight).  On private streams the four K1 launches
   // interleave CTA by CTA, all end together, and the waves then march in lockstep -- K1 x 4 (19 ms), every beta pass at once
   // (8), every alpha pass (5), the statistics (3): 9.5 ms per wave (HFBGPU_TRACE_KERNELS).  Experiment behind
   // HFBGPU_SMALL_FIFO=1: the K1 launches of such waves through ONE stream in submission order, so that wave w's recursions
   // -- a handful of warps bound by the latency of their T-step chains -- run under K1 of wave w + 1; with it
   // HFBGPU_K1_RESERVE=1 (K1 leaves one SM per 8 utterances of the wave free) and -DBW_RING_MINB=9 (ring-window beta kernel
   // at 165 instead of 224 registers, no spills, so that a K1 CTA, a beta and an alpha warp fit one SM's register file).
   // Measured on config #5: the waves do stagger, but K1 then takes 10.9 ms instead of 5.6 next to the ~200 recursion warps
   // of its predecessors, whatever the two other switches say: 47 M frames/s against 61 M in lockstep.  With the 96-register
   // ("lean") build of K1, which this switch also selects, and the tables built on the wave's own stream, K1 keeps its 5.8 ms
   // beside the recursion warps -- but then THEY slow down (beta 13-18 ms instead of 6.5: a latency-bound chain that shares
   // its SM's issue slots with 14 busy warps): 53-58 M.  What is left is spatial separation (an SM partition per kernel
   // family, i.e. green contexts).  Left off.
   static const bool smallFifoOn = getenv("HFBGPU_SMALL_FIFO") && atoi(getenv("HFBGPU_SMALL_FIFO")) != 0;
   const bool smallWave = ((nU + 7) / 8) * 4 <= c->smCount;
   cudaStream_t sg = (!tm && (getenv("HFBGPU_GMM_STREAM") || (smallFifoOn && smallWave && c->useV3))) ? c->gmmStream : st;
   // ---- K0: tables (on the wave's own stream: with a shared K1 stream the tables of wave w + 1 are built under K1 of wave w,
   // so that K1 of wave w + 1 is ready to go the moment its predecessor ends -- and is placed before the recursion warps)
   const bool tr = tm || c->trace;
   if (c->trace) cudaEventRecord(S.ev[5], st);
   prep_kernel<<<nU, 128, 0, st>>>(c->dm, W);
   c->stats.launches++; c->stats.launchesMisc++;
   if (tr) cudaEventRecord(S.ev[0], st);
   if (sg != st) { CK(cudaEventRecord(S.evIn, st)); CK(cudaStreamWaitEvent(sg, S.evIn, 0)); }
   // ---- K1
   int gk = c->opt.gmmKernel;
   if (gk == 0) gk = (c->useV3 || gmm_tc_available(c->tc)) ? 2 : 1;
   // K4 on the tensor cores (decided here: K1 then also leaves every frame's expanded operand row for it)
   const bool tcStats = !c->upd && c->useV3 && c->opt.gmmKernel != 1 && stats_tc_supported(c->tc3, c->dm.D) && !getenv("HFBGPU_STATS5") &&
                        !getenv("HFBGPU_NO_STATS_PRE") && c->dm.D + 1 <= 40 && !feat2 && !getenv("HFBGPU_STATS3") && c->opt.uFlags != 0;
   // (single-Gaussian sets: their K1 is bound by its load / store queue and the extra row stores cost it 0.3 ms on config #2,
   // more than the statistics kernel gains -- there the padded pre-pass, a streaming kernel, writes the rows)
   const bool expRows = tcStats && (c->tc3.MP > 1 || !getenv("HFBGPU_NO_PAD")) && !getenv("HFBGPU_NO_EXPA");
   if (gk == 2 && c->useV3) {
      int nl = 0;
      // HFBGPU_K1_RESERVE=1 (experiment, see "small waves" above): K1 leaves one SM per 8 utterances of a small wave free
      int smK1 = c->smCount;
      {
         static const bool reserveOn = getenv("HFBGPU_K1_RESERVE") && atoi(getenv("HFBGPU_K1_RESERVE")) != 0;
         const int need = (nU + 7) / 8;
         if (reserveOn && !tm && need * 4 <= c->smCount) smK1 = std::max(2, (c->smCount - need) & ~1);
      }
      if ((rc = gmm_tc3_launch(c->tc3, S.tcw, c->dm, W, waveFrames, (const int2 *)(base + oIt), (int)w.tcItems.size(),
                               (const int2 *)(base + oIt4), (int)w.tcItems4.size(), smK1, sg, &nl, expRows,
                               /* lean (96-register) build */ sg != st && smallFifoOn && smallWave))) return rc;
      c->stats.launches += nl; c->stats.launchesGmm += nl;
   } else if (gk == 2) {
      int nl = 0;
      if ((rc = gmm_tc_launch(c->tc, S.tcw, c->dm, W, waveFrames, (const int2 *)(base + oIt), (int)w.tcItems.size(),
                              (const int2 *)(base + oIt2), (int)w.tcItems2.size(),
                              (const int2 *)(base + oIt4), (int)w.tcItems4.size(),
                              c->smCount, sg, &nl, tm ? S.evX : nullptr))) return rc;
      S.hasX = tm;
      c->stats.launches += nl; c->stats.launchesGmm += nl;
   } else if (w.tiles > 0) {
      size_t smem = sizeof(float) * ((size_t)c->dm.D * GT_FR + (size_t)GT_FR * (GT_SL + 1));
      gmm_fp32_kernel<<<(unsigned)w.tiles, 128, smem, sg>>>(c->dm, W);
      c->stats.launches++; c->stats.launchesGmm++;
   }
   if (sg != st) { CK(cudaEventRecord(S.evGmm, sg)); CK(cudaStreamWaitEvent(st, S.evGmm, 0)); }
   if (tr) cudaEventRecord(S.ev[1], st);
   // ---- K2 / K3
   if (w.maxQ > 0) {
      // experiment (HFBGPU_REC_STREAM=1): the latency-bound recursions of every wave on a high-priority stream of their
      // own, so that their CTAs are placed first and the throughput kernels of the other waves fill in around them.
      // Measured on B200, cfg3: 153.4 -> 155.4 M frames/s from device features, no change end to end; left off.
      // (Round 2, measured in timing mode: alpha_l2r_kernel on a second stream WHILE gmm_tc4_kernel runs -- what a
      // software pipeline "alpha of wave w under the GMM of wave w + 1" would do -- takes 3.99 ms for the pair against
      // 2.03 + 0.77 one after the other: next to the 448-thread GMM CTA only four alpha warps fit an SM and they share
      // its issue slots and its power budget.  Likewise alpha of one wave next to the beta pass of another, beta capped
      // at 6 warps per SM so that both fit: 3.05 ms for the pair against 1.08 + 0.76, with or without a common
      // shared-memory carve-out -- two latency-bound kernels that stream 5 GB between them lengthen each other's
      // dependent loads.  The waves stay on their own streams.)
      cudaStream_t sr = (!tm && S.recStream && getenv("HFBGPU_REC_STREAM")) ? S.recStream : st;
      if (sr != st) { CK(cudaEventRecord(S.evIn, st)); CK(cudaStreamWaitEvent(sr, S.evIn, 0)); }
      int nt = std::min(256, std::max(32, (w.maxQ + 31) & ~31));
      const int ntGeneric = std::min(1024, std::max(32, (w.maxQ + 31) & ~31));   // long transcriptions: still 1 model/thread
      size_t rsm = rec_smem_bytes(w.maxS, w.maxQ), asm_ = alpha_warp_smem_bytes(w.maxS, w.maxQ);
      if (rsm > (size_t)c->maxSmemOptin || asm_ > (size_t)c->maxSmemOptin) {
         g_lastError = "utterance too long for the shared-memory window"; return HFB_EUNSUPPORTED;
      }
      const bool exact = getenv("HFBGPU_EXACT_LADD") != nullptr;
      const bool noFast = getenv("HFBGPU_NO_FAST") != nullptr;        // test hooks: generic kernels only /
      const int forceRedo = getenv("HFBGPU_FORCE_REDO") ? 1 : 0;      // fast alpha gives up immediately
      const bool fastOk = !noFast && w.maxN <= 8;                      // alpha: any Q (32-model sliding window)
      const bool betaFastOk = fastOk && w.maxQ <= 256;                 // beta: one thread per model
      const bool l2r = fastOk && !exact && c->hm.l2r && !getenv("HFBGPU_NO_L2R");   // standard topology
      if (l2r && w.maxQ <= 1024) {
         const size_t fsm = beta_fast_smem_bytes(w.maxQ);
         const bool pruning = c->opt.pruneInit < 0.5 * HFB_NOPRUNE;
         if (w.maxQ <= 32 * BW_NM && !getenv("HFBGPU_NO_BETA_WARP")) {
            // HFBGPU_BETA_ALU: the FP64 <-> FP32 conversions of the log-add as integer operations (experiment)
            if (getenv("HFBGPU_BETA_ALU")) beta_l2r_warp_kernel<true, false><<<nU, 32, 0, sr>>>(c->dm, W);
            else beta_l2r_warp_kernel<false, false><<<nU, 32, 0, sr>>>(c->dm, W);
         }
         else if (pruning && !getenv("HFBGPU_NO_RING")) {
            // long transcriptions under a beam: one warp per utterance, 128-model ring window that slides down with the
            // beam; the one-thread-per-label kernel redoes the utterances whose beam outgrew it
            beta_l2r_warp_kernel<false, true><<<nU, 32, 0, sr>>>(c->dm, W);
            beta_l2r_kernel<1024><<<nU, ntGeneric, fsm, sr>>>(c->dm, W, 1);
            c->stats.launches++; c->stats.launchesBeta++;
         }
         else if (w.maxQ <= 128) beta_l2r_kernel<128><<<nU, nt, fsm, sr>>>(c->dm, W, 0);  // 72 registers, 7 CTAs/SM
         else if (w.maxQ <= 256) beta_l2r_kernel<256><<<nU, nt, fsm, sr>>>(c->dm, W, 0);
         else if (pruning && !getenv("HFBGPU_NO_SLIDE")) {
            // long transcriptions under a beam: 256-model sliding window, the one-thread-per-label kernel redoes overflows
            beta_l2r_slide_kernel<<<nU, 256, fsm, sr>>>(c->dm, W);
            beta_l2r_kernel<1024><<<nU, ntGeneric, fsm, sr>>>(c->dm, W, 1);
            c->stats.launches++; c->stats.launchesBeta++;
         } else beta_l2r_kernel<1024><<<nU, ntGeneric, fsm, sr>>>(c->dm, W, 0);
         c->stats.launchesL2R++;
      } else if (betaFastOk) {
         const size_t fsm = beta_fast_smem_bytes(w.maxQ);
         if (w.maxN <= 5) {
            if (exact) beta_fast_kernel<true, 3><<<nU, nt, fsm, sr>>>(c->dm, W);
            else beta_fast_kernel<false, 3><<<nU, nt, fsm, sr>>>(c->dm, W);
         } else {
            if (exact) beta_fast_kernel<true, 6><<<nU, nt, fsm, sr>>>(c->dm, W);
            else beta_fast_kernel<false, 6><<<nU, nt, fsm, sr>>>(c->dm, W);
         }
      } else if (exact) beta_kernel<true><<<nU, ntGeneric, rsm, sr>>>(c->dm, W);
      else beta_kernel<false><<<nU, ntGeneric, rsm, sr>>>(c->dm, W);
      if (tr) cudaEventRecord(S.ev[2], st);
      if (fastOk) {                                    // register/shuffle kernel; generic one redoes overflows
         if (l2r) { alpha_l2r_kernel<<<nU, 32, 0, sr>>>(c->dm, W, forceRedo); c->stats.launchesL2R++; }
         else if (w.maxN <= 5) {
            if (exact) alpha_fast_kernel<true, 3><<<nU, 32, 0, sr>>>(c->dm, W, forceRedo);
            else alpha_fast_kernel<false, 3><<<nU, 32, 0, sr>>>(c->dm, W, forceRedo);
         } else {
            if (exact) alpha_fast_kernel<true, 6><<<nU, 32, 0, sr>>>(c->dm, W, forceRedo);
            else alpha_fast_kernel<false, 6><<<nU, 32, 0, sr>>>(c->dm, W, forceRedo);
         }
         c->stats.launches++; c->stats.launchesAlpha++;
      }
      if (exact) alpha_warp_kernel<true><<<nU, 32, asm_, sr>>>(c->dm, W, fastOk ? 1 : 0);
      else alpha_warp_kernel<false><<<nU, 32, asm_, sr>>>(c->dm, W, fastOk ? 1 : 0);
      if (sr != st) { CK(cudaEventRecord(S.evGmm, sr)); CK(cudaStreamWaitEvent(st, S.evGmm, 0)); }
      if (tr) cudaEventRecord(S.ev[3], st);
      c->stats.launches += 2; c->stats.launchesBeta++; c->stats.launchesAlpha++;
      // ---- K4
      if (c->upd) {
         // two-model re-estimation: statistics of the update set from this set's alignment (hfb_kernels2.cuh)
         if (w.totalP > 0) {
            stats_two_kernel<<<(unsigned)((w.totalP + ST_WARPS - 1) / ST_WARPS), 32 * ST_WARPS, stats_two_smem_bytes(c->dm.D), st>>>(
               c->dm, c->upd->dm, W, (const int *)(base + oLabUp), (c->opt.flags & HFB_OPT_ALIGN_COMP_LEVEL) ? 1 : 0);
            c->stats.launches++; c->stats.launchesStats++;
         }
      } else if (w.totalP > 0 && c->opt.uFlags != 0) {
         const int Dd = c->dm.D;
         // mixture sets: occupancy-weighted sums on the tensor cores (mma.sync 3xTF32); else the FP32 kernel
         // (single-Gaussian sets: only with the tcgen05 kernel, where the sums are one contraction per tile)
         if ((c->hm.maxM >= 4 || tcStats) && Dd + 1 <= 40 && !W.feat2 && !getenv("HFBGPU_STATS3")) {
            // positions bucketed by tied state (counting sort), then S5_CAP sorted positions per warp
            const int Jm = c->hm.J;
            const size_t nIdx = ((size_t)3 * (Jm + 2) + 3) & ~(size_t)3;
            const bool pre = !getenv("HFBGPU_NO_STATS_PRE");
            // list capacity: 48 entries per frame of the wave (measured need on the bench configurations: 10-25)
            long long vCap = 0;
            if (pre) {
               long long frames = 0;
               for (const auto &u : w.utt) frames += u.T;
               const char *e = getenv("HFBGPU_STATS_PRE_CAP");
               vCap = e ? atoll(e) : 48 * frames;
               if (vCap < 1) vCap = 1;
               if ((rc = S.dValid.reserve((size_t)vCap))) return rc;
            }
            if ((rc = S.dStateIdx.reserve(nIdx + 4 + 2 * (size_t)w.totalP))) return rc;
            int *cnt = S.dStateIdx.p, *off = cnt + (Jm + 2), *fill = off + (Jm + 2);
            unsigned long long *vCursor = (unsigned long long *)(cnt + nIdx);
            int *vcnt = cnt + nIdx + 4, *posIdx = vcnt + w.totalP;
            PosRec *list = (PosRec *)(sc + sl.posList);
            CK(cudaMemsetAsync(cnt, 0, (nIdx + 4) * sizeof(int), st));
            statpos_count_kernel<<<nU, 128, 0, st>>>(W, cnt);
            statpos_scan_kernel<<<1, 1024, 0, st>>>(cnt, off, fill, Jm);
            int *overflow = cnt + nIdx + 2;             // zeroed with the counters above
            statpos_scatter_kernel<<<nU, 128, 0, st>>>(W, off, fill, list, vCursor, vCap, posIdx, overflow);
            if (pre) {
               const int gy = std::max(1, std::min(16, (w.maxQ + SPRE_WARPS - 1) / SPRE_WARPS));
               stats_pre_kernel<<<dim3(nU, gy), 32 * SPRE_WARPS, 0, st>>>(c->dm, W, list, posIdx, S.dValid.p, vcnt);
               c->stats.launches++; c->stats.launchesStats++;
            }
            // positions per warp: more of them = longer runs of one state per warp (fewer flushes, the prefetch pipeline
            // stays primed: 16 -> 64 is 1.75 -> 1.67 ms on config #3), as long as two waves of CTAs remain
            int cap = S5_CAP;
            while (cap < 64 && w.totalP >= (long long)2 * cap * S4_WARPS * 3 * c->smCount * 2) cap *= 2;
            const unsigned nWarps = (unsigned)((w.totalP + cap - 1) / cap);
            const unsigned grid = (nWarps + S4_WARPS - 1) / S4_WARPS;
            // the sums on tcgen05 (hfb_stats_tc.cuh); stats5_kernel then only runs for a wave whose frame lists overflowed
            const int *only = nullptr;
            // (only behind gmm_tc3_kernel: its per-frame flags say which frames are outside the FP16 operand range)
            if (pre && tcStats) {
               stats_tc_launch(c->tc3, c->dm, W, list, off + Jm, S.dValid.p, vcnt, S.tcw.dFlag3, overflow, w.totalP, st,
                               expRows ? S.tcw.dExpA : nullptr);
               c->stats.launches++; c->stats.launchesStats++;
               only = overflow;
            }
            if (Dd + 1 <= 32) stats5_kernel<4><<<grid, 32 * S4_WARPS, S4_WARPS * stats5_warp_bytes<4>(Dd), st>>>(c->dm, W, c->dCentre.p, list, off + Jm, S.dValid.p, vcnt, cap, only);
            else stats5_kernel<5><<<grid, 32 * S4_WARPS, S4_WARPS * stats5_warp_bytes<5>(Dd), st>>>(c->dm, W, c->dCentre.p, list, off + Jm, S.dValid.p, vcnt, cap, only);
            c->stats.launches += 3; c->stats.launchesStats += 3;
         } else
            stats3_kernel<<<(unsigned)((w.totalP + ST_WARPS - 1) / ST_WARPS), 32 * ST_WARPS, stats_smem_bytes(c->dm.D), st>>>(c->dm, W);
         c->stats.launches++; c->stats.launchesStats++;
      }
      if (tr) cudaEventRecord(S.ev[4], st);
   } else if (tm) {
      for (int k = 2; k <= 4; k++) cudaEventRecord(S.ev[k], st);
   }
   CK(cudaGetLastError());

   // ---- results back (asynchronous; consumed by finish_wave)
   if ((size_t)nU > S.hOutCap) {
      if (S.hOut) cudaFreeHost(S.hOut);
      S.hOutCap = (size_t)nU * 2;
      CK(cudaMallocHost(&S.hOut, S.hOutCap * sizeof(UttOut)));
   }
   CK(cudaMemcpyAsync(S.hOut, W.out, (size_t)nU * sizeof(UttOut), cudaMemcpyDeviceToHost, st));
   c->stats.d2hBytes += (int64_t)nU * sizeof(UttOut);
   if (wantBeams) {
      if ((size_t)waveFrames * 4 > S.hBeamsCap) {
         if (S.hBeams) cudaFreeHost(S.hBeams);
         S.hBeamsCap = (size_t)waveFrames * 8;
         CK(cudaMallocHost(&S.hBeams, S.hBeamsCap * sizeof(short)));
      }
      CK(cudaMemcpyAsync(S.hBeams, S.dBeams.p, (size_t)waveFrames * 4 * sizeof(short), cudaMemcpyDeviceToHost, st));
      c->stats.d2hBytes += (int64_t)waveFrames * 8;
   }
   return HFB_OK;
}

// A wave that failed part-way (out of memory, unsupported shape, launch error) must not stay "in flight": whatever it
// enqueued is drained and the slot is released, so that a later wait / finish never reads results that were not produced.
static int launch_wave(hfbgpu_ctx *c, hfbgpu_ctx::Slot &S, const int32_t *labBase, const int32_t *labUpBase, const float *feat, const float *feat2,
                       bool featOnDevice, long long waveFrame0, long long waveFrames, bool wantBeams,
                       const hfb_compressed *cf = nullptr, int cfUtt0 = 0)
{
   const int rc = launch_wave_impl(c, S, labBase, labUpBase, feat, feat2, featOnDevice, waveFrame0, waveFrames, wantBeams, cf, cfUtt0);
   if (rc) {
      cudaStreamSynchronize(S.stream);
      cudaGetLastError();
      S.busy = false; S.res = nullptr; S.hasX = false;
      memset(&S.beams, 0, sizeof(S.beams));
   }
   return rc;
}

static int finish_wave(hfbgpu_ctx *c, hfbgpu_ctx::Slot &S)
{
   if (!S.busy) return HFB_OK;
   S.busy = false;
   if (!S.res || !S.hOut || !S.w || S.w->utt.size() > S.hOutCap) { g_lastError = "internal: wave slot without results"; return HFB_ECUDA; }
   hfb_utt_result *res = S.res;
   const hfb_beams *beams = &S.beams;
   WaveTables &w = *S.w;
   const int nU = (int)w.utt.size();
   CK(cudaStreamSynchronize(S.stream));
   if (c->timing) {
      float ms;
      cudaEventElapsedTime(&ms, S.ev[0], S.ev[1]); c->stats.msGmm += ms;
      if (S.hasX) { cudaEventElapsedTime(&ms, S.ev[0], S.evX); c->stats.msExpand += ms; S.hasX = false; }
      cudaEventElapsedTime(&ms, S.ev[1], S.ev[2]); c->stats.msBeta += ms;
      cudaEventElapsedTime(&ms, S.ev[2], S.ev[3]); c->stats.msAlpha += ms;
      cudaEventElapsedTime(&ms, S.ev[3], S.ev[4]); c->stats.msStats += ms;
   }
   if (c->trace && !c->timing) {
      float t5, t0, t1, t2, t3, t4;
      cudaEventElapsedTime(&t5, c->evRef, S.ev[5]); cudaEventElapsedTime(&t0, c->evRef, S.ev[0]);
      cudaEventElapsedTime(&t1, c->evRef, S.ev[1]); cudaEventElapsedTime(&t2, c->evRef, S.ev[2]);
      cudaEventElapsedTime(&t3, c->evRef, S.ev[3]); cudaEventElapsedTime(&t4, c->evRef, S.ev[4]);
      fprintf(stderr, "[hfbgpu trace] slot %d utts %4d | gmm stream reached %9.3f prep done %9.3f | gmm done %9.3f beta done %9.3f "
                      "alpha done %9.3f stats done %9.3f ms\n", (int)(&S - c->slot), nU, t5, t0, t1, t2, t3, t4);
   }
   const long long waveFrames = S.waveFrames;
   for (int k = 0; k < nU; k++) {
      hfb_utt_result &r = res[w.uttIndex[k]];
      const UttOut &o = S.hOut[k];
      r.status = o.status; r.retries = o.retries; r.pr = o.pr; r.pruneThresh = o.thresh;
      const UttDesc &u = w.utt[k];
      if (o.status == 0) c->stats.gmmPairs += (int64_t)o.pairs;
      if (S.wantBeams) {
         const long long dst = S.waveFrame0 + u.frameBase;                         // frame index in the batch
         const short *lo = S.hBeams + u.frameBase, *hi = lo + waveFrames, *s = hi + waveFrames, *e = s + waveFrames;
         for (int t = 0; t < u.T; t++) {
            bool okb = (o.status == 0 || o.status == HFB_UTT_EALPHA);
            if (beams->qLo) beams->qLo[dst + t] = okb ? (int16_t)(lo[t] + 1) : 0;
            if (beams->qHi) beams->qHi[dst + t] = okb ? (int16_t)(hi[t] + 1) : 0;
            if (beams->sq) beams->sq[dst + t] = (o.status == 0) ? (int16_t)(s[t] + 1) : 0;
            if (beams->eq) beams->eq[dst + t] = (o.status == 0) ? (int16_t)(e[t] + 1) : 0;
            if (o.status == 0) { c->stats.betaCells += hi[t] - lo[t] + 1; c->stats.alphaCells += e[t] - s[t] + 1; }
         }
      }
   }
   return HFB_OK;
}

// Enqueues the batch (one or more waves) without waiting for it.  `splitWaves` = how many
// waves to aim at (the blocking call splits a batch over the streams; the asynchronous one keeps
// a batch in one wave so that consecutive calls overlap instead).
static int submit_impl(hfbgpu_ctx *c, const hfb_batch *b, hfb_utt_result *res, const hfb_beams *beams,
                       bool featOnDevice, int splitWaves, const float *feat2 = nullptr, const hfb_compressed *cf = nullptr)
{
   if (!c || !b || !res) return HFB_EINVAL;
   if (b->numUtt < 0 || (b->numUtt > 0 && (!b->frameOff || (!b->feat && !cf) || !b->labOff || !b->lab))) return HFB_EINVAL;
   if (cf && (featOnDevice || feat2 || (b->numUtt > 0 && (!cf->feat || !cf->scaleA || !cf->scaleB)))) return HFB_EINVAL;
   CK(cudaSetDevice(c->device));
   CK(cudaStreamSynchronize(c->stream));               // accumulator zeroing / model uploads are done
   const HostModel &h = c->hm;
   // two-model re-estimation: `lab` indexes the update set, `labAlign` this context's (alignment) set
   if (c->upd && b->numUtt > 0 && !b->labAlign) { g_lastError = "two-model re-estimation: hfb_batch.labAlign is missing"; return HFB_EINVAL; }
   const int32_t *labA = c->upd ? b->labAlign : b->lab;
   if (c->upd && b->numUtt > 0) {
      const HostModel &hu = c->upd->hm;
      for (int i = b->labOff[0]; i < b->labOff[b->numUtt]; i++) {
         const int pa = labA[i], pu = b->lab[i];
         if (pu < 0 || pu >= hu.P) { g_lastError = "two-model re-estimation: update-set label out of range (HError 2321)"; return HFB_EINVAL; }
         if (pa >= 0 && pa < h.P && h.hmmN[pa] != hu.hmmN[pu]) {
            g_lastError = "Num states differ in align and update models (HError 999, HFB.c:549-551)"; return HFB_EINVAL;
         }
         if ((c->opt.flags & HFB_OPT_ALIGN_COMP_LEVEL) && pa >= 0 && pa < h.P)
            for (int j = 0; j < h.hmmN[pa] - 2; j++) {
               const int sa = h.hmmState[h.hmmStateOff[pa] + j], su = hu.hmmState[hu.hmmStateOff[pu] + j];
               if (h.stateMixOff[sa + 1] - h.stateMixOff[sa] != hu.stateMixOff[su + 1] - hu.stateMixOff[su]) {
                  g_lastError = "Cannot align at the component level if number of components is different (HError 999, HFB.c:1523-1524)";
                  return HFB_EINVAL;
               }
            }
      }
   }
   // validated for the WHOLE batch before anything is launched: a rejected batch leaves the accumulators untouched
   for (int u = 0; u < b->numUtt; u++) {
      const long long T = b->frameOff[u + 1] - b->frameOff[u];
      const long long Q = (long long)b->labOff[u + 1] - b->labOff[u];
      if (T < 0 || Q < 0) return HFB_EINVAL;
      if (T > HFB_MAX_FRAMES || Q > HFB_MAX_LABELS) {
         g_lastError = "utterance beyond the int16 beam range (more than 32767 frames or 32766 labels): hand it to the reference's FBFile";
         return HFB_EUNSUPPORTED;
      }
   }
   const bool wantBeams = beams && (beams->qLo || beams->qHi || beams->sq || beams->eq);
   // timing mode serialises the waves so that the per-kernel events do not overlap
   const int ns = c->timing ? 1 : c->numSlots;
   const int want = std::max(1, std::min(ns, splitWaves));
   const long long totalFrames = b->numUtt ? b->frameOff[b->numUtt] - b->frameOff[0] : 0;
   const long long targetFrames = (b->numUtt >= 64 * want) ? (totalFrames + want - 1) / want : totalFrames;
   const size_t wsPerSlot = c->workspaceBytes / (size_t)ns;
   const int maxWaveUtts = c->timing ? 16384 : std::min(16384, c->waveUtts);
   int u0 = 0, rcAll = HFB_OK;
   while (u0 < b->numUtt) {
      hfbgpu_ctx::Slot &S = c->slot[c->nextSlot % ns];
      int rc = finish_wave(c, S);
      if (rc) { rcAll = rc; break; }
      if (!S.w) S.w = new WaveTables();
      WaveTables &w = *S.w;
      w.clear();
      w.lab0 = b->labOff[u0];
      const long long waveFrame0 = b->frameOff[u0];
      size_t bytes = 0;
      int u1 = u0;
      while (u1 < b->numUtt) {
         long long f0 = b->frameOff[u1], f1 = b->frameOff[u1 + 1];
         int T = (int)(f1 - f0), Q = b->labOff[u1 + 1] - b->labOff[u1];
         size_t perFrame = 0;
         for (int q = 0; q < Q; q++) {
            int p = labA[b->labOff[u1] + q];
            int N = (p >= 0 && p < h.P) ? h.hmmN[p] : 2;
            perFrame += (size_t)N * 8 + (size_t)(N - 2) * 12 + 8;
         }
         size_t need = (size_t)T * perFrame;
         if (u1 > u0 && (bytes + need > wsPerSlot || u1 - u0 >= maxWaveUtts || f0 - waveFrame0 >= targetFrames)) break;
         bytes += add_utterance(h, w, u1, T, labA + b->labOff[u1], Q, b->labOff[u1] - w.lab0, f0 - waveFrame0,
                                (c->useV3 && c->opt.gmmKernel != 1) ? c->tc3.globalSlots : 0);
         u1++;
      }
      if (rcAll) break;
      const long long waveFrames = b->frameOff[u1] - waveFrame0;
      S.res = res; S.ticket = c->submitSeq;
      if (wantBeams) S.beams = *beams; else memset(&S.beams, 0, sizeof(S.beams));
      rc = launch_wave(c, S, labA + w.lab0, c->upd ? b->lab + w.lab0 : nullptr, b->feat, feat2, featOnDevice, waveFrame0, waveFrames, wantBeams, cf, u0);
      if (rc) { rcAll = rc; break; }
      u0 = u1; c->nextSlot++;
      if (c->timing) { rc = finish_wave(c, S); if (rc) { rcAll = rc; break; } }
   }
   return rcAll;
}

static int wait_impl(hfbgpu_ctx *c)
{
   if (!c) return HFB_EINVAL;
   CK(cudaSetDevice(c->device));
   int rcAll = HFB_OK;
   for (int i = 0; i < hfbgpu_ctx::NSLOT; i++) {        // oldest first
      int rc = finish_wave(c, c->slot[(c->nextSlot + i) % hfbgpu_ctx::NSLOT]);
      if (rc && !rcAll) rcAll = rc;
   }
   return rcAll;
}


// ------------------------------------------------------------------------------------------
// Device groups: one host thread, one context, N GPUs (SURVEY 8b "device list", 8e).
// hfbgpu_create_multi builds one child context per device; every batch is cut into N contiguous utterance ranges of
// equal sum(T * Q) and submitted to the children back to back (their kernels run concurrently, each on its own GPU);
// the per-GPU FP64 accumulators are combined ON THE DEVICE by a kernel on child 0 that reads the peers' buffers over
// NVLink (peer access; a staged cudaMemcpyPeer where that is unavailable) -- the in-library equivalent of the one
// all-reduce per pass, and of what `HERest -p 0` does with the per-process dumps (HTrain.c:1626-1687).
// ------------------------------------------------------------------------------------------
#define HFB_MAX_GROUP 16
struct PeerPtrs { const double *p[HFB_MAX_GROUP]; };

__global__ void acc_sum_peers_kernel(double *__restrict__ dst, PeerPtrs src, int n, long long count)
{
   for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
      double a = dst[i];
      for (int k = 0; k < n; k++) a += src.p[k][i];
      dst[i] = a;
   }
}

static bool is_group(const hfbgpu_ctx *c) { return c && !c->kids.empty(); }

static int group_reduce(hfbgpu_ctx *g)
{
   int rcAll = HFB_OK;
   for (auto *k : g->kids) { int rc = wait_impl(k); if (rc && !rcAll) rcAll = rc; }
   if (rcAll) return rcAll;
   if (g->reduced || g->kids.size() < 2) { g->reduced = true; return HFB_OK; }
   hfbgpu_ctx *k0 = g->kids[0];
   const long long count = k0->L.count;
   for (auto *k : g->kids) { CK(cudaSetDevice(k->device)); CK(cudaStreamSynchronize(k->stream)); }
   CK(cudaSetDevice(k0->device));
   PeerPtrs pp;
   int n = 0;
   std::vector<double *> staged;
   for (size_t i = 1; i < g->kids.size(); i++) {
      hfbgpu_ctx *k = g->kids[i];
      int can = 0;
      cudaDeviceCanAccessPeer(&can, k0->device, k->device);
      if (can) {
         cudaError_t e = cudaDeviceEnablePeerAccess(k->device, 0);
         if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
         cudaGetLastError();
      }
      if (can) pp.p[n++] = accp(k);
      else {                                            // no NVLink / PCIe peer path: stage a copy on device 0
         double *tmp = nullptr;
         if (cudaMalloc(&tmp, (size_t)count * sizeof(double)) != cudaSuccess) { cudaGetLastError(); for (auto *t : staged) cudaFree(t); return HFB_ENOMEM; }
         staged.push_back(tmp);
         CK(cudaMemcpyPeerAsync(tmp, k0->device, accp(k), k->device, (size_t)count * sizeof(double), k0->stream));
         pp.p[n++] = tmp;
      }
   }
   const int grid = std::max(1, std::min(k0->smCount * 8, (int)((count + 255) / 256)));
   acc_sum_peers_kernel<<<grid, 256, 0, k0->stream>>>(accp(k0), pp, n, count);
   k0->stats.launches++; k0->stats.launchesMisc++;
   CK(cudaGetLastError());
   CK(cudaStreamSynchronize(k0->stream));
   for (auto *t : staged) cudaFree(t);
   for (size_t i = 1; i < g->kids.size(); i++) {        // their content now lives in child 0
      hfbgpu_ctx *k = g->kids[i];
      CK(cudaSetDevice(k->device));
      CK(cudaMemsetAsync(accp(k), 0, (size_t)count * sizeof(double), k->stream));
      CK(cudaStreamSynchronize(k->stream));
   }
   g->reduced = true;
   return HFB_OK;
}

// contiguous utterance ranges of (nearly) equal sum(T * Q)
static void group_split(const hfb_batch *b, int n, std::vector<int> &cut)
{
   cut.assign(n + 1, 0);
   std::vector<double> pre((size_t)b->numUtt + 1, 0.0);
   for (int u = 0; u < b->numUtt; u++)
      pre[u + 1] = pre[u] + (double)(b->frameOff[u + 1] - b->frameOff[u]) * (double)std::max(1, b->labOff[u + 1] - b->labOff[u]);
   int u = 0;
   for (int k = 1; k < n; k++) {
      const double want = pre[b->numUtt] * k / n;
      while (u < b->numUtt && pre[u + 1] <= want) u++;
      cut[k] = u;
   }
   cut[n] = b->numUtt;
}

static int group_submit(hfbgpu_ctx *g, const hfb_batch *b, const float *feat2, hfb_utt_result *res, const hfb_beams *beams, int mode,
                        const hfb_compressed *cf = nullptr)
{
   // mode 0 = hfbgpu_submit (asynchronous), 1 = hfbgpu_accumulate (blocking), 2 = hfbgpu_accumulate_retrain
   if (!b || !res) return HFB_EINVAL;
   if (b->numUtt < 0 || (b->numUtt > 0 && (!b->frameOff || (!b->feat && !cf) || !b->labOff || !b->lab))) return HFB_EINVAL;
   const int n = (int)g->kids.size();
   std::vector<int> cut;
   group_split(b, n, cut);
   g->submitSeq++;
   g->reduced = false;
   std::vector<int64_t> tk((size_t)n, 0);
   int rcAll = HFB_OK;
   for (int k = 0; k < n; k++) {
      hfbgpu_ctx *kid = g->kids[k];
      tk[k] = kid->submitSeq;
      if (cut[k + 1] == cut[k]) continue;
      hfb_batch sub = *b;                               // offsets stay absolute: the children index feat / lab with them
      sub.numUtt = cut[k + 1] - cut[k]; sub.frameOff = b->frameOff + cut[k]; sub.labOff = b->labOff + cut[k];
      kid->submitSeq++;
      hfb_compressed subc;
      if (cf) {                                         // A, B are indexed by utterance: move them with the range
         const size_t cols = (size_t)(kid->qual.enabled ? kid->qual.numStatic : kid->hm.D);
         subc = *cf; subc.scaleA = cf->scaleA + (size_t)cut[k] * cols; subc.scaleB = cf->scaleB + (size_t)cut[k] * cols;
      }
      int rc = submit_impl(kid, &sub, res + cut[k], beams, false, mode == 0 ? 1 : (int)hfbgpu_ctx::NSLOT, feat2, cf ? &subc : nullptr);
      tk[k] = kid->submitSeq;
      if (rc && !rcAll) rcAll = rc;
   }
   if (g->kidTickets.size() < 64) g->kidTickets.resize(64);
   g->kidTickets[(size_t)(g->submitSeq % 64)] = tk;
   if (mode != 0)
      for (auto *kid : g->kids) { int rc = wait_impl(kid); if (rc && !rcAll) rcAll = rc; }
   return rcAll;
}

extern "C" int hfbgpu_create_multi(hfbgpu_ctx **out, const hfb_model *m, const hfb_options *opt, const int32_t *devices, int32_t numDevices)
{
   if (!out || !m || !opt || !devices || numDevices < 1 || numDevices > HFB_MAX_GROUP) return HFB_EINVAL;
   *out = nullptr;
   for (int i = 0; i < numDevices; i++)
      for (int j = 0; j < i; j++) if (devices[i] == devices[j]) { g_lastError = "device listed twice"; return HFB_EINVAL; }
   hfbgpu_ctx *g = new hfbgpu_ctx();
   g->opt = *opt;
   memset(&g->stats, 0, sizeof(g->stats));
   for (int i = 0; i < numDevices; i++) {
      hfb_options o = *opt;
      o.device = devices[i];
      hfbgpu_ctx *kid = nullptr;
      int rc = hfbgpu_create(&kid, m, &o);
      if (rc) { for (auto *k : g->kids) hfbgpu_destroy(k); delete g; return rc; }
      g->kids.push_back(kid);
   }
   g->device = devices[0];
   g->L = g->kids[0]->L;
   *out = g;
   return HFB_OK;
}

extern "C" int hfbgpu_num_devices(hfbgpu_ctx *c) { return !c ? 0 : (is_group(c) ? (int)c->kids.size() : 1); }

// Sums the per-device accumulators into the first device's buffer (no-op for a single-device context).
extern "C" int hfbgpu_reduce_accs(hfbgpu_ctx *c)
{
   if (!c) return HFB_EINVAL;
   return is_group(c) ? group_reduce(c) : wait_impl(c);
}

extern "C" int hfbgpu_accumulate(hfbgpu_ctx *c, const hfb_batch *b, hfb_utt_result *res, const hfb_beams *beams)
{
   if (is_group(c)) return group_submit(c, b, nullptr, res, beams, 1);
   int rc = submit_impl(c, b, res, beams, false, hfbgpu_ctx::NSLOT);
   int rc2 = wait_impl(c);
   return rc ? rc : rc2;
}

extern "C" int hfbgpu_accumulate_device(hfbgpu_ctx *c, const hfb_batch *b, hfb_utt_result *res, const hfb_beams *beams)
{
   if (is_group(c)) { g_lastError = "device-resident features belong to one GPU: use host features with a device group"; return HFB_EUNSUPPORTED; }
   int rc = submit_impl(c, b, res, beams, true, hfbgpu_ctx::NSLOT);
   int rc2 = wait_impl(c);
   return rc ? rc : rc2;
}

// HERest -r: state / component occupancies from batch->feat, mean and variance statistics from feat2
extern "C" int hfbgpu_accumulate_retrain(hfbgpu_ctx *c, const hfb_batch *b, const float *feat2, hfb_utt_result *res,
                                         const hfb_beams *beams, int featOnDevice)
{
   if (!feat2) return HFB_EINVAL;
   if (is_group(c)) return featOnDevice ? HFB_EUNSUPPORTED : group_submit(c, b, feat2, res, beams, 2);
   int rc = submit_impl(c, b, res, beams, featOnDevice != 0, hfbgpu_ctx::NSLOT, feat2);
   int rc2 = wait_impl(c);
   return rc ? rc : rc2;
}

extern "C" int hfbgpu_submit(hfbgpu_ctx *c, const hfb_batch *b, hfb_utt_result *res, const hfb_beams *beams, int featOnDevice)
{
   if (is_group(c)) return featOnDevice ? HFB_EUNSUPPORTED : group_submit(c, b, nullptr, res, beams, 0);
   if (c) c->submitSeq++;
   return submit_impl(c, b, res, beams, featOnDevice != 0, 1);
}

// `_C` compressed parameter files: the batch's features are the files' 16-bit integers + the vectors A, B of every file
extern "C" int hfbgpu_submit_compressed(hfbgpu_ctx *c, const hfb_batch *b, const hfb_compressed *cf, hfb_utt_result *res, const hfb_beams *beams)
{
   if (!cf) return HFB_EINVAL;
   if (is_group(c)) return group_submit(c, b, nullptr, res, beams, 0, cf);
   if (c) c->submitSeq++;
   return submit_impl(c, b, res, beams, false, 1, nullptr, cf);
}

extern "C" int hfbgpu_accumulate_compressed(hfbgpu_ctx *c, const hfb_batch *b, const hfb_compressed *cf, hfb_utt_result *res, const hfb_beams *beams)
{
   if (!cf) return HFB_EINVAL;
   if (is_group(c)) return group_submit(c, b, nullptr, res, beams, 1, cf);
   int rc = submit_impl(c, b, res, beams, false, hfbgpu_ctx::NSLOT, nullptr, cf);
   int rc2 = wait_impl(c);
   return rc ? rc : rc2;
}

extern "C" int64_t hfbgpu_last_ticket(hfbgpu_ctx *c) { return c ? c->submitSeq : 0; }

// Completes the batches submitted up to and including `ticket` and leaves younger ones in flight.
extern "C" int hfbgpu_wait_ticket(hfbgpu_ctx *c, int64_t ticket)
{
   if (!c) return HFB_EINVAL;
   if (is_group(c)) {
      if (ticket <= c->submitSeq - 64 || ticket > c->submitSeq) return hfbgpu_wait(c);
      const std::vector<int64_t> &tk = c->kidTickets[(size_t)(ticket % 64)];
      int rcAll = HFB_OK;
      for (size_t k = 0; k < c->kids.size(); k++) { int rc = hfbgpu_wait_ticket(c->kids[k], tk[k]); if (rc && !rcAll) rcAll = rc; }
      return rcAll;
   }
   CK(cudaSetDevice(c->device));
   int rcAll = HFB_OK;
   for (int i = 0; i < hfbgpu_ctx::NSLOT; i++) {        // oldest first
      hfbgpu_ctx::Slot &S = c->slot[(c->nextSlot + i) % hfbgpu_ctx::NSLOT];
      if (!S.busy || S.ticket > ticket) continue;
      int rc = finish_wave(c, S);
      if (rc && !rcAll) rcAll = rc;
   }
   return rcAll;
}

extern "C" int hfbgpu_wait(hfbgpu_ctx *c)
{
   if (is_group(c)) {
      int rcAll = HFB_OK;
      for (auto *k : c->kids) { int rc = wait_impl(k); if (rc && !rcAll) rcAll = rc; }
      return rcAll;
   }
   return wait_impl(c);
}

// ------------------------------------------------------------------------------------------
// M-step on the device
// ------------------------------------------------------------------------------------------
extern "C" int hfbgpu_mstep(hfbgpu_ctx *c, const hfb_mstep_options *opt, hfb_mstep_result *out)
{
   if (!c || !opt || !out || !out->mean || !out->var || !out->gConst || !out->mixWeight || !out->transP || !opt->varFloor)
      return HFB_EINVAL;
   if (is_group(c)) { int rc = group_reduce(c); return rc ? rc : hfbgpu_mstep(c->kids[0], opt, out); }
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }
   if (c->upd) return hfbgpu_mstep(c->upd, opt, out);   // two-model re-estimation: the update set is what gets re-estimated
   const HostModel &h = c->hm;
   const int D = h.D, G = h.G, J = h.J, nT = h.numTrans, nM = h.numMeanAcc, nV = h.numVarAcc;
   const int sumM = h.stateMixOff[J];
   const size_t sumNN = h.transLogA.size();
   cudaStream_t st = c->stream;
   CK(cudaStreamSynchronize(st));
   // host-side index tables: use count of every variance vector, first Gaussian that uses it
   std::vector<int> varUse(nV, 0), first(G, 0), seen(nV, 0);
   for (int g = 0; g < G; g++) { varUse[h.varId[g]]++; if (!seen[h.varId[g]]) { seen[h.varId[g]] = 1; first[g] = 1; } }
   int maxN = 1;
   for (int n : h.transN) maxN = std::max(maxN, n);
   DevBuf<int> dInts;        // transOn | stateOn | meanOn | varOn | counters | varUse | first | transN
   DevBuf<float> dFl;        // vFloor | mean | var | gConst | mixWeight | transP
   const size_t nFlags = (size_t)nT + J + nM + nV + 4;
   int rc;
   if ((rc = dInts.reserve(nFlags + nV + G + nT)) ||
       (rc = dFl.reserve((size_t)D + 2 * (size_t)G * D + G + sumM + sumNN))) { dInts.release(); dFl.release(); return rc; }
   int *transOn = dInts.p, *stateOn = transOn + nT, *meanOn = stateOn + J, *varOn = meanOn + nM, *counters = varOn + nV;
   int *dVarUse = counters + 4, *dFirst = dVarUse + nV, *dTransN = dFirst + G;
   float *vFloor = dFl.p, *oMean = vFloor + D, *oVar = oMean + (size_t)G * D, *oGc = oVar + (size_t)G * D,
         *oW = oGc + G, *oT = oW + sumM;
   cudaError_t e = cudaSuccess;
   auto chk = [&](cudaError_t x) { if (e == cudaSuccess) e = x; };
   chk(cudaMemsetAsync(dInts.p, 0, nFlags * sizeof(int), st));
   chk(cudaMemcpyAsync(dVarUse, varUse.data(), (size_t)nV * sizeof(int), cudaMemcpyHostToDevice, st));
   chk(cudaMemcpyAsync(dFirst, first.data(), (size_t)G * sizeof(int), cudaMemcpyHostToDevice, st));
   chk(cudaMemcpyAsync(dTransN, h.transN.data(), (size_t)nT * sizeof(int), cudaMemcpyHostToDevice, st));
   chk(cudaMemcpyAsync(vFloor, opt->varFloor, (size_t)D * sizeof(float), cudaMemcpyHostToDevice, st));
   MStepDev S;
   S.acc = c->dAcc.p; S.minEgs = opt->minEgs; S.uFlags = c->opt.uFlags; S.maxM = h.maxM; S.mixWeightFloor = opt->mixWeightFloor;
   S.vFloor = vFloor; S.transOn = transOn; S.stateOn = stateOn; S.meanOn = meanOn; S.varOn = varOn;
   S.varUse = dVarUse; S.gaussFirstOfVar = dFirst; S.counters = counters;
   S.mean = oMean; S.var = oVar; S.gConst = oGc; S.mixWeight = oW; S.transP = oT;
   mstep_enable_kernel<<<(h.P + 127) / 128, 128, 0, st>>>(c->dm, S);
   mstep_trans_kernel<<<nT, ((maxN + 31) / 32) * 32, 0, st>>>(c->dm, S, dTransN);
   mstep_state_kernel<<<(J + 127) / 128, 128, 0, st>>>(c->dm, S);
   mstep_gauss_kernel<<<(G + 127) / 128, 128, 0, st>>>(c->dm, S);
   c->stats.launches += 4; c->stats.launchesMisc += 4;
   chk(cudaGetLastError());
   int hc[4] = {0, 0, 0, 0};
   chk(cudaMemcpyAsync(out->mean, oMean, (size_t)G * D * sizeof(float), cudaMemcpyDeviceToHost, st));
   chk(cudaMemcpyAsync(out->var, oVar, (size_t)G * D * sizeof(float), cudaMemcpyDeviceToHost, st));
   chk(cudaMemcpyAsync(out->gConst, oGc, (size_t)G * sizeof(float), cudaMemcpyDeviceToHost, st));
   chk(cudaMemcpyAsync(out->mixWeight, oW, (size_t)sumM * sizeof(float), cudaMemcpyDeviceToHost, st));
   chk(cudaMemcpyAsync(out->transP, oT, sumNN * sizeof(float), cudaMemcpyDeviceToHost, st));
   chk(cudaMemcpyAsync(hc, counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
   chk(cudaStreamSynchronize(st));
   dInts.release(); dFl.release();
   if (e != cudaSuccess) { g_lastError = std::string("hfbgpu_mstep: ") + cudaGetErrorString(e); return HFB_ECUDA; }
   out->nFloorVar = hc[0]; out->nFloorVarMix = hc[1]; out->nCopied = hc[2]; out->nNoOcc = hc[3];
   return HFB_OK;
}

// ------------------------------------------------------------------------------------------
// Parameter-kind qualifiers (hfb_feat.cuh)
// ------------------------------------------------------------------------------------------
extern "C" int hfbgpu_set_qualifiers(hfbgpu_ctx *c, const hfb_qualifiers *q)
{
   if (!c) return HFB_EINVAL;
   if (is_group(c)) {
      int rcAll = HFB_OK;
      for (auto *k : c->kids) { int rc = hfbgpu_set_qualifiers(k, q); if (rc && !rcAll) rcAll = rc; }
      return rcAll;
   }
   { int rc = wait_impl(c); if (rc) return rc; }
   if (!q) { c->qual = FeatQual(); return HFB_OK; }
   const int orders = 1 + (q->delWin > 0) + (q->accWin > 0) + (q->thirdWin > 0);
   if (q->numStatic < 1 || q->delWin < 0 || q->accWin < 0 || q->thirdWin < 0 || q->delWin > 64 || q->accWin > 64 ||
       q->thirdWin > 64 || (q->accWin > 0 && q->delWin == 0) || (q->thirdWin > 0 && q->accWin == 0) ||
       q->zeroMeanCols < 0 || q->zeroMeanCols > q->numStatic || (q->suppressEnergy && (q->delWin == 0 || q->numStatic < 2)) ||
       (q->suppressEnergy && q->zeroMeanCols > q->numStatic - 1)) {
      g_lastError = "inconsistent qualifier description"; return HFB_EINVAL;
   }
   if (q->numStatic * orders - (q->suppressEnergy ? 1 : 0) != c->hm.D) {
      g_lastError = "qualifiers do not expand to the model's vector size"; return HFB_EINVAL;
   }
   FeatQual f;
   f.numStatic = q->numStatic; f.win[0] = q->delWin; f.win[1] = q->accWin; f.win[2] = q->thirdWin;
   f.simpleDiffs = q->simpleDiffs != 0; f.zeroMeanCols = q->zeroMeanCols; f.suppressEnergy = q->suppressEnergy != 0; f.enabled = 1;
   c->qual = f;
   return HFB_OK;
}

extern "C" int hfbgpu_expand_features(hfbgpu_ctx *c, const float *src, const int64_t *frameOff, int32_t numUtt, float *dst)
{
   if (!c || !src || !frameOff || !dst || numUtt < 0) return HFB_EINVAL;
   if (is_group(c)) return hfbgpu_expand_features(c->kids[0], src, frameOff, numUtt, dst);
   if (!c->qual.enabled) { g_lastError = "hfbgpu_set_qualifiers has not been called"; return HFB_EINVAL; }
   if (numUtt == 0) return HFB_OK;
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }
   const int D = c->hm.D, ns = c->qual.numStatic;
   std::vector<UttDesc> utt((size_t)numUtt);
   int maxT = 0;
   for (int u = 0; u < numUtt; u++) {
      const long long T = frameOff[u + 1] - frameOff[u];
      if (T < 1 || T > 0x7fffffff) return HFB_EINVAL;
      memset(&utt[u], 0, sizeof(UttDesc));
      utt[u].T = (int)T; utt[u].featOff = frameOff[u] - frameOff[0];
      maxT = std::max(maxT, (int)T);
   }
   const size_t frames = (size_t)(frameOff[numUtt] - frameOff[0]);
   DevBuf<float> dSrc, dDst;
   DevBuf<UttDesc> dUtt;
   int rc;
   if ((rc = dSrc.reserve(frames * ns)) || (rc = dDst.reserve(frames * D)) || (rc = dUtt.reserve((size_t)numUtt))) {
      dSrc.release(); dDst.release(); dUtt.release(); return rc;
   }
   cudaStream_t st = c->stream;
   cudaError_t e = cudaMemcpyAsync(dSrc.p, src + (size_t)frameOff[0] * ns, frames * ns * sizeof(float), cudaMemcpyHostToDevice, st);
   if (e == cudaSuccess) e = cudaMemcpyAsync(dUtt.p, utt.data(), utt.size() * sizeof(UttDesc), cudaMemcpyHostToDevice, st);
   int nl = 0;
   if (e == cudaSuccess) { feat_expand_launch(c->qual, dUtt.p, numUtt, maxT, dSrc.p, dDst.p, D, st, &nl); e = cudaGetLastError(); }
   if (e == cudaSuccess) e = cudaMemcpyAsync(dst + (size_t)frameOff[0] * D, dDst.p, frames * D * sizeof(float), cudaMemcpyDeviceToHost, st);
   if (e == cudaSuccess) e = cudaStreamSynchronize(st);
   c->stats.launches += nl; c->stats.launchesMisc += nl;
   dSrc.release(); dDst.release(); dUtt.release();
   if (e != cudaSuccess) { g_lastError = cudaGetErrorString(e); cudaGetLastError(); return HFB_ECUDA; }
   return HFB_OK;
}

// The decompression alone: what ReadAsTable yields for `_C` files (HParm.c:3489-3494)
extern "C" int hfbgpu_decompress_features(hfbgpu_ctx *c, const hfb_compressed *cf, const int64_t *frameOff, int32_t numUtt, int32_t cols, float *dst)
{
   if (!c || !cf || !cf->feat || !cf->scaleA || !cf->scaleB || !frameOff || !dst || numUtt < 0 || cols < 1) return HFB_EINVAL;
   if (is_group(c)) return hfbgpu_decompress_features(c->kids[0], cf, frameOff, numUtt, cols, dst);
   if (numUtt == 0) return HFB_OK;
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }
   std::vector<UttDesc> utt((size_t)numUtt);
   int maxT = 0;
   for (int u = 0; u < numUtt; u++) {
      const long long T = frameOff[u + 1] - frameOff[u];
      if (T < 0 || T * cols > 0x7fffffffLL) return HFB_EINVAL;
      memset(&utt[u], 0, sizeof(UttDesc));
      utt[u].T = (int)T; utt[u].featOff = frameOff[u] - frameOff[0];
      maxT = std::max(maxT, (int)T);
   }
   const size_t frames = (size_t)(frameOff[numUtt] - frameOff[0]);
   DevBuf<short> dSrc;
   DevBuf<float> dDst, dAB;
   DevBuf<UttDesc> dUtt;
   int rc;
   if ((rc = dSrc.reserve(frames * cols + 8)) || (rc = dDst.reserve(frames * cols + 4)) || (rc = dAB.reserve(2 * (size_t)numUtt * cols)) ||
       (rc = dUtt.reserve((size_t)numUtt))) {
      dSrc.release(); dDst.release(); dAB.release(); dUtt.release(); return rc;
   }
   cudaStream_t st = c->stream;
   const size_t nab = (size_t)numUtt * cols;
   cudaError_t e = cudaMemcpyAsync(dSrc.p, cf->feat + (size_t)frameOff[0] * cols, frames * cols * sizeof(int16_t), cudaMemcpyHostToDevice, st);
   if (e == cudaSuccess) e = cudaMemcpyAsync(dAB.p, cf->scaleA, nab * sizeof(float), cudaMemcpyHostToDevice, st);
   if (e == cudaSuccess) e = cudaMemcpyAsync(dAB.p + nab, cf->scaleB, nab * sizeof(float), cudaMemcpyHostToDevice, st);
   if (e == cudaSuccess) e = cudaMemcpyAsync(dUtt.p, utt.data(), utt.size() * sizeof(UttDesc), cudaMemcpyHostToDevice, st);
   if (e == cudaSuccess) { feat_decompress_launch(dUtt.p, numUtt, maxT, dSrc.p, dAB.p, dAB.p + nab, dDst.p, cols, st); e = cudaGetLastError(); }
   if (e == cudaSuccess) e = cudaMemcpyAsync(dst + (size_t)frameOff[0] * cols, dDst.p, frames * cols * sizeof(float), cudaMemcpyDeviceToHost, st);
   if (e == cudaSuccess) e = cudaStreamSynchronize(st);
   c->stats.launches++; c->stats.launchesMisc++;
   dSrc.release(); dDst.release(); dAB.release(); dUtt.release();
   if (e != cudaSuccess) { g_lastError = cudaGetErrorString(e); cudaGetLastError(); return HFB_ECUDA; }
   return HFB_OK;
}

extern "C" void *hfbgpu_host_alloc(size_t bytes)
{
   void *p = nullptr;
   if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { cudaGetLastError(); return nullptr; }
   return p;
}
extern "C" void hfbgpu_host_free(void *p) { if (p) cudaFreeHost(p); }

// ------------------------------------------------------------------------------------------
// OutP alone
// ------------------------------------------------------------------------------------------
extern "C" int hfbgpu_state_loglik(hfbgpu_ctx *c, const float *feat, int32_t T, const int32_t *states, int32_t n,
                                   float *out, float *mixOut)
{
   if (!c || !feat || !states || !out || T < 1 || n < 1) return HFB_EINVAL;
   if (is_group(c)) return hfbgpu_state_loglik(c->kids[0], feat, T, states, n, out, mixOut);
   if (mixOut) { g_lastError = "per-mixture output is not exported by the GPU path"; return HFB_EUNSUPPORTED; }
   CK(cudaSetDevice(c->device));
   { int rc = wait_impl(c); if (rc) return rc; }        // borrows slot 0's buffers: never under an in-flight wave
   const HostModel &h = c->hm;
   hfbgpu_ctx::Slot &S0 = c->slot[0];
   cudaStream_t st0 = S0.stream;
   for (int i = 0; i < n; i++) if (states[i] < 0 || states[i] >= h.J) return HFB_EINVAL;
   UttDesc u;
   memset(&u, 0, sizeof(u));
   const int nPad = (n + 3) & ~3;                       // row stride of the output-probability matrix
   u.T = T; u.J = nPad; u.Jt = n; u.P = n;
   UttOut o;
   memset(&o, 0, sizeof(o));
   std::vector<int2> items, items2, items4;
   for (int t0 = 0; t0 < T; t0 += TC_BM) items.push_back(make_int2(0, t0));
   for (int t0 = 0; t0 < T; t0 += 2 * TC_BM) items2.push_back(make_int2(0, t0));
   for (int t0 = 0; t0 < T; t0 += 4 * TC_BM) items4.push_back(make_int2(0, t0));
   const int nTiles = ((T + GT_FR - 1) / GT_FR) * ((n + GT_SL - 1) / GT_SL);
   const int tilePre[2] = {0, nTiles};
   std::vector<unsigned char> blob;
   const std::vector<int2> tiv((size_t)n, make_int2(0, T - 1));                   // every tile is needed in every frame
   size_t oUtt = blob_put(blob, &u, 1), oOut = blob_put(blob, &o, 1), oSs = blob_put(blob, states, (size_t)n),
          oTp = blob_put(blob, tilePre, 2), oIt = blob_put(blob, items), oIt2 = blob_put(blob, items2), oIt4 = blob_put(blob, items4),
          oTf = blob_put(blob, tiv);
   int rc;
   if ((rc = S0.dTables.reserve(blob.size())) || (rc = S0.dFeat.reserve((size_t)T * h.D + 4)) ||
       (rc = S0.dB.reserve((size_t)T * nPad + 1)))
      return rc;
   CK(cudaMemcpyAsync(S0.dTables.p, blob.data(), blob.size(), cudaMemcpyHostToDevice, st0));
   CK(cudaMemcpyAsync(S0.dFeat.p, feat, (size_t)T * h.D * sizeof(float), cudaMemcpyHostToDevice, st0));
   Wave W;
   memset(&W, 0, sizeof(W));
   W.utt = (UttDesc *)(S0.dTables.p + oUtt); W.out = (UttOut *)(S0.dTables.p + oOut); W.numUtt = 1;
   W.slotState = (int *)(S0.dTables.p + oSs); W.tilePre = (const int *)(S0.dTables.p + oTp);
   W.feat = S0.dFeat.p; W.b = S0.dB.p;
   W.tileIv = (int2 *)(S0.dTables.p + oTf);
   int gk = c->opt.gmmKernel;
   const bool v3 = c->useV3 && c->tc3.MP > 1;           // single-Gaussian sets: the tensor-core kernel needs slot = state
   if (gk == 0) gk = (v3 || gmm_tc_available(c->tc)) ? 2 : 1;
   if (gk == 2 && !v3 && !gmm_tc_available(c->tc)) gk = 1;
   if (gk == 2 && v3) {
      int nl = 0;
      if ((rc = gmm_tc3_launch(c->tc3, S0.tcw, c->dm, W, T, (const int2 *)(S0.dTables.p + oIt), (int)items.size(),
                               (const int2 *)(S0.dTables.p + oIt4), (int)items4.size(), c->smCount, st0, &nl))) return rc;
      c->stats.launches += nl; c->stats.launchesGmm += nl;
   } else if (gk == 2) {
      int nl = 0;
      if ((rc = gmm_tc_launch(c->tc, S0.tcw, c->dm, W, T, (const int2 *)(S0.dTables.p + oIt), (int)items.size(),
                              (const int2 *)(S0.dTables.p + oIt2), (int)items2.size(),
                              (const int2 *)(S0.dTables.p + oIt4), (int)items4.size(), c->smCount, st0, &nl))) return rc;
      c->stats.launches += nl; c->stats.launchesGmm += nl;
   } else {
      size_t smem = sizeof(float) * ((size_t)c->dm.D * GT_FR + (size_t)GT_FR * (GT_SL + 1));
      gmm_fp32_kernel<<<(unsigned)nTiles, 128, smem, st0>>>(c->dm, W);
      c->stats.launches++; c->stats.launchesGmm++;
   }
   CK(cudaGetLastError());
   CK(cudaMemcpy2DAsync(out, (size_t)n * sizeof(float), S0.dB.p, (size_t)nPad * sizeof(float), (size_t)n * sizeof(float), (size_t)T,
                        cudaMemcpyDeviceToHost, st0));
   CK(cudaStreamSynchronize(st0));
   return HFB_OK;
}

