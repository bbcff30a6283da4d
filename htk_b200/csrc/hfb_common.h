// hfb_common.h -- structures shared by the host API (hfbgpu.cu) and the kernels.
//
// Device-side conventions: model numbers q = 0..Q-1, state numbers i = 0..N-1
// (0 = non-emitting entry, N-1 = non-emitting exit), frames t = 0..T-1.  The
// reference (HTKLib/HFB.c) is 1-based throughout; beams are converted to its
// numbering only when they are copied out to the caller.
#pragma once
#include <stdint.h>
#include "../../include/hfbgpu.h"

#define HFB_MAXN 16                 // max states per HMM handled by the recursion kernels

struct DevModel {
   int D, Dp;                       // vector size, padded to a multiple of 4
   int G, J, P, numTrans, maxM;
   const float *mean, *ivar;        // [G][Dp]
   const float *gconst;             // [G]
   const int *meanId, *varId;       // [G]
   const int *stateMixOff;          // [J+1]
   const int *mixGauss;             // [sumM]
   const float *mixLogWt;           // [sumM]
   const float *transLogA;          // all matrices, row-major N*N each
   hfb_acc_layout L;
};

struct UttDesc {
   int T, Q;
   int S;                           // sum of N_q: doubles per beta column
   int P;                           // sum of (N_q - 2): emitting positions
   int J;                           // distinct tied states (output-probability slots)
   int modOff;                      // offset of model 0 in the per-model arrays
   int slotOff;                     // offset into slotState[]
   int posOff;                      // offset into posSlot[] / posState[]
   long long featOff;               // first frame in the feature matrix
   long long bOff;                  // floats : [T][J] state log-likelihoods
   long long betaOff;               // doubles: [T][S]
   long long occOff;                // doubles: [T][P] log occupancies / initx
   long long frameBase;             // first frame in the per-frame beam arrays
};

struct UttOut {                     // mirrors hfb_utt_result
   int status;
   int retries;
   double pr;
   double thresh;
};

struct PosRef { int utt, q, j; };   // one emitting state position (stats kernel work item)
struct GmmTile { int utt, t0, s0; };// one [frames x slots] tile of the FP32 GMM kernel

struct Wave {                       // everything the kernels of one wave need
   const UttDesc *utt;              // [numUtt in wave]
   UttOut *out;
   int numUtt;
   // per-model arrays (concatenated over the wave's utterances)
   const int *mN;                   // states of the model
   const int *mTrans;               // offset of its matrix in transLogA
   const int *mSoff;                // offset of its states inside a beta column
   const int *mPoff;                // offset of its emitting states inside an occ row
   const int *mDms;                 // minimum duration (qDms)
   const int *mPre;                 // sum of mDms over preceding models
   const int *mSuf;                 // sum of mDms over following models
   const int *mHmm;                 // physical HMM index
   const long long *mTrAcc;         // offset of its TrAcc.tran block in the accumulators
   const long long *mTrOcc;         // offset of its TrAcc.occ block
   int *mTmin, *mTmax;              // first / last frame inside the alpha beam
   const int *slotState;            // tied state of each slot
   const int *posSlot;              // slot of each emitting position
   const int *posState;             // tied state of each emitting position
   const float *feat;               // [frames][D]
   float *b;
   double *beta;
   double *occ;
   short *qLo, *qHi, *sq, *eq;      // per-frame beams (0-based)
   double *acc;
   // options
   double pruneInit, pruneInc, pruneLim;
   double minFrwdP;
   int uFlags;
};
