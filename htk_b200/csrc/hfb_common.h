// hfb_common.h -- structures shared by the host API (hfbgpu.cu) and the kernels.
//
// Device-side conventions: model numbers q = 0..Q-1, state numbers i = 0..N-1
// (0 = non-emitting entry, N-1 = non-emitting exit), frames t = 0..T-1.  The
// reference (HTKLib/HFB.c) is 1-based throughout; beams are converted to its
// numbering only when they are copied out to the caller.
#pragma once
#include <stdint.h>
#include <vector_types.h>
#include "../../include/hfbgpu.h"

#define HFB_MAXN 16                 // max states per HMM handled by the recursion kernels
#define HFB_UTT_BETAWIDE 30000       // internal: beam outgrew the sliding beta window, the wide kernel redoes the utterance

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
   // per physical HMM / per transition matrix (read by the table-building kernel)
   const int *hmmN, *hmmStateOff, *hmmState, *hmmTrans;
   const int *transOffF;            // offset of the matrix in transLogA
   const int *transMinDur;          // SetMinDurs, HFB.c:106-155
   const long long *tranAccOff, *tranOccOff;   // offsets of TrAcc.tran / TrAcc.occ in the accumulators
   hfb_acc_layout L;
};

struct UttDesc {
   int T, Q;
   int S;                           // sum of N_q: doubles per beta column
   int P;                           // sum of (N_q - 2): emitting positions
   int J;                           // row stride of b[t][slot] = distinct tied states rounded up to a multiple of 4
                                    // (16-byte output stores of the tensor-core kernel); set by prep_kernel
   int labOff;                      // offset of the transcription in the wave's label array
   int modOff;                      // offset of model 0 in the per-model arrays
   int slotOff;                     // offset into slotState[]
   int posOff;                      // offset into posSlot[] / posState[]
   int Jt;                          // distinct tied states (output-probability slots in use); set by prep_kernel
   long long featOff;               // first frame in the feature matrix
   long long bOff;                  // floats : [T][J] state log-likelihoods
   long long betaOff;               // doubles: [T][S]
   long long occOff;                // doubles: [T][P] alpha of the emitting states (inside the alpha beam)
   long long aentOff;               // doubles: [T][Q] alpha of the entry states (inside the alpha beam)
   long long frameBase;             // first frame in the per-frame beam arrays
};

struct UttOut {                     // hfb_utt_result + what the host wants back
   int status;
   int retries;
   double pr;
   double thresh;
   int J;
   int redo;                        // fast alpha kernel gave up (window > 32 models): generic kernel redoes it
   long long pairs;                 // (frame, distinct state) pairs inside the slots' taper intervals: what K1 evaluates
};

struct Wave {                       // everything the kernels of one wave need
   UttDesc *utt;                    // [numUtt in wave]
   UttOut *out;
   int numUtt;
   const int *lab;                  // physical-HMM index per label, concatenated
   const int *posPre;               // [numUtt+1] prefix sums of P (emitting positions)
   const int *tilePre;              // [numUtt+1] prefix sums of FP32-GMM tiles
   int totalPos;
   // per-model arrays (concatenated over the wave's utterances), written by prep_kernel
   int *mN;                         // states of the model
   int *mTrans;                     // offset of its matrix in transLogA
   int *mSoff;                      // offset of its states inside a beta column
   int *mPoff;                      // offset of its emitting states inside an occ row
   int *mDms;                       // minimum duration (qDms)
   int *mPre;                       // sum of mDms over preceding models
   int *mSuf;                       // sum of mDms over following models
   int *mHmm;                       // physical HMM index
   long long *mTrAcc;               // offset of its TrAcc.tran block in the accumulators
   long long *mTrOcc;               // offset of its TrAcc.occ block
   int *mTmin, *mTmax;              // first / last frame inside the alpha beam
   int *slotState;                  // tied state of each slot
   // Frames in which a slot's output probability can be needed at all: the union, over the positions that share the
   // slot, of the beam taper of their model (SetBeamTaper, HFB.c:1116-1145) widened by the look-back / look-ahead of
   // SetBeta (models qLo-1 .. qHi of frame t+1, HFB.c:1207-1215).  The tensor-core kernel skips (tile, frame block)
   // combinations outside [first, last], as the reference's Setotprob does (HFB.c:1014-1016).  Written by prep_kernel.
   int *slotFirst, *slotLast;       // per slot (indexed like slotState)
   int2 *tileIv;                    // (first, last) per group of `spt` consecutive slots (indexed u.slotOff + tile)
   int spt;                         // slots per tensor-core tile (TC_BN / MP); 0 = no tensor-core kernel
   int globalSlots;                 // > 0: small single-Gaussian set -- slot j of every utterance IS tied state j (J = globalSlots)
   int noTaperSkip;                 // HFBGPU_NO_TAPER_SKIP: intervals cover the whole utterance
   int *posSlot;                    // slot of each emitting position
   int *posState;                   // tied state of each emitting position
   int *posQ;                       // model (label) index of each emitting position
   const float *feat;               // [frames][D]
   const float *feat2;              // single-pass retraining (HERest -r): second parameterisation, or NULL
   float *b;
   double *beta;
   double *occ;
   double *aent;
   short *qLo, *qHi, *sq, *eq;      // per-frame beams (0-based)
   double *acc;
   // options
   double pruneInit, pruneInc, pruneLim;
   double minFrwdP;
   int uFlags;
};
