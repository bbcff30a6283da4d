/* hfb_oracle.h -- CPU restatement of HTK's Baum-Welch E-step.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this.  The product (libhfbgpu) never links, imports or executes it.
 *
 * Parity status: PINNED.  The restatement is checked against the reference's own
 * HERest (built unmodified into oracle/_ref by oracle/Makefile) on HTKDemo and on
 * synthetic tied-state / tee-model sets: accumulator dumps (HER1.acc), per-utterance
 * log-likelihoods and per-frame beta/alpha beams.  See tests/test_oracle_vs_ref.py and
 * the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py).
 */
#ifndef HFB_ORACLE_H_
#define HFB_ORACLE_H_

#include "../include/hfbgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* SetMinDurs (HTKLib/HFB.c:91-155): minimum emitting-state path per transition matrix */
int hfbo_min_durs(const hfb_model *m, int32_t *out);

/* FBFile over a batch (HTKLib/HFB.c:1923 driven like HTKTools/HERest.c:502-534).
 *  acc        flat accumulators in hfbgpu_acc_layout() order
 *  accDouble  0: float elements, summed exactly like the reference (HTrain.h:211-232)
 *             1: double elements (higher-precision oracle for long workloads)
 *  threads    1 = strictly sequential like the reference; >1 = utterances in parallel
 *             with per-thread double accumulators (requires accDouble=1)
 */
int hfbo_accumulate(const hfb_model *m, const hfb_options *opt, const hfb_batch *b,
                    hfb_utt_result *res, const hfb_beams *beams,
                    void *acc, int accDouble, int threads);

/* ShStrP/MOutP (HTKLib/HFB.c:898-988, HModel.c:5420-5431): out[T][n], mixOut optional */
int hfbo_state_loglik(const hfb_model *m, const float *feat, int32_t T,
                      const int32_t *states, int32_t n, float *out, float *mixOut);

/* per-(t, emitting state position) occupancies of ONE utterance, for localisation:
 * occ[T][P] with P = sum over labels of (N_q - 2); 0 outside the alpha beam.       */
int hfbo_utt_occupancy(const hfb_model *m, const hfb_options *opt,
                       const float *feat, int32_t T, const int32_t *lab, int32_t Q,
                       float *occ, hfb_utt_result *res);

#ifdef __cplusplus
}
#endif
#endif
