// gmm_tc.cuh -- tcgen05 (5th-gen tensor core) GMM log-likelihood path.  Placeholder
// interface; the kernel lands in a later commit.  Until then gmm_tc_available() is false
// and the FP32 CUDA-core kernel is the one that runs.
#pragma once
#include <vector>
#include "hfb_common.h"

struct GmmTcModel {
   bool ready = false;
};

static inline int gmm_tc_prepare(GmmTcModel &, const hfb_model *, cudaStream_t) { return HFB_OK; }
static inline void gmm_tc_release(GmmTcModel &) {}
static inline bool gmm_tc_available(const GmmTcModel &t) { return t.ready; }
static inline int gmm_tc_launch(GmmTcModel &, const DevModel &, const Wave &, const std::vector<UttDesc> &,
                                cudaStream_t, int *) { return HFB_EUNSUPPORTED; }
