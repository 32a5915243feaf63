// kernels_tc.cuh -- bf16 tensor-core (tcgen05 / TMEM) path of the fused CIF step.  (stub: filled in next)
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>
#include "../../include/flowchain_b200.h"
#include "plan.h"

namespace fc {

// host pointers into the canonical blob, per step
struct TcStepSrc {
    const float* cpl_W[kMaxBlocks][2][kMaxHidden + 2];
    const float* cpl_b[kMaxBlocks][2][kMaxHidden + 2];
    const float* dens_W[2][2][3];
    const float* dens_b[2][2][3];
};

inline void tc_pack(const fc_model_desc&, std::vector<StepPlan>&, const std::vector<TcStepSrc>&, std::vector<unsigned char>& img) {
    img.clear();
}

inline int tc_launch_separate(const fc_model_desc&, const StepPlan*, const unsigned char*, size_t, const std::vector<StepPlan>&,
                              const RowArgs&, int, cudaStream_t, uint64_t*, std::string* msg) {
    *msg = "the bf16 tensor-core path is not built in this library";
    return 1;
}

}  // namespace fc
