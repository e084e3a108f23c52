// drl_update.cuh -- arguments shared by the two implementations of the minibatch gradient
// (FP32 CUDA-core path in update_ops.cu, tcgen05 tensor-core path in update_tc.cu).
#pragma once
#include "drl_common.cuh"

namespace drl {

struct GradArgs {
    const float* packed;
    const float* rec;
    const uint32_t* idx;
    uint32_t mb_start, mb_count;
    const float* adv_stats;   // [2] mean, std
    float clip_coef, ent_coef, vf_coef;
    float* grad_part;         // [gridDim.x][ppad]
    float* loss_part;         // [gridDim.x][LOSS_TERMS]
    int ppad;
    long long* dbg;           // optional cycle stamps (diagnostics), else nullptr
};

// fixed-order fold of `grid` per-CTA partial gradients / loss sums (update_ops.cu)
int launch_grad_reduce(const GradArgs& g, int grid, int P, float* grad_out, float* loss_terms_out, cudaStream_t st);

// tensor-core implementation (update_tc.cu)
// grid_out != nullptr: skip the fold of the partials and report the number of partials instead
int launch_grad_tc(const drl_net_t* net, const GradArgs& g, int P, float* grad_out, float* loss_terms_out, cudaStream_t st,
                   int* grid_out = nullptr);

}  // namespace drl
