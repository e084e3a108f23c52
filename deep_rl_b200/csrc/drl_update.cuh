// drl_update.cuh -- arguments shared by the two implementations of the minibatch gradient
// (FP32 CUDA-core path in update_ops.cu, tcgen05 tensor-core path in update_tc.cu).
#pragma once
#include "drl_common.cuh"

namespace drl {

constexpr int DRL_MAX_STEPS_PER_LAUNCH = 8;

struct AdamArgs {
    float* params; const float* grad; float* m; float* v; float* packed; float* norm_out;
    int P;
    float grad_scale, max_norm, beta2, om_beta1, om_beta2, eps, neg_step_size, bc2_sqrt;
};

// Optional in-kernel tail of the tensor-core gradient kernel (single GPU): after a grid-wide barrier the CTAs fold
// the partial gradients slice-wise, clip by the global norm and apply Adam -- the whole minibatch step is one launch.
struct TailArgs {
    int enabled;
    AdamArgs a;
    float* grad_out; float* loss_terms_out;
    double* cta_sumsq;        // [gridDim.x]
    uint32_t* ctr;            // [4] grid barriers 1-3, depart (zero between launches)
    // multi-GPU (world > 1): symmetric peer buffers, layout = CommLayout
    int world, rank;
    unsigned char* peer[DRL_MAX_RANKS];
    uint32_t seq;
    int* error_flag;
    // graph-replayable launch (drl_ppo_minibatch_update_ctl): Adam scalars and sequence number come from device memory
    const drl_ctrl_t* ctrl;
    int ordinal;
    // one launch = nsteps consecutive, equally sized minibatches (GradArgs.mb_start + s * mb_count, adv_stats[2s], loss terms
    // row s, optimizer step s): by-value Adam scalars per step (ctrl == nullptr)
    int nsteps;
    float neg_step_size_s[DRL_MAX_STEPS_PER_LAUNCH], bc2_sqrt_s[DRL_MAX_STEPS_PER_LAUNCH];
};

// Symmetric buffer of one rank (push protocol): two generations (seq & 1) of [DRL_MAX_RANKS] gradient copies -- rank s WRITES its
// folded slices into slot s of every rank's buffer -- followed by two generations of arrival flags [DRL_MAX_RANKS][FLAG_CTAS]
// (one flag per source rank and CTA slice, holding the sequence number of the step that wrote the slice).
constexpr int FLAG_CTAS = 160;
struct CommLayout {
    size_t xgrad[2], flags[2], rank_stride, total;
};
__host__ __device__ inline CommLayout comm_layout(int64_t P) {
    CommLayout c;
    const size_t g = ((size_t)P * 4 + 255) / 256 * 256;
    c.rank_stride = g;
    c.xgrad[0] = 0; c.xgrad[1] = (size_t)DRL_MAX_RANKS * g;
    c.flags[0] = 2 * (size_t)DRL_MAX_RANKS * g; c.flags[1] = c.flags[0] + sizeof(uint32_t) * DRL_MAX_RANKS * FLAG_CTAS;
    c.total = c.flags[1] + sizeof(uint32_t) * DRL_MAX_RANKS * FLAG_CTAS;
    return c;
}

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
    TailArgs tail;
};

// fixed-order fold of `grid` per-CTA partial gradients / loss sums (update_ops.cu)
int launch_grad_reduce(const GradArgs& g, int grid, int P, float* grad_out, float* loss_terms_out, cudaStream_t st);

// tensor-core implementation (update_tc.cu)
// grid_out != nullptr: skip the fold of the partials and report the number of partials instead
int launch_grad_tc(const drl_net_t* net, const GradArgs& g, int P, float* grad_out, float* loss_terms_out, cudaStream_t st,
                   int* grid_out = nullptr);
// tensor-core gradient + in-kernel fold / clip / Adam tail (g.tail filled by the caller), one cooperative launch
int launch_grad_tc_fused(const drl_net_t* net, GradArgs& g, cudaStream_t st);

}  // namespace drl
