// drl_pack.cuh -- canonical (state_dict order) <-> packed (kernel) parameter index map, and the
// layout of the caller-owned update workspace.
#pragma once
#include <cuda_bf16.h>

#include "drl_common.cuh"

namespace drl {

int check_net(const drl_net_t* net);

// Store canonical parameter `i` (value v) at its packed position(s).  The H x H matrices are kept
// twice: W2T for the forward pass, W2P for the backward pass.
template <int O, int A>
__device__ __forceinline__ void packed_store(float* __restrict__ packed, int i, float v) {
    using P = Packed<O, A>;
    const int net = i >= P::C_ACTOR ? 1 : 0;
    int r = i - net * P::C_ACTOR;
    if (r < H * O) {
        const int o = r / O, k = r % O;
        packed[P::W1T + (net * O + k) * H + perm_pos(o)] = v;
        // fp32 container, bf16-rounded value: the tensor-core rollout computes layer 1 on CUDA cores with exactly the operand
        // values the update's layer-1 GEMM sees (TC_W1B below), so that log-probs recorded there and re-evaluated here agree
        packed[P::TC_W1 + (net * H + o) * P::OW + k] = __bfloat162float(__float2bfloat16_rn(v));
        __nv_bfloat16* w1b = reinterpret_cast<__nv_bfloat16*>(packed + P::TC_W1B) + net * (2 * H * 8);
        w1b[o * 8 + k] = __float2bfloat16_rn(v);             // chunk 0, k < O  (x_hi)
        w1b[H * 8 + o * 8 + k] = __float2bfloat16_rn(v);     // chunk 1, k + 8  (x_lo)
        return;
    }
    r -= H * O;
    if (r < H) {
        packed[P::B1 + net * H + perm_pos(r)] = v;
        packed[P::TC_B1 + net * H + r] = __bfloat162float(__float2bfloat16_rn(v));
        reinterpret_cast<__nv_bfloat16*>(packed + P::TC_W1B)[net * (2 * H * 8) + r * 8 + O] = __float2bfloat16_rn(v);   // ones column
        return;
    }
    r -= H;
    if (r < H * H) {
        const int o = r / H, k = r % H;
        packed[P::W2T + (net * H + k) * H + perm_pos(o)] = v;
        packed[P::W2P + (net * H + o) * H + perm_pos(k)] = v;
        // bf16 SW128 tile image: row o, 16-byte chunk k/8 stored at chunk position (k/8) ^ (o & 7)
        unsigned char* tc = reinterpret_cast<unsigned char*>(packed + P::TC_W2);
        const int off = net * (H * H * 2) + o * 128 + ((((k >> 3) ^ (o & 7)) << 4)) + (k & 7) * 2;
        *reinterpret_cast<__nv_bfloat16*>(tc + off) = __float2bfloat16_rn(v);
        return;
    }
    r -= H * H;
    if (r < H) {
        packed[P::B2 + net * H + perm_pos(r)] = v;
        packed[P::TC_B2 + net * H + r] = v;
        return;
    }
    r -= H;
    const int nout = net == 0 ? A : 1;
    if (r < nout * H) {
        const int a = r / H, k = r % H;
        packed[P::W4 + (net * A + a) * H + perm_pos(k)] = v;
        packed[P::TC_W4 + (net == 0 ? a : A) * H + k] = v;
        return;
    }
    r -= nout * H;
    packed[P::B4 + net * A + r] = v;
    packed[P::TC_B4 + (net == 0 ? r : A)] = v;
}

// ---- update workspace (caller-owned, zero-initialised once) ----
constexpr int MAX_GRAD_CTAS = 160;   // >= SM count (148); the grad kernel runs one persistent CTA per SM
constexpr int MAX_MINIBATCHES = 16;
constexpr int STAT_PARTS = 512;      // max partial-sum CTAs per minibatch in drl_adv_stats
constexpr int LOSS_TERMS = 8;

struct WorkspaceLayout {
    size_t counters, stat_partials, cta_sumsq, loss_partials, grad_partials, debug, ev_partials, stage, total;
    int ppad;
};
// hidden = 256 adds the staging buffers of update256.cu: h1 and dz2 of 2 nets x 4096 tiles x 64 KB = 1 GB
inline WorkspaceLayout workspace_layout(int64_t P, int hidden = 64) {
    WorkspaceLayout w;
    w.ppad = (int)((P + 3) / 4 * 4);
    w.counters = 0;
    w.stat_partials = 64;
    // squared-norm shares of the fused clip+Adam steps: their own region, the statistics of a later epoch may run on
    // another stream while a minibatch step is in flight
    w.cta_sumsq = w.stat_partials + sizeof(double) * MAX_MINIBATCHES * STAT_PARTS * 2;
    w.ev_partials = w.cta_sumsq + sizeof(double) * 1024;   // drl_explained_variance: [STAT_PARTS][4] doubles (counter word 12);
                                                           // every region up to here has a size independent of P
    w.loss_partials = w.ev_partials + sizeof(double) * STAT_PARTS * 4;
    w.grad_partials = w.loss_partials + sizeof(float) * MAX_GRAD_CTAS * LOSS_TERMS;
    w.debug = (w.grad_partials + sizeof(float) * (size_t)MAX_GRAD_CTAS * w.ppad + 1023) / 1024 * 1024;   // 4 KB of cycle stamps
    w.stage = w.debug + 4096;                       // (-DDRL_TC_STAMPS builds; hidden = 64: the LAST 4 KB of the workspace)
    w.total = w.stage + (hidden == 256 ? (size_t)2 * 2 * 4096 * 65536 : 0);
    return w;
}
inline size_t workspace_bytes_for(int64_t P, int hidden = 64) { return workspace_layout(P, hidden).total; }

}  // namespace drl
