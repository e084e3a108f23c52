// replay_ops.cu -- replay-buffer sampling and gather: the random-index gather family of the reference's
// off-policy scripts, over flat struct-of-arrays storage in HBM.
//   deep_rl/dqn.py:116-122   batch_inds = np.random.randint(global_step, size=batch_size); b_x = x[batch_inds], b_next = x[batch_inds + 1]
//   deep_rl/per.py:127-135   probabilities = p^alpha / sum(p^alpha); batch_inds = torch.multinomial(priorities, batch, replacement=True)
//   deep_rl/per.py:144-146   priorities[batch_inds] = |td_errors|; max_priority = max(max(priorities), max_priority)
// The reference's RNG streams (numpy MT19937, torch CPU mt19937) are replaced by the build's Philox contract: draw
// i of call `draw_ctr` is word 0 of philox(i, draw_lo, draw_hi, TAG_REPLAY); the prioritized sampler is an inverse-CDF draw over
// prefix sums whose association is fixed (1024-element blocks: lane-strided partial sums, butterfly fold; blocks and the
// in-block search sequential, all in float64), so that the CPU oracle reproduces every index bit for bit.
// All HBM-bound: the gather moves 2 x OP x 4 + 9 bytes per sampled transition in 32-byte sectors.
#include "drl_common.cuh"

namespace drl {

constexpr uint32_t TAG_REPLAY = 3;
constexpr int PRI_BLOCK = 1024;      // priorities per prefix-sum block

__device__ __forceinline__ uint32_t replay_word(uint64_t seed, uint32_t i, uint64_t draw_ctr) {
    return philox_seeded(seed, i, (uint32_t)draw_ctr, (uint32_t)(draw_ctr >> 32), TAG_REPLAY).x;
}

// idx[i] = floor(u32 * size / 2^32): the multiply-shift map of a 32-bit word onto [0, size)
__global__ void __launch_bounds__(256) replay_uniform_kernel(uint32_t* __restrict__ idx, uint32_t batch, uint32_t size, uint64_t seed,
                                                              uint64_t draw_ctr) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < batch) idx[i] = (uint32_t)(((uint64_t)replay_word(seed, i, draw_ctr) * (uint64_t)size) >> 32);
}

// one thread per sampled transition and 16-byte observation chunk would split the rows; a thread per transition keeps every
// row's loads together (two rows of OP floats, act, rew, term at idx and idx + 1)
template <int OP>
__global__ void __launch_bounds__(256) replay_gather_kernel(const float* __restrict__ obs, const int32_t* __restrict__ act,
                                                             const float* __restrict__ rew, const uint8_t* __restrict__ term,
                                                             const uint32_t* __restrict__ idx, uint32_t batch, float* __restrict__ b_obs,
                                                             float* __restrict__ b_next, int32_t* __restrict__ b_act,
                                                             float* __restrict__ b_rew, uint8_t* __restrict__ b_term) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const size_t s = idx[i];
    const float4* o = reinterpret_cast<const float4*>(obs + s * OP);
    float4 v[2 * (OP / 4)];
#pragma unroll
    for (int q = 0; q < 2 * (OP / 4); ++q) v[q] = __ldg(o + q);          // rows s and s + 1 are adjacent
    const int32_t a = __ldg(act + s);
    const float r = __ldg(rew + s + 1);
    const uint8_t t = __ldg(term + s + 1);
    float4* bo = reinterpret_cast<float4*>(b_obs + (size_t)i * OP);
    float4* bn = reinterpret_cast<float4*>(b_next + (size_t)i * OP);
#pragma unroll
    for (int q = 0; q < OP / 4; ++q) { bo[q] = v[q]; bn[q] = v[OP / 4 + q]; }
    b_act[i] = a; b_rew[i] = r; b_term[i] = t;
}

// ---- prioritized sampling ----
// pass 1: per 1024-block sums of w = priorities (sampling weights, per.py:129) and of w^alpha (per.py:128), float64.
// One warp per block: lane l adds elements l, l + 32, ... in order, then the 32 lane sums fold in the xor-butterfly order.
__global__ void __launch_bounds__(256) priority_block_sums_kernel(const float* __restrict__ pri, uint32_t size, float alpha,
                                                                   double* __restrict__ bsum, double* __restrict__ bsum_alpha) {
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const uint32_t nblk = (size + PRI_BLOCK - 1) / PRI_BLOCK;
    if (warp >= nblk) return;
    const uint32_t lo = warp * PRI_BLOCK, hi = min(size, lo + PRI_BLOCK);
    double s = 0.0, sa = 0.0;
    for (uint32_t i = lo + lane; i < hi; i += 32) {
        const float p = __ldg(pri + i);
        s += (double)p;
        sa += (double)powf(p, alpha);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        sa += __shfl_xor_sync(0xffffffffu, sa, o);
    }
    if (lane == 0) { bsum[warp] = s; bsum_alpha[warp] = sa; }
}

// pass 2 (one warp): exclusive prefix of the block sums in sequential (left to right) order; totals at [nblk].  The warp stages
// 2048 sums at a time in shared memory with coalesced loads, lane 0 adds them up in order -- the association of a plain loop,
// without 4096 dependent global-memory round trips.
__global__ void __launch_bounds__(32) priority_block_prefix_kernel(double* __restrict__ bsum, double* __restrict__ bsum_alpha, uint32_t nblk) {
    __shared__ double sh[2048], sha[2048];
    const int lane = threadIdx.x;
    double run = 0.0, runa = 0.0;
    for (uint32_t b0 = 0; b0 < nblk; b0 += 2048) {
        const uint32_t nb = min(2048u, nblk - b0);
        for (uint32_t i = lane; i < nb; i += 32) { sh[i] = bsum[b0 + i]; sha[i] = bsum_alpha[b0 + i]; }
        __syncwarp();
        if (lane == 0) {
            for (uint32_t i = 0; i < nb; ++i) {
                const double s = sh[i];
                sh[i] = run;
                run += s;
                runa += sha[i];
            }
        }
        __syncwarp();
        for (uint32_t i = lane; i < nb; i += 32) bsum[b0 + i] = sh[i];
        __syncwarp();
    }
    if (lane == 0) { bsum[nblk] = run; bsum_alpha[nblk] = runa; }
}

// pass 3: draw i: target = u * total (u = (w + 0.5) / 2^32 in float64); the last block whose exclusive prefix is <= target, then a
// sequential walk through that block (cumulative float64 sum in index order) to the first element whose inclusive sum exceeds it
__global__ void __launch_bounds__(128) priority_sample_kernel(const float* __restrict__ pri, uint32_t size, float alpha,
                                                               const double* __restrict__ bprefix, const double* __restrict__ bsum_alpha,
                                                               uint32_t nblk, uint32_t batch, uint64_t seed, uint64_t draw_ctr,
                                                               uint32_t* __restrict__ idx, float* __restrict__ prob) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const double total = bprefix[nblk];
    const double target = u01_f64(replay_word(seed, i, draw_ctr)) * total;
    uint32_t lo = 0, hi = nblk;            // invariant: bprefix[lo] <= target; answer block in [lo, hi)
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (bprefix[mid] <= target) lo = mid; else hi = mid;
    }
    const uint32_t e0 = lo * PRI_BLOCK, e1 = min(size, e0 + PRI_BLOCK);
    double run = bprefix[lo];
    uint32_t pick = e1 - 1;
    for (uint32_t e = e0; e < e1; ++e) {
        run += (double)__ldg(pri + e);
        if (run > target) { pick = e; break; }
    }
    idx[i] = pick;
    if (prob != nullptr) prob[i] = (float)((double)powf(__ldg(pri + pick), alpha) / bsum_alpha[nblk]);
}

// priorities[idx] = |td|; max_priority = max(max over the written values, max_priority).  Duplicated indices: the reference's
// index_put keeps the LAST occurrence; so does this kernel (a thread only writes if no later batch element has the same index).
__global__ void __launch_bounds__(256) priority_update_kernel(float* __restrict__ pri, const uint32_t* __restrict__ idx,
                                                               const float* __restrict__ td, uint32_t batch, float* __restrict__ max_pri) {
    __shared__ float smax[8];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.0f;
    if (i < batch) {
        v = fabsf(td[i]);
        const uint32_t me = idx[i];
        bool last = true;
        for (uint32_t k = i + 1; k < batch; ++k)
            if (__ldg(idx + k) == me) { last = false; break; }
        if (last) pri[me] = v;
        else v = 0.0f;                 // an overwritten duplicate never reaches the buffer, so it cannot raise the maximum either
    }
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 16)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 8));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 4)); v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    if ((threadIdx.x & 31) == 0) smax[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = smax[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, smax[w]);
        atomicMax(reinterpret_cast<int*>(max_pri), __float_as_int(m));      // non-negative floats order like their bit patterns
    }
}

}  // namespace drl

using namespace drl;

extern "C" {

int drl_replay_sample_uniform(uint32_t* idx_out, uint32_t batch, uint32_t size, uint64_t seed, uint64_t draw_ctr, void* stream) {
    DRL_REQUIRE(idx_out, "drl_replay_sample_uniform: idx_out is NULL");
    DRL_REQUIRE(size > 0, "drl_replay_sample_uniform: empty buffer");
    if (batch == 0) return DRL_OK;
    replay_uniform_kernel<<<(batch + 255) / 256, 256, 0, as_stream(stream)>>>(idx_out, batch, size, seed, draw_ctr);
    DRL_LAUNCH_CHECK("replay_uniform_kernel");
    return DRL_OK;
}

int drl_replay_gather(const float* obs, const int32_t* act, const float* rew, const uint8_t* term, const uint32_t* idx, uint32_t batch,
                      int32_t obs_stride, float* b_obs, float* b_next_obs, int32_t* b_act, float* b_rew, uint8_t* b_term, void* stream) {
    DRL_REQUIRE(obs && act && rew && term && idx && b_obs && b_next_obs && b_act && b_rew && b_term, "drl_replay_gather: NULL pointer");
    DRL_REQUIRE(obs_stride == 4 || obs_stride == 8, "drl_replay_gather: obs_stride=%d (4 or 8)", obs_stride);
    if (batch == 0) return DRL_OK;
    const int blocks = (int)((batch + 255) / 256);
    if (obs_stride == 4)
        replay_gather_kernel<4><<<blocks, 256, 0, as_stream(stream)>>>(obs, act, rew, term, idx, batch, b_obs, b_next_obs, b_act, b_rew, b_term);
    else
        replay_gather_kernel<8><<<blocks, 256, 0, as_stream(stream)>>>(obs, act, rew, term, idx, batch, b_obs, b_next_obs, b_act, b_rew, b_term);
    DRL_LAUNCH_CHECK("replay_gather_kernel");
    return DRL_OK;
}

size_t drl_replay_scratch_bytes(uint32_t size) { return sizeof(double) * 2 * ((size_t)(size + PRI_BLOCK - 1) / PRI_BLOCK + 1); }

int drl_replay_sample_priority(const float* priorities, uint32_t size, float alpha, uint32_t batch, uint64_t seed, uint64_t draw_ctr,
                               uint32_t* idx_out, float* prob_out, void* scratch, size_t scratch_bytes, void* stream) {
    DRL_REQUIRE(priorities && idx_out && scratch, "drl_replay_sample_priority: NULL pointer");
    DRL_REQUIRE(size > 0, "drl_replay_sample_priority: empty buffer");
    DRL_REQUIRE(scratch_bytes >= drl_replay_scratch_bytes(size), "drl_replay_sample_priority: scratch %zu < %zu bytes", scratch_bytes,
                drl_replay_scratch_bytes(size));
    const uint32_t nblk = (size + PRI_BLOCK - 1) / PRI_BLOCK;
    double* bsum = reinterpret_cast<double*>(scratch);
    double* bsum_alpha = bsum + nblk + 1;
    cudaStream_t st = as_stream(stream);
    priority_block_sums_kernel<<<(nblk + 7) / 8, 256, 0, st>>>(priorities, size, alpha, bsum, bsum_alpha);
    DRL_LAUNCH_CHECK("priority_block_sums_kernel");
    priority_block_prefix_kernel<<<1, 32, 0, st>>>(bsum, bsum_alpha, nblk);
    DRL_LAUNCH_CHECK("priority_block_prefix_kernel");
    if (batch > 0) {
        priority_sample_kernel<<<(batch + 127) / 128, 128, 0, st>>>(priorities, size, alpha, bsum, bsum_alpha, nblk, batch, seed, draw_ctr,
                                                                    idx_out, prob_out);
        DRL_LAUNCH_CHECK("priority_sample_kernel");
    }
    return DRL_OK;
}

int drl_replay_update_priorities(float* priorities, const uint32_t* idx, const float* td_errors, uint32_t batch, float* max_priority_inout,
                                 void* stream) {
    DRL_REQUIRE(priorities && idx && td_errors && max_priority_inout, "drl_replay_update_priorities: NULL pointer");
    DRL_REQUIRE(batch <= 16384, "drl_replay_update_priorities: batch=%u > 16384 (the last-duplicate-wins scan is quadratic)", batch);
    if (batch == 0) return DRL_OK;
    priority_update_kernel<<<(batch + 255) / 256, 256, 0, as_stream(stream)>>>(priorities, idx, td_errors, batch, max_priority_inout);
    DRL_LAUNCH_CHECK("priority_update_kernel");
    return DRL_OK;
}

}  // extern "C"
