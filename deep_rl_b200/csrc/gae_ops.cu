// gae_ops.cu -- GAE/returns reverse-time scan (deep_rl/ppo.py:144-151), minibatch permutation
// (ppo.py:155) and per-minibatch advantage statistics (ppo.py:169).  All HBM-bound, coalesced SoA.
#include "drl_pack.cuh"

namespace drl {

// ---------------------------------------------------------------------------------------------
// GAE: one thread per env walks t = T-1 .. 0; lanes of a warp are 32 consecutive envs, so every
// load/store is a fully coalesced 128-byte (f32) or 32-byte (u8) row segment.  The arithmetic keeps
// the reference's grouping and fp32 rounding (-fmad=false: nothing below is contracted):
//     adv[t] = rew[t+1] + gamma*(1 - done[t+1]) * (val[t+1] + lambda*last) - val[t]
// Optionally packs the per-sample record the update kernels gather: one 32-byte (O<=4) or 64-byte
// sector-aligned row {obs.., logp, adv, val, act}, so a random minibatch gather costs one/two sectors
// instead of six.  ret = adv + val is recomputed by the consumer (bit-identical fp32 add).
// ---------------------------------------------------------------------------------------------
template <int OP, int RW>
__global__ void __launch_bounds__(128) gae_kernel(const float* __restrict__ rew, const uint8_t* __restrict__ done,
                                                   const float* __restrict__ val, const float* __restrict__ obs,
                                                   const uint8_t* __restrict__ act, const float* __restrict__ logp,
                                                   int T, int N, float gamma, float lam, float* __restrict__ adv,
                                                   float* __restrict__ ret, float* __restrict__ rec) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float last = 0.0f;
    float v1 = val[(size_t)T * N + n];
    adv[(size_t)T * N + n] = 0.0f;
    ret[(size_t)T * N + n] = 0.0f + v1;
    // The recurrence is a chain of two dependent roundings per step; everything it reads is independent of it.  Each thread
    // therefore fetches a block of GB steps (all loads in flight at once), then runs the chain over registers: with few envs
    // per GPU the kernel is bound by load latency, not bandwidth.
    constexpr int GB = 8;
    for (int tb = T - 1; tb >= 0; tb -= GB) {
        float r_[GB], v_[GB], lp_[GB];
        uint8_t d_[GB], a_[GB];
        float4 o_[GB][OP / 4];
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const int t = tb - j;
            if (t >= 0) {
                const size_t i0 = (size_t)t * N + n, i1 = i0 + N;
                d_[j] = done[i1];
                r_[j] = rew[i1];
                v_[j] = val[i0];
                if (rec != nullptr) {
                    lp_[j] = logp[i0];
                    a_[j] = act[i0];
                    const float4* o4 = reinterpret_cast<const float4*>(obs + i0 * OP);
#pragma unroll
                    for (int q = 0; q < OP / 4; ++q) o_[j][q] = o4[q];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < GB; ++j) {
            const int t = tb - j;
            if (t >= 0) {
                const size_t i0 = (size_t)t * N + n;
                const float d = (float)d_[j];
                const float v0 = v_[j];
                const float nd = 1.0f - d;
                const float a = gamma * nd;
                const float ll = lam * last;
                const float b = v1 + ll;
                const float c = a * b;
                const float dd = r_[j] + c;
                const float ad = dd - v0;
                adv[i0] = ad;
                ret[i0] = ad + v0;
                if (rec != nullptr) {
                    float4* r4 = reinterpret_cast<float4*>(rec + i0 * RW);
                    r4[0] = o_[j][0];
                    if (RW == 16) {
                        r4[1] = o_[j][OP / 4 - 1];
                        r4[2] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    r4[RW / 4 - 1] = make_float4(lp_[j], ad, v0, __int_as_float((int)a_[j]));
                }
                last = ad;
                v1 = v0;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Permutation: idx[i] = cycle-walked 6-round alternating Feistel network on ceil(log2 B) bits.  Round keys are
// Philox words (seed; epoch_ctr, rank), round function = murmur3 finalizer of (half ^ key).
// Integer-only => bit-exact with the CPU oracle's restatement.  HBM-bound: 4 bytes written per index.
// ---------------------------------------------------------------------------------------------
constexpr int PERM_ROUNDS = 6;

__device__ __forceinline__ uint32_t fmix32(uint32_t h) {
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__device__ __forceinline__ uint32_t perm_index(uint32_t i, uint32_t B, uint32_t a, uint32_t b, const uint32_t (&keys)[8]) {
    uint32_t x = i;
    do {
        uint32_t lb = a, rb = b;
        uint32_t L = x >> rb, R = x & ((1u << rb) - 1u);
#pragma unroll
        for (uint32_t r = 0; r < PERM_ROUNDS; ++r) {
            const uint32_t nR = L ^ (fmix32(R ^ keys[r]) & ((1u << lb) - 1u));
            L = R;
            R = nR;
            const uint32_t t = lb; lb = rb; rb = t;
        }
        x = (L << rb) | R;
    } while (x >= B);
    return x;
}

__global__ void __launch_bounds__(256) permutation_kernel(uint32_t* __restrict__ idx, uint32_t B, uint32_t a, uint32_t b,
                                                           uint64_t seed, uint32_t epoch_ctr, uint32_t rank,
                                                           const drl_ctrl_t* __restrict__ ctrl) {
    if (ctrl != nullptr) epoch_ctr += ctrl->epoch_ctr;     // graph-replayable launch: epoch_ctr is the offset inside the update
    const uint4 k0 = philox_seeded(seed, epoch_ctr, rank, 0u, TAG_PERM);
    const uint4 k1 = philox_seeded(seed, epoch_ctr, rank, 1u, TAG_PERM);
    const uint32_t keys[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride)
        idx[i] = perm_index(i, B, a, b, keys);
}

// ---------------------------------------------------------------------------------------------
// Advantage statistics of every minibatch of one epoch in one launch.  grid = (STAT_PARTS, nmb);
// each CTA sums its slice in fp64, the last CTA to finish folds the partials in a fixed order
// (deterministic) and writes mean and unbiased std as fp32.
// ---------------------------------------------------------------------------------------------
template <int RW>
__global__ void __launch_bounds__(256) adv_stats_kernel(const float* __restrict__ rec, const uint32_t* __restrict__ idx,
                                                         uint32_t B, uint32_t mb_size, float* __restrict__ stats_out,
                                                         double* __restrict__ partials, uint32_t* __restrict__ counter) {
    const uint32_t mb = blockIdx.y, part = blockIdx.x, nmb = gridDim.y, nparts = gridDim.x;
    const uint32_t lo = mb * mb_size;
    const uint32_t hi = min(B, lo + mb_size);
    double s = 0.0, ss = 0.0;
    // four independent (index -> record) gathers in flight per thread: the loop is latency-bound otherwise
    const uint32_t step = nparts * 256u;
    uint32_t i = lo + part * 256u + threadIdx.x;
    for (; i + 3u * step < hi; i += 4u * step) {
        uint32_t sx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) sx[j] = idx ? __ldg(idx + i + j * step) : i + j * step;
        float av[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) av[j] = __ldg(rec + (size_t)sx[j] * RW + (RW - 3));
#pragma unroll
        for (int j = 0; j < 4; ++j) { const double a = (double)av[j]; s += a; ss = fma(a, a, ss); }
    }
    for (; i < hi; i += step) {
        const uint32_t sidx = idx ? idx[i] : i;
        const double a = (double)rec[(size_t)sidx * RW + (RW - 3)];
        s += a;
        ss = fma(a, a, ss);
    }
    __shared__ double sh[2][8];
    __shared__ bool is_last;
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh[0][warp] = s; sh[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0.0, tss = 0.0;
        for (int w = 0; w < 8; ++w) { ts += sh[0][w]; tss += sh[1][w]; }
        partials[((size_t)mb * STAT_PARTS + part) * 2 + 0] = ts;
        partials[((size_t)mb * STAT_PARTS + part) * 2 + 1] = tss;
        __threadfence();
        const uint32_t done = atomicAdd(counter, 1u);
        is_last = (done == nmb * nparts - 1);
    }
    __syncthreads();
    if (is_last && threadIdx.x < nmb) {
        __threadfence();
        const uint32_t k = threadIdx.x;
        double ts = 0.0, tss = 0.0;
        for (uint32_t p = 0; p < nparts; ++p) {
            ts += __ldcg(&partials[((size_t)k * STAT_PARTS + p) * 2 + 0]);
            tss += __ldcg(&partials[((size_t)k * STAT_PARTS + p) * 2 + 1]);
        }
        const uint32_t klo = k * mb_size, khi = min(B, klo + mb_size);
        const double cnt = (double)(khi - klo);
        const double mean = ts / cnt;
        const double var = (tss - ts * mean) / (cnt - 1.0);   // unbiased, torch.std default
        stats_out[2 * k + 0] = (float)mean;
        stats_out[2 * k + 1] = (float)sqrt(var > 0.0 ? var : 0.0);
        if (k == 0) *counter = 0;   // re-arm for the next launch (stream-ordered)
    }
}

// Same statistics without the gather: walk the samples in NATURAL order (coalesced 4-byte reads of the advantage
// plane), invert the keyed Feistel permutation to find each sample's position and hence its minibatch.
// HBM traffic 4 B/sample instead of a 32-byte sector per gathered sample.
constexpr int STATS_MAX_MB = 8;

__device__ __forceinline__ uint32_t perm_position(uint32_t s, uint32_t B, uint32_t a, uint32_t b, const uint32_t (&keys)[8]) {
    uint32_t x = s;
    do {
        uint32_t lb = a, rb = b;
        uint32_t L = x >> rb, R = x & ((1u << rb) - 1u);
#pragma unroll
        for (int r = PERM_ROUNDS - 1; r >= 0; --r) {
            const uint32_t t = lb; lb = rb; rb = t;
            const uint32_t R0 = L;
            const uint32_t L0 = R ^ (fmix32(R0 ^ keys[r]) & ((1u << lb) - 1u));
            L = L0; R = R0;
        }
        x = (L << rb) | R;
    } while (x >= B);
    return x;
}

__global__ void __launch_bounds__(256) adv_stats_perm_kernel(const float* __restrict__ adv, uint32_t B, uint32_t mb_size, uint32_t nmb,
                                                              uint32_t a, uint32_t b, uint64_t seed, uint32_t epoch_ctr, uint32_t rank,
                                                              float* __restrict__ stats_out, double* __restrict__ partials,
                                                              uint32_t* __restrict__ counter, const drl_ctrl_t* __restrict__ ctrl) {
    if (ctrl != nullptr) epoch_ctr += ctrl->epoch_ctr;
    const uint4 k0 = philox_seeded(seed, epoch_ctr, rank, 0u, TAG_PERM);
    const uint4 k1 = philox_seeded(seed, epoch_ctr, rank, 1u, TAG_PERM);
    const uint32_t keys[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
    double sm_[STATS_MAX_MB], ss_[STATS_MAX_MB];
#pragma unroll
    for (int m = 0; m < STATS_MAX_MB; ++m) { sm_[m] = 0.0; ss_[m] = 0.0; }
    const uint32_t stride = gridDim.x * blockDim.x;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < B; s += stride) {
        const double v = (double)__ldg(adv + s);
        const uint32_t mb = perm_position(s, B, a, b, keys) / mb_size;
#pragma unroll
        for (int m = 0; m < STATS_MAX_MB; ++m)
            if (mb == (uint32_t)m) { sm_[m] += v; ss_[m] = fma(v, v, ss_[m]); }
    }
    __shared__ double sh[2][STATS_MAX_MB][8];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int m = 0; m < STATS_MAX_MB; ++m) {
        const double t0 = warp_sum(sm_[m]), t1 = warp_sum(ss_[m]);
        if (lane == 0) { sh[0][m][warp] = t0; sh[1][m][warp] = t1; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * STATS_MAX_MB) {
        const int which = threadIdx.x / STATS_MAX_MB, m = threadIdx.x % STATS_MAX_MB;
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[which][m][w];
        partials[((size_t)m * STAT_PARTS + blockIdx.x) * 2 + which] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t done = atomicAdd(counter, 1u);
        is_last = (done == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {   // one warp per minibatch folds the partials (lane-strided, then a fixed shuffle tree)
        __threadfence();
        const uint32_t k = threadIdx.x >> 5;
        if (k < nmb) {
            double ts = 0.0, tss = 0.0;
            for (uint32_t p = threadIdx.x & 31; p < gridDim.x; p += 32) {
                ts += __ldcg(&partials[((size_t)k * STAT_PARTS + p) * 2 + 0]);
                tss += __ldcg(&partials[((size_t)k * STAT_PARTS + p) * 2 + 1]);
            }
            ts = warp_sum(ts);
            tss = warp_sum(tss);
            if ((threadIdx.x & 31) == 0) {
                const uint32_t klo = k * mb_size, khi = min(B, klo + mb_size);
                const double cnt = (double)(khi - klo);
                const double mean = ts / cnt;
                const double var = (tss - ts * mean) / (cnt - 1.0);
                stats_out[2 * k + 0] = (float)mean;
                stats_out[2 * k + 1] = (float)sqrt(var > 0.0 ? var : 0.0);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *counter = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// Explained variance (ppo.py:194-195): 1 - Var(values - returns) / Var(values), unbiased variances over all n slots.
// Each CTA sums its grid-strided slice in fp64, the last CTA folds the partials in a fixed order.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) explained_variance_kernel(const float* __restrict__ values, const float* __restrict__ returns,
                                                                  int64_t n, float* __restrict__ out, double* __restrict__ partials,
                                                                  uint32_t* __restrict__ counter) {
    double acc[4] = {0.0, 0.0, 0.0, 0.0};    // sum y, sum y^2, sum d, sum d^2 with d = y - r (fp32 subtraction like torch)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const float yf = __ldg(values + i), rf = __ldg(returns + i);
        const double y = (double)yf, d = (double)(yf - rf);
        acc[0] += y; acc[1] = fma(y, y, acc[1]); acc[2] += d; acc[3] = fma(d, d, acc[3]);
    }
    __shared__ double sh[4][8];
    __shared__ bool is_last;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const double t = warp_sum(acc[q]);
        if (lane == 0) sh[q][warp] = t;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sh[threadIdx.x][w];
        partials[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    __syncthreads();
    if (is_last && threadIdx.x < 32) {
        __threadfence();
        double t[4] = {0.0, 0.0, 0.0, 0.0};
        for (uint32_t p = threadIdx.x; p < gridDim.x; p += 32)
#pragma unroll
            for (int q = 0; q < 4; ++q) t[q] += __ldcg(&partials[(size_t)p * 4 + q]);
#pragma unroll
        for (int q = 0; q < 4; ++q) t[q] = warp_sum(t[q]);
        if (threadIdx.x == 0) {
            const double cnt = (double)n;
            const double var_y = (t[1] - t[0] * (t[0] / cnt)) / (cnt - 1.0);
            const double var_d = (t[3] - t[2] * (t[2] / cnt)) / (cnt - 1.0);
            out[0] = var_y == 0.0 ? __int_as_float(0x7FC00000) : (float)(1.0 - var_d / var_y);
            *counter = 0;
        }
    }
}

}  // namespace drl

using namespace drl;

extern "C" {

int drl_gae(const drl_rollout_buf_t* buf, const drl_net_t* net, int32_t T, int32_t N, float gamma, float gae_lambda,
            float* adv_out, float* ret_out, float* rec_out, void* stream) {
    int rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(buf && buf->rew && buf->done && buf->val && adv_out && ret_out, "drl_gae: NULL pointer");
    DRL_REQUIRE(T > 0 && N > 0, "drl_gae: T=%d N=%d", T, N);
    DRL_REQUIRE(rec_out == nullptr || (buf->obs && buf->act && buf->logp), "drl_gae: record packing needs obs/act/logp");
    const int blocks = (N + 127) / 128;
    cudaStream_t st = as_stream(stream);
    if (net->obs_stride == 4)
        gae_kernel<4, 8><<<blocks, 128, 0, st>>>(buf->rew, buf->done, buf->val, buf->obs, buf->act, buf->logp, T, N, gamma,
                                                gae_lambda, adv_out, ret_out, rec_out);
    else
        gae_kernel<8, 16><<<blocks, 128, 0, st>>>(buf->rew, buf->done, buf->val, buf->obs, buf->act, buf->logp, T, N, gamma,
                                                 gae_lambda, adv_out, ret_out, rec_out);
    DRL_LAUNCH_CHECK("gae_kernel");
    return DRL_OK;
}

int drl_explained_variance(const float* values, const float* returns, int64_t n, float* out, void* workspace,
                           size_t workspace_bytes, void* stream) {
    DRL_REQUIRE(values && returns && out && workspace, "drl_explained_variance: NULL pointer");
    DRL_REQUIRE(n >= 2, "drl_explained_variance: n=%lld", (long long)n);
    const WorkspaceLayout w = workspace_layout(0);        // the regions used here do not depend on the parameter count
    DRL_REQUIRE(workspace_bytes >= w.total, "drl_explained_variance: workspace %zu < %zu bytes", workspace_bytes, w.total);
    uint32_t* counter = reinterpret_cast<uint32_t*>((char*)workspace + w.counters) + 12;
    double* partials = reinterpret_cast<double*>((char*)workspace + w.ev_partials);
    int64_t blocks = (n + 4095) / 4096;
    if (blocks > STAT_PARTS) blocks = STAT_PARTS;
    explained_variance_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(values, returns, n, out, partials, counter);
    DRL_LAUNCH_CHECK("explained_variance_kernel");
    return DRL_OK;
}

static int permutation_impl(uint32_t* idx_out, uint32_t B, uint64_t seed, uint32_t epoch_ctr, uint32_t rank, void* stream,
                            const drl_ctrl_t* ctrl) {
    DRL_REQUIRE(idx_out, "drl_permutation: idx_out is NULL");
    DRL_REQUIRE(B > 0 && B <= 0x80000000u, "drl_permutation: B=%u out of range", B);
    DRL_REQUIRE(rank < (1u << 24), "drl_permutation: rank=%u out of range", rank);
    uint32_t k = 2;
    while (k < 32 && (1ull << k) < (unsigned long long)B) ++k;
    const uint32_t a = k / 2, b = k - a;
    long long blocks = ((long long)B + 2047) / 2048;          // ~8 indices per thread amortise the two Philox blocks
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    permutation_kernel<<<(int)blocks, 256, 0, as_stream(stream)>>>(idx_out, B, a, b, seed, epoch_ctr, rank, ctrl);
    DRL_LAUNCH_CHECK("permutation_kernel");
    return DRL_OK;
}
int drl_permutation(uint32_t* idx_out, uint32_t B, uint64_t seed, uint32_t epoch_ctr, uint32_t rank, void* stream) {
    return permutation_impl(idx_out, B, seed, epoch_ctr, rank, stream, nullptr);
}
int drl_permutation_ctl(uint32_t* idx_out, uint32_t B, uint64_t seed, const drl_ctrl_t* ctrl, uint32_t epoch_off, uint32_t rank,
                        void* stream) {
    DRL_REQUIRE(ctrl != nullptr, "drl_permutation_ctl: ctrl is NULL");
    return permutation_impl(idx_out, B, seed, epoch_off, rank, stream, ctrl);
}

int drl_adv_stats(const drl_net_t* net, const float* rec, const uint32_t* idx, uint32_t B, uint32_t mb_size,
                  float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(rec && stats_out && workspace, "drl_adv_stats: NULL pointer");
    DRL_REQUIRE(B > 0 && mb_size > 0, "drl_adv_stats: B=%u mb_size=%u", B, mb_size);
    const uint32_t nmb = (B + mb_size - 1) / mb_size;
    DRL_REQUIRE(nmb <= (uint32_t)MAX_MINIBATCHES, "drl_adv_stats: %u minibatches > %d", nmb, MAX_MINIBATCHES);
    const WorkspaceLayout w = workspace_layout(drl_param_count(net));
    DRL_REQUIRE(workspace_bytes >= w.total, "drl_adv_stats: workspace %zu < %zu bytes", workspace_bytes, w.total);
    uint32_t* counter = reinterpret_cast<uint32_t*>((char*)workspace + w.counters);
    double* partials = reinterpret_cast<double*>((char*)workspace + w.stat_partials);
    uint32_t nparts = (mb_size + 4095u) / 4096u;               // >= 16 gathers per thread, up to STAT_PARTS CTAs per minibatch
    if (nparts > (uint32_t)STAT_PARTS) nparts = STAT_PARTS;
    if (nparts == 0) nparts = 1;
    dim3 grid(nparts, nmb);
    if (net->obs_dim <= 4) adv_stats_kernel<8><<<grid, 256, 0, as_stream(stream)>>>(rec, idx, B, mb_size, stats_out, partials, counter);
    else adv_stats_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(rec, idx, B, mb_size, stats_out, partials, counter);
    DRL_LAUNCH_CHECK("adv_stats_kernel");
    return DRL_OK;
}

static int adv_stats_perm_impl(const drl_net_t* net, const float* adv, uint32_t B, uint32_t mb_size, uint64_t seed, uint32_t epoch_ctr,
                               uint32_t rank, float* stats_out, void* workspace, size_t workspace_bytes, void* stream,
                               const drl_ctrl_t* ctrl) {
    int rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(adv && stats_out && workspace, "drl_adv_stats_perm: NULL pointer");
    DRL_REQUIRE(B > 0 && mb_size > 0 && B <= 0x80000000u, "drl_adv_stats_perm: B=%u mb_size=%u", B, mb_size);
    const uint32_t nmb = (B + mb_size - 1) / mb_size;
    if (nmb > (uint32_t)STATS_MAX_MB) { set_error("drl_adv_stats_perm: %u minibatches > %d (use drl_adv_stats)", nmb, STATS_MAX_MB); return DRL_ERR_UNSUPPORTED; }
    const WorkspaceLayout w = workspace_layout(drl_param_count(net));
    DRL_REQUIRE(workspace_bytes >= w.total, "drl_adv_stats_perm: workspace %zu < %zu bytes", workspace_bytes, w.total);
    uint32_t* counter = reinterpret_cast<uint32_t*>((char*)workspace + w.counters);
    double* partials = reinterpret_cast<double*>((char*)workspace + w.stat_partials);
    uint32_t k = 2;
    while (k < 32 && (1ull << k) < (unsigned long long)B) ++k;
    uint32_t blocks = (B + 2047u) / 2048u;
    if (blocks > (uint32_t)STAT_PARTS) blocks = STAT_PARTS;
    adv_stats_perm_kernel<<<blocks, 256, 0, as_stream(stream)>>>(adv, B, mb_size, nmb, k / 2, k - k / 2, seed, epoch_ctr, rank,
                                                                 stats_out, partials, counter, ctrl);
    DRL_LAUNCH_CHECK("adv_stats_perm_kernel");
    return DRL_OK;
}
int drl_adv_stats_perm(const drl_net_t* net, const float* adv, uint32_t B, uint32_t mb_size, uint64_t seed, uint32_t epoch_ctr,
                       uint32_t rank, float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    return adv_stats_perm_impl(net, adv, B, mb_size, seed, epoch_ctr, rank, stats_out, workspace, workspace_bytes, stream, nullptr);
}
int drl_adv_stats_perm_ctl(const drl_net_t* net, const float* adv, uint32_t B, uint32_t mb_size, uint64_t seed, const drl_ctrl_t* ctrl,
                           uint32_t epoch_off, uint32_t rank, float* stats_out, void* workspace, size_t workspace_bytes, void* stream) {
    DRL_REQUIRE(ctrl != nullptr, "drl_adv_stats_perm_ctl: ctrl is NULL");
    return adv_stats_perm_impl(net, adv, B, mb_size, seed, epoch_off, rank, stats_out, workspace, workspace_bytes, stream, ctrl);
}

}  // extern "C"
