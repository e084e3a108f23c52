// rollout_tc.cu -- tensor-core variant of the fused rollout (deep_rl/ppo.py:110-141) for large env counts.
// Same inputs, outputs, Philox streams and fp64 env physics as rollout_kernel (rollout.cu); the two 64x64
// hidden layers run on tcgen05 (bf16 operands, fp32 TMEM accumulators) with the same numerics as the tensor-core
// update (update_tc.cu), so log-probs recorded here and re-evaluated there agree to bf16 rounding.
//
// One CTA = 128 environments (the 128 TMEM lanes) for all T steps; 16 compute warps + 1 MMA-issuer warp.
// Compute thread (warp w, lane l): env row r = 32*(w&3)+l, net = w>>3, hidden units [32*((w>>2)&1), +32).
// Warps 0-3 (actor, first half) additionally own the env state (float64 registers).  Per step:
//   S0  owners: observation from state -> obs[t] in HBM and the fp32 obs tile in shared memory
//   --  row-window barrier (the 4 warps that share 32 rows)
//   S1  all: layer 1 on CUDA cores for 32 units -> bf16 SW128 tile; hand the forward GEMM to the issuer warp
//   S2  all: wait, tcgen05.ld z2, tanh, partial head dot products -> exchange buffer
//   --  row-window barrier
//   S3  owners: logits / value, value store, Philox inverse-CDF sample, log-prob, fp64 env step with auto-reset
//       and episode statistics, reward / done stores
#include "drl_env.cuh"
#include "drl_pack.cuh"
#include "drl_tc_common.cuh"

namespace drl {

int check_env(const drl_env_t* env);
drl_ep_log_t log_or_empty(const drl_ep_log_t* log);

enum : uint32_t { RB_FWD = 1, RB_ROW0 = 2 };   // named barriers: issuer hand-off, 4 row-window barriers (2..5)

template <int O, int A>
struct RoTcSmem {
    using P = Packed<O, A>;
    static constexpr int W_BYTES = (P::TC_END - P::TC_W2) * 4;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_H1 = (OFF_W + W_BYTES + 1023) / 1024 * 1024;   // (actor, critic) x 16 KB
    static constexpr int OFF_OBS = OFF_H1 + 32768;                           // fp32 [128][OW]
    static constexpr int OFF_XCH = OFF_OBS + TC_TILE * P::OW * 4;            // fp32 [8][128] head partial sums
    static constexpr int OFF_BAR = OFF_XCH + 8 * TC_TILE * 4;
    static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(drl_env_t env, const float* __restrict__ packed, int T,
                                                                   uint64_t step0, drl_rollout_buf_t buf, drl_ep_log_t log,
                                                                   int rows) {
    using SP = EnvSpec<KIND>;
    constexpr int O = SP::O, A = SP::A, OP = SP::OP;
    using P = Packed<O, A>;
    using S = RoTcSmem<O, A>;
    constexpr int OW = P::OW;
    static_assert(OW == OP, "obs stride");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* tW2 = sm + S::OFF_W;
    const float* sW1 = reinterpret_cast<const float*>(sm + S::OFF_W + 2 * H * H * 2);
    const float* sB1 = sW1 + 2 * H * OW;
    const float* sB2 = sB1 + 2 * H;
    const float* sW4 = sB2 + 2 * H;
    const float* sB4 = sW4 + (A + 1) * H;
    unsigned char* tH1 = sm + S::OFF_H1;
    float* obs_s = reinterpret_cast<float*>(sm + S::OFF_OBS);
    float* xch = reinterpret_cast<float*>(sm + S::OFF_XCH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);   // 0 weights, 1 fwd
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == TC_COMPUTE / 32;

    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_fence_init();
    }
    if (is_mma_warp) umma::tmem_alloc(slot, 128);      // the issuer warp always stays: it also frees the columns
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        mbar_expect_tx(bars, (uint32_t)S::W_BYTES);
        const char* src = reinterpret_cast<const char*>(packed + P::TC_W2);
        for (uint32_t off = 0; off < (uint32_t)S::W_BYTES; off += 16384u) {
            const uint32_t n = (uint32_t)S::W_BYTES - off < 16384u ? (uint32_t)S::W_BYTES - off : 16384u;
            bulk_g2s(sm + S::OFF_W + off, src + off, n, bars);
        }
    }
    const uint32_t tmem = *slot;
    mbar_wait(bars, 0);
    // `rows` (32, 64 or 128) environments per CTA: with few environments the per-step latency, not the throughput, sets
    // the pace, so they are spread over more CTAs and the warps of the unused row windows leave.
    const uint32_t nthr = (uint32_t)(rows / 32) * 4u * 32u + 32u;      // participants of the issuer hand-off barrier
    if (!is_mma_warp && (warp & 3) * 32 >= rows) return;

    if (is_mma_warp) {
        const uint32_t aW2 = smem_u32(tW2), aH1 = smem_u32(tH1);
        constexpr uint32_t ID_FWD = umma::make_idesc(128, 64, false, false);
        for (int t = 0; t <= T; ++t) {
            named_bar_sync(RB_FWD, nthr);
            umma::fence_after_sync();
            if (umma::elect_one()) {
#pragma unroll
                for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        umma::mma(tmem + n2 * 64, umma::make_desc(aH1 + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aW2 + n2 * 8192 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, kb > 0);
                umma::commit(bars + 1);
            }
            __syncwarp();
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::tmem_dealloc(tmem, 128);
        return;
    }

    // =========================== compute warps ===========================
    const int rw = warp & 3, half = (warp >> 2) & 1, net = (warp >> 3) & 1;
    const int r = rw * 32 + lane;
    const int u0 = half * HU;
    const uint32_t trow = tmem + ((uint32_t)(rw * 32) << 16);
    const int wrow0 = net == 0 ? 0 : A;
    const int nheads = net == 0 ? A : 1;
    const bool owner_warp = warp < 4;                         // actor, first half: owns the env state
    const int N = env.num_envs;
    const int n = blockIdx.x * rows + r;
    const bool own = owner_warp && n < N;
    const uint32_t gid = env.env_gid0 + (uint32_t)n;

    EnvLane e;
    e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0.0; e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    if (own) env_load(e, env, n);

    for (int t = 0; t <= T; ++t) {
        // ---- S0: observation of the current state ----
        if (owner_warp) {
            float obs[OP];
#pragma unroll
            for (int i = 0; i < OP; ++i) obs[i] = 0.0f;
            if (own) {
                env_observation<KIND>(e.s, obs);
                float4* o4 = reinterpret_cast<float4*>(buf.obs + ((size_t)t * N + n) * OP);
#pragma unroll
                for (int q = 0; q < OP / 4; ++q) o4[q] = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
            }
#pragma unroll
            for (int q = 0; q < OP / 4; ++q)
                *reinterpret_cast<float4*>(obs_s + r * OW + 4 * q) = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
        }
        named_bar_sync(RB_ROW0 + rw, 128);

        // ---- S1: layer 1 (32 units of this thread's net) -> bf16 tile, hand the forward GEMM ----
        {
            float x[OW];
#pragma unroll
            for (int q = 0; q < OW / 4; ++q) {
                const float4 v4 = *reinterpret_cast<const float4*>(obs_s + r * OW + 4 * q);
                x[4 * q] = v4.x; x[4 * q + 1] = v4.y; x[4 * q + 2] = v4.z; x[4 * q + 3] = v4.w;
            }
            float h[HU];
            const float* w1 = sW1 + (net * H + u0) * OW;
            const float* b1 = sB1 + net * H + u0;
#pragma unroll
            for (int k4 = 0; k4 < HU / 4; ++k4) {
                const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * k4);
                const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = 4 * k4 + j;
                    float z = bv[j];
#pragma unroll
                    for (int q = 0; q < OW / 4; ++q) {
                        const float4 w = *reinterpret_cast<const float4*>(w1 + kk * OW + 4 * q);
                        z = fmaf(x[4 * q + 3], w.w, fmaf(x[4 * q + 2], w.z, fmaf(x[4 * q + 1], w.y, fmaf(x[4 * q], w.x, z))));
                    }
                    h[kk] = tanh_mufu(z);
                }
            }
            store_half_row_sw128(tH1 + net * 16384, r, half * 4, h);
        }
        umma::fence_proxy_async();
        umma::fence_before_sync();
        named_bar_arrive(RB_FWD, nthr);

        // ---- S2: layer-2 epilogue and partial head dot products ----
        mbar_wait(bars + 1, (uint32_t)t & 1u);
        umma::fence_after_sync();
        {
            float h[HU];
            umma::ld32(trow + net * 64 + u0, h);
            const float* b2 = sB2 + net * H + u0;
#pragma unroll
            for (int k4 = 0; k4 < HU / 4; ++k4) {
                const float4 bb = *reinterpret_cast<const float4*>(b2 + 4 * k4);
                h[4 * k4 + 0] = tanh_mufu(h[4 * k4 + 0] + bb.x);
                h[4 * k4 + 1] = tanh_mufu(h[4 * k4 + 1] + bb.y);
                h[4 * k4 + 2] = tanh_mufu(h[4 * k4 + 2] + bb.z);
                h[4 * k4 + 3] = tanh_mufu(h[4 * k4 + 3] + bb.w);
            }
#pragma unroll
            for (int a = 0; a < A; ++a) {
                if (a < nheads) {
                    const float* w = sW4 + (wrow0 + a) * H + u0;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int k4 = 0; k4 < HU / 4; ++k4) {
                        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * k4);
                        s0 = fmaf(h[4 * k4 + 0], ww.x, s0);
                        s1 = fmaf(h[4 * k4 + 1], ww.y, s1);
                        s2 = fmaf(h[4 * k4 + 2], ww.z, s2);
                        s3 = fmaf(h[4 * k4 + 3], ww.w, s3);
                    }
                    // slots: actor half h -> h*A + a (a < A <= 3), critic half h -> 6 + h
                    const int sl = net == 0 ? half * A + a : 6 + half;
                    xch[sl * TC_TILE + r] = (s0 + s1) + (s2 + s3);
                }
            }
        }
        umma::fence_before_sync();
        named_bar_sync(RB_ROW0 + rw, 128);

        // ---- S3: owners: value store, sample, env step ----
        if (own) {
            const size_t i0 = (size_t)t * N + n;
            buf.val[i0] = (xch[6 * TC_TILE + r] + xch[7 * TC_TILE + r]) + sB4[A];
            if (t < T) {
                float l[A];
#pragma unroll
                for (int a = 0; a < A; ++a) l[a] = (xch[a * TC_TILE + r] + xch[(A + a) * TC_TILE + r]) + sB4[a];
                const uint64_t step = step0 + (uint64_t)t;
                const uint4 rr = philox_seeded(env.seed, gid, (uint32_t)step, (uint32_t)(step >> 32), TAG_ACTION);
                float lp;
                const int act = sample_categorical<A>(l, u01_f32(rr.x), lp);
                buf.act[i0] = (uint8_t)act;
                buf.logp[i0] = lp;
                float reward;
                const bool done = env_step<KIND>(e, act, reward, env.seed, gid, step, env.max_episode_steps, log);
                buf.rew[i0 + N] = reward;
                buf.done[i0 + N] = done ? 1 : 0;
            }
        }
    }
    if (own) env_store(e, env, n);
    umma::fence_before_sync();
    __syncthreads();
}

template <int KIND>
static int launch_rollout_tc_kind(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                                  const drl_ep_log_t& log, cudaStream_t st) {
    using SP = EnvSpec<KIND>;
    const int smem = RoTcSmem<SP::O, SP::A>::TOTAL;
    DRL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    // envs per CTA: 128 when that fills the GPU, otherwise fewer rows per CTA over more CTAs (latency-bound regime)
    const int sms = sm_count();
    int rows = 128;
    if ((env.num_envs + 127) / 128 < sms) rows = (env.num_envs + 63) / 64 <= sms ? 64 : 128;
    if ((env.num_envs + 63) / 64 < sms && (env.num_envs + 31) / 32 <= sms) rows = 32;
    const char* ov = getenv("DRL_ROLLOUT_ROWS");
    if (ov) { const int v = atoi(ov); if (v == 32 || v == 64 || v == 128) rows = v; }
    const int blocks = (env.num_envs + rows - 1) / rows;
    rollout_tc_kernel<KIND><<<blocks, TC_THREADS, smem, st>>>(env, packed, T, step0, buf, log, rows);
    DRL_LAUNCH_CHECK("rollout_tc_kernel");
    return DRL_OK;
}

int launch_rollout_tc(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                      const drl_ep_log_t& log, cudaStream_t st) {
    if (env.kind == DRL_ENV_CARTPOLE) return launch_rollout_tc_kind<DRL_ENV_CARTPOLE>(env, packed, T, step0, buf, log, st);
    return launch_rollout_tc_kind<DRL_ENV_ACROBOT>(env, packed, T, step0, buf, log, st);
}

}  // namespace drl
