// rollout256.cu -- fused rollout (deep_rl/ppo.py:110-141) for the 256-wide actor-critic.
// One CTA = 128 environments for all T steps; the ACTOR's W2 (128 KB bf16) stays in shared memory.  Per step:
//   S0   owner threads (one per env, float64 state in registers): observation -> obs[t] in HBM and the [obs_hi|1|obs_lo]
//        operand tile; layer 1 is the same K = 16 GEMM as in the update (identical operands => identical z1)
//   P0   all 512 compute threads: h1 = tanh(z1) -> four bf16 K-chunks; the layer-2 GEMM (M128 N256 K256) is issued chunk by
//        chunk behind them
//   P1   h2 = tanh(z2 + b2), partial head sums (64 units per thread), exchange among the four threads of a row
//   S3   owners: logits, Philox inverse-CDF sample, log-prob, fp64 env step with auto-reset and episode statistics, stores
// The critic is not needed to advance the environments: values V(obs[t]) for all (T+1) x N stored observations are computed
// afterwards by one batched forward pass of the critic (mlp256_kernel in forward mode, update256.cu) -- a plain GEMM-shaped
// launch instead of a second 128 KB weight image per step.
#include "drl_env.cuh"
#include "drl_h256.cuh"
#include "drl_tc_common.cuh"

namespace drl {

int check_env(const drl_env_t* env);
drl_ep_log_t log_or_empty(const drl_ep_log_t* log);

namespace h256 {

int launch_forward256(const drl_net_t* net, const float* packed, const float* obs, int64_t n, float* logits, float* value,
                      int only_net, cudaStream_t st);

enum : uint32_t { RB_H1 = 1, RB_L1 = 5, RB_QUAD = 6 };
constexpr uint32_t RB_ALL = TC_COMPUTE + 32;

template <int O, int A>
struct RoSmem256 {
    static constexpr int OFF_W2 = 0;
    static constexpr int OFF_ACT = OFF_W2 + HH * HH * 2;
    static constexpr int OFF_W1B = OFF_ACT + TILE_BYTES;
    static constexpr int OFF_B2 = OFF_W1B + 2 * HH * 8 * 2;
    static constexpr int OFF_W4 = OFF_B2 + HH * 4;
    static constexpr int OFF_B4 = OFF_W4 + 3 * HH * 4;
    static constexpr int OFF_OBS = OFF_B4 + 128;
    static constexpr int OFF_XCH = OFF_OBS + 4096;
    static constexpr int OFF_BAR = OFF_XCH + 4 * 3 * 128 * 4;
    static constexpr int TOTAL = OFF_BAR + 64 + 1024;
    static_assert(TOTAL <= 232448, "shared memory");
};

template <int KIND>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout256_kernel(drl_env_t env, const float* __restrict__ packed, int T, uint64_t step0,
                                                                   drl_rollout_buf_t buf, drl_ep_log_t log,
                                                                   const drl_ctrl_t* __restrict__ ctrl) {
    if (ctrl != nullptr) step0 = ctrl->env_step;
    using SP = EnvSpec<KIND>;
    constexpr int O = SP::O, A = SP::A, OP = SP::OP;
    using P = Packed256<O, A>;
    using S = RoSmem256<O, A>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* tW2 = sm + S::OFF_W2;
    unsigned char* tACT = sm + S::OFF_ACT;
    unsigned char* tW1B = sm + S::OFF_W1B;
    const float* sB2 = reinterpret_cast<const float*>(sm + S::OFF_B2);
    const float* sW4 = reinterpret_cast<const float*>(sm + S::OFF_W4);
    const float* sB4 = reinterpret_cast<const float*>(sm + S::OFF_B4);
    unsigned char* tOBS = sm + S::OFF_OBS;
    float* xch = reinterpret_cast<float*>(sm + S::OFF_XCH);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);   // 0 weights, 1 layer 1, 2 layer 2
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == TC_COMPUTE / 32;

    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1); mbar_init(bars + 2, 1);
        mbar_fence_init();
    }
    if (is_mma_warp) umma::tmem_alloc(slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        mbar_expect_tx(bars, (uint32_t)(HH * HH * 2 + 2 * HH * 8 * 2 + HH * 4 + A * HH * 4 + 16));
        const char* w2 = reinterpret_cast<const char*>(packed + P::W2);
        for (uint32_t off = 0; off < (uint32_t)(HH * HH * 2); off += 32768u) bulk_g2s(tW2 + off, w2 + off, 32768u, bars);
        bulk_g2s(tW1B, packed + P::W1B, 2 * HH * 8 * 2, bars);
        bulk_g2s(sm + S::OFF_B2, packed + P::B2, HH * 4, bars);
        bulk_g2s(sm + S::OFF_W4, packed + P::W4, A * HH * 4, bars);
        bulk_g2s(sm + S::OFF_B4, packed + P::B4, 16u, bars);
    }
    const uint32_t tmem = *slot;
    mbar_wait(bars, 0);
    constexpr uint32_t C_Z1 = 0, C_Z2 = 256;

    if (is_mma_warp) {
        const uint32_t aW2 = smem_u32(tW2), aACT = smem_u32(tACT), aW1B = smem_u32(tW1B), aOBS = smem_u32(tOBS);
        constexpr uint32_t ID_L1 = umma::make_idesc(128, 256, false, false);
        constexpr uint32_t ID_FWD = umma::make_idesc(128, 256, false, false);
        for (int t = 0; t < T; ++t) {
            named_bar_sync(RB_L1, RB_ALL);
            umma::fence_after_sync();
            if (umma::elect_one()) {
                umma::mma(tmem + C_Z1, umma::make_desc(aOBS, 2048, 128, umma::LAYOUT_NONE),
                          umma::make_desc(aW1B, 4096, 128, umma::LAYOUT_NONE), ID_L1, 0u);
                umma::commit(bars + 1);
            }
            __syncwarp();
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                named_bar_sync(RB_H1 + c, RB_ALL);
                umma::fence_after_sync();
                if (umma::elect_one()) {
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        umma::mma(tmem + C_Z2, umma::make_desc(aACT + c * SLOT_BYTES + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aW2 + c * 32768 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, (c > 0 || kb > 0) ? 1u : 0u);
                    if (c == NCH - 1) umma::commit(bars + 2);
                }
                __syncwarp();
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::tmem_dealloc(tmem, 512);
        return;
    }

    const int q = warp & 3, j = warp >> 2;
    const int r = q * 32 + lane;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
    const bool owner_warp = j == 0;
    const int N = env.num_envs;
    const int n = blockIdx.x * TC_TILE + r;
    const bool own = owner_warp && n < N;
    const uint32_t gid = env.env_gid0 + (uint32_t)n;
    const uint32_t off0 = umma::sw128_off(r, 2 * j), off1 = umma::sw128_off(r, 2 * j + 1);

    EnvLane e;
    e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0.0; e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    if (own) env_load(e, env, n);

    for (int t = 0; t <= T; ++t) {
        // ---- S0: observation of the current state -> HBM; for t < T also the layer-1 operand tile ----
        if (owner_warp) {
            float obs[OP];
#pragma unroll
            for (int i = 0; i < OP; ++i) obs[i] = 0.0f;
            if (own) {
                env_observation<KIND>(e.s, obs);
                float4* o4 = reinterpret_cast<float4*>(buf.obs + ((size_t)t * N + n) * OP);
#pragma unroll
                for (int qq = 0; qq < OP / 4; ++qq) o4[qq] = make_float4(obs[4 * qq], obs[4 * qq + 1], obs[4 * qq + 2], obs[4 * qq + 3]);
            }
            if (t < T) {
                float o16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o16[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(obs[i]));
                    o16[i] = hi;
                    o16[8 + i] = obs[i] - hi;
                }
                o16[O] = 1.0f;
                umma::store_row_ns16(tOBS, TC_TILE, r, o16);
                umma::fence_proxy_async();
            }
        }
        if (t == T) break;            // the last observation needs no action (its value comes from the critic pass)
        umma::fence_before_sync();
        named_bar_arrive(RB_L1, RB_ALL);

        // ---- P0: h1 chunks ----
        mbar_wait(bars + 1, (uint32_t)t & 1u);
        umma::fence_after_sync();
#pragma unroll 1
        for (int c = 0; c < NCH; ++c) {
            float z[16];
            umma::ld16(trow + C_Z1 + 64 * c + 16 * j, z);
#pragma unroll
            for (int x = 0; x < 16; ++x) z[x] = tanh_mufu(z[x]);
            uint4 q0, q1;
            q0.x = umma::pack_bf16(z[0], z[1]); q0.y = umma::pack_bf16(z[2], z[3]); q0.z = umma::pack_bf16(z[4], z[5]); q0.w = umma::pack_bf16(z[6], z[7]);
            q1.x = umma::pack_bf16(z[8], z[9]); q1.y = umma::pack_bf16(z[10], z[11]); q1.z = umma::pack_bf16(z[12], z[13]); q1.w = umma::pack_bf16(z[14], z[15]);
            *reinterpret_cast<uint4*>(tACT + c * SLOT_BYTES + off0) = q0;
            *reinterpret_cast<uint4*>(tACT + c * SLOT_BYTES + off1) = q1;
            umma::fence_proxy_async();
            umma::fence_before_sync();
            named_bar_arrive(RB_H1 + c, RB_ALL);
        }

        // ---- P1: layer-2 epilogue and partial head sums ----
        mbar_wait(bars + 2, (uint32_t)t & 1u);
        umma::fence_after_sync();
        {
            float ps[A];
#pragma unroll
            for (int a = 0; a < A; ++a) ps[a] = 0.0f;
#pragma unroll 1
            for (int c = 0; c < NCH; ++c) {
                float z[16];
                umma::ld16(trow + C_Z2 + 64 * c + 16 * j, z);
#pragma unroll
                for (int e4 = 0; e4 < 4; ++e4) {
                    const float4 bb = *reinterpret_cast<const float4*>(sB2 + 64 * c + 16 * j + 4 * e4);
                    z[4 * e4 + 0] = tanh_mufu(z[4 * e4 + 0] + bb.x);
                    z[4 * e4 + 1] = tanh_mufu(z[4 * e4 + 1] + bb.y);
                    z[4 * e4 + 2] = tanh_mufu(z[4 * e4 + 2] + bb.z);
                    z[4 * e4 + 3] = tanh_mufu(z[4 * e4 + 3] + bb.w);
                }
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    const float* w = sW4 + a * HH + 64 * c + 16 * j;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int e4 = 0; e4 < 4; ++e4) {
                        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * e4);
                        s0 = fmaf(z[4 * e4 + 0], ww.x, s0);
                        s1 = fmaf(z[4 * e4 + 1], ww.y, s1);
                        s2 = fmaf(z[4 * e4 + 2], ww.z, s2);
                        s3 = fmaf(z[4 * e4 + 3], ww.w, s3);
                    }
                    ps[a] += (s0 + s1) + (s2 + s3);
                }
            }
#pragma unroll
            for (int a = 0; a < A; ++a) xch[(j * 3 + a) * TC_TILE + r] = ps[a];
        }
        umma::fence_before_sync();
        named_bar_sync(RB_QUAD + q, 128);

        // ---- S3: owners: sample, env step ----
        if (own) {
            const size_t i0 = (size_t)t * N + n;
            float l[A];
#pragma unroll
            for (int a = 0; a < A; ++a)
                l[a] = ((xch[(0 * 3 + a) * TC_TILE + r] + xch[(1 * 3 + a) * TC_TILE + r]) +
                        (xch[(2 * 3 + a) * TC_TILE + r] + xch[(3 * 3 + a) * TC_TILE + r])) + sB4[a];
            if (buf.logits != nullptr) {
#pragma unroll
                for (int a = 0; a < A; ++a) buf.logits[i0 * A + a] = l[a];
            }
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
        named_bar_sync(RB_QUAD + q, 128);      // the exchange buffer is rewritten in the next step
    }
    if (own) env_store(e, env, n);
    umma::fence_before_sync();
    __syncthreads();
}

template <int KIND>
static int launch_rollout256_kind(const drl_env_t& env, const drl_net_t* net, const float* packed, int T, uint64_t step0,
                                  const drl_rollout_buf_t& buf, const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    using SP = EnvSpec<KIND>;
    const int smem = RoSmem256<SP::O, SP::A>::TOTAL;
    DRL_CUDA(cudaFuncSetAttribute(rollout256_kernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int blocks = (env.num_envs + TC_TILE - 1) / TC_TILE;
    rollout256_kernel<KIND><<<blocks, TC_THREADS, smem, st>>>(env, packed, T, step0, buf, log, ctrl);
    DRL_LAUNCH_CHECK("rollout256_kernel");
    // values of every stored observation, V(obs[t]) for t = 0..T: one batched forward pass of the critic
    return launch_forward256(net, packed, buf.obs, (int64_t)(T + 1) * env.num_envs, nullptr, buf.val, 1, st);
}

int launch_rollout256(const drl_env_t& env, const drl_net_t* net, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                      const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    if (env.kind == DRL_ENV_CARTPOLE) return launch_rollout256_kind<DRL_ENV_CARTPOLE>(env, net, packed, T, step0, buf, log, st, ctrl);
    if (env.kind == DRL_ENV_MOUNTAINCAR) return launch_rollout256_kind<DRL_ENV_MOUNTAINCAR>(env, net, packed, T, step0, buf, log, st, ctrl);
    return launch_rollout256_kind<DRL_ENV_ACROBOT>(env, net, packed, T, step0, buf, log, st, ctrl);
}

}  // namespace h256
}  // namespace drl
