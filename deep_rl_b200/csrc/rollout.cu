// rollout.cu -- the fused rollout kernel: T x (actor + critic forward, Philox categorical sample,
// log-prob, env step with auto-reset + episode statistics, SoA stores), deep_rl/ppo.py:110-141.
//
// One launch covers the whole rollout: a warp owns 8*SUB environments for all T steps, so there is no
// grid-wide synchronisation; env state lives in registers (float64, like gym), weights in shared
// memory (36.9 KB, staged once by TMA bulk copy).  Per env-step the kernel writes obs 16/32 B +
// act 1 B + logp 4 B + val 4 B + rew 4 B + done 1 B to the [T+1][N] planes and reads nothing from HBM.
#include "drl_env.cuh"
#include "drl_mlp.cuh"
#include "drl_pack.cuh"

namespace drl {

int check_env(const drl_env_t* env);
drl_ep_log_t log_or_empty(const drl_ep_log_t* log);
int launch_rollout_tc(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                      const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl);   // rollout_tc.cu

namespace h256 {
int launch_rollout256(const drl_env_t& env, const drl_net_t* net, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                      const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl);     // rollout256.cu
}

constexpr int RO_WARPS = 4;

template <int KIND, int SUB, int TL>
__global__ void __launch_bounds__(RO_WARPS * 32) rollout_kernel(drl_env_t env, const float* __restrict__ packed, int T,
                                                                uint64_t step0, drl_rollout_buf_t buf, drl_ep_log_t log, const drl_ctrl_t* __restrict__ ctrl) {
    if (ctrl != nullptr) step0 = ctrl->env_step;      // graph-replayable launch: the counter lives in device memory
    using S = EnvSpec<KIND>;
    constexpr int O = S::O, A = S::A, OP = S::OP;
    using P = Packed<O, A>;
    constexpr int EPW = TL * SUB;                         // envs per warp
    constexpr int WS = SUB * OBS_S + H1_S + SUB * OUT_S;  // per-warp scratch floats

    extern __shared__ __align__(128) float smem[];
    float* sw = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P::FWD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* obs_s = smem + P::FWD + 4 + warp * WS;
    float* h1_s = obs_s + SUB * OBS_S;
    float* out_s = h1_s + H1_S;
    stage_params(sw, packed, P::FWD, bar);

    const int N = env.num_envs;
    const int env0 = (blockIdx.x * RO_WARPS + warp) * EPW;
    if (env0 >= N) return;
    const int n = env0 + lane;
    const bool own = lane < EPW && n < N;
    const uint32_t gid = env.env_gid0 + (uint32_t)n;

    EnvLane e;
    e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0.0; e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    if (own) env_load(e, env, n);

    for (int t = 0; t <= T; ++t) {
        // ---- observation of the current state: to HBM (obs[t]) and to the warp's input tile ----
        if (lane < EPW) {
            float obs[OP];
#pragma unroll
            for (int i = 0; i < OP; ++i) obs[i] = 0.0f;
            if (own) {
                env_observation<KIND>(e.s, obs);
                float4* o4 = reinterpret_cast<float4*>(buf.obs + ((size_t)t * N + n) * OP);
#pragma unroll
                for (int q = 0; q < OP / 4; ++q) o4[q] = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
            }
            float* dst = obs_s + (lane / TL) * OBS_S + (lane % TL);
#pragma unroll
            for (int i = 0; i < O; ++i) dst[i * TL] = obs[i];
        }
        __syncwarp();

        // ---- actor + critic forward for the warp's SUB tiles ----
#pragma unroll
        for (int sub = 0; sub < SUB; ++sub) {
            float h2[TL][UPL];
            mlp_forward_tile<O, A, TL>(sw, obs_s + sub * OBS_S, h1_s, out_s + sub * TL * OUT_W, lane, h2);
        }

        // ---- per-env tail: value store, sample, env step ----
        if (own) {
            const size_t i0 = (size_t)t * N + n;
            buf.val[i0] = out_s[lane * OUT_W + 3];
            if (t < T) {
                float l[A];
#pragma unroll
                for (int a = 0; a < A; ++a) l[a] = out_s[lane * OUT_W + a];
                if (buf.logits != nullptr) {
#pragma unroll
                    for (int a = 0; a < A; ++a) buf.logits[i0 * A + a] = l[a];
                }
                const uint64_t step = step0 + (uint64_t)t;
                const uint4 r = philox_seeded(env.seed, gid, (uint32_t)step, (uint32_t)(step >> 32), TAG_ACTION);
                float lp;
                const int act = sample_categorical<A>(l, u01_f32(r.x), lp);
                buf.act[i0] = (uint8_t)act;
                buf.logp[i0] = lp;
                float reward;
                const bool done = env_step<KIND>(e, act, reward, env.seed, gid, step, env.max_episode_steps, log);
                buf.rew[i0 + N] = reward;
                buf.done[i0 + N] = done ? 1 : 0;
            }
        }
        __syncwarp();
    }
    if (own) env_store(e, env, n);
}

template <int KIND, int SUB, int TL>
int launch_rollout(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                   const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    using S = EnvSpec<KIND>;
    constexpr int WS = SUB * OBS_S + H1_S + SUB * OUT_S;
    (void)TL;
    const size_t smem = sizeof(float) * (Packed<S::O, S::A>::FWD + 4 + RO_WARPS * WS);
    DRL_CUDA(cudaFuncSetAttribute(rollout_kernel<KIND, SUB, TL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int epc = TL * SUB * RO_WARPS;
    const int blocks = (env.num_envs + epc - 1) / epc;
    rollout_kernel<KIND, SUB, TL><<<blocks, RO_WARPS * 32, smem, st>>>(env, packed, T, step0, buf, log, ctrl);
    DRL_LAUNCH_CHECK("rollout_kernel");
    return DRL_OK;
}

// envs per warp: small N wants as many warps as possible (the kernel is latency-bound at one warp per
// scheduler), large N amortises the per-env tail (sampling, fp64 env step) over more tiles.
// Returns envs per warp: 4 (one 4-env tile), 8, 16 or 32 (1, 2 or 4 8-env tiles).
int pick_envs_per_warp(int N) {
    const char* ov = getenv("DRL_ROLLOUT_EPW");
    if (ov) { int v = atoi(ov); if (v == 4 || v == 8 || v == 16 || v == 32) return v; }
    const int sms = sm_count();
    if (N <= sms * 8 * 4) return 4;              // <= 8 warps per SM even at 4 envs per warp
    const int warps_at_8 = (N + 7) / 8;
    if (warps_at_8 <= sms * 24) return 8;
    if (warps_at_8 <= sms * 48) return 16;
    return 32;
}

template <int KIND>
int dispatch_rollout(int epw, const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                     const drl_ep_log_t& l, cudaStream_t st, const drl_ctrl_t* ctrl) {
    if (epw == 4) return launch_rollout<KIND, 1, 4>(env, packed, T, step0, buf, l, st, ctrl);
    if (epw == 8) return launch_rollout<KIND, 1, 8>(env, packed, T, step0, buf, l, st, ctrl);
    if (epw == 16) return launch_rollout<KIND, 2, 8>(env, packed, T, step0, buf, l, st, ctrl);
    return launch_rollout<KIND, 4, 8>(env, packed, T, step0, buf, l, st, ctrl);
}

}  // namespace drl

using namespace drl;

static int rollout_impl(const drl_env_t* env, const drl_net_t* net, const float* packed, int32_t T, uint64_t step0,
                        const drl_rollout_buf_t* buf, const drl_ep_log_t* log, uint32_t flags, void* stream, const drl_ctrl_t* ctrl) {
    int rc = check_env(env);
    if (rc != DRL_OK) return rc;
    rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(packed && buf, "drl_rollout: NULL pointer");
    DRL_REQUIRE(buf->obs && buf->act && buf->logp && buf->val && buf->rew && buf->done, "drl_rollout: NULL buffer plane");
    DRL_REQUIRE(T > 0, "drl_rollout: T=%d", T);
    DRL_REQUIRE(net->obs_dim == drl_env_obs_dim(env->kind) && net->num_actions == drl_env_num_actions(env->kind),
                "drl_rollout: net shape does not match env kind %d", env->kind);
    const drl_ep_log_t l = log_or_empty(log);
    cudaStream_t st = as_stream(stream);
    if (net->hidden == 256) {
        DRL_REQUIRE(flags & DRL_ROLLOUT_TENSOR_CORES, "drl_rollout: hidden=256 exists on the tensor-core path only");
        return h256::launch_rollout256(*env, net, packed, T, step0, *buf, l, st, ctrl);
    }
    if (flags & DRL_ROLLOUT_TENSOR_CORES) return launch_rollout_tc(*env, packed, T, step0, *buf, l, st, ctrl);
    const int epw = pick_envs_per_warp(env->num_envs);
    if (env->kind == DRL_ENV_CARTPOLE) return dispatch_rollout<DRL_ENV_CARTPOLE>(epw, *env, packed, T, step0, *buf, l, st, ctrl);
    if (env->kind == DRL_ENV_MOUNTAINCAR) return dispatch_rollout<DRL_ENV_MOUNTAINCAR>(epw, *env, packed, T, step0, *buf, l, st, ctrl);
    return dispatch_rollout<DRL_ENV_ACROBOT>(epw, *env, packed, T, step0, *buf, l, st, ctrl);
}

extern "C" int drl_rollout(const drl_env_t* env, const drl_net_t* net, const float* packed, int32_t T, uint64_t step0,
                           const drl_rollout_buf_t* buf, const drl_ep_log_t* log, uint32_t flags, void* stream) {
    return rollout_impl(env, net, packed, T, step0, buf, log, flags, stream, nullptr);
}

extern "C" int drl_rollout_ctl(const drl_env_t* env, const drl_net_t* net, const float* packed, int32_t T, const drl_ctrl_t* ctrl,
                               const drl_rollout_buf_t* buf, const drl_ep_log_t* log, uint32_t flags, void* stream) {
    DRL_REQUIRE(ctrl != nullptr, "drl_rollout_ctl: ctrl is NULL");
    return rollout_impl(env, net, packed, T, 0, buf, log, flags, stream, ctrl);
}
