// env_ops.cu -- stand-alone batched environment kernels (one thread per env, state in HBM).
// Replaces TorchWrapper.step/reset + gym CartPoleEnv/AcrobotEnv + TimeLimit + RecordEpisodeStatistics
// (deep_rl/ppo.py:10-22,79) and the auto-reset of ppo.py:127-129.  HBM-bound: per env-step it reads
// state 32 B + action 4 B + counters 12 B and writes state 32 B + obs 16/32 B + rew 4 B + done 1 B +
// counters 12 B, all coalesced SoA.
#include "drl_env.cuh"

namespace drl {

template <int KIND>
__global__ void __launch_bounds__(256) env_reset_kernel(drl_env_t env, float* __restrict__ obs_out) {
    constexpr int OP = EnvSpec<KIND>::OP;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= env.num_envs) return;
    EnvLane e;
    env_reset_state<KIND>(e.s, env.seed, env.env_gid0 + (uint32_t)n, 0xFFFFFFFFFFFFFFFFull);
    e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    env_store(e, env, n);
    float obs[OP];
    env_observation<KIND>(e.s, obs);
    float4* o4 = reinterpret_cast<float4*>(obs_out + (size_t)n * OP);
#pragma unroll
    for (int q = 0; q < OP / 4; ++q) o4[q] = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
}

template <int KIND>
__global__ void __launch_bounds__(256) env_observe_kernel(drl_env_t env, float* __restrict__ obs_out) {
    constexpr int OP = EnvSpec<KIND>::OP;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= env.num_envs) return;
    double s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = env.state[(size_t)i * env.num_envs + n];
    float obs[OP];
    env_observation<KIND>(s, obs);
    float4* o4 = reinterpret_cast<float4*>(obs_out + (size_t)n * OP);
#pragma unroll
    for (int q = 0; q < OP / 4; ++q) o4[q] = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
}

// The finished-episode log is aggregated per BLOCK here (one atomicAdd on the counter and one per sum for up to 256 envs): with
// millions of envs and an untrained policy ~5 % of them finish in every step, and even warp-aggregated atomics on three
// addresses (drl_env.cuh) would then bound the kernel instead of HBM.
template <int KIND>
__global__ void __launch_bounds__(256) env_step_kernel(drl_env_t env, uint64_t step, const int32_t* __restrict__ actions,
                                                        float* __restrict__ obs_out, float* __restrict__ rew_out,
                                                        uint8_t* __restrict__ done_out, drl_ep_log_t log) {
    constexpr int OP = EnvSpec<KIND>::OP;
    __shared__ uint32_t w_cnt[8], w_base[8];
    __shared__ double w_ret[8], w_len[8];
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = n < env.num_envs;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t gid = env.env_gid0 + (uint32_t)n;
    EnvLane e;
    e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0.0; e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    float reward = 0.0f;
    bool done = false;
    if (live) {
        env_load(e, env, n);
        const bool term = env_physics<KIND>(e.s, actions[n], reward);
        e.elapsed += 1;                                   // TimeLimit + RecordEpisodeStatistics, as env_after_physics
        done = term || (e.elapsed >= env.max_episode_steps);
        e.ep_ret = e.ep_ret + reward;
        e.ep_len += 1;
    }
    if (log.count != nullptr) {
        const unsigned dmask = __ballot_sync(0xffffffffu, done);
        double sr = done ? (double)e.ep_ret : 0.0, sl = done ? (double)e.ep_len : 0.0;
        sr = warp_sum(sr); sl = warp_sum(sl);
        if (lane == 0) { w_cnt[warp] = (uint32_t)__popc(dmask); w_ret[warp] = sr; w_len[warp] = sl; }
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            double tr = 0.0, tl = 0.0;
            for (int w = 0; w < 8; ++w) { w_base[w] = tot; tot += w_cnt[w]; tr += w_ret[w]; tl += w_len[w]; }
            if (tot > 0) {
                const uint32_t base = atomicAdd(log.count, tot);
                if (log.sum_ret) atomicAdd(log.sum_ret, tr);
                if (log.sum_len) atomicAdd(log.sum_len, tl);
                for (int w = 0; w < 8; ++w) w_base[w] += base;
            }
        }
        __syncthreads();
        if (done) {
            const uint32_t slot = w_base[warp] + (uint32_t)__popc(dmask & ((1u << lane) - 1u));
            if (slot < log.cap && log.entries) {
                drl_ep_entry_t en;
                en.step = step; en.env = gid; en.ret = e.ep_ret; en.len = e.ep_len; en.pad = 0u;
                log.entries[slot] = en;
            }
        }
    }
    if (!live) return;
    if (done) {
        env_reset_state<KIND>(e.s, env.seed, gid, step);
        e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    }
    env_store(e, env, n);
    float obs[OP];
    env_observation<KIND>(e.s, obs);
    float4* o4 = reinterpret_cast<float4*>(obs_out + (size_t)n * OP);
#pragma unroll
    for (int q = 0; q < OP / 4; ++q) o4[q] = make_float4(obs[4 * q], obs[4 * q + 1], obs[4 * q + 2], obs[4 * q + 3]);
    rew_out[n] = reward;
    done_out[n] = done ? 1 : 0;
}

int check_env(const drl_env_t* env) {
    if (env == nullptr) { set_error("env is NULL"); return DRL_ERR_ARG; }
    if (env->kind != DRL_ENV_CARTPOLE && env->kind != DRL_ENV_ACROBOT && env->kind != DRL_ENV_MOUNTAINCAR) { set_error("unknown env kind %d", env->kind); return DRL_ERR_ARG; }
    if (env->num_envs <= 0) { set_error("num_envs=%d", env->num_envs); return DRL_ERR_ARG; }
    if (!env->state || !env->elapsed || !env->ep_ret || !env->ep_len) { set_error("env state pointer is NULL"); return DRL_ERR_ARG; }
    if (env->max_episode_steps <= 0) { set_error("max_episode_steps=%d", env->max_episode_steps); return DRL_ERR_ARG; }
    return DRL_OK;
}

drl_ep_log_t log_or_empty(const drl_ep_log_t* log) {
    drl_ep_log_t l;
    memset(&l, 0, sizeof(l));
    if (log) l = *log;
    return l;
}

}  // namespace drl

using namespace drl;

extern "C" {

int drl_env_reset(const drl_env_t* env, float* obs_out, void* stream) {
    int rc = check_env(env);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(obs_out, "drl_env_reset: obs_out is NULL");
    const int blocks = (env->num_envs + 255) / 256;
    if (env->kind == DRL_ENV_CARTPOLE) env_reset_kernel<DRL_ENV_CARTPOLE><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    else if (env->kind == DRL_ENV_MOUNTAINCAR) env_reset_kernel<DRL_ENV_MOUNTAINCAR><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    else env_reset_kernel<DRL_ENV_ACROBOT><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    DRL_LAUNCH_CHECK("env_reset_kernel");
    return DRL_OK;
}

int drl_env_observe(const drl_env_t* env, float* obs_out, void* stream) {
    int rc = check_env(env);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(obs_out, "drl_env_observe: obs_out is NULL");
    const int blocks = (env->num_envs + 255) / 256;
    if (env->kind == DRL_ENV_CARTPOLE) env_observe_kernel<DRL_ENV_CARTPOLE><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    else if (env->kind == DRL_ENV_MOUNTAINCAR) env_observe_kernel<DRL_ENV_MOUNTAINCAR><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    else env_observe_kernel<DRL_ENV_ACROBOT><<<blocks, 256, 0, as_stream(stream)>>>(*env, obs_out);
    DRL_LAUNCH_CHECK("env_observe_kernel");
    return DRL_OK;
}

int drl_env_step(const drl_env_t* env, uint64_t step, const int32_t* actions, float* obs_out, float* rew_out,
                 uint8_t* done_out, const drl_ep_log_t* log, void* stream) {
    int rc = check_env(env);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(actions && obs_out && rew_out && done_out, "drl_env_step: NULL pointer");
    const int blocks = (env->num_envs + 255) / 256;
    const drl_ep_log_t l = log_or_empty(log);
    if (env->kind == DRL_ENV_CARTPOLE)
        env_step_kernel<DRL_ENV_CARTPOLE><<<blocks, 256, 0, as_stream(stream)>>>(*env, step, actions, obs_out, rew_out, done_out, l);
    else if (env->kind == DRL_ENV_MOUNTAINCAR)
        env_step_kernel<DRL_ENV_MOUNTAINCAR><<<blocks, 256, 0, as_stream(stream)>>>(*env, step, actions, obs_out, rew_out, done_out, l);
    else
        env_step_kernel<DRL_ENV_ACROBOT><<<blocks, 256, 0, as_stream(stream)>>>(*env, step, actions, obs_out, rew_out, done_out, l);
    DRL_LAUNCH_CHECK("env_step_kernel");
    return DRL_OK;
}

}  // extern "C"
