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

template <int KIND>
__global__ void __launch_bounds__(256) env_step_kernel(drl_env_t env, uint64_t step, const int32_t* __restrict__ actions,
                                                        float* __restrict__ obs_out, float* __restrict__ rew_out,
                                                        uint8_t* __restrict__ done_out, drl_ep_log_t log) {
    constexpr int OP = EnvSpec<KIND>::OP;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= env.num_envs) return;
    EnvLane e;
    env_load(e, env, n);
    float reward;
    const bool done = env_step<KIND>(e, actions[n], reward, env.seed, env.env_gid0 + (uint32_t)n, step,
                                     env.max_episode_steps, log);
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
