// drl_env.cuh -- device-side classic-control environments (gym==0.21 semantics).
// State is float64 in registers/HBM exactly like gym's; only the observation is cast to float32.
// Replaces: gym CartPoleEnv/AcrobotEnv.step + TimeLimit + RecordEpisodeStatistics + the manual
// auto-reset of deep_rl/ppo.py:127-129.
#pragma once
#include "drl_common.cuh"

namespace drl {

template <int KIND> struct EnvSpec;
template <> struct EnvSpec<DRL_ENV_CARTPOLE> { static constexpr int O = 4, A = 2, OP = 4; };
template <> struct EnvSpec<DRL_ENV_ACROBOT>  { static constexpr int O = 6, A = 3, OP = 8; };
template <> struct EnvSpec<DRL_ENV_MOUNTAINCAR> { static constexpr int O = 2, A = 3, OP = 4; };

struct EnvLane {          // one environment, held in registers by its owning lane
    double s[4];
    int elapsed;
    float ep_ret;
    int ep_len;
};

// CartPole-v1 physics (euler), returns terminated; reward is always 1.0.
__device__ __forceinline__ bool cartpole_physics(double (&s)[4], int action) {
    const double gravity = 9.8, masspole = 0.1, total_mass = 0.1 + 1.0, length = 0.5;
    const double polemass_length = 0.1 * 0.5, force_mag = 10.0, tau = 0.02;
    const double theta_thr = 12.0 * 2.0 * 3.14159265358979323846 / 360.0, x_thr = 2.4;
    double x = s[0], x_dot = s[1], th = s[2], th_dot = s[3];
    double force = action == 1 ? force_mag : -force_mag;
    double sinth, costh;
    sincos(th, &sinth, &costh);
    double temp = (force + polemass_length * (th_dot * th_dot) * sinth) / total_mass;
    double thacc = (gravity * sinth - costh * temp) / (length * (4.0 / 3.0 - masspole * (costh * costh) / total_mass));
    double xacc = temp - polemass_length * thacc * costh / total_mass;
    x = x + tau * x_dot;
    x_dot = x_dot + tau * xacc;
    th = th + tau * th_dot;
    th_dot = th_dot + tau * thacc;
    s[0] = x; s[1] = x_dot; s[2] = th; s[3] = th_dot;
    return (x < -x_thr) || (x > x_thr) || (th < -theta_thr) || (th > theta_thr);
}

// MountainCar-v0 (gym 0.21 mountain_car.py): state (position, velocity) in s[0], s[1]; reward is always -1.  Returns terminated.
__device__ __forceinline__ bool mountaincar_physics(double (&s)[4], int action) {
    double position = s[0], velocity = s[1];
    velocity = velocity + ((double)(action - 1) * 0.001 + cos(3.0 * position) * (-0.0025));
    velocity = fmin(fmax(velocity, -0.07), 0.07);
    position = position + velocity;
    position = fmin(fmax(position, -1.2), 0.6);
    if (position == -1.2 && velocity < 0.0) velocity = 0.0;
    s[0] = position; s[1] = velocity;
    return position >= 0.5 && velocity >= 0.0;
}

// Acrobot-v1 "book" equations of motion.
__device__ __forceinline__ void acrobot_dsdt(const double (&y)[4], double torque, double (&d)[4]) {
    const double m1 = 1.0, m2 = 1.0, l1 = 1.0, lc1 = 0.5, lc2 = 0.5, I1 = 1.0, I2 = 1.0, g = 9.8;
    const double PI = 3.14159265358979323846;
    double th1 = y[0], th2 = y[1], dth1 = y[2], dth2 = y[3];
    double s2, c2;
    sincos(th2, &s2, &c2);
    double d1 = m1 * (lc1 * lc1) + m2 * (l1 * l1 + lc2 * lc2 + 2.0 * l1 * lc2 * c2) + I1 + I2;
    double d2 = m2 * (lc2 * lc2 + l1 * lc2 * c2) + I2;
    double phi2 = m2 * lc2 * g * cos(th1 + th2 - PI / 2.0);
    double phi1 = -m2 * l1 * lc2 * (dth2 * dth2) * s2 - 2.0 * m2 * l1 * lc2 * dth2 * dth1 * s2 +
                  (m1 * lc1 + m2 * l1) * g * cos(th1 - PI / 2.0) + phi2;
    double ddth2 = (torque + d2 / d1 * phi1 - m2 * l1 * lc2 * (dth1 * dth1) * s2 - phi2) /
                   (m2 * (lc2 * lc2) + I2 - (d2 * d2) / d1);
    double ddth1 = -(d2 * ddth2 + phi1) / d1;
    d[0] = dth1; d[1] = dth2; d[2] = ddth1; d[3] = ddth2;
}

__device__ __forceinline__ double wrap_pi(double x) {
    const double PI = 3.14159265358979323846;
    const double diff = PI - (-PI);
    while (x > PI) x = x - diff;
    while (x < -PI) x = x + diff;
    return x;
}

// Acrobot-v1 step: RK4 over dt = 0.2, wrap angles, clip velocities.  Returns terminated.
__device__ __forceinline__ bool acrobot_physics(double (&s)[4], int action, float& reward) {
    const double PI = 3.14159265358979323846;
    const double dt = 0.2, dt2 = dt / 2.0;
    double torque = (double)action - 1.0;
    double k1[4], k2[4], k3[4], k4[4], yt[4];
    acrobot_dsdt(s, torque, k1);
#pragma unroll
    for (int i = 0; i < 4; ++i) yt[i] = s[i] + dt2 * k1[i];
    acrobot_dsdt(yt, torque, k2);
#pragma unroll
    for (int i = 0; i < 4; ++i) yt[i] = s[i] + dt2 * k2[i];
    acrobot_dsdt(yt, torque, k3);
#pragma unroll
    for (int i = 0; i < 4; ++i) yt[i] = s[i] + dt * k3[i];
    acrobot_dsdt(yt, torque, k4);
    double ns[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ns[i] = s[i] + dt / 6.0 * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    ns[0] = wrap_pi(ns[0]);
    ns[1] = wrap_pi(ns[1]);
    ns[2] = fmin(fmax(ns[2], -4.0 * PI), 4.0 * PI);
    ns[3] = fmin(fmax(ns[3], -9.0 * PI), 9.0 * PI);
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = ns[i];
    bool terminal = (-cos(ns[0]) - cos(ns[1] + ns[0])) > 1.0;
    reward = terminal ? 0.0f : -1.0f;
    return terminal;
}

template <int KIND>
__device__ __forceinline__ void env_observation(const double (&s)[4], float (&obs)[EnvSpec<KIND>::OP]) {
    if constexpr (KIND == DRL_ENV_CARTPOLE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) obs[i] = (float)s[i];
    } else if constexpr (KIND == DRL_ENV_MOUNTAINCAR) {
        obs[0] = (float)s[0]; obs[1] = (float)s[1]; obs[2] = 0.0f; obs[3] = 0.0f;
    } else {
        double s1, c1, s2, c2;
        sincos(s[0], &s1, &c1);
        sincos(s[1], &s2, &c2);
        obs[0] = (float)c1; obs[1] = (float)s1; obs[2] = (float)c2; obs[3] = (float)s2;
        obs[4] = (float)s[2]; obs[5] = (float)s[3]; obs[6] = 0.0f; obs[7] = 0.0f;
    }
}

// Reset draw (mirrored by the CPU oracle): four 32-bit uniforms of one Philox block.
template <int KIND>
__device__ __forceinline__ void env_reset_state(double (&s)[4], uint64_t seed, uint32_t gid, uint64_t step) {
    uint4 o = philox_seeded(seed, gid, (uint32_t)step, (uint32_t)(step >> 32), TAG_RESET);
    if constexpr (KIND == DRL_ENV_MOUNTAINCAR) {   // position ~ U(-0.6, -0.4), velocity 0
        s[0] = -0.6 + 0.2 * u01_f64(o.x);
        s[1] = 0.0; s[2] = 0.0; s[3] = 0.0;
        return;
    }
    const double half = KIND == DRL_ENV_CARTPOLE ? 0.05 : 0.1;
    s[0] = -half + (2.0 * half) * u01_f64(o.x);
    s[1] = -half + (2.0 * half) * u01_f64(o.y);
    s[2] = -half + (2.0 * half) * u01_f64(o.z);
    s[3] = -half + (2.0 * half) * u01_f64(o.w);
}

// The physics of one step for a given action (no wrappers): new state in s, reward out, returns terminated.  Depends on the state
// and the action only, so it can be evaluated for every possible action before the action is known (rollout_tc.cu).
template <int KIND>
__device__ __forceinline__ bool env_physics(double (&s)[4], int action, float& reward) {
    if constexpr (KIND == DRL_ENV_CARTPOLE) { reward = 1.0f; return cartpole_physics(s, action); }
    else if constexpr (KIND == DRL_ENV_MOUNTAINCAR) { reward = -1.0f; return mountaincar_physics(s, action); }
    else { return acrobot_physics(s, action, reward); }
}

// One env step with TimeLimit, episode statistics and auto-reset.  Returns done; `reward` out.
template <int KIND>
__device__ __forceinline__ bool env_after_physics(EnvLane& e, bool term, float reward, uint64_t seed, uint32_t gid, uint64_t step,
                                                  int max_episode_steps, const drl_ep_log_t& log);

template <int KIND>
__device__ __forceinline__ bool env_step(EnvLane& e, int action, float& reward, uint64_t seed, uint32_t gid,
                                         uint64_t step, int max_episode_steps, const drl_ep_log_t& log) {
    const bool term = env_physics<KIND>(e.s, action, reward);
    return env_after_physics<KIND>(e, term, reward, seed, gid, step, max_episode_steps, log);
}

// TimeLimit (done = terminated or elapsed >= max), RecordEpisodeStatistics, finished-episode log and auto-reset, after the physics
template <int KIND>
__device__ __forceinline__ bool env_after_physics(EnvLane& e, bool term, float reward, uint64_t seed, uint32_t gid, uint64_t step,
                                                  int max_episode_steps, const drl_ep_log_t& log) {
    e.elapsed += 1;
    bool done = term || (e.elapsed >= max_episode_steps);
    e.ep_ret = e.ep_ret + reward;
    e.ep_len += 1;
    // Finished-episode log, aggregated per warp: one atomicAdd on the shared counter (and one per sum) for all lanes of the
    // warp that finished in this step instead of one per lane -- an untrained policy ends ~10 % of all episodes every step,
    // and per-lane atomics on three addresses then serialise the whole grid.  Every lane that called env_step takes part
    // (the callers' branches are warp-uniform up to the set of lanes that own an environment).
    if (log.count != nullptr) {
        const unsigned active = __activemask();
        const unsigned dmask = __ballot_sync(active, done);
        if (dmask != 0u) {
            const int lane = (int)(threadIdx.x & 31u);
            const int leader = __ffs(dmask) - 1;
            double sr = 0.0, sl = 0.0;
            for (unsigned bits = dmask; bits != 0u; bits &= bits - 1u) {      // warp-uniform loop over the finished lanes
                const int b = __ffs(bits) - 1;
                sr += (double)__shfl_sync(active, e.ep_ret, b);
                sl += (double)__shfl_sync(active, e.ep_len, b);
            }
            uint32_t base = 0;
            if (lane == leader) {
                base = atomicAdd(log.count, (uint32_t)__popc(dmask));
                if (log.sum_ret) atomicAdd(log.sum_ret, sr);
                if (log.sum_len) atomicAdd(log.sum_len, sl);
            }
            base = __shfl_sync(active, base, leader);
            if (done) {
                const uint32_t slot = base + (uint32_t)__popc(dmask & ((1u << lane) - 1u));
                if (slot < log.cap && log.entries) {      // one 24-byte record: a 16-byte and an 8-byte store
                    drl_ep_entry_t en;
                    en.step = step; en.env = gid; en.ret = e.ep_ret; en.len = e.ep_len; en.pad = 0u;
                    log.entries[slot] = en;
                }
            }
        }
    }
    if (done) {
        env_reset_state<KIND>(e.s, seed, gid, step);
        e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    }
    return done;
}

__device__ __forceinline__ void env_load(EnvLane& e, const drl_env_t& env, int n) {
    const int N = env.num_envs;
#pragma unroll
    for (int i = 0; i < 4; ++i) e.s[i] = env.state[(size_t)i * N + n];
    e.elapsed = env.elapsed[n]; e.ep_ret = env.ep_ret[n]; e.ep_len = env.ep_len[n];
}
__device__ __forceinline__ void env_store(const EnvLane& e, const drl_env_t& env, int n) {
    const int N = env.num_envs;
#pragma unroll
    for (int i = 0; i < 4; ++i) env.state[(size_t)i * N + n] = e.s[i];
    env.elapsed[n] = e.elapsed; env.ep_ret[n] = e.ep_ret; env.ep_len[n] = e.ep_len;
}

}  // namespace drl
