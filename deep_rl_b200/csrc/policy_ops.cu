// policy_ops.cu -- stand-alone batched forward (ActorCritic.get_value / get_action_distribution,
// deep_rl/ppo.py:49-54) and the Philox categorical sampler (get_action, ppo.py:56-59).
#include "drl_mlp.cuh"
#include "drl_pack.cuh"

namespace drl {

namespace h256 {
int launch_forward256(const drl_net_t* net, const float* packed, const float* obs, int64_t n, float* logits, float* value,
                      int only_net, cudaStream_t st);     // update256.cu
}

constexpr int FWD_WARPS = 4;

template <int O, int A, int OP>
__global__ void __launch_bounds__(FWD_WARPS * 32) policy_forward_kernel(const float* __restrict__ packed,
                                                                       const float* __restrict__ obs, int64_t n,
                                                                       float* __restrict__ logits_out,
                                                                       float* __restrict__ value_out) {
    using P = Packed<O, A>;
    extern __shared__ __align__(128) float smem[];
    float* sw = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + P::FWD);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* obs_s = smem + P::FWD + 4 + warp * (OBS_S + H1_S + OUT_S);
    float* h1_s = obs_s + OBS_S;
    float* out_s = h1_s + H1_S;
    stage_params(sw, packed, P::FWD, bar);

    const int64_t tiles = (n + TILE - 1) / TILE;
    for (int64_t tile = (int64_t)blockIdx.x * FWD_WARPS + warp; tile < tiles; tile += (int64_t)gridDim.x * FWD_WARPS) {
        const int64_t s = tile * TILE + lane;
        if (lane < TILE) {
            float x[OP];
#pragma unroll
            for (int i = 0; i < OP; ++i) x[i] = 0.0f;
            if (s < n) {
                const float4* p4 = reinterpret_cast<const float4*>(obs + s * OP);
#pragma unroll
                for (int q = 0; q < OP / 4; ++q) {
                    const float4 v = p4[q];
                    x[4 * q] = v.x; x[4 * q + 1] = v.y; x[4 * q + 2] = v.z; x[4 * q + 3] = v.w;
                }
            }
#pragma unroll
            for (int i = 0; i < O; ++i) obs_s[i * TILE + lane] = x[i];
        }
        __syncwarp();
        float h2[TILE][UPL];
        mlp_forward_tile<O, A>(sw, obs_s, h1_s, out_s, lane, h2);
        if (lane < TILE && s < n) {
#pragma unroll
            for (int a = 0; a < A; ++a) logits_out[s * A + a] = out_s[lane * OUT_W + a];
            value_out[s] = out_s[lane * OUT_W + 3];
        }
        __syncwarp();
    }
}

template <int A>
__global__ void __launch_bounds__(256) sample_kernel(const float* __restrict__ logits, int64_t n, uint64_t seed,
                                                      uint32_t env_gid0, uint64_t step, int32_t* __restrict__ act_out,
                                                      float* __restrict__ logp_out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float l[A];
#pragma unroll
    for (int a = 0; a < A; ++a) l[a] = logits[i * A + a];
    const uint4 r = philox_seeded(seed, env_gid0 + (uint32_t)i, (uint32_t)step, (uint32_t)(step >> 32), TAG_ACTION);
    float lp;
    const int act = sample_categorical<A>(l, u01_f32(r.x), lp);
    act_out[i] = act;
    if (logp_out) logp_out[i] = lp;
}

template <int O, int A, int OP>
int launch_forward(const float* packed, const float* obs, int64_t n, float* logits, float* value, cudaStream_t st) {
    const size_t smem = sizeof(float) * (Packed<O, A>::FWD + 4 + FWD_WARPS * (OBS_S + H1_S + OUT_S));
    // per call: the attribute belongs to the current device / context, and the header promises re-entrancy across threads
    DRL_CUDA(cudaFuncSetAttribute(policy_forward_kernel<O, A, OP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t tiles = (n + TILE - 1) / TILE;
    int64_t blocks = (tiles + FWD_WARPS - 1) / FWD_WARPS;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    policy_forward_kernel<O, A, OP><<<(int)blocks, FWD_WARPS * 32, smem, st>>>(packed, obs, n, logits, value);
    DRL_LAUNCH_CHECK("policy_forward_kernel");
    return DRL_OK;
}

}  // namespace drl

using namespace drl;

extern "C" {

int drl_policy_forward(const drl_net_t* net, const float* packed, const float* obs, int64_t n, float* logits_out,
                       float* value_out, void* stream) {
    int rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(packed && obs && logits_out && value_out, "drl_policy_forward: NULL pointer");
    if (n <= 0) return DRL_OK;
    if (net->hidden == 256)      // 256-wide nets exist on the tensor-core path only (bf16 operands, fp32 accumulation)
        return h256::launch_forward256(net, packed, obs, n, logits_out, value_out, -1, as_stream(stream));
    if (net->obs_dim == 4) return launch_forward<4, 2, 4>(packed, obs, n, logits_out, value_out, as_stream(stream));
    if (net->obs_dim == 2) return launch_forward<2, 3, 4>(packed, obs, n, logits_out, value_out, as_stream(stream));
    return launch_forward<6, 3, 8>(packed, obs, n, logits_out, value_out, as_stream(stream));
}

int drl_sample(const float* logits, int64_t n, int32_t num_actions, uint64_t seed, uint32_t env_gid0, uint64_t step,
               int32_t* act_out, float* logp_out, void* stream) {
    DRL_REQUIRE(logits && act_out, "drl_sample: NULL pointer");
    DRL_REQUIRE(num_actions == 2 || num_actions == 3, "drl_sample: num_actions=%d unsupported", num_actions);
    if (n <= 0) return DRL_OK;
    const int blocks = (int)((n + 255) / 256);
    if (num_actions == 2) sample_kernel<2><<<blocks, 256, 0, as_stream(stream)>>>(logits, n, seed, env_gid0, step, act_out, logp_out);
    else sample_kernel<3><<<blocks, 256, 0, as_stream(stream)>>>(logits, n, seed, env_gid0, step, act_out, logp_out);
    DRL_LAUNCH_CHECK("sample_kernel");
    return DRL_OK;
}

}  // extern "C"
