// abi_misc.cu -- error plumbing, shape queries and the canonical->packed parameter transform.
#include <stdarg.h>

#include "drl_env.cuh"
#include "drl_h256.cuh"
#include "drl_pack.cuh"

namespace drl {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return DRL_ERR_CUDA;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int check_net(const drl_net_t* net) {
    if (net == nullptr) { set_error("net is NULL"); return DRL_ERR_ARG; }
    if (net->hidden != H && net->hidden != h256::HH) {
        set_error("hidden=%d unsupported (this build: %d and %d)", net->hidden, H, h256::HH);
        return DRL_ERR_UNSUPPORTED;
    }
    const bool cart = net->obs_dim == 4 && net->num_actions == 2 && net->obs_stride == 4;
    const bool acro = net->obs_dim == 6 && net->num_actions == 3 && net->obs_stride == 8;
    const bool mcar = net->obs_dim == 2 && net->num_actions == 3 && net->obs_stride == 4;
    if (!cart && !acro && !mcar) {
        set_error("unsupported net shape O=%d A=%d OP=%d (CartPole 4/2/4, Acrobot 6/3/8 or MountainCar 2/3/4)", net->obs_dim,
                  net->num_actions, net->obs_stride);
        return DRL_ERR_UNSUPPORTED;
    }
    return DRL_OK;
}

template <int O, int A>
__global__ void pack_params_kernel(const float* __restrict__ params, float* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Packed<O, A>::C_ALL) packed_store<O, A>(packed, i, params[i]);
}
template <int O, int A>
__global__ void pack_params256_kernel(const float* __restrict__ params, float* __restrict__ packed) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < h256::Packed256<O, A>::C_ALL) h256::packed_store256<O, A>(packed, i, params[i]);
}

}  // namespace drl

using namespace drl;

extern "C" {

int drl_abi_version(void) { return DRL_ABI_VERSION; }
const char* drl_last_error(void) { return g_err; }

int drl_env_obs_dim(int32_t kind) { return kind == DRL_ENV_CARTPOLE ? 4 : kind == DRL_ENV_ACROBOT ? 6 : kind == DRL_ENV_MOUNTAINCAR ? 2 : DRL_ERR_ARG; }
int drl_env_num_actions(int32_t kind) { return kind == DRL_ENV_CARTPOLE ? 2 : (kind == DRL_ENV_ACROBOT || kind == DRL_ENV_MOUNTAINCAR) ? 3 : DRL_ERR_ARG; }
int drl_env_obs_stride(int32_t kind) { return (kind == DRL_ENV_CARTPOLE || kind == DRL_ENV_MOUNTAINCAR) ? 4 : kind == DRL_ENV_ACROBOT ? 8 : DRL_ERR_ARG; }

int64_t drl_param_count(const drl_net_t* net) {
    if (check_net(net) != DRL_OK) return -1;
    if (net->hidden == h256::HH)
        return net->obs_dim == 4 ? h256::Packed256<4, 2>::C_ALL : net->obs_dim == 6 ? h256::Packed256<6, 3>::C_ALL : h256::Packed256<2, 3>::C_ALL;
    return net->obs_dim == 4 ? Packed<4, 2>::C_ALL : net->obs_dim == 6 ? Packed<6, 3>::C_ALL : Packed<2, 3>::C_ALL;
}
int64_t drl_packed_count(const drl_net_t* net) {
    if (check_net(net) != DRL_OK) return -1;
    if (net->hidden == h256::HH)
        return net->obs_dim == 4 ? h256::Packed256<4, 2>::TOTAL : net->obs_dim == 6 ? h256::Packed256<6, 3>::TOTAL : h256::Packed256<2, 3>::TOTAL;
    return net->obs_dim == 4 ? Packed<4, 2>::TOTAL : net->obs_dim == 6 ? Packed<6, 3>::TOTAL : Packed<2, 3>::TOTAL;
}
int drl_record_width(const drl_net_t* net) {
    if (check_net(net) != DRL_OK) return -1;
    return net->obs_dim <= 4 ? 8 : 16;
}
size_t drl_workspace_bytes(const drl_net_t* net) {
    if (check_net(net) != DRL_OK) return 0;
    return workspace_bytes_for(drl_param_count(net), net->hidden);
}

int drl_pack_params(const drl_net_t* net, const float* params, float* packed_out, void* stream) {
    int rc = check_net(net);
    if (rc != DRL_OK) return rc;
    DRL_REQUIRE(params && packed_out, "drl_pack_params: NULL pointer");
    const int P = (int)drl_param_count(net);
    const int blocks = (P + 255) / 256;
    if (net->hidden == h256::HH) {
        if (net->obs_dim == 4) pack_params256_kernel<4, 2><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
        else if (net->obs_dim == 6) pack_params256_kernel<6, 3><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
        else pack_params256_kernel<2, 3><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
        DRL_LAUNCH_CHECK("pack_params256_kernel");
        return DRL_OK;
    }
    if (net->obs_dim == 4) pack_params_kernel<4, 2><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
    else if (net->obs_dim == 6) pack_params_kernel<6, 3><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
    else pack_params_kernel<2, 3><<<blocks, 256, 0, as_stream(stream)>>>(params, packed_out);
    DRL_LAUNCH_CHECK("pack_params_kernel");
    return DRL_OK;
}

}  // extern "C"
