// umma_selftest.cu -- diagnostic entry point: runs each tcgen05.mma operand-layout combination the update
// kernel uses on caller-provided matrices, so the descriptors / swizzles / TMEM addressing can be checked
// against a plain fp32 matmul (tests/test_gpu_umma.py).
//   mode 0: D[128x64]  = A[128x64] . B[64x64]^T      A K-major SW128, B K-major SW128      (forward z2 = h1 . W2^T)
//   mode 1: D[128x64]  = A[128x64] . B[64x64]        A K-major SW128, B MN-major SW128     (dh1 = dz2 . W2)
//   mode 2: D[128x128] = A[128x128]^T . B[128x128]   A, B MN-major SW128 tile pairs        (dW2 = dz2^T . h1)
//   mode 3: D[128x16]  = A[128x128]^T . B[128x16]    A MN-major SW128 pair, B MN-major NS16 (dW1 / dW4 / db)
#include "drl_tc_common.cuh"

namespace drl {

__global__ void __launch_bounds__(128) umma_selftest_kernel(int mode, int variant, const float* __restrict__ a,
                                                             const float* __restrict__ b, float* __restrict__ d) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // SW128 tiles need 1024-byte alignment
    unsigned char* tA = smem;               // 2 x 16 KB
    unsigned char* tB = smem + 32768;       // 2 x 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 65536 + 8);
    const int t = threadIdx.x, warp = t >> 5;

    float v[64];
    if (mode <= 1) {
#pragma unroll
        for (int k = 0; k < 64; ++k) v[k] = a[t * 64 + k];
        umma::store_row_sw128(tA, t, v);
        if (t < 64) {
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = b[t * 64 + k];
            umma::store_row_sw128(tB, t, v);
        }
    } else {
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int k = 0; k < 64; ++k) v[k] = a[t * 128 + half * 64 + k];
            umma::store_row_sw128(tA + half * 16384, t, v);
        }
        if (mode == 2) {
            for (int half = 0; half < 2; ++half) {
#pragma unroll
                for (int k = 0; k < 64; ++k) v[k] = b[t * 128 + half * 64 + k];
                umma::store_row_sw128(tB + half * 16384, t, v);
            }
        } else {
            float w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = b[t * 16 + k];
            umma::store_row_ns16(tB, 128, t, w);
        }
    }
    if (t == 0) {
        mbar_init(bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc(slot, 128);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *slot;

    if (t == 0) {
        const uint32_t aA = smem_u32(tA), aB = smem_u32(tB);
        const bool swap = variant == 1;
        if (mode == 0) {
            const uint32_t idesc = umma::make_idesc(128, 64, false, false);
            for (int kb = 0; kb < 4; ++kb)
                umma::mma(tmem, umma::make_desc(aA + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                          umma::make_desc(aB + kb * 32, 16, 1024, umma::LAYOUT_SW128), idesc, kb > 0);
        } else if (mode == 1) {
            const uint32_t idesc = umma::make_idesc(128, 64, false, true);
            const uint32_t lbo = swap ? 1024 : 8192, sbo = swap ? 8192 : 1024;
            for (int kb = 0; kb < 4; ++kb)
                umma::mma(tmem, umma::make_desc(aA + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                          umma::make_desc(aB + kb * 2048, lbo, sbo, umma::LAYOUT_SW128), idesc, kb > 0);
        } else if (mode == 2) {
            const uint32_t idesc = umma::make_idesc(128, 128, true, true);
            const uint32_t lbo = swap ? 1024 : 16384, sbo = swap ? 16384 : 1024;
            for (int kb = 0; kb < 8; ++kb)
                umma::mma(tmem, umma::make_desc(aA + kb * 2048, lbo, sbo, umma::LAYOUT_SW128),
                          umma::make_desc(aB + kb * 2048, lbo, sbo, umma::LAYOUT_SW128), idesc, kb > 0);
        } else {
            const uint32_t idesc = umma::make_idesc(128, 16, true, true);
            const uint32_t lbo = swap ? 2048 : 128, sbo = swap ? 128 : 2048;
            for (int kb = 0; kb < 8; ++kb)
                umma::mma(tmem, umma::make_desc(aA + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                          umma::make_desc(aB + kb * 256, lbo, sbo, umma::LAYOUT_NONE), idesc, kb > 0);
        }
        umma::commit(bar);
    }
    mbar_wait(bar, 0);
    umma::fence_after_sync();

    const int ncols = mode == 2 ? 128 : (mode == 3 ? 16 : 64);
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    if (ncols == 16) {
        float o[16];
        umma::ld16(lane_addr, o);
#pragma unroll
        for (int j = 0; j < 16; ++j) d[t * 16 + j] = o[j];
    } else {
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            float o[32];
            umma::ld32(lane_addr + c0, o);
#pragma unroll
            for (int j = 0; j < 32; ++j) d[t * ncols + c0 + j] = o[j];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 128);
}

__global__ void __launch_bounds__(256) tanh_selftest_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) y[i] = tanh_mufu(x[i]);
}

}  // namespace drl

using namespace drl;

extern "C" int drl_selftest_tanh(const float* x, float* y, int64_t n, void* stream) {
    DRL_REQUIRE(x && y, "drl_selftest_tanh: NULL pointer");
    if (n <= 0) return DRL_OK;
    tanh_selftest_kernel<<<(unsigned)((n + 255) / 256), 256, 0, as_stream(stream)>>>(x, y, n);
    DRL_LAUNCH_CHECK("tanh_selftest_kernel");
    return DRL_OK;
}

extern "C" int drl_selftest_umma(int32_t mode, int32_t variant, const float* a, const float* b, float* d_out, void* stream) {
    DRL_REQUIRE(mode >= 0 && mode <= 3, "drl_selftest_umma: mode=%d", mode);
    DRL_REQUIRE(a && b && d_out, "drl_selftest_umma: NULL pointer");
    const int smem = 65536 + 64 + 1024;
    DRL_CUDA(cudaFuncSetAttribute(umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    umma_selftest_kernel<<<1, 128, smem, as_stream(stream)>>>(mode, variant, a, b, d_out);
    DRL_LAUNCH_CHECK("umma_selftest_kernel");
    return DRL_OK;
}
