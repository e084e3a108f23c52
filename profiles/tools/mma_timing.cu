// mma_timing.cu -- how long do the tcgen05.mma groups of ppo_grad_tc_kernel take on the tensor pipe?
// One CTA, one issuing thread: issue a group, commit to an mbarrier, spin until it completes, print clock64 deltas.
// Build + run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deep_rl_b200/csrc -o /tmp/mma_timing profiles/tools/mma_timing.cu && /tmp/mma_timing
#include <cstdio>
#include "drl_umma.cuh"

using namespace drl;

constexpr int NG = 9, REP = 6;

__global__ void __launch_bounds__(128) timing_kernel(long long* out, long long* obs) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 163840);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 163840 + 32);
    for (int i = threadIdx.x; i < 163840 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u + i % 7;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 1, 1); mbar_init(bar + 2, 1); mbar_fence_init(); }
    if (threadIdx.x < 32) umma::tmem_alloc(slot, 512);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *slot;
    if (threadIdx.x < 32 && umma::elect_one()) {
        const uint32_t aW2 = smem_u32(sm), aH1 = aW2 + 16384, aDZ = aH1 + 32768, aH2 = aDZ + 32768, aOBS = aH2 + 32768, aDOUT = aOBS + 4096,
                       aW1B = aDOUT + 4096;
        constexpr uint32_t ID_FWD = umma::make_idesc(128, 64, false, false), ID_DH1 = umma::make_idesc(128, 64, false, true),
                           ID_W2 = umma::make_idesc(128, 128, true, true), ID_N16 = umma::make_idesc(128, 16, true, true);
        uint32_t phase = 0;
        for (int gk = 0; gk < NG; ++gk) {
            for (int rep = 0; rep < REP; ++rep) {
                const long long t0 = clock64();
                if (gk == 8) obs[rep * 4] = t0;
                if (gk == 0) {          // fwd: 2 nets x 4 K-blocks, M128 N64 K16, SW128 K-major
                    for (int n2 = 0; n2 < 2; ++n2)
                        for (int kb = 0; kb < 4; ++kb)
                            umma::mma(tmem + n2 * 64, umma::make_desc(aH1 + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                      umma::make_desc(aW2 + n2 * 8192 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, kb > 0);
                } else if (gk == 1) {   // dh1: B MN-major
                    for (int n2 = 0; n2 < 2; ++n2)
                        for (int kb = 0; kb < 4; ++kb)
                            umma::mma(tmem + 128 + n2 * 64, umma::make_desc(aDZ + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                      umma::make_desc(aW2 + n2 * 8192 + kb * 2048, 8192, 1024, umma::LAYOUT_SW128), ID_DH1, kb > 0);
                } else if (gk == 2) {   // dW2: M128 N128 K128, both MN-major
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 256, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aH1 + kb * 2048, 16384, 1024, umma::LAYOUT_SW128), ID_W2, kb > 0);
                } else if (gk == 3) {   // one N16 group (db2 / dW4 / dW1)
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 384, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aOBS + kb * 256, 128, 2048, umma::LAYOUT_NONE), ID_N16, kb > 0);
                } else if (gk == 4) {   // layer 1: 2 MMAs, K16, no swizzle
                    for (int n2 = 0; n2 < 2; ++n2)
                        umma::mma(tmem + n2 * 64, umma::make_desc(aOBS, 2048, 128, umma::LAYOUT_NONE),
                                  umma::make_desc(aW1B + n2 * 2048, 1024, 128, umma::LAYOUT_NONE), ID_FWD, 0u);
                } else if (gk == 5) {   // whole backward group: dh1 + dW2 + db2 + dW4 = 32 MMAs
                    for (int n2 = 0; n2 < 2; ++n2)
                        for (int kb = 0; kb < 4; ++kb)
                            umma::mma(tmem + 128 + n2 * 64, umma::make_desc(aDZ + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                      umma::make_desc(aW2 + n2 * 8192 + kb * 2048, 8192, 1024, umma::LAYOUT_SW128), ID_DH1, kb > 0);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 256, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aH1 + kb * 2048, 16384, 1024, umma::LAYOUT_SW128), ID_W2, kb > 0);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 384, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aOBS + kb * 256, 128, 2048, umma::LAYOUT_NONE), ID_N16, kb > 0);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 400, umma::make_desc(aH2 + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aDOUT + kb * 256, 128, 2048, umma::LAYOUT_NONE), ID_N16, kb > 0);
                } else if (gk == 6) {   // one single M128 N64 K16 MMA
                    umma::mma(tmem, umma::make_desc(aH1, 16, 1024, umma::LAYOUT_SW128), umma::make_desc(aW2, 16, 1024, umma::LAYOUT_SW128), ID_FWD, 0u);
                } else if (gk == 8) {   // pipeline order of the update kernel: l1 + dh1 | commit A | dW2 + db2 + dW4 | commit B | fwd | commit C
                    for (int n2 = 0; n2 < 2; ++n2)
                        umma::mma(tmem + n2 * 64, umma::make_desc(aOBS, 2048, 128, umma::LAYOUT_NONE),
                                  umma::make_desc(aW1B + n2 * 2048, 1024, 128, umma::LAYOUT_NONE), ID_FWD, 0u);
                    for (int n2 = 0; n2 < 2; ++n2)
                        for (int kb = 0; kb < 4; ++kb)
                            umma::mma(tmem + 128 + n2 * 64, umma::make_desc(aDZ + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                      umma::make_desc(aW2 + n2 * 8192 + kb * 2048, 8192, 1024, umma::LAYOUT_SW128), ID_DH1, kb > 0);
                    umma::commit(bar + 1);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 256, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aH1 + kb * 2048, 16384, 1024, umma::LAYOUT_SW128), ID_W2, kb > 0);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 384, umma::make_desc(aDZ + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aOBS + kb * 256, 128, 2048, umma::LAYOUT_NONE), ID_N16, kb > 0);
                    for (int kb = 0; kb < 8; ++kb)
                        umma::mma(tmem + 400, umma::make_desc(aH2 + kb * 2048, 16384, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aDOUT + kb * 256, 128, 2048, umma::LAYOUT_NONE), ID_N16, kb > 0);
                    umma::commit(bar + 2);
                    for (int n2 = 0; n2 < 2; ++n2)
                        for (int kb = 0; kb < 4; ++kb)
                            umma::mma(tmem + n2 * 64, umma::make_desc(aH1 + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                      umma::make_desc(aW2 + n2 * 8192 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, kb > 0);
                } else {                // nothing: commit + wait round trip only
                }
                const long long t1 = clock64();
                umma::commit(bar);
                long long ta = 0, tb = 0;
                if (gk == 8) {
                    mbar_wait(bar + 1, phase); ta = clock64() - t0;
                    mbar_wait(bar + 2, phase); tb = clock64() - t0;
                }
                mbar_wait(bar, phase);
                phase ^= 1u;
                const long long t2 = clock64();
                out[(gk * REP + rep) * 2] = gk == 8 ? ta * 100000 + tb : t1 - t0;
                out[(gk * REP + rep) * 2 + 1] = t2 - t0;
            }
        }
    }
    else if (threadIdx.x >= 32 && (threadIdx.x & 31) == 0) {
        // observers: warp w waits for commit A (w=1), B (w=2), C (w=3) of the pipeline experiment and stamps its completion
        const int w = threadIdx.x >> 5;
        uint64_t* b = w == 1 ? bar + 1 : (w == 2 ? bar + 2 : bar);
        uint32_t ph = 0;
        if (w == 3) for (int i = 0; i < 8 * REP; ++i) { mbar_wait(b, ph); ph ^= 1u; }   // skip the single-group experiments
        for (int rep = 0; rep < REP; ++rep) {
            mbar_wait(b, ph);
            ph ^= 1u;
            obs[rep * 4 + w] = clock64();
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) umma::tmem_dealloc(tmem, 512);
}

int main() {
    long long *d, *dobs;
    cudaMalloc(&d, NG * REP * 2 * sizeof(long long));
    cudaMalloc(&dobs, REP * 4 * sizeof(long long));
    const int smem = 163840 + 64 + 1024;
    cudaFuncSetAttribute(timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    timing_kernel<<<1, 128, smem>>>(d, dobs);
    long long h[NG * REP * 2];
    cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
    const char* names[NG] = {"fwd 8x(M128 N64 K16) K/K", "dh1 8x(M128 N64 K16) K/MN", "dW2 8x(M128 N128 K16) MN/MN", "N16 8x(M128 N16 K16) MN/MN",
                             "layer1 2x(M128 N64 K16) nosw", "bwd group (32 MMAs)", "single MMA", "empty commit", "pipeline (A*1e5+B / C)"};
    for (int gk = 0; gk < NG; ++gk) {
        printf("%-32s issue/complete cycles:", names[gk]);
        for (int rep = 0; rep < REP; ++rep) printf("  %lld/%lld", h[(gk * REP + rep) * 2], h[(gk * REP + rep) * 2 + 1]);
        printf("\n");
    }
    long long ho[REP * 4];
    cudaMemcpy(ho, dobs, sizeof(ho), cudaMemcpyDeviceToHost);
    printf("pipeline seen by observer warps (cycles after issue start): l1+dh1 commit / weight-grad commit / fwd commit\n");
    for (int rep = 0; rep < REP; ++rep) printf("   %lld / %lld / %lld\n", ho[rep * 4 + 1] - ho[rep * 4], ho[rep * 4 + 2] - ho[rep * 4], ho[rep * 4 + 3] - ho[rep * 4]);
    return 0;
}
