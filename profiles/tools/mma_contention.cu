// mma_contention.cu -- does background work of the compute warps slow a tcgen05.mma group down?
// 16 "compute" warps run a background loop (none / tcgen05.ld / MUFU / st.shared / ld.shared / FFMA) while warp 16 issues
// the dh1 group (8 x M128 N64 K16) and the fwd group of ppo_grad_tc_kernel and measures issue -> completion.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I deep_rl_b200/csrc -o /tmp/mma_contention profiles/tools/mma_contention.cu && /tmp/mma_contention
#include <cstdio>
#include "drl_umma.cuh"

using namespace drl;
constexpr int NMODE = 13, REP = 4;

__global__ void __launch_bounds__(544) contention_kernel(int mode, long long* out, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 163840);
    uint32_t* slot = reinterpret_cast<uint32_t*>(sm + 163840 + 32);
    volatile int* stop = reinterpret_cast<volatile int*>(sm + 163840 + 64);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 163840 / 4; i += 544) {
        uint32_t v = 0x3c003c00u + i % 7;
        if (mode == 9) v = 0x322b322bu + (i % 5) * 0x00010001u;                       // ~1e-8: the magnitude of real dz2 values
        if (mode == 10) { v = (uint32_t)i * 2654435761u; v ^= v >> 13; v &= 0xBF7FBF7Fu; }   // random finite bf16 pairs
        if (mode == 11) v = 0x00010001u * (1 + i % 3);                                // bf16 denormals
        reinterpret_cast<uint32_t*>(sm)[i] = v;
    }
    if (tid == 0) { mbar_init(bar, 1); mbar_fence_init(); *stop = 0; }
    if (warp == 1) umma::tmem_alloc(slot, 512);
    umma::fence_proxy_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *slot;
    if (warp == 16) {
        if (umma::elect_one()) {
            const uint32_t aW2 = smem_u32(sm), aH1 = aW2 + 16384, aDZ = aH1 + 32768;
            constexpr uint32_t ID_FWD = umma::make_idesc(128, 64, false, false), ID_DH1 = umma::make_idesc(128, 64, false, true);
            uint32_t phase = 0;
            for (int i = 0; i < 2000; ++i) asm volatile("nanosleep.u32 20;");   // let the background loops get going
            for (int gk = 0; gk < 2; ++gk)
                for (int rep = 0; rep < REP; ++rep) {
                    const long long t0 = clock64();
                    *reinterpret_cast<volatile long long*>(sm + 163840 + 72) = t0;
                    if (gk == 0) {
                        for (int n2 = 0; n2 < 2; ++n2)
                            for (int kb = 0; kb < 4; ++kb)
                                umma::mma(tmem + n2 * 64, umma::make_desc(aH1 + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                          umma::make_desc(aW2 + n2 * 8192 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, kb > 0);
                    } else {
                        for (int n2 = 0; n2 < 2; ++n2)
                            for (int kb = 0; kb < 4; ++kb)
                                umma::mma(tmem + 128 + n2 * 64, umma::make_desc(aDZ + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                          umma::make_desc(aW2 + n2 * 8192 + kb * 2048, 8192, 1024, umma::LAYOUT_SW128), ID_DH1, kb > 0);
                    }
                    const long long t1 = clock64();
                    umma::commit(bar);
                    mbar_wait(bar, phase);
                    phase ^= 1u;
                    const long long t2 = clock64();
                    out[(gk * REP + rep) * 2] = t1 - t0;
                    out[(gk * REP + rep) * 2 + 1] = t2 - t0;
                    for (int i = 0; i < 50; ++i) asm volatile("nanosleep.u32 20;");
                }
            *stop = 1;
        }
    } else {
        // background load of the 16 compute warps
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        float acc = (float)tid;
        unsigned char* scratch = sm + 98304 + (size_t)tid * 16;     // 8 KB region away from the operand tiles
        if (mode == 12) {      // all 512 compute threads wait on the commit barrier, like Z(k) of the update kernel
            volatile long long* t0s = reinterpret_cast<volatile long long*>(sm + 163840 + 72);
            uint32_t ph = 0;
            for (int i = 0; i < 2 * REP; ++i) {
                mbar_wait(bar, ph);
                ph ^= 1u;
                const long long t = clock64();
                if (tid == 0) out[16 + i] = t - *t0s;
                if (tid == 511) out[32 + i] = t - *t0s;
            }
        }
        while (*stop == 0) {
            if (mode == 1) {
                float v[32];
                umma::ld32(trow + 256 + (warp >> 2) * 32, v);
                acc += v[0] + v[31];
            } else if (mode == 2) {
#pragma unroll
                for (int i = 0; i < 32; ++i) { float y; asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(acc + i)); acc = y; }
            } else if (mode == 3) {
#pragma unroll
                for (int i = 0; i < 8; ++i) asm volatile("st.shared.v4.u32 [%0], {%1, %2, %1, %2};" ::"r"(smem_u32(scratch + (i & 3) * 8704)), "r"(tid), "r"(i) : "memory");
            } else if (mode == 4) {
#pragma unroll
                for (int i = 0; i < 8; ++i) { uint32_t q0, q1, q2, q3; asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(smem_u32(scratch + (i & 3) * 8704)) : "memory"); acc += (float)(q0 + q3); }
            } else if (mode == 5) {
#pragma unroll
                for (int i = 0; i < 64; ++i) acc = fmaf(acc, 1.0001f, 0.5f);
            } else if (mode == 6 || (mode == 7 && (warp & 3) != 0) || (mode == 8 && (warp & 3) == 0)) {
                // 32 INDEPENDENT tanh per thread, like the layer epilogues of the update kernel (saturates the MUFU queue)
                float y[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) asm volatile("tanh.approx.f32 %0, %1;" : "=f"(y[i]) : "f"(acc + i));
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += y[i];
            } else {
                asm volatile("nanosleep.u32 100;");
            }
        }
        if (acc == 12345.678f) sink[tid] = acc;
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; float* sink;
    cudaMalloc(&d, 64 * sizeof(long long));
    cudaMalloc(&sink, 544 * 4);
    const int smem = 163840 + 128 + 1024;
    cudaFuncSetAttribute(contention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const char* names[NMODE] = {"idle (nanosleep)", "tcgen05.ld loop", "MUFU tanh loop", "st.shared.v4 loop", "ld.shared.v4 loop", "FFMA loop", "MUFU burst, all warps", "MUFU burst, SMSP 1-3", "MUFU burst, SMSP 0", "idle, operands ~1e-8", "idle, random operands", "idle, denormal operands", "512 threads wait on the barrier"};
    for (int mode = 0; mode < NMODE; ++mode) {
        contention_kernel<<<1, 544, smem>>>(mode, d, sink);
        long long h[64];
        cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error: %s\n", cudaGetErrorString(e)); return 1; }
        printf("%-20s fwd issue/complete:", names[mode]);
        for (int rep = 0; rep < REP; ++rep) printf(" %lld/%lld", h[rep * 2], h[rep * 2 + 1]);
        printf("   dh1:");
        for (int rep = 0; rep < REP; ++rep) printf(" %lld/%lld", h[(REP + rep) * 2], h[(REP + rep) * 2 + 1]);
        printf("\n");
        if (mode == 12) {
            printf("   seen by waiting compute thread 0 / 511:");
            for (int i = 0; i < 2 * REP; ++i) printf(" %lld/%lld", h[16 + i], h[32 + i]);
            printf("\n");
        }
    }
    return 0;
}
