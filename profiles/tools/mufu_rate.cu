// mufu_rate.cu -- throughput of the activation candidates on one SM: tanh.approx.f32, tanh.approx.bf16x2, tanh.approx.f16x2,
// ex2.approx.f32 + rcp.approx (the fp32-path tanh), and an FMA-pipe-only rational tanh.  One CTA of W warps, each thread runs
// a dependent-free stream of U independent chains; prints cycles per element per SM.
//   nvcc -arch=sm_100a -O3 -o /tmp/mufu_rate profiles/tools/mufu_rate.cu && /tmp/mufu_rate
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cstdio>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float x[8];
    for (int i = 0; i < 8; ++i) x[i] = 0.001f * (threadIdx.x + 1) + 0.1f * i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) {
                asm volatile("tanh.approx.f32 %0, %0;" : "+f"(x[i]));
            } else if (MODE == 1) {      // bf16x2: two elements per instruction
                unsigned u = __float_as_uint(x[i]);
                asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(u));
                x[i] = __uint_as_float(u);
            } else if (MODE == 2) {
                unsigned u = __float_as_uint(x[i]);
                asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(u));
                x[i] = __uint_as_float(u);
            } else if (MODE == 3) {      // ex2 + rcp
                float t, r;
                float y = x[i] * 2.885390081777927f;
                asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(y));
                float d = t + 1.0f;
                asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(d));
                x[i] = fmaf(-2.0f, r, 1.0f);
            } else {                     // FMA pipe only: odd rational minimax on [-4.97, 4.97] with a Newton reciprocal seeded by bit tricks
                float v = fminf(fmaxf(x[i], -4.97f), 4.97f);
                float s = v * v;
                float p = fmaf(s, fmaf(s, fmaf(s, 4.89352455891786e-03f * 0.0f + 2.0e-5f, 1.2e-3f), 5.1e-2f), 1.0f);     // placeholder degrees: cost model only
                float q = fmaf(s, fmaf(s, fmaf(s, 1.1e-4f, 6.3e-3f), 3.8e-1f), 1.0f);
                float r0 = __uint_as_float(0x7EF311C7u - __float_as_uint(q));
                r0 = r0 * fmaf(-q, r0, 2.0f);
                r0 = r0 * fmaf(-q, r0, 2.0f);
                x[i] = v * p * r0;
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 1 << 20); cudaMallocManaged(&cyc, 1024);
    const int iters = 2000;
    const char* names[5] = {"tanh.approx.f32", "tanh.approx.bf16x2 (2 elements/instr)", "tanh.approx.f16x2 (2 elements/instr)", "ex2+rcp tanh (fp32 path)", "FMA-pipe rational tanh"};
    for (int warps = 4; warps <= 16; warps *= 2)
        for (int mode = 0; mode < 5; ++mode) {
            auto run = [&](int m) {
                if (m == 0) k<0><<<1, warps * 32>>>(out, cyc, iters);
                if (m == 1) k<1><<<1, warps * 32>>>(out, cyc, iters);
                if (m == 2) k<2><<<1, warps * 32>>>(out, cyc, iters);
                if (m == 3) k<3><<<1, warps * 32>>>(out, cyc, iters);
                if (m == 4) k<4><<<1, warps * 32>>>(out, cyc, iters);
            };
            run(mode); cudaDeviceSynchronize();
            run(mode); cudaDeviceSynchronize();
            const double instr = (double)iters * 8 * warps * 32;
            const double elems = instr * ((mode == 1 || mode == 2) ? 2 : 1);
            printf("%2d warps  %-40s %8.3f cycles per 32 elements per SM  (%.1f elements/clk/SM)\n", warps, names[mode], cyc[0] / elems * 32, elems / cyc[0]);
        }
    return 0;
}
