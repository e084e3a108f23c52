// drl_mlp.cuh -- warp-tile forward pass of the two tanh MLPs of ActorCritic (deep_rl/ppo.py:31-54)
// on FP32 CUDA cores.  One warp = one tile of 8 samples; lanes 0-15 carry the actor trunk, lanes
// 16-31 the critic trunk; each lane owns 4 hidden units (u + 16 j) x 8 samples = 32 accumulators.
// Weights come from the packed layout staged in shared memory (drl_common.cuh: Packed<O,A>).
#pragma once
#include "drl_common.cuh"

namespace drl {

constexpr int OBS_S = 64;    // floats: obs tile  [i < 8][e < 8]
constexpr int H1_S = 1024;   // floats: activation tile [net][k < 64][e < 8]
constexpr int OUT_S = 32;    // floats: [e < 8][4] = logits[0..A), value at [3]
constexpr int OUT_W = 4;

// Reduce part[e][a] over the 16 lanes of a half-warp by recursive halving (8A shuffles for an 8-sample tile
// instead of 32A).  On return every lane of the half holds the full sums for sample e = u >> (TILE == 8 ? 1 : 2).
template <int A, int TL>
__device__ __forceinline__ void half_warp_reduce(const float (&part)[TL][A], int u, float (&out)[A]) {
    static_assert(TL == 8 || TL == 4, "tile of 4 or 8 samples");
    const bool hi3 = (u >> 3) & 1, hi2 = (u >> 2) & 1, hi1 = (u >> 1) & 1;
#pragma unroll
    for (int a = 0; a < A; ++a) {
        if constexpr (TL == 8) {
            float r1[4], r2[2];
#pragma unroll
            for (int x = 0; x < 4; ++x) {
                float keep = hi3 ? part[4 + x][a] : part[x][a];
                float send = hi3 ? part[x][a] : part[4 + x][a];
                r1[x] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                float keep = hi2 ? r1[2 + x] : r1[x];
                float send = hi2 ? r1[x] : r1[2 + x];
                r2[x] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            float keep = hi1 ? r2[1] : r2[0];
            float send = hi1 ? r2[0] : r2[1];
            float r3 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
            out[a] = r3 + __shfl_xor_sync(0xffffffffu, r3, 1);
        } else {
            float r1[2];
#pragma unroll
            for (int x = 0; x < 2; ++x) {
                float keep = hi3 ? part[2 + x][a] : part[x][a];
                float send = hi3 ? part[x][a] : part[2 + x][a];
                r1[x] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            float keep = hi2 ? r1[1] : r1[0];
            float send = hi2 ? r1[0] : r1[1];
            float r2 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            float r3 = r2 + __shfl_xor_sync(0xffffffffu, r2, 2);
            out[a] = r3 + __shfl_xor_sync(0xffffffffu, r3, 1);
        }
    }
}

// Forward one TL-sample tile (TL = 8, or 4 when few envs must be spread over many warps).
//   in : obs_s[i][e] filled and visible to the warp (caller did __syncwarp)
//   out: h1_s[net][k][e] = tanh(layer 1), h2[e][j] = tanh(layer 2) for this lane's units (registers),
//        out_s[e][0..A) = logits, out_s[e][3] = value; ends with __syncwarp().
template <int O, int A, int TL = TILE>
__device__ __forceinline__ void mlp_forward_tile(const float* __restrict__ sw, const float* __restrict__ obs_s,
                                                 float* __restrict__ h1_s, float* __restrict__ out_s, int lane,
                                                 float (&h2)[TL][UPL]) {
    using P = Packed<O, A>;
    const int net = lane >> 4, u = lane & 15;
    float acc[TL][UPL];

    // ---- layer 1: z1 = obs . W1^T + b1 ----
    {
        const float4 b = *reinterpret_cast<const float4*>(sw + P::B1 + net * H + 4 * u);
#pragma unroll
        for (int e = 0; e < TL; ++e) { acc[e][0] = b.x; acc[e][1] = b.y; acc[e][2] = b.z; acc[e][3] = b.w; }
#pragma unroll
        for (int i = 0; i < O; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(sw + P::W1T + (net * O + i) * H + 4 * u);
            float xs[TL];
#pragma unroll
            for (int q = 0; q < TL / 4; ++q) {
                const float4 x4 = *reinterpret_cast<const float4*>(obs_s + i * TL + 4 * q);
                xs[4 * q] = x4.x; xs[4 * q + 1] = x4.y; xs[4 * q + 2] = x4.z; xs[4 * q + 3] = x4.w;
            }
#pragma unroll
            for (int e = 0; e < TL; ++e) {
                acc[e][0] = fmaf(xs[e], w.x, acc[e][0]);
                acc[e][1] = fmaf(xs[e], w.y, acc[e][1]);
                acc[e][2] = fmaf(xs[e], w.z, acc[e][2]);
                acc[e][3] = fmaf(xs[e], w.w, acc[e][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            float* row = h1_s + (net * H + u + 16 * j) * TL;
#pragma unroll
            for (int q = 0; q < TL / 4; ++q)
                *reinterpret_cast<float4*>(row + 4 * q) = make_float4(tanh_fast(acc[4 * q][j]), tanh_fast(acc[4 * q + 1][j]),
                                                                      tanh_fast(acc[4 * q + 2][j]), tanh_fast(acc[4 * q + 3][j]));
        }
    }
    __syncwarp();

    // ---- layer 2: z2 = h1 . W2^T + b2 ----
    {
        const float4 b = *reinterpret_cast<const float4*>(sw + P::B2 + net * H + 4 * u);
#pragma unroll
        for (int e = 0; e < TL; ++e) { acc[e][0] = b.x; acc[e][1] = b.y; acc[e][2] = b.z; acc[e][3] = b.w; }
        const float* wp = sw + P::W2T + net * H * H + 4 * u;
        const float* ap = h1_s + net * H * TL;
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(wp + k * H);
            float xs[TL];
#pragma unroll
            for (int q = 0; q < TL / 4; ++q) {
                const float4 x4 = *reinterpret_cast<const float4*>(ap + k * TL + 4 * q);
                xs[4 * q] = x4.x; xs[4 * q + 1] = x4.y; xs[4 * q + 2] = x4.z; xs[4 * q + 3] = x4.w;
            }
#pragma unroll
            for (int e = 0; e < TL; ++e) {
                acc[e][0] = fmaf(xs[e], w.x, acc[e][0]);
                acc[e][1] = fmaf(xs[e], w.y, acc[e][1]);
                acc[e][2] = fmaf(xs[e], w.z, acc[e][2]);
                acc[e][3] = fmaf(xs[e], w.w, acc[e][3]);
            }
        }
#pragma unroll
        for (int e = 0; e < TL; ++e)
#pragma unroll
            for (int j = 0; j < UPL; ++j) h2[e][j] = tanh_fast(acc[e][j]);
    }

    // ---- heads: logits = h2 . W4a^T + b4a (actor half), value = h2 . w4c + b4c (critic half) ----
    {
        float part[TL][A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float4 w = *reinterpret_cast<const float4*>(sw + P::W4 + (net * A + a) * H + 4 * u);
#pragma unroll
            for (int e = 0; e < TL; ++e)
                part[e][a] = fmaf(h2[e][3], w.w, fmaf(h2[e][2], w.z, fmaf(h2[e][1], w.y, h2[e][0] * w.x)));
        }
        float red[A];
        half_warp_reduce<A, TL>(part, u, red);
        constexpr int SH = TL == 8 ? 1 : 2;   // lanes per sample after the reduction: 2 or 4
        if ((u & ((1 << SH) - 1)) == 0) {
            const int e = u >> SH;
            if (net == 0) {
#pragma unroll
                for (int a = 0; a < A; ++a) out_s[e * OUT_W + a] = red[a] + sw[P::B4 + a];
            } else {
                out_s[e * OUT_W + 3] = red[0] + sw[P::B4 + A];
            }
        }
    }
    __syncwarp();
}

}  // namespace drl
