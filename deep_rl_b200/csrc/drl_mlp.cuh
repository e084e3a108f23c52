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

// Reduce part[e][a] over the 16 lanes of a half-warp by recursive halving: 8A shuffles instead of
// 32A.  On return lanes u (both parities) of each half hold the full sums for sample e = u >> 1.
template <int A>
__device__ __forceinline__ void half_warp_reduce(const float (&part)[TILE][A], int u, float (&out)[A]) {
    const bool hi3 = (u >> 3) & 1, hi2 = (u >> 2) & 1, hi1 = (u >> 1) & 1;
    float r1[4][A], r2[2][A];
#pragma unroll
    for (int a = 0; a < A; ++a) {
#pragma unroll
        for (int x = 0; x < 4; ++x) {
            float keep = hi3 ? part[4 + x][a] : part[x][a];
            float send = hi3 ? part[x][a] : part[4 + x][a];
            r1[x][a] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            float keep = hi2 ? r1[2 + x][a] : r1[x][a];
            float send = hi2 ? r1[x][a] : r1[2 + x][a];
            r2[x][a] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        float keep = hi1 ? r2[1][a] : r2[0][a];
        float send = hi1 ? r2[0][a] : r2[1][a];
        float r3 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        out[a] = r3 + __shfl_xor_sync(0xffffffffu, r3, 1);
    }
}

// Forward one 8-sample tile.
//   in : obs_s[i][e] filled and visible to the warp (caller did __syncwarp)
//   out: h1_s[net][k][e] = tanh(layer 1), h2[e][j] = tanh(layer 2) for this lane's units (registers),
//        out_s[e][0..A) = logits, out_s[e][3] = value; ends with __syncwarp().
template <int O, int A>
__device__ __forceinline__ void mlp_forward_tile(const float* __restrict__ sw, const float* __restrict__ obs_s,
                                                 float* __restrict__ h1_s, float* __restrict__ out_s, int lane,
                                                 float (&h2)[TILE][UPL]) {
    using P = Packed<O, A>;
    const int net = lane >> 4, u = lane & 15;
    float acc[TILE][UPL];

    // ---- layer 1: z1 = obs . W1^T + b1 ----
    {
        const float4 b = *reinterpret_cast<const float4*>(sw + P::B1 + net * H + 4 * u);
#pragma unroll
        for (int e = 0; e < TILE; ++e) { acc[e][0] = b.x; acc[e][1] = b.y; acc[e][2] = b.z; acc[e][3] = b.w; }
#pragma unroll
        for (int i = 0; i < O; ++i) {
            const float4 w = *reinterpret_cast<const float4*>(sw + P::W1T + (net * O + i) * H + 4 * u);
            const float4 x0 = *reinterpret_cast<const float4*>(obs_s + i * TILE);
            const float4 x1 = *reinterpret_cast<const float4*>(obs_s + i * TILE + 4);
            const float xs[TILE] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int e = 0; e < TILE; ++e) {
                acc[e][0] = fmaf(xs[e], w.x, acc[e][0]);
                acc[e][1] = fmaf(xs[e], w.y, acc[e][1]);
                acc[e][2] = fmaf(xs[e], w.z, acc[e][2]);
                acc[e][3] = fmaf(xs[e], w.w, acc[e][3]);
            }
        }
#pragma unroll
        for (int j = 0; j < UPL; ++j) {
            float* row = h1_s + (net * H + u + 16 * j) * TILE;
            *reinterpret_cast<float4*>(row) = make_float4(tanh_fast(acc[0][j]), tanh_fast(acc[1][j]),
                                                          tanh_fast(acc[2][j]), tanh_fast(acc[3][j]));
            *reinterpret_cast<float4*>(row + 4) = make_float4(tanh_fast(acc[4][j]), tanh_fast(acc[5][j]),
                                                              tanh_fast(acc[6][j]), tanh_fast(acc[7][j]));
        }
    }
    __syncwarp();

    // ---- layer 2: z2 = h1 . W2^T + b2 ----
    {
        const float4 b = *reinterpret_cast<const float4*>(sw + P::B2 + net * H + 4 * u);
#pragma unroll
        for (int e = 0; e < TILE; ++e) { acc[e][0] = b.x; acc[e][1] = b.y; acc[e][2] = b.z; acc[e][3] = b.w; }
        const float* wp = sw + P::W2T + net * H * H + 4 * u;
        const float* ap = h1_s + net * H * TILE;
#pragma unroll 8
        for (int k = 0; k < H; ++k) {
            const float4 w = *reinterpret_cast<const float4*>(wp + k * H);
            const float4 x0 = *reinterpret_cast<const float4*>(ap + k * TILE);
            const float4 x1 = *reinterpret_cast<const float4*>(ap + k * TILE + 4);
            const float xs[TILE] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
            for (int e = 0; e < TILE; ++e) {
                acc[e][0] = fmaf(xs[e], w.x, acc[e][0]);
                acc[e][1] = fmaf(xs[e], w.y, acc[e][1]);
                acc[e][2] = fmaf(xs[e], w.z, acc[e][2]);
                acc[e][3] = fmaf(xs[e], w.w, acc[e][3]);
            }
        }
#pragma unroll
        for (int e = 0; e < TILE; ++e)
#pragma unroll
            for (int j = 0; j < UPL; ++j) h2[e][j] = tanh_fast(acc[e][j]);
    }

    // ---- heads: logits = h2 . W4a^T + b4a (actor half), value = h2 . w4c + b4c (critic half) ----
    {
        float part[TILE][A];
#pragma unroll
        for (int a = 0; a < A; ++a) {
            const float4 w = *reinterpret_cast<const float4*>(sw + P::W4 + (net * A + a) * H + 4 * u);
#pragma unroll
            for (int e = 0; e < TILE; ++e)
                part[e][a] = fmaf(h2[e][3], w.w, fmaf(h2[e][2], w.z, fmaf(h2[e][1], w.y, h2[e][0] * w.x)));
        }
        float red[A];
        half_warp_reduce<A>(part, u, red);
        if ((u & 1) == 0) {
            const int e = u >> 1;
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
