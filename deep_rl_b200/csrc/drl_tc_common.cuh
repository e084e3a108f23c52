// drl_tc_common.cuh -- pieces shared by the two tcgen05 kernels (update_tc.cu, rollout_tc.cu): thread geometry,
// named barriers, SFU tanh, half-row bf16 tile stores.
#pragma once
#include "drl_umma.cuh"

namespace drl {

constexpr int TC_COMPUTE = 512;              // 16 compute warps: thread = (row window rw, half, net)
constexpr int TC_THREADS = TC_COMPUTE + 32;  // + the MMA-issuer warp
constexpr int HU = 32;                       // hidden units per compute thread
constexpr int TC_TILE = 128;                 // samples / envs per tile = TMEM lanes

// tanh on the SFU (one MUFU op, |abs err| ~ 5e-4): below the bf16 rounding the activations get anyway
__device__ __forceinline__ float tanh_mufu(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// bar.sync / bar.arrive are warp-aligned instructions: every caller is a whole warp, and the __syncwarp() makes sure the warp has
// reconverged after lane-divergent code (live / padding rows, env physics branches) before it executes them
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
    __syncwarp();
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t count) {
    __syncwarp();
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}
// four 16-byte chunks (32 bf16) of row `row` of a SW128 tile, starting at chunk c0
__device__ __forceinline__ void store_half_row_sw128(unsigned char* tile, int row, int c0, const float (&v)[HU]) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        uint4 q;
        q.x = umma::pack_bf16(v[8 * c + 0], v[8 * c + 1]);
        q.y = umma::pack_bf16(v[8 * c + 2], v[8 * c + 3]);
        q.z = umma::pack_bf16(v[8 * c + 4], v[8 * c + 5]);
        q.w = umma::pack_bf16(v[8 * c + 6], v[8 * c + 7]);
        *reinterpret_cast<uint4*>(tile + umma::sw128_off(row, c0 + c)) = q;
    }
}

}  // namespace drl
