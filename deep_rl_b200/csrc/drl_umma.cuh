// drl_umma.cuh -- thin wrappers over the sm_100a tensor-core path: tcgen05.mma (UMMA) with operands in
// shared memory, fp32 accumulators in tensor memory (TMEM), tcgen05.ld for the epilogue, mbarrier commits.
// Operand tiles use the canonical UMMA shared-memory layouts (bf16):
//   SW128 tile  : rows of 128 bytes (64 bf16), 8-row groups of 1024 bytes, 16-byte chunk c of row r stored at
//                 chunk position c ^ (r & 7).  Read as K-major (rows = M/N, row = 64 K elements) or as
//                 MN-major (rows = K, row = 64 M/N elements) -- the same bytes serve both, which is how one
//                 activation tile feeds the forward GEMM (K-major) and the weight-gradient GEMM (MN-major).
//   NS16 tile   : MN-major, no swizzle, 16 M/N elements: [2 chunks][K rows][8 bf16]; chunk stride = SBO,
//                 8-row group stride = LBO = 128 bytes.
#pragma once
#include <cuda_bf16.h>

#include "drl_common.cuh"

namespace drl {
namespace umma {

constexpr uint32_t LAYOUT_NONE = 0, LAYOUT_SW128 = 2;

__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | ((uint64_t)layout << 61);
}

// kind::f16, bf16 x bf16 -> fp32.  a_mn / b_mn: operand is MN-major (transposed) instead of K-major.
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// byte offset of 16-byte chunk `chunk` of row `row` inside a SW128 tile
__device__ __forceinline__ uint32_t sw128_off(int row, int chunk) { return (uint32_t)row * 128u + (uint32_t)((chunk ^ (row & 7)) << 4); }

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_slot, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {      // the same warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// true in exactly one lane of a converged warp; lets the compiler emit the tcgen05 sequence without a per-lane loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mma(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 8 consecutive 32-bit columns of this thread's TMEM lane, raw words (tcgen05.st / tcgen05.ld): tensor memory as a per-thread
// scratch pad for values that must survive a register-hungry phase
__device__ __forceinline__ void st8_raw(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void ld8_raw(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int N>
__device__ __forceinline__ void ldn(uint32_t taddr, float (&v)[N]) {
    static_assert(N == 8 || N == 16 || N == 32, "tcgen05.ld width");
    if constexpr (N == 8) ld8(taddr, v);
    else if constexpr (N == 16) ld16(taddr, v);
    else ld32(taddr, v);
}

__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}
__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); }

// store 64 fp32 values as one bf16 row of a SW128 tile (8 conflict-free 16-byte stores)
__device__ __forceinline__ void store_row_sw128(unsigned char* tile, int row, const float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        uint4 q;
        q.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
        q.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
        q.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
        q.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
        *reinterpret_cast<uint4*>(tile + sw128_off(row, c)) = q;
    }
}
__device__ __forceinline__ void load_row_sw128(const unsigned char* tile, int row, float (&v)[64]) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const uint4 q = *reinterpret_cast<const uint4*>(tile + sw128_off(row, c));
        v[8 * c + 0] = bf16_lo(q.x); v[8 * c + 1] = bf16_hi(q.x);
        v[8 * c + 2] = bf16_lo(q.y); v[8 * c + 3] = bf16_hi(q.y);
        v[8 * c + 4] = bf16_lo(q.z); v[8 * c + 5] = bf16_hi(q.z);
        v[8 * c + 6] = bf16_lo(q.w); v[8 * c + 7] = bf16_hi(q.w);
    }
}
// store 16 fp32 values as row `row` of an NS16 tile with `rows` K rows
__device__ __forceinline__ void store_row_ns16(unsigned char* tile, int rows, int row, const float (&v)[16]) {
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        uint4 q;
        q.x = pack_bf16(v[8 * c + 0], v[8 * c + 1]);
        q.y = pack_bf16(v[8 * c + 2], v[8 * c + 3]);
        q.z = pack_bf16(v[8 * c + 4], v[8 * c + 5]);
        q.w = pack_bf16(v[8 * c + 6], v[8 * c + 7]);
        *reinterpret_cast<uint4*>(tile + (size_t)c * rows * 16 + (size_t)row * 16) = q;
    }
}

}  // namespace umma
}  // namespace drl
