// drl_h256.cuh -- parameter layouts and constants of the 256-wide actor-critic (BASELINE config C5: the reference's
// ActorCritic of deep_rl/ppo.py:31-47 with hidden = 256 instead of 64).  At this width the two hidden layers are real
// contractions (2 x 256 x 256 MACs per net and sample), W2 alone is 128 KB in bf16 and its gradient fills the whole
// tensor memory of an SM, so the kernels differ from the 64-wide ones (update256.cu, rollout256.cu):
//   * one net per CTA (W2 of that net resident in shared memory for forward and backward),
//   * activations stream through four 16 KB K-chunks ([128 samples x 64 units] bf16, SW128),
//   * h1 and dz2 are staged in global memory (bf16, in the exact shared-memory image a 128-sample tile has), and a second
//     kernel -- a split-K tcgen05 GEMM -- turns them into dW2.
#pragma once
#include <cuda_bf16.h>

#include "drl_common.cuh"

namespace drl {
namespace h256 {

constexpr int HH = 256;              // hidden width
constexpr int NCH = 4;               // 64-unit chunks per layer
constexpr int SLOT_BYTES = 16384;    // one [128 x 64] bf16 SW128 chunk
constexpr int TILE_BYTES = NCH * SLOT_BYTES;   // one [128 x 256] activation tile (h1 or dz2) in its shared-memory image

// Packed (kernel) layout, in floats.  net 0 = actor, 1 = critic.
//   W2   : bf16 [2][4 chunks c][256 rows o][64 i]: SW128 K-major image, 16-byte unit u of row o stored at u ^ (o & 7)   (128 KB per net)
//   W1B  : bf16 [2][2 K-chunks][256 rows n][8]: K-major no-swizzle B operand of the layer-1 GEMM (K = 16):
//          k < O: W1[n][k], k == O: b1[n], k in [8, 8 + O): W1[n][k - 8] again (multiplies the low half of the observation)
//   B2   : fp32 [2][256]
//   W4   : fp32 [A + 1][256] (actor rows, then the critic row), B4 : fp32 [4] (actor biases, critic bias at [A])
//   W1F  : fp32 [2][256][OW] and B1F : fp32 [2][256], bf16-ROUNDED values (CUDA-core layer 1 of the rollout = the GEMM's operands)
template <int O, int A>
struct Packed256 {
    static constexpr int OW = O <= 4 ? 4 : 8;
    static constexpr int W2 = 0;
    static constexpr int W1B = W2 + 2 * HH * HH / 2;
    static constexpr int B2 = W1B + 2 * 2 * HH * 8 / 2;
    static constexpr int W4 = B2 + 2 * HH;
    static constexpr int B4 = W4 + (A + 1) * HH;
    static constexpr int W1F = B4 + 4;
    static constexpr int B1F = W1F + 2 * HH * OW;
    static constexpr int TOTAL = B1F + 2 * HH;
    // canonical (state_dict) layout
    static constexpr int C_NET = HH * O + HH + HH * HH + HH;      // trunk parameters per net
    static constexpr int C_ACTOR = C_NET + A * HH + A;
    static constexpr int C_ALL = C_ACTOR + C_NET + HH + 1;
    static constexpr int W2_OFF = HH * O + HH;                    // offset of W2 inside a net's canonical block
};

// byte offset of element (row o, column i) inside a net's W2 image
__host__ __device__ inline int w2_image_off(int o, int i) {
    const int c = i >> 6, u = (i & 63) >> 3;
    return c * (HH * 128) + o * 128 + ((u ^ (o & 7)) << 4) + (i & 7) * 2;
}

template <int O, int A>
__device__ __forceinline__ void packed_store256(float* __restrict__ packed, int idx, float v) {
    using P = Packed256<O, A>;
    const int net = idx >= P::C_ACTOR ? 1 : 0;
    int r = idx - net * P::C_ACTOR;
    const __nv_bfloat16 vb = __float2bfloat16_rn(v);
    if (r < HH * O) {
        const int o = r / O, k = r % O;
        __nv_bfloat16* w1b = reinterpret_cast<__nv_bfloat16*>(packed + P::W1B) + net * (2 * HH * 8);
        w1b[o * 8 + k] = vb;                  // K-chunk 0: multiplies obs_hi
        w1b[HH * 8 + o * 8 + k] = vb;         // K-chunk 1: multiplies obs_lo
        packed[P::W1F + (net * HH + o) * P::OW + k] = __bfloat162float(vb);
        return;
    }
    r -= HH * O;
    if (r < HH) {
        reinterpret_cast<__nv_bfloat16*>(packed + P::W1B)[net * (2 * HH * 8) + r * 8 + O] = vb;      // multiplies the ones column
        packed[P::B1F + net * HH + r] = __bfloat162float(vb);
        return;
    }
    r -= HH;
    if (r < HH * HH) {
        unsigned char* img = reinterpret_cast<unsigned char*>(packed + P::W2) + (size_t)net * (HH * HH * 2);
        *reinterpret_cast<__nv_bfloat16*>(img + w2_image_off(r / HH, r % HH)) = vb;
        return;
    }
    r -= HH * HH;
    if (r < HH) { packed[P::B2 + net * HH + r] = v; return; }
    r -= HH;
    const int nout = net == 0 ? A : 1;
    if (r < nout * HH) { packed[P::W4 + ((net == 0 ? 0 : A) + r / HH) * HH + r % HH] = v; return; }
    r -= nout * HH;
    packed[P::B4 + (net == 0 ? r : A)] = v;
}

// column of the [obs_hi | 1 | obs_lo] operand tile (16 columns) that carries output gradient `a` of the current net: the
// tile has free columns (their W1B rows are zero), so dW4 = h2^T . dout comes out of an N = 16 GEMM against the same tile
__host__ __device__ constexpr int dout_col(int O, int a) { return O <= 4 ? O + 1 + a : (a == 0 ? 7 : 13 + a); }

struct Grad256Args {
    const float* packed;
    const float* rec;          // [B][RW] sample records (gradient mode) or observations [n][OP] (forward mode)
    const uint32_t* idx;       // permutation (nullable)
    uint32_t mb_start, mb_count;      // this launch covers samples [mb_start, mb_start + mb_count) of the permuted order
    uint32_t mb_total;         // size of the whole minibatch (the mean's denominator)
    const float* adv_stats;
    float clip_coef, ent_coef, vf_coef;
    unsigned char* stage_h1;   // [2 nets][tiles][TILE_BYTES]
    unsigned char* stage_dz;
    uint32_t stage_tiles;      // tiles per net in the staging buffers
    float* grad_part;          // [grid][ppad], accumulated (read-modify-write, row owned by the CTA)
    float* loss_part;          // [grid][LOSS_TERMS], accumulated
    int ppad;
    float* logits_out;         // forward mode: [n][A]
    float* value_out;          // forward mode: [n]
    int only_net;              // -1: even CTAs the actor, odd CTAs the critic; 0 / 1: every CTA that net
    int first;                 // 1: first launch of this minibatch: the CTA WRITES its partial sums instead of adding to them
    long long* dbg;            // optional cycle stamps of CTA 0 (DRL_TC_DEBUG=1), else nullptr
};

constexpr int STAGE_TILES = 4096;   // 128-sample tiles per net the staging buffers hold (524,288 samples per pair of launches)

}  // namespace h256
}  // namespace drl
