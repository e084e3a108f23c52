// rollout_tc.cu -- tensor-core variant of the fused rollout (deep_rl/ppo.py:110-141) for large env counts.
// Same inputs, outputs, Philox streams and fp64 env physics as rollout_kernel (rollout.cu); the two 64x64
// hidden layers run on tcgen05 (bf16 operands, fp32 TMEM accumulators) with the same numerics as the tensor-core
// update (update_tc.cu), so log-probs recorded here and re-evaluated there agree to bf16 rounding.
//
// One CTA = 128 / REP environments for all T steps; 16 compute warps + 1 MMA-issuer warp.  The GEMM tile always has the
// 128 rows of the TMEM lanes; with REP > 1 every environment occupies REP rows (one per lane quadrant group), so that
// REP times as many threads share its per-step work -- with few environments per GPU the per-step latency, not the
// throughput, sets the pace, and a warp can only read the TMEM lanes of its own quadrant.
// Compute thread (warp w, lane l): quadrant q = w&3 (TMEM lanes [32q, 32q+32)), half = (w>>2)&1, net = w>>3;
//   env row  e = 32*(q mod 4/REP) + l,  replica = q div (4/REP),  hidden units [32*half + replica*32/REP, +32/REP).
// The threads with half = net = replica = 0 additionally own the env state (float64 registers).  Per step:
//   S0  owners: observation from state -> obs[t] in HBM and the fp32 obs tile in shared memory
//   --  group barrier (the 4*REP warps that share 32 environments)
//   S1  all: layer 1 on CUDA cores for 32/REP units -> bf16 SW128 tile (all REP rows of the env); hand the forward GEMM.
//       REP = 1 (throughput regime): layer 1 is a K = 16 GEMM as in update_tc.cu -- the owners write the bf16 [obs_hi|1|obs_lo]
//       row in S0, the issuer runs z1 = [obs_hi|1|obs_lo] . [W1|b1|W1]^T, S1 is tcgen05.ld + tanh + tile store
//   S2  all: wait, tcgen05.ld z2, tanh, partial head dot products -> exchange buffer
//   --  group barrier
//   ENV  four extra warps (one per 32 env rows, lane = env): the fp64 physics of the step for EVERY action, from the state the
//       owner published in S0 -- the physics depends on (state, action) only, so the candidates are computed while the compute
//       warps run the MLP (S1, GEMM, S2) instead of after the sample
//   S3  owners: logits / value, value store, Philox inverse-CDF sample, log-prob, pick the candidate of the sampled action,
//       TimeLimit / episode statistics / auto-reset, reward / done stores
#include "drl_env.cuh"
#include "drl_pack.cuh"
#include "drl_tc_common.cuh"

namespace drl {

int check_env(const drl_env_t* env);
drl_ep_log_t log_or_empty(const drl_ep_log_t* log);

// per-step cycle stamps of CTA 0 (owner warp and one other warp), only in a -DDRL_ROLLOUT_STAMPS build (profiles/tools/ro_stamps.py)
#ifdef DRL_ROLLOUT_STAMPS
__device__ long long g_ro_dbg[1024];
#define RO_STAMP(ev) do { if (blockIdx.x == 0 && lane == 0 && t >= 8 && t < 12 && (warp == 0 || warp == 5)) g_ro_dbg[(warp == 0 ? 0 : 512) + (t - 8) * 16 + (ev)] = clock64(); } while (0)
#else
#define RO_STAMP(ev) do { } while (0)
#endif
enum : uint32_t { RB_FWD = 1, RB_ROW0 = 2, RB_ROW1 = 6, RB_L1 = 10 };   // named barriers: issuer hand-off; per row group: state published (2..5), step evaluated (6..9);
                                                                        // REP = 1: the owners' [obs|1] rows are written (owners arrive, issuer syncs)

template <int O, int A>
struct RoTcSmem {
    using P = Packed<O, A>;
    static constexpr int W_BYTES = (P::TC_END - P::TC_W2) * 4;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_H1 = (OFF_W + W_BYTES + 1023) / 1024 * 1024;   // (actor, critic) x 16 KB
    static constexpr int OFF_OBS = OFF_H1 + 32768;                           // fp32 [128][OW]
    static constexpr int OBS_BYTES = TC_TILE * P::OW * 4 > 4096 ? TC_TILE * P::OW * 4 : 4096;   // REP = 1: the bf16 NS16 operand tile [128][16] instead
    static constexpr int OFF_XCH = OFF_OBS + OBS_BYTES;                      // fp32 [32 slots][128] head partial sums
    static constexpr int OFF_ST = OFF_XCH + 32 * TC_TILE * 4;                // published env state: double [128][4]
    static constexpr int OFF_CAND = OFF_ST + TC_TILE * 32;                   // candidates: double [3 actions][128][4] + {reward, term} [3][128]
    static constexpr int OFF_BAR = OFF_CAND + 3 * TC_TILE * 32 + 3 * TC_TILE * 8;
    static constexpr int TOTAL = OFF_BAR + 64 + 1024;
};

constexpr int RO_ENV_WARPS = 4;
constexpr int RO_THREADS = TC_THREADS + 32 * RO_ENV_WARPS;

template <int KIND, int REP>
__global__ void __launch_bounds__(RO_THREADS, 1) rollout_tc_kernel(drl_env_t env, const float* __restrict__ packed, int T,
                                                                   uint64_t step0, drl_rollout_buf_t buf, drl_ep_log_t log, const drl_ctrl_t* __restrict__ ctrl) {
    if (ctrl != nullptr) step0 = ctrl->env_step;      // graph-replayable launch: the counter lives in device memory
    using SP = EnvSpec<KIND>;
    constexpr int O = SP::O, A = SP::A, OP = SP::OP;
    using P = Packed<O, A>;
    using S = RoTcSmem<O, A>;
    constexpr int OW = P::OW;
    constexpr int ROWS = TC_TILE / REP;       // environments per CTA
    constexpr int QG = 4 / REP;               // lane quadrants per replica
    constexpr int UPT = HU / REP;             // hidden units per thread
    // Candidate physics for every action on the env warps pays when the physics is short next to the MLP (CartPole's Euler step,
    // MountainCar): Acrobot's RK4 step costs about as much as the whole MLP pass, three of them per env would become the
    // critical path, so there the owner evaluates the sampled action only.
    constexpr bool SPECULATE = KIND != DRL_ENV_ACROBOT;
    static_assert(OW == OP, "obs stride");
    static_assert(REP == 1 || REP == 2 || REP == 4, "replication factor");
    extern __shared__ unsigned char smem_raw[];
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* tW2 = sm + S::OFF_W;
    const float* sW1 = reinterpret_cast<const float*>(sm + S::OFF_W + 2 * H * H * 2);
    const float* sB1 = sW1 + 2 * H * OW;
    const float* sB2 = sB1 + 2 * H;
    const float* sW4 = sB2 + 2 * H;
    const float* sB4 = sW4 + (A + 1) * H;
    unsigned char* tH1 = sm + S::OFF_H1;
    float* obs_s = reinterpret_cast<float*>(sm + S::OFF_OBS);
    float* xch = reinterpret_cast<float*>(sm + S::OFF_XCH);
    double* st_s = reinterpret_cast<double*>(sm + S::OFF_ST);
    double* cand_s = reinterpret_cast<double*>(sm + S::OFF_CAND);
    float2* cand_rt = reinterpret_cast<float2*>(sm + S::OFF_CAND + 3 * TC_TILE * 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::OFF_BAR);   // 0 weights, 1 fwd, 2 layer 1 (REP = 1)
    uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 3);
    unsigned char* tW1B = reinterpret_cast<unsigned char*>(const_cast<float*>(sB4 + 4));   // bf16 layer-1 B tiles [W1|b1|W1] (drl_pack.cuh)
    unsigned char* tOBS = sm + S::OFF_OBS;                           // REP = 1: NS16 [obs_hi|1|obs_lo] rows

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_mma_warp = warp == TC_COMPUTE / 32;

    if (tid == 0) {
        mbar_init(bars, 1);
        mbar_init(bars + 1, 1);
        mbar_init(bars + 2, 1);
        mbar_fence_init();
    }
    if (is_mma_warp) umma::tmem_alloc(slot, 128);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (tid == 0) {
        mbar_expect_tx(bars, (uint32_t)S::W_BYTES);
        const char* src = reinterpret_cast<const char*>(packed + P::TC_W2);
        for (uint32_t off = 0; off < (uint32_t)S::W_BYTES; off += 16384u) {
            const uint32_t n = (uint32_t)S::W_BYTES - off < 16384u ? (uint32_t)S::W_BYTES - off : 16384u;
            bulk_g2s(sm + S::OFF_W + off, src + off, n, bars);
        }
    }
    const uint32_t tmem = *slot;
    mbar_wait(bars, 0);

    if (is_mma_warp) {
        const uint32_t aW2 = smem_u32(tW2), aH1 = smem_u32(tH1);
        constexpr uint32_t ID_FWD = umma::make_idesc(128, 64, false, false);
        const uint32_t aW1B = smem_u32(tW1B), aOBS = smem_u32(tOBS);
        constexpr uint32_t ID_L1 = umma::make_idesc(128, 64, false, false);
        for (int t = 0; t <= T; ++t) {
            if (REP == 1) {       // layer 1: both operands K-major without swizzle, K = 16
                named_bar_sync(RB_L1, 128 + 32);
                umma::fence_after_sync();
                if (umma::elect_one()) {
#pragma unroll
                    for (int n2 = 0; n2 < 2; ++n2)
                        umma::mma(tmem + n2 * 64, umma::make_desc(aOBS, 2048, 128, umma::LAYOUT_NONE),
                                  umma::make_desc(aW1B + n2 * 2048, 1024, 128, umma::LAYOUT_NONE), ID_L1, 0u);
                    umma::commit(bars + 2);
                }
                __syncwarp();
            }
            named_bar_sync(RB_FWD, TC_THREADS);
            umma::fence_after_sync();
            if (umma::elect_one()) {
#pragma unroll
                for (int n2 = 0; n2 < 2; ++n2)
#pragma unroll
                    for (int kb = 0; kb < 4; ++kb)
                        umma::mma(tmem + n2 * 64, umma::make_desc(aH1 + n2 * 16384 + kb * 32, 16, 1024, umma::LAYOUT_SW128),
                                  umma::make_desc(aW2 + n2 * 8192 + kb * 32, 16, 1024, umma::LAYOUT_SW128), ID_FWD, kb > 0);
                umma::commit(bars + 1);
            }
            __syncwarp();
        }
        umma::fence_before_sync();
        __syncthreads();
        umma::tmem_dealloc(tmem, 128);
        return;
    }

    if (warp > TC_COMPUTE / 32) {
        // =========================== env warps: warp e <-> env rows 32e .. 32e+31 (row group e) ===========================
        const int eg = warp - TC_COMPUTE / 32 - 1;
        if (eg < ROWS / 32) {
            const int er_e = eg * 32 + lane;
            const bool live = blockIdx.x * ROWS + er_e < env.num_envs;
            for (int t = 0; t <= T; ++t) {
                named_bar_sync(RB_ROW0 + eg, 128 * REP + 32);            // the owners have published the state of step t
                if (SPECULATE && t < T && live) {
                    const double2* sp = reinterpret_cast<const double2*>(st_s + er_e * 4);
                    const double2 s01 = sp[0], s23 = sp[1];
#pragma unroll
                    for (int a = 0; a < A; ++a) {
                        double cs[4] = {s01.x, s01.y, s23.x, s23.y};
                        float creward;
                        const bool cterm = env_physics<KIND>(cs, a, creward);
                        double2* cp = reinterpret_cast<double2*>(cand_s + ((size_t)a * TC_TILE + er_e) * 4);
                        cp[0] = make_double2(cs[0], cs[1]);
                        cp[1] = make_double2(cs[2], cs[3]);
                        cand_rt[a * TC_TILE + er_e] = make_float2(creward, cterm ? 1.0f : 0.0f);
                    }
                }
                named_bar_arrive(RB_ROW1 + eg, 128 * REP + 32);          // candidates ready (the owners sync on this after S2)
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        return;
    }

    // =========================== compute warps ===========================
    const int q = warp & 3, half = (warp >> 2) & 1, net = (warp >> 3) & 1;
    const int grp = q % QG, replica = q / QG;
    const int er = grp * 32 + lane;                             // env row inside the CTA
    const int u0 = half * HU + replica * UPT;
    const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);    // this warp's TMEM lane quadrant
    const int wrow0 = net == 0 ? 0 : A;
    const int nheads = net == 0 ? A : 1;
    const bool owner_warp = net == 0 && half == 0 && replica == 0;   // owns the env state
    const int N = env.num_envs;
    const int n = blockIdx.x * ROWS + er;
    const bool own = owner_warp && n < N;
    const uint32_t gid = env.env_gid0 + (uint32_t)n;
    constexpr int NSLOT = 2 * REP;                              // head partial sums per env and output

    EnvLane e;
    e.s[0] = e.s[1] = e.s[2] = e.s[3] = 0.0; e.elapsed = 0; e.ep_ret = 0.0f; e.ep_len = 0;
    if (own) env_load(e, env, n);
    float lp_diff = 0.0f, lp_sum = 1.0f;     // (l_act - max) and the softmax denominator of the previous step's draw

    for (int t = 0; t <= T; ++t) {
        // ---- S0: observation of the current state ----
        RO_STAMP(0);

        if (owner_warp) {
            float obs[OP];
#pragma unroll
            for (int i = 0; i < OP; ++i) obs[i] = 0.0f;
            if (own) {
                env_observation<KIND>(e.s, obs);
                float4* o4 = reinterpret_cast<float4*>(buf.obs + ((size_t)t * N + n) * OP);
#pragma unroll
                for (int qq = 0; qq < OP / 4; ++qq) o4[qq] = make_float4(obs[4 * qq], obs[4 * qq + 1], obs[4 * qq + 2], obs[4 * qq + 3]);
                double2* sp = reinterpret_cast<double2*>(st_s + er * 4);        // for the env warps' candidate steps
                sp[0] = make_double2(e.s[0], e.s[1]);
                sp[1] = make_double2(e.s[2], e.s[3]);
            }
            if (REP == 1) {
                // the update's layer-1 operand row [obs_hi | 1 | obs_lo] (update_tc.cu loader): the issuer runs the same K = 16 GEMM
                float o16[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) o16[i] = 0.0f;
#pragma unroll
                for (int i = 0; i < O; ++i) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(obs[i]));
                    o16[i] = hi;
                    o16[8 + i] = obs[i] - hi;
                }
                o16[O] = 1.0f;
                umma::store_row_ns16(tOBS, TC_TILE, er, o16);
                umma::fence_proxy_async();
                umma::fence_before_sync();
                named_bar_arrive(RB_L1, 128 + 32);
            } else {
                // layer-1 input as the update's GEMM sees it: obs = hi + lo with both halves rounded to bf16 (update_tc.cu loader)
#pragma unroll
                for (int i = 0; i < OP; ++i) {
                    const float hi = __bfloat162float(__float2bfloat16_rn(obs[i]));
                    obs[i] = hi + __bfloat162float(__float2bfloat16_rn(obs[i] - hi));
                }
#pragma unroll
                for (int qq = 0; qq < OP / 4; ++qq)
                    *reinterpret_cast<float4*>(obs_s + er * OW + 4 * qq) = make_float4(obs[4 * qq], obs[4 * qq + 1], obs[4 * qq + 2], obs[4 * qq + 3]);
            }
        }
        RO_STAMP(1);
        named_bar_sync(RB_ROW0 + grp, 128 * REP + 32);
        RO_STAMP(2);

        // ---- S1: layer 1 (UPT units of this thread's net) -> bf16 tile rows of every replica, hand the forward GEMM ----
        if (REP == 1) {
            mbar_wait(bars + 2, (uint32_t)t & 1u);
            umma::fence_after_sync();
            float h[UPT];
            umma::ldn<UPT>(trow + net * 64 + u0, h);
#pragma unroll
            for (int j = 0; j < UPT; ++j) h[j] = tanh_mufu(h[j]);
#pragma unroll
            for (int c = 0; c < UPT / 8; ++c) {
                uint4 ch;
                ch.x = umma::pack_bf16(h[8 * c + 0], h[8 * c + 1]);
                ch.y = umma::pack_bf16(h[8 * c + 2], h[8 * c + 3]);
                ch.z = umma::pack_bf16(h[8 * c + 4], h[8 * c + 5]);
                ch.w = umma::pack_bf16(h[8 * c + 6], h[8 * c + 7]);
                *reinterpret_cast<uint4*>(tH1 + net * 16384 + umma::sw128_off(q * 32 + lane, u0 / 8 + c)) = ch;
            }
        } else {
            float x[OW];
#pragma unroll
            for (int qq = 0; qq < OW / 4; ++qq) {
                const float4 v4 = *reinterpret_cast<const float4*>(obs_s + er * OW + 4 * qq);
                x[4 * qq] = v4.x; x[4 * qq + 1] = v4.y; x[4 * qq + 2] = v4.z; x[4 * qq + 3] = v4.w;
            }
            float h[UPT];
            const float* w1 = sW1 + (net * H + u0) * OW;
            const float* b1 = sB1 + net * H + u0;
#pragma unroll
            for (int k4 = 0; k4 < UPT / 4; ++k4) {
                const float4 bb = *reinterpret_cast<const float4*>(b1 + 4 * k4);
                const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int kk = 4 * k4 + j;
                    float z = bv[j];
#pragma unroll
                    for (int qq = 0; qq < OW / 4; ++qq) {
                        const float4 w = *reinterpret_cast<const float4*>(w1 + kk * OW + 4 * qq);
                        z = fmaf(x[4 * qq + 3], w.w, fmaf(x[4 * qq + 2], w.z, fmaf(x[4 * qq + 1], w.y, fmaf(x[4 * qq], w.x, z))));
                    }
                    h[kk] = tanh_mufu(z);
                }
            }
            uint4 ch[UPT / 8];
#pragma unroll
            for (int c = 0; c < UPT / 8; ++c) {
                ch[c].x = umma::pack_bf16(h[8 * c + 0], h[8 * c + 1]);
                ch[c].y = umma::pack_bf16(h[8 * c + 2], h[8 * c + 3]);
                ch[c].z = umma::pack_bf16(h[8 * c + 4], h[8 * c + 5]);
                ch[c].w = umma::pack_bf16(h[8 * c + 6], h[8 * c + 7]);
            }
#pragma unroll
            for (int rp = 0; rp < REP; ++rp) {      // the GEMM row of every replica of this env needs all 64 units
                const int row = (rp * QG + grp) * 32 + lane;
#pragma unroll
                for (int c = 0; c < UPT / 8; ++c)
                    *reinterpret_cast<uint4*>(tH1 + net * 16384 + umma::sw128_off(row, u0 / 8 + c)) = ch[c];
            }
        }
        RO_STAMP(3);
        umma::fence_proxy_async();
        umma::fence_before_sync();
        named_bar_arrive(RB_FWD, TC_THREADS);
        RO_STAMP(4);

        // ---- owners, while the GEMM runs: the pieces of S3 that do not depend on this step's logits.  The owner threads are the
        // critical path of a step (S3 -> S0 -> group barrier); here they would only spin on the mbarrier. ----
        uint4 rr = make_uint4(0u, 0u, 0u, 0u);
        if (own) {
            if (t > 0) buf.logp[(size_t)(t - 1) * N + n] = __fsub_rn(lp_diff, logf(lp_sum));   // log-prob of step t-1: (l_act - max) - log(sum)
            if (t < T) {
                const uint64_t step = step0 + (uint64_t)t;
                rr = philox_seeded(env.seed, gid, (uint32_t)step, (uint32_t)(step >> 32), TAG_ACTION);
            }
        }

        // ---- S2: layer-2 epilogue and partial head dot products ----
        mbar_wait(bars + 1, (uint32_t)t & 1u);
        umma::fence_after_sync();
        RO_STAMP(5);
        {
            float h[UPT];
            umma::ldn<UPT>(trow + net * 64 + u0, h);
            const float* b2 = sB2 + net * H + u0;
#pragma unroll
            for (int k4 = 0; k4 < UPT / 4; ++k4) {
                const float4 bb = *reinterpret_cast<const float4*>(b2 + 4 * k4);
                h[4 * k4 + 0] = tanh_mufu(h[4 * k4 + 0] + bb.x);
                h[4 * k4 + 1] = tanh_mufu(h[4 * k4 + 1] + bb.y);
                h[4 * k4 + 2] = tanh_mufu(h[4 * k4 + 2] + bb.z);
                h[4 * k4 + 3] = tanh_mufu(h[4 * k4 + 3] + bb.w);
            }
#pragma unroll
            for (int a = 0; a < A; ++a) {
                if (a < nheads) {
                    const float* w = sW4 + (wrow0 + a) * H + u0;
                    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                    for (int k4 = 0; k4 < UPT / 4; ++k4) {
                        const float4 ww = *reinterpret_cast<const float4*>(w + 4 * k4);
                        s0 = fmaf(h[4 * k4 + 0], ww.x, s0);
                        s1 = fmaf(h[4 * k4 + 1], ww.y, s1);
                        s2 = fmaf(h[4 * k4 + 2], ww.z, s2);
                        s3 = fmaf(h[4 * k4 + 3], ww.w, s3);
                    }
                    // slots: actor output a, part j = half*REP + replica -> j*A + a; critic part j -> NSLOT*A + j
                    const int j = half * REP + replica;
                    const int sl = net == 0 ? j * A + a : NSLOT * A + j;
                    xch[sl * TC_TILE + er] = (s0 + s1) + (s2 + s3);
                }
            }
        }
        RO_STAMP(6);
        umma::fence_before_sync();
        named_bar_sync(RB_ROW1 + grp, 128 * REP + 32);
        RO_STAMP(7);

        // ---- S3: owners: value store, sample, env step ----
        if (own) {
            const size_t i0 = (size_t)t * N + n;
            float v = xch[(NSLOT * A) * TC_TILE + er];
#pragma unroll
            for (int j = 1; j < NSLOT; ++j) v += xch[(NSLOT * A + j) * TC_TILE + er];
            buf.val[i0] = v + sB4[A];
            if (t < T) {
                float l[A];
#pragma unroll
                for (int a = 0; a < A; ++a) {
                    float sacc = xch[a * TC_TILE + er];
#pragma unroll
                    for (int j = 1; j < NSLOT; ++j) sacc += xch[(j * A + a) * TC_TILE + er];
                    l[a] = sacc + sB4[a];
                }
                if (buf.logits != nullptr) {
#pragma unroll
                    for (int a = 0; a < A; ++a) buf.logits[i0 * A + a] = l[a];
                }
                RO_STAMP(8);
                const uint64_t step = step0 + (uint64_t)t;
                RO_STAMP(9);
                const int act = sample_categorical_split<A>(l, u01_f32(rr.x), lp_diff, lp_sum);   // log-prob finished during the next GEMM wait
                RO_STAMP(10);
                buf.act[i0] = (uint8_t)act;
                float reward;
                bool term;
                if (SPECULATE) {          // the env warp has evaluated every action: take the sampled one
                    const double2* cp = reinterpret_cast<const double2*>(cand_s + ((size_t)act * TC_TILE + er) * 4);
                    const double2 c01 = cp[0], c23 = cp[1];
                    e.s[0] = c01.x; e.s[1] = c01.y; e.s[2] = c23.x; e.s[3] = c23.y;
                    const float2 rt = cand_rt[act * TC_TILE + er];
                    reward = rt.x; term = rt.y != 0.0f;
                } else {
                    term = env_physics<KIND>(e.s, act, reward);
                }
                const bool done = env_after_physics<KIND>(e, term, reward, env.seed, gid, step, env.max_episode_steps, log);
                buf.rew[i0 + N] = reward;
                buf.done[i0 + N] = done ? 1 : 0;
                RO_STAMP(11);
            }
        }
    }
    if (own) env_store(e, env, n);
    umma::fence_before_sync();
    __syncthreads();
}

template <int KIND, int REP>
static int launch_rollout_tc_rep(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                                 const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    using SP = EnvSpec<KIND>;
    const int smem = RoTcSmem<SP::O, SP::A>::TOTAL;
    DRL_CUDA(cudaFuncSetAttribute(rollout_tc_kernel<KIND, REP>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    constexpr int ROWS = TC_TILE / REP;
    const int blocks = (env.num_envs + ROWS - 1) / ROWS;
    rollout_tc_kernel<KIND, REP><<<blocks, RO_THREADS, smem, st>>>(env, packed, T, step0, buf, log, ctrl);
    DRL_LAUNCH_CHECK("rollout_tc_kernel");
    return DRL_OK;
}

template <int KIND>
static int launch_rollout_tc_kind(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                                  const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    // envs per CTA: 128 when that fills the GPU, otherwise fewer envs per CTA over more CTAs, each env spread over 2 or 4
    // GEMM rows (latency-bound regime)
    const int sms = sm_count();
    int rows = 128;
    if ((env.num_envs + 127) / 128 < sms) rows = (env.num_envs + 63) / 64 <= sms ? 64 : 128;
    if ((env.num_envs + 63) / 64 < sms && (env.num_envs + 31) / 32 <= sms) rows = 32;
    const char* ov = getenv("DRL_ROLLOUT_ROWS");
    if (ov) { const int v = atoi(ov); if (v == 32 || v == 64 || v == 128) rows = v; }
    if (rows == 32) return launch_rollout_tc_rep<KIND, 4>(env, packed, T, step0, buf, log, st, ctrl);
    if (rows == 64) return launch_rollout_tc_rep<KIND, 2>(env, packed, T, step0, buf, log, st, ctrl);
    return launch_rollout_tc_rep<KIND, 1>(env, packed, T, step0, buf, log, st, ctrl);
}

int launch_rollout_tc(const drl_env_t& env, const float* packed, int T, uint64_t step0, const drl_rollout_buf_t& buf,
                      const drl_ep_log_t& log, cudaStream_t st, const drl_ctrl_t* ctrl) {
    if (env.kind == DRL_ENV_CARTPOLE) return launch_rollout_tc_kind<DRL_ENV_CARTPOLE>(env, packed, T, step0, buf, log, st, ctrl);
    if (env.kind == DRL_ENV_MOUNTAINCAR) return launch_rollout_tc_kind<DRL_ENV_MOUNTAINCAR>(env, packed, T, step0, buf, log, st, ctrl);
    return launch_rollout_tc_kind<DRL_ENV_ACROBOT>(env, packed, T, step0, buf, log, st, ctrl);
}

}  // namespace drl

#ifdef DRL_ROLLOUT_STAMPS
extern "C" int drl_debug_rollout_stamps(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, drl::g_ro_dbg, sizeof(long long) * 1024) == cudaSuccess ? 0 : 1;
}
#endif
