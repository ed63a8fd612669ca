// Persistent recurrent kernels of the encoder's training step (batch <= 32 rows per GPU).
//
// The bidirectional GRU's time loop (models/Encoder.py:55-60, nn.GRU over the packed source batch) is a chain of T tiny steps —
// 32 x 512 x 1536 each — whose cost as one kernel launch per step is the launch itself: 9 us per step forward, 14 us backward,
// of which the arithmetic is well under 1 us.  Here ONE launch walks all T steps of both directions:
//   * grid (H / 8, 2): a CTA owns 8 hidden units of one direction for the whole sequence and keeps ITS slice of the recurrent
//     matrix in shared memory (forward: 24 rows of W_hh, k-fast; backward: 8 columns of W_hh, all 3H rows) — the weights are read
//     from L2 once per sequence instead of once per step;
//   * per step the CTAs of a direction exchange the state (forward: h_t, 32 x H; backward: dgh_t, 32 x 3H) through a global
//     exchange buffer laid out exactly as the consumers want it in shared memory ([32][K + 4] floats, already rounded to bf16 in
//     bf16 mode), so a step's input is one linear, fully coalesced copy (ld.global.cg: L1 is never consulted), and meet at a flag
//     barrier: every CTA publishes its step counter with st.release, one warp polls its peers' flags with ld.acquire — no atomic,
//     whose 64 same-address operations serialise in L2.  All CTAs are co-resident by construction (2·H/8 <= number of SMs, one
//     CTA per SM — checked by the launcher, which otherwise reports "unsupported" and the caller keeps the per-step kernels);
//   * the 32 x K x {24, 8} product of a step runs on the tensor cores like linear_rows32_kernel's inner loop (mma.sync m16n8k8
//     TF32: one product of bf16-rounded operands in bf16 mode, error-compensated 3xTF32 in FP32 mode), the K range dealt to the
//     warps, partial sums met in shared memory, and the GRU cell (forward) or its derivative (backward) finished by one thread
//     per (row, unit) — which keeps the previous state / the BPTT carry dh·z in a register across the steps.
// A poll that does not see its peers within ~1 s gives up: the CTA finishes the sequence without waiting (never a hang) and
// poisons its outputs with NaN, so a broken exchange cannot pass for a result.
#include "enc_seq.cuh"
#include "seq_common.cuh"
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

namespace vag {
namespace {

constexpr int ES_UNITS = 8;           // hidden units per CTA
#ifndef ES_FWD_WARPS_N
#define ES_FWD_WARPS_N 8
#endif
constexpr int ES_FWD_WARPS = ES_FWD_WARPS_N;   // forward: K = H dealt to the warps
constexpr int ES_BWD_WARPS = 16;      // backward: K = 3H dealt to 16 warps
constexpr int ES_FWD_RPW = 32 / ES_FWD_WARPS;  // forward: rows of the state tile each warp copies
constexpr int ES_FWD_MAXLD = 5 * ES_FWD_RPW;   // forward: float4 loads per lane and step (<= 160 float4 per row)
constexpr int ES_BWD_MAXLD = 14;      // backward: float4 loads per lane and chunk (2 rows per warp, <= 224 float4 per row)

#ifdef ES_PROFILE
#define ES_T(i) do { if (prof) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define ES_T(i) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------------------------
template <bool RB>
__global__ void __launch_bounds__(ES_FWD_WARPS * 32, 1) enc_seq_fwd_kernel(const EncSeqFwd a) {
    extern __shared__ __align__(16) float es_smem[];
    const int d = blockIdx.y, u0 = blockIdx.x * ES_UNITS, n_cta = gridDim.x;
    const int H = a.H, B = a.B, T = a.T;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int P = H + 4;                                    // pitch of weights, state tile and exchange buffer: conflict-free fragments
    const int KW = a.kw;                                    // contraction indices per warp
    float* ws = es_smem;                                    // [24][P]   row n = gate·8 + unit
    float* xs = ws + 24 * P;                                // [32][P]   the previous state of this direction
    float* part = xs + 32 * P;                              // [8 warps][32][25]
    const float* w_hh = a.w_hh[d];
    for (int i = tid; i < 24 * (H / 4); i += blockDim.x) {
        const int n = i / (H / 4), q = i - n * (H / 4);
        float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + (int64_t)((n >> 3) * H + u0 + (n & 7)) * H + 4 * q));
        if (RB) { v.x = es_rbf16(v.x); v.y = es_rbf16(v.y); v.z = es_rbf16(v.z); v.w = es_rbf16(v.w); }
        *reinterpret_cast<float4*>(ws + n * P + 4 * q) = v;
    }
    for (int n = tid; n < 24; n += blockDim.x) *reinterpret_cast<float4*>(ws + n * P + H) = make_float4(0.f, 0.f, 0.f, 0.f);
    // epilogue thread = (row, unit)
    const int erow = tid >> 3, eu = tid & 7;
    const bool ework = tid < 256 && erow < B;
    float bias[3] = {0.f, 0.f, 0.f};
    int len = 0;
    if (ework) {
        len = a.lengths[erow];
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) bias[gt] = a.b_hh[d] ? a.b_hh[d][gt * H + u0 + eu] : 0.f;
    }
    int* flags = a.bar + d * ES_MAX_CTAS;
    __shared__ int gave_up;
    if (tid == 0) gave_up = 0;
    __syncthreads();
    const int k_lo = wid * KW, k_n = max(0, min(KW, H - k_lo));     // this warp's contraction range
    const int qrow = P >> 2;                                         // float4 per row of the exchange buffer
    float hp = 0.f;                                                  // the state this thread's (row, unit) carries through time
#ifdef ES_PROFILE
    const bool prof = tid == 0 && blockIdx.x == 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? s : T - 1 - s;
        const float* x_in = a.xch + ((size_t)(s & 1) * 2 + d) * 32 * P;
        float* x_out = a.xch + ((size_t)((s + 1) & 1) * 2 + d) * 32 * P;
        const int64_t o3 = (((int64_t)d * T + t) * B + erow) * 3 * H + u0 + eu;
        float gi[3] = {0.f, 0.f, 0.f};
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gi[gt] = __ldg(a.gi + o3 + gt * H);
        }
        float dacc[3][2][4];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 4; ++u) dacc[nt][mt][u] = 0.f;
        if (s > 0) {                                        // step 0 starts from the zero state: nothing to wait for or to contract
            es_wait(flags, n_cta, s, &gave_up);
            ES_T(0);
            // the whole [32][P] state of this direction, ES_FWD_RPW rows per warp, every load issued before the first store
            float4 v[ES_FWD_MAXLD];
#pragma unroll
            for (int i = 0; i < ES_FWD_MAXLD; ++i) {
                const int r = i / 5, q = (i % 5) * 32 + lane;
                if (q < qrow) v[i] = __ldcg(reinterpret_cast<const float4*>(x_in + (wid * ES_FWD_RPW + r) * P) + q);
            }
            ES_T(5);
#pragma unroll
            for (int i = 0; i < ES_FWD_MAXLD; ++i) {
                const int r = i / 5, q = (i % 5) * 32 + lane;
                if (q < qrow) reinterpret_cast<float4*>(xs + (wid * ES_FWD_RPW + r) * P)[q] = v[i];
            }
            __syncthreads();
            ES_T(6);
#pragma unroll 4
            for (int ks = 0; ks < (k_n >> 3); ++ks) {
                const int kk = k_lo + ks * 8 + t4;
                uint32_t ah[2][4], al[2][4];
                es_load_a<RB>(ah, al, xs, P, kk, g);
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    const float* wb = ws + (nt * 8 + g) * P + kk;
                    const float bf[2] = {wb[0], wb[4]};
                    uint32_t bh[2], bl[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        bh[u] = RB ? __float_as_uint(bf[u]) : es_tf32(bf[u]);
                        bl[u] = RB ? 0u : es_tf32(bf[u] - __uint_as_float(bh[u]));
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (!RB) {
                            es_mma(dacc[nt][mt], al[mt], bh);
                            es_mma(dacc[nt][mt], ah[mt], bl);
                        }
                        es_mma(dacc[nt][mt], ah[mt], bh);
                    }
                }
            }
        }
        ES_T(1);
        // partial sums → shared memory ([warp][row][25])
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    part[(wid * 32 + mt * 16 + g + (u >> 1) * 8) * 25 + nt * 8 + 2 * t4 + (u & 1)] = dacc[nt][mt][u];
        __syncthreads();
        ES_T(2);
        float sv[3] = {0.f, 0.f, 0.f};
        if (ework) {
            float pre[3];
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < ES_FWD_WARPS; ++w) v += part[(w * 32 + erow) * 25 + gt * 8 + eu];
                pre[gt] = v + bias[gt];
            }
            const bool live = len > t;
            if (gave_up) pre[0] = __int_as_float(0x7fc00000);
            if (live) {
                const float r = sigmoidf_precise(gi[0] + pre[0]);
                const float z = sigmoidf_precise(gi[1] + pre[1]);
                const float n = tanhf(gi[2] + r * pre[2]);
                hp = (1.0f - z) * n + z * hp;
            }
            x_out[erow * P + u0 + eu] = RB ? es_rbf16(hp) : hp;     // masked rows pass their state on unchanged
            sv[0] = live ? pre[0] : 0.f; sv[1] = live ? pre[1] : 0.f; sv[2] = live ? pre[2] : 0.f;
        }
        ES_T(3);
        // the release only has to cover the exchanged state: what later KERNELS read (gh, ctx_out) is stored after it, in the shadow
        // of the next step's wait
        if (s + 1 < T) es_arrive(flags + blockIdx.x, s + 1);          // also the CTA barrier that frees xs / part for the next step
        if (ework) {
            if (len > t) a.ctx_out[((int64_t)erow * T + t) * 2 * H + (int64_t)d * H + u0 + eu] = hp;
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) a.gh[o3 + gt * H] = sv[gt];
        }
        ES_T(4);
    }
#ifdef ES_PROFILE
    if (prof) printf("fwd d%d T%d: wait %lld  issue %lld  land+store %lld  mma %lld  reduce-sync %lld  epilogue %lld  arrive %lld (cycles, summed)\n", d, T, tacc[0], tacc[5], tacc[6], tacc[1], tacc[2], tacc[3], tacc[4]);
#endif
}

// ------------------------------------------------------------------------------------------------------------------------------
// backward (BPTT)
// ------------------------------------------------------------------------------------------------------------------------------
template <bool RB>
__global__ void __launch_bounds__(ES_BWD_WARPS * 32, 1) enc_seq_bwd_kernel(const EncSeqBwd a) {
    extern __shared__ __align__(16) float es_smem[];
    const int d = blockIdx.y, u0 = blockIdx.x * ES_UNITS, n_cta = gridDim.x;
    const int H = a.H, B = a.B, T = a.T, K = 3 * H;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int P = K + 4;                                    // exchange-buffer pitch
    const int n_chunk = a.n_chunk, CW = K / n_chunk, CP = CW + 4;   // the state is staged in n_chunk column ranges of CW
    const int KW = a.kw;                                    // contraction indices per warp and chunk (multiple of 8)
    float* wt = es_smem;                                    // [3H][8]: W_hh[k][u0 + n]
    float* xs = wt + (size_t)K * 8;                         // [32][CP]
    float* part = xs + 32 * CP;                             // [16 warps][32][9]
    const float* w_hh = a.w_hh[d];
    for (int i = tid; i < K * 2; i += blockDim.x) {
        const int k = i >> 1, q = i & 1;
        float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + (int64_t)k * H + u0 + 4 * q));
        if (RB) { v.x = es_rbf16(v.x); v.y = es_rbf16(v.y); v.z = es_rbf16(v.z); v.w = es_rbf16(v.w); }
        *reinterpret_cast<float4*>(wt + k * 8 + 4 * q) = v;
    }
    for (int r = tid; r < 32; r += blockDim.x) *reinterpret_cast<float4*>(xs + r * CP + CW) = make_float4(0.f, 0.f, 0.f, 0.f);
    const int erow = tid >> 3, eu = tid & 7;
    const bool ework = tid < 256 && erow < B;
    const int len = ework ? a.lengths[erow] : 0;
    float carry = 0.f;
    int* flags = a.bar + d * ES_MAX_CTAS;
    __shared__ int gave_up;
    if (tid == 0) gave_up = 0;
    __syncthreads();
    const int k_lo = wid * KW, k_n = max(0, min(KW, CW - k_lo));   // this warp's range inside a chunk
    const int qrow = CW >> 2;                                       // float4 per row and chunk
#ifdef ES_PROFILE
    const bool prof = tid == 0 && blockIdx.x == 0;
    long long tacc[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? T - 1 - s : s, tp = d == 0 ? t - 1 : t + 1;
        // epilogue operands (saved by the forward pass / the incoming gradient): requested before the wait
        float gi[3] = {0.f, 0.f, 0.f}, gh[3] = {0.f, 0.f, 0.f}, hp = 0.f, dc = 0.f;
        const int64_t o3 = (((int64_t)d * T + t) * B + erow) * 3 * H + u0 + eu;
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) {
                gi[gt] = __ldg(a.gi + o3 + gt * H);
                gh[gt] = __ldg(a.gh + o3 + gt * H);
            }
            if (tp >= 0 && tp < T) hp = __ldg(a.ctx + ((int64_t)erow * T + tp) * 2 * H + (int64_t)d * H + u0 + eu);
            dc = __ldg(a.dctx + ((int64_t)erow * T + t) * 2 * H + (int64_t)d * H + u0 + eu);
        }
        float dacc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) dacc[mt][u] = 0.f;
        const float* x_in = a.xch + ((size_t)(s & 1) * 2 + d) * 32 * P;
        float* x_out = a.xch + ((size_t)((s + 1) & 1) * 2 + d) * 32 * P;
        if (s > 0) {
            es_wait(flags, n_cta, s, &gave_up);
            ES_T(0);
            // chunk c = columns [c·CW, (c+1)·CW) of the previous step's dgh: warp w copies rows 2w and 2w+1
            float4 nxt[ES_BWD_MAXLD];
            auto issue = [&](int c) {
#pragma unroll
                for (int i = 0; i < ES_BWD_MAXLD; ++i) {
                    const int r = i / (ES_BWD_MAXLD / 2), q = (i % (ES_BWD_MAXLD / 2)) * 32 + lane;
                    if (q < qrow) nxt[i] = __ldcg(reinterpret_cast<const float4*>(x_in + (wid * 2 + r) * P + c * CW) + q);
                }
            };
            issue(0);
            for (int c = 0; c < n_chunk; ++c) {
                if (c > 0) __syncthreads();                   // everybody is done with the previous chunk's tile
#pragma unroll
                for (int i = 0; i < ES_BWD_MAXLD; ++i) {
                    const int r = i / (ES_BWD_MAXLD / 2), q = (i % (ES_BWD_MAXLD / 2)) * 32 + lane;
                    if (q < qrow) reinterpret_cast<float4*>(xs + (wid * 2 + r) * CP)[q] = nxt[i];
                }
                __syncthreads();
                ES_T(5);
                if (c + 1 < n_chunk) issue(c + 1);            // the next chunk's loads fly during this chunk's MMAs
#pragma unroll 3
                for (int ks = 0; ks < (k_n >> 3); ++ks) {
                    const int kk = k_lo + ks * 8 + t4;
                    uint32_t ah[2][4], al[2][4];
                    es_load_a<RB>(ah, al, xs, CP, kk, g);
                    const int kg = c * CW + kk;
                    const float bf[2] = {wt[kg * 8 + g], wt[(kg + 4) * 8 + g]};
                    uint32_t bh[2], bl[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        bh[u] = RB ? __float_as_uint(bf[u]) : es_tf32(bf[u]);
                        bl[u] = RB ? 0u : es_tf32(bf[u] - __uint_as_float(bh[u]));
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (!RB) {
                            es_mma(dacc[mt], al[mt], bh);
                            es_mma(dacc[mt], ah[mt], bl);
                        }
                        es_mma(dacc[mt], ah[mt], bh);
                    }
                }
            }
        }
        ES_T(1);
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) part[(wid * 32 + mt * 16 + g + (u >> 1) * 8) * 9 + 2 * t4 + (u & 1)] = dacc[mt][u];
        __syncthreads();
        ES_T(2);
        float sv[4] = {0.f, 0.f, 0.f, 0.f};                          // dr, dz, dn, dn·r of this step (zeros for masked rows)
        if (ework) {
            float gsum = 0.f;
#pragma unroll
            for (int w = 0; w < ES_BWD_WARPS; ++w) gsum += part[(w * 32 + erow) * 9 + eu];
            gsum += carry + dc;
            if (gave_up) gsum = __int_as_float(0x7fc00000);
            float* px = x_out + erow * P + u0 + eu;
            if (len <= t) {
                px[0] = 0.f; px[H] = 0.f; px[2 * H] = 0.f;
                carry = 0.f;
            } else {
                const float r = sigmoidf_precise(gi[0] + gh[0]);
                const float z = sigmoidf_precise(gi[1] + gh[1]);
                const float hn = gh[2];
                const float n = tanhf(gi[2] + r * hn);
                const float dn_pre = gsum * (1.f - z) * (1.f - n * n);
                const float dz_pre = gsum * (hp - n) * z * (1.f - z);
                const float dr_pre = dn_pre * hn * r * (1.f - r);
                sv[0] = dr_pre; sv[1] = dz_pre; sv[2] = dn_pre; sv[3] = dn_pre * r;
                px[0] = RB ? es_rbf16(dr_pre) : dr_pre;       // the next step's operand, rounded once by its producer in bf16 mode
                px[H] = RB ? es_rbf16(dz_pre) : dz_pre;
                px[2 * H] = RB ? es_rbf16(sv[3]) : sv[3];
                carry = gsum * z;
            }
        }
        ES_T(3);
        if (s + 1 < T) es_arrive(flags + blockIdx.x, s + 1);          // covers the exchange copy; the saved tensors follow
        if (ework) {
            float* pa = a.dgi + o3;
            float* pb = a.dgh + o3;
            a.hprev_all[(((int64_t)d * T + t) * B + erow) * H + u0 + eu] = hp;
            pa[0] = sv[0]; pa[H] = sv[1]; pa[2 * H] = sv[2];
            pb[0] = sv[0]; pb[H] = sv[1]; pb[2 * H] = sv[3];
        }
        ES_T(4);
    }
#ifdef ES_PROFILE
    if (prof) printf("bwd d%d T%d: wait %lld  land+store %lld  mma %lld  reduce-sync %lld  epilogue %lld  arrive %lld (cycles, summed)\n", d, T, tacc[0], tacc[5], tacc[1], tacc[2], tacc[3], tacc[4]);
#endif
}

// ------------------------------------------------------------------------------------------------------------------------------
// bf16 mode with bf16 STORAGE: the operands of a bf16-mode contraction are bf16 by definition, so the exchange buffer, the state
// tile and the weight slice hold bf16 (half the L2 → SM bytes of a step, which is what bounds the backward exchange: 197 → 99 KB
// per CTA) and the product runs as mma.sync m16n8k16 bf16 (half the MMA and LDS instructions).  Same arithmetic as the float-
// storage path: bf16-rounded operands, exact products, FP32 accumulation (the summation order inside a k-step differs).
// ------------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void es_mma_b16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t es_pack_b16(float lo, float hi) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(lo)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(hi)) << 16);
}
// A fragments (both 16-row tiles) of one k-step of 16 from a [32][pitch] bf16 tile; kk = first column of this lane's pair
__device__ __forceinline__ void es_load_a_b16(uint32_t (&a)[2][4], const __nv_bfloat16* xs, int pitch, int kk, int g) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const __nv_bfloat16* xa = xs + (mt * 16 + g) * pitch + kk;
        a[mt][0] = *reinterpret_cast<const uint32_t*>(xa);
        a[mt][1] = *reinterpret_cast<const uint32_t*>(xa + 8 * pitch);
        a[mt][2] = *reinterpret_cast<const uint32_t*>(xa + 8);
        a[mt][3] = *reinterpret_cast<const uint32_t*>(xa + 8 * pitch + 8);
    }
}
constexpr int ES_B16_FWD_LD = 3;      // 16-byte loads per lane and row (<= 96 chunks of 8 bf16 per row)
constexpr int ES_B16_BWD_LD = 7;      // (<= 224 chunks per row)

__global__ void __launch_bounds__(ES_FWD_WARPS * 32, 1) enc_seq_fwd_b16_kernel(const EncSeqFwd a) {
    extern __shared__ __align__(16) float es_smem[];
    const int d = blockIdx.y, u0 = blockIdx.x * ES_UNITS, n_cta = gridDim.x;
    const int H = a.H, B = a.B, T = a.T;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int P = H + 8;                                    // bf16 elements per row: (P / 2) mod 32 = 4 words, conflict-free fragments
    const int KW = a.kw;                                    // contraction indices per warp (multiple of 16)
    __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(es_smem);      // [24][P]   row n = gate·8 + unit
    __nv_bfloat16* xs = ws + 24 * P;                                    // [32][P]
    float* part = reinterpret_cast<float*>(xs + 32 * P);                // [8 warps][32][25]
    const float* w_hh = a.w_hh[d];
    for (int i = tid; i < 24 * (H / 4); i += blockDim.x) {
        const int n = i / (H / 4), q = i - n * (H / 4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(w_hh + (int64_t)((n >> 3) * H + u0 + (n & 7)) * H + 4 * q));
        *reinterpret_cast<uint2*>(ws + n * P + 4 * q) = make_uint2(es_pack_b16(v.x, v.y), es_pack_b16(v.z, v.w));
    }
    for (int n = tid; n < 24; n += blockDim.x) *reinterpret_cast<uint4*>(ws + n * P + H) = make_uint4(0u, 0u, 0u, 0u);
    const int erow = tid >> 3, eu = tid & 7;
    const bool ework = tid < 256 && erow < B;
    float bias[3] = {0.f, 0.f, 0.f};
    int len = 0;
    if (ework) {
        len = a.lengths[erow];
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) bias[gt] = a.b_hh[d] ? a.b_hh[d][gt * H + u0 + eu] : 0.f;
    }
    int* flags = a.bar + d * ES_MAX_CTAS;
    __shared__ int gave_up;
    if (tid == 0) gave_up = 0;
    __syncthreads();
    const int k_lo = wid * KW, k_n = max(0, min(KW, H - k_lo));
    const int qrow = P >> 3;                                // 16-byte chunks per row
    __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(a.xch);
    float hp = 0.f;
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? s : T - 1 - s;
        const __nv_bfloat16* x_in = xch + ((size_t)(s & 1) * 2 + d) * 32 * P;
        __nv_bfloat16* x_out = xch + ((size_t)((s + 1) & 1) * 2 + d) * 32 * P;
        const int64_t o3 = (((int64_t)d * T + t) * B + erow) * 3 * H + u0 + eu;
        float gi[3] = {0.f, 0.f, 0.f};
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gi[gt] = __ldg(a.gi + o3 + gt * H);
        }
        float dacc[3][2][4];
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 4; ++u) dacc[nt][mt][u] = 0.f;
        if (s > 0) {
            es_wait(flags, n_cta, s, &gave_up);
            uint4 v[ES_B16_FWD_LD * ES_FWD_RPW];
#pragma unroll
            for (int i = 0; i < ES_B16_FWD_LD * ES_FWD_RPW; ++i) {
                const int r = i / ES_B16_FWD_LD, q = (i % ES_B16_FWD_LD) * 32 + lane;
                if (q < qrow) v[i] = __ldcg(reinterpret_cast<const uint4*>(x_in + (wid * ES_FWD_RPW + r) * P) + q);
            }
#pragma unroll
            for (int i = 0; i < ES_B16_FWD_LD * ES_FWD_RPW; ++i) {
                const int r = i / ES_B16_FWD_LD, q = (i % ES_B16_FWD_LD) * 32 + lane;
                if (q < qrow) reinterpret_cast<uint4*>(xs + (wid * ES_FWD_RPW + r) * P)[q] = v[i];
            }
            __syncthreads();
#pragma unroll 2
            for (int ks = 0; ks < (k_n >> 4); ++ks) {
                const int kk = k_lo + ks * 16 + 2 * t4;
                uint32_t af[2][4];
                es_load_a_b16(af, xs, P, kk, g);
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    const __nv_bfloat16* wb = ws + (nt * 8 + g) * P + kk;
                    const uint32_t bf[2] = {*reinterpret_cast<const uint32_t*>(wb), *reinterpret_cast<const uint32_t*>(wb + 8)};
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) es_mma_b16(dacc[nt][mt], af[mt], bf);
                }
            }
        }
#pragma unroll
        for (int nt = 0; nt < 3; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    part[(wid * 32 + mt * 16 + g + (u >> 1) * 8) * 25 + nt * 8 + 2 * t4 + (u & 1)] = dacc[nt][mt][u];
        __syncthreads();
        float sv[3] = {0.f, 0.f, 0.f};
        if (ework) {
            float pre[3];
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) {
                float v = 0.f;
#pragma unroll
                for (int w = 0; w < ES_FWD_WARPS; ++w) v += part[(w * 32 + erow) * 25 + gt * 8 + eu];
                pre[gt] = v + bias[gt];
            }
            const bool live = len > t;
            if (gave_up) pre[0] = __int_as_float(0x7fc00000);
            if (live) {
                const float r = sigmoidf_precise(gi[0] + pre[0]);
                const float z = sigmoidf_precise(gi[1] + pre[1]);
                const float n = tanhf(gi[2] + r * pre[2]);
                hp = (1.0f - z) * n + z * hp;
            }
            x_out[erow * P + u0 + eu] = __float2bfloat16_rn(hp);
            sv[0] = live ? pre[0] : 0.f; sv[1] = live ? pre[1] : 0.f; sv[2] = live ? pre[2] : 0.f;
        }
        if (s + 1 < T) es_arrive(flags + blockIdx.x, s + 1);
        if (ework) {
            if (len > t) a.ctx_out[((int64_t)erow * T + t) * 2 * H + (int64_t)d * H + u0 + eu] = hp;
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) a.gh[o3 + gt * H] = sv[gt];
        }
    }
}

__global__ void __launch_bounds__(ES_BWD_WARPS * 32, 1) enc_seq_bwd_b16_kernel(const EncSeqBwd a) {
    extern __shared__ __align__(16) float es_smem[];
    const int d = blockIdx.y, u0 = blockIdx.x * ES_UNITS, n_cta = gridDim.x;
    const int H = a.H, B = a.B, T = a.T, K = 3 * H;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int P = K + 8;                                    // bf16 elements per row of the exchange buffer, the tile and the slab
    const int KW = a.kw;                                    // contraction indices per warp (multiple of 16)
    __nv_bfloat16* wt = reinterpret_cast<__nv_bfloat16*>(es_smem);      // [8][P]: W_hh[k][u0 + n], k contiguous
    __nv_bfloat16* xs = wt + 8 * P;                                     // [32][P]
    float* part = reinterpret_cast<float*>(xs + 32 * P);                // [16 warps][32][9]
    const float* w_hh = a.w_hh[d];
    for (int k = tid; k < K; k += blockDim.x) {
        const float4 v0 = __ldg(reinterpret_cast<const float4*>(w_hh + (int64_t)k * H + u0));
        const float4 v1 = __ldg(reinterpret_cast<const float4*>(w_hh + (int64_t)k * H + u0 + 4));
        const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int n = 0; n < 8; ++n) wt[n * P + k] = __float2bfloat16_rn(vv[n]);
    }
    for (int n = tid; n < 8; n += blockDim.x) *reinterpret_cast<uint4*>(wt + n * P + K) = make_uint4(0u, 0u, 0u, 0u);
    const int erow = tid >> 3, eu = tid & 7;
    const bool ework = tid < 256 && erow < B;
    const int len = ework ? a.lengths[erow] : 0;
    float carry = 0.f;
    int* flags = a.bar + d * ES_MAX_CTAS;
    __shared__ int gave_up;
    if (tid == 0) gave_up = 0;
    __syncthreads();
    const int k_lo = wid * KW, k_n = max(0, min(KW, K - k_lo));
    const int qrow = P >> 3;
    __nv_bfloat16* xch = reinterpret_cast<__nv_bfloat16*>(a.xch);
    for (int s = 0; s < T; ++s) {
        const int t = d == 0 ? T - 1 - s : s, tp = d == 0 ? t - 1 : t + 1;
        float gi[3] = {0.f, 0.f, 0.f}, gh[3] = {0.f, 0.f, 0.f}, hp = 0.f, dc = 0.f;
        const int64_t o3 = (((int64_t)d * T + t) * B + erow) * 3 * H + u0 + eu;
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) {
                gi[gt] = __ldg(a.gi + o3 + gt * H);
                gh[gt] = __ldg(a.gh + o3 + gt * H);
            }
            if (tp >= 0 && tp < T) hp = __ldg(a.ctx + ((int64_t)erow * T + tp) * 2 * H + (int64_t)d * H + u0 + eu);
            dc = __ldg(a.dctx + ((int64_t)erow * T + t) * 2 * H + (int64_t)d * H + u0 + eu);
        }
        float dacc[2][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) dacc[mt][u] = 0.f;
        const __nv_bfloat16* x_in = xch + ((size_t)(s & 1) * 2 + d) * 32 * P;
        __nv_bfloat16* x_out = xch + ((size_t)((s + 1) & 1) * 2 + d) * 32 * P;
        if (s > 0) {
            es_wait(flags, n_cta, s, &gave_up);
            uint4 v[2 * ES_B16_BWD_LD];                      // warp w copies rows 2w and 2w+1
#pragma unroll
            for (int i = 0; i < 2 * ES_B16_BWD_LD; ++i) {
                const int r = i / ES_B16_BWD_LD, q = (i % ES_B16_BWD_LD) * 32 + lane;
                if (q < qrow) v[i] = __ldcg(reinterpret_cast<const uint4*>(x_in + (wid * 2 + r) * P) + q);
            }
#pragma unroll
            for (int i = 0; i < 2 * ES_B16_BWD_LD; ++i) {
                const int r = i / ES_B16_BWD_LD, q = (i % ES_B16_BWD_LD) * 32 + lane;
                if (q < qrow) reinterpret_cast<uint4*>(xs + (wid * 2 + r) * P)[q] = v[i];
            }
            __syncthreads();
#pragma unroll 3
            for (int ks = 0; ks < (k_n >> 4); ++ks) {
                const int kk = k_lo + ks * 16 + 2 * t4;
                uint32_t af[2][4];
                es_load_a_b16(af, xs, P, kk, g);
                const __nv_bfloat16* wb = wt + g * P + kk;
                const uint32_t bf[2] = {*reinterpret_cast<const uint32_t*>(wb), *reinterpret_cast<const uint32_t*>(wb + 8)};
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) es_mma_b16(dacc[mt], af[mt], bf);
            }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) part[(wid * 32 + mt * 16 + g + (u >> 1) * 8) * 9 + 2 * t4 + (u & 1)] = dacc[mt][u];
        __syncthreads();
        float sv[4] = {0.f, 0.f, 0.f, 0.f};
        if (ework) {
            float gsum = 0.f;
#pragma unroll
            for (int w = 0; w < ES_BWD_WARPS; ++w) gsum += part[(w * 32 + erow) * 9 + eu];
            gsum += carry + dc;
            if (gave_up) gsum = __int_as_float(0x7fc00000);
            __nv_bfloat16* px = x_out + erow * P + u0 + eu;
            if (len <= t) {
                px[0] = __float2bfloat16_rn(0.f); px[H] = px[0]; px[2 * H] = px[0];
                carry = 0.f;
            } else {
                const float r = sigmoidf_precise(gi[0] + gh[0]);
                const float z = sigmoidf_precise(gi[1] + gh[1]);
                const float hn = gh[2];
                const float n = tanhf(gi[2] + r * hn);
                const float dn_pre = gsum * (1.f - z) * (1.f - n * n);
                const float dz_pre = gsum * (hp - n) * z * (1.f - z);
                const float dr_pre = dn_pre * hn * r * (1.f - r);
                sv[0] = dr_pre; sv[1] = dz_pre; sv[2] = dn_pre; sv[3] = dn_pre * r;
                px[0] = __float2bfloat16_rn(dr_pre);
                px[H] = __float2bfloat16_rn(dz_pre);
                px[2 * H] = __float2bfloat16_rn(sv[3]);
                carry = gsum * z;
            }
        }
        if (s + 1 < T) es_arrive(flags + blockIdx.x, s + 1);
        if (ework) {
            float* pa = a.dgi + o3;
            float* pb = a.dgh + o3;
            a.hprev_all[(((int64_t)d * T + t) * B + erow) * H + u0 + eu] = hp;
            pa[0] = sv[0]; pa[H] = sv[1]; pa[2 * H] = sv[2];
            pb[0] = sv[0]; pb[H] = sv[1]; pb[2 * H] = sv[3];
        }
    }
}

bool es_b16_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VAG_ENC_SEQ_B16"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
int es_b16_fwd_kw(int H) { return ((H + ES_FWD_WARPS - 1) / ES_FWD_WARPS + 15) / 16 * 16; }
int es_b16_bwd_kw(int H) { return ((3 * H + ES_BWD_WARPS - 1) / ES_BWD_WARPS + 15) / 16 * 16; }
size_t es_b16_fwd_smem(int H) { return (size_t)(24 + 32) * (H + 8) * 2 + sizeof(float) * (size_t)ES_FWD_WARPS * 32 * 25; }
size_t es_b16_bwd_smem(int H) { return (size_t)(8 + 32) * (3 * H + 8) * 2 + sizeof(float) * (size_t)ES_BWD_WARPS * 32 * 9; }
bool es_b16_fwd_ok(int H) { return es_b16_enabled() && (H % 16) == 0 && (H + 8) / 8 <= 32 * ES_B16_FWD_LD && es_b16_fwd_smem(H) <= 200 * 1024; }
bool es_b16_bwd_ok(int H) { return es_b16_enabled() && (H % 16) == 0 && (3 * H + 8) / 8 <= 32 * ES_B16_BWD_LD && es_b16_bwd_smem(H) <= 200 * 1024; }

int es_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}
bool es_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("VAG_ENC_SEQ"); v = (e && e[0] == '0') ? 0 : 1; }
    return v == 1;
}
int es_fwd_kw(int H) { return ((H + ES_FWD_WARPS - 1) / ES_FWD_WARPS + 7) / 8 * 8; }
int es_bwd_chunks(int H) { return (size_t)32 * (3 * H + 4) * 4 > 100 * 1024 ? 2 : 1; }
int es_bwd_kw(int H) { return ((3 * H / es_bwd_chunks(H) + ES_BWD_WARPS - 1) / ES_BWD_WARPS + 7) / 8 * 8; }
size_t es_fwd_smem(int H) { return sizeof(float) * ((size_t)24 * (H + 4) + (size_t)32 * (H + 4) + (size_t)ES_FWD_WARPS * 32 * 25); }
size_t es_bwd_smem(int H) {
    return sizeof(float) * ((size_t)3 * H * 8 + (size_t)32 * (3 * H / es_bwd_chunks(H) + 4) + (size_t)ES_BWD_WARPS * 32 * 9);
}
bool es_common_ok(int B, int T, int H) {
    return es_enabled() && B >= 1 && B <= 32 && T >= 1 && H >= 8 && (H % 8) == 0 && 2 * (H / ES_UNITS) <= es_num_sms() && H / ES_UNITS <= ES_MAX_CTAS;
}
int es_prepare(void* scratch, size_t scratch_bytes, int H, cudaStream_t st) {
    if (!scratch || scratch_bytes < enc_seq_scratch_bytes(H)) {
        set_error("enc_seq: scratch too small");
        return VAG_ERR_WORKSPACE;
    }
    // flags = 0 (step counters), exchange buffer = 0 (rows >= B and the pad columns are never written)
    VAG_CUDA(cudaMemsetAsync(scratch, 0, enc_seq_scratch_bytes(H), st));
    return VAG_OK;
}

}  // namespace

size_t enc_seq_scratch_bytes(int H) {      // flags + the larger (backward) exchange buffer
    return 2 * ES_MAX_CTAS * sizeof(int) + sizeof(float) * (size_t)4 * 32 * (3 * H + 4) + 256;
}
bool enc_seq_fwd_ok(int B, int T, int H) {
    // every row of the exchange buffer must fit the per-lane load budget
    return es_common_ok(B, T, H) && (H + 4) / 4 <= 32 * 5 && es_fwd_smem(H) <= 200 * 1024;
}
bool enc_seq_bwd_ok(int B, int T, int H) {
    const int nc = es_bwd_chunks(H), cw = 3 * H / nc;
    return es_common_ok(B, T, H) && (3 * H) % nc == 0 && (cw % 8) == 0 && cw / 4 <= 32 * (ES_BWD_MAXLD / 2) && es_bwd_smem(H) <= 200 * 1024;
}

int enc_seq_fwd(EncSeqFwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st) {
    if (!enc_seq_fwd_ok(a.B, a.T, a.H)) {
        set_error("enc_seq_fwd: unsupported shape");
        return VAG_ERR_UNSUPPORTED;
    }
    VAG_TRY(es_prepare(scratch, scratch_bytes, a.H, st));
    a.bar = reinterpret_cast<int*>(scratch);
    a.xch = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 2 * ES_MAX_CTAS * sizeof(int));
    const dim3 grid(a.H / ES_UNITS, 2);
    if (round_bf16 && es_b16_fwd_ok(a.H)) {
        a.kw = es_b16_fwd_kw(a.H);
        const size_t smem16 = es_b16_fwd_smem(a.H);
        static size_t configured16 = 0;
        if (smem16 > configured16) {
            VAG_CUDA(cudaFuncSetAttribute(enc_seq_fwd_b16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            configured16 = smem16;
        }
        enc_seq_fwd_b16_kernel<<<grid, ES_FWD_WARPS * 32, smem16, st>>>(a);
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    a.kw = es_fwd_kw(a.H);
    const size_t smem = es_fwd_smem(a.H);
    static size_t configured[2] = {0, 0};
    if (smem > configured[round_bf16]) {
        if (round_bf16) VAG_CUDA(cudaFuncSetAttribute(enc_seq_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else VAG_CUDA(cudaFuncSetAttribute(enc_seq_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[round_bf16] = smem;
    }
    if (round_bf16) enc_seq_fwd_kernel<true><<<grid, ES_FWD_WARPS * 32, smem, st>>>(a);
    else enc_seq_fwd_kernel<false><<<grid, ES_FWD_WARPS * 32, smem, st>>>(a);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int enc_seq_bwd(EncSeqBwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st) {
    if (!enc_seq_bwd_ok(a.B, a.T, a.H)) {
        set_error("enc_seq_bwd: unsupported shape");
        return VAG_ERR_UNSUPPORTED;
    }
    VAG_TRY(es_prepare(scratch, scratch_bytes, a.H, st));
    a.bar = reinterpret_cast<int*>(scratch);
    a.xch = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + 2 * ES_MAX_CTAS * sizeof(int));
    const dim3 grid(a.H / ES_UNITS, 2);
    if (round_bf16 && es_b16_bwd_ok(a.H)) {
        a.n_chunk = 1;
        a.kw = es_b16_bwd_kw(a.H);
        const size_t smem16 = es_b16_bwd_smem(a.H);
        static size_t configured16 = 0;
        if (smem16 > configured16) {
            VAG_CUDA(cudaFuncSetAttribute(enc_seq_bwd_b16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
            configured16 = smem16;
        }
        enc_seq_bwd_b16_kernel<<<grid, ES_BWD_WARPS * 32, smem16, st>>>(a);
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    a.n_chunk = es_bwd_chunks(a.H);
    a.kw = es_bwd_kw(a.H);
    const size_t smem = es_bwd_smem(a.H);
    static size_t configured[2] = {0, 0};
    if (smem > configured[round_bf16]) {
        if (round_bf16) VAG_CUDA(cudaFuncSetAttribute(enc_seq_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else VAG_CUDA(cudaFuncSetAttribute(enc_seq_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[round_bf16] = smem;
    }
    if (round_bf16) enc_seq_bwd_kernel<true><<<grid, ES_BWD_WARPS * 32, smem, st>>>(a);
    else enc_seq_bwd_kernel<false><<<grid, ES_BWD_WARPS * 32, smem, st>>>(a);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

}  // namespace vag
