// Persistent forward time loop of the decoder's training step (teacher forcing, batch <= 32 rows per GPU).
//
// One teacher-forced decoder step (layers/NMT_Decoder.py:118-131) is a chain of five small dependent products around the attention:
//   A  gh1 = h2'·W_hh1ᵀ  → GRU cell 1 (its input side gi1 is known for all steps: teacher forcing)            → h1
//   B  q = h1·W_attn_hᵀ   and   gh2 = h1·W_hh2ᵀ
//   C  attention of every sentence over its source positions (scores, masked soft-max, context)                → α, c
//   D  x2 = c·W_c2hᵀ
//   E  gi2 = x2·W_ih2ᵀ    → GRU cell 2 (with gh2 and h1)                                                       → h2
// As launches that is 5 kernels x Tt steps of 5-13 us each, almost all of it launch and dependency latency.  Here ONE launch
// walks all Tt steps: H/4 CTAs (128 at H = 512, one per SM, co-resident), CTA i owns hidden units 4i … 4i+3 (12 gate rows of each
// GRU matrix), 8 columns of q and 4 columns of x2, keeps ITS slices of the five matrices in shared memory for the whole sequence
// (107 KB), and the CTAs exchange h1 / q / c / x2 / h2 through global buffers laid out as the consumers stage them, with the
// flag barrier of seq_common.cuh between the phases (5 per step).  CTA b < B also runs the attention of sentence b.
// Everything the backward pass reads (gh1, h1, q, α, c, x2, gi2, gh2, h2 for all steps) is written exactly where the per-step
// kernels put it, so vag_decoder_seq_bwd_f32 does not know the difference.  Arithmetic: the products like linear_rows32_kernel
// (mma.sync m16n8k8 TF32: 3xTF32 in FP32 mode, one product of bf16-rounded operands in bf16 mode — the operands are rounded once,
// by their producer, when they enter the exchange buffer), attention and cells in FP32.
#include "dec_seq.cuh"
#include "seq_common.cuh"
#include <stdio.h>
#include <stdlib.h>

namespace vag {
namespace {

constexpr int DS_WARPS = 8;
constexpr int DS_UNITS = 4;           // hidden units per CTA
constexpr int DS_MAXLD = 20;          // float4 loads per lane and tile copy (4 rows per warp, <= 160 float4 per row)
constexpr float kDsTwoLog2e = 2.8853900817779268f;
#ifdef DS_PROFILE
#define DS_T(i) do { if (prof) { const long long now_ = clock64(); tacc[i] += now_ - tlast; tlast = now_; } } while (0)
#else
#define DS_T(i) do { } while (0)
#endif

// v0/(1+2^u0) + … + v3/(1+2^u3) with ONE reciprocal (attention.cu's sum4_v_over_one_plus_exp2: u clamped to 30, beyond which the
// term is below half an ulp of the O(1) sum it joins, and the product of the four denominators stays < 2^124)
__device__ __forceinline__ float ds_sum4(const float4& v, float u0, float u1, float u2, float u3) {
    float e0, e1, e2, e3, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fminf(u0, 30.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fminf(u1, 30.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(fminf(u2, 30.f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e3) : "f"(fminf(u3, 30.f)));
    const float a0 = e0 + 1.0f, a1 = e1 + 1.0f, a2 = e2 + 1.0f, a3 = e3 + 1.0f;
    const float p01 = a0 * a1, p23 = a2 * a3;
    const float n01 = fmaf(v.x, a1, v.y * a0), n23 = fmaf(v.z, a3, v.w * a2);
    const float num = fmaf(n01, p23, n23 * p01);
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p01 * p23));
    return num * r;
}

// [32][P] tile ← global rows (pitch ld floats, `rows` valid rows, `cols` floats per row — a multiple of 4 — the rest zero): every load
// of the warp's 4 rows is issued before the first store.  Ends with a CTA barrier.
__device__ __forceinline__ void ds_copy_tile(float* xs, int P, const float* src, int64_t ld, int rows, int cols, int wid, int lane) {
    float4 v[DS_MAXLD];
    const int qrow = cols >> 2;
#pragma unroll
    for (int i = 0; i < DS_MAXLD; ++i) {
        const int r = wid * 4 + i / 5, q = (i % 5) * 32 + lane;
        v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < qrow && r < rows) v[i] = __ldcg(reinterpret_cast<const float4*>(src + (int64_t)r * ld) + q);
    }
#pragma unroll
    for (int i = 0; i < DS_MAXLD; ++i) {
        const int r = wid * 4 + i / 5, q = (i % 5) * 32 + lane;
        if (q < qrow) reinterpret_cast<float4*>(xs + r * P)[q] = v[i];
    }
    __syncthreads();
}

// dacc[nt] += tile[32][k_lo … k_lo+k_n) · rows (nt·8 + g) of ws (pitch WP, k-fast; rows >= n_rows wrap onto real rows: their
// columns are never read)
template <int NT, bool RB>
__device__ __forceinline__ void ds_mma(float (&dacc)[NT][2][4], const float* xs, int P, int xk0, const float* ws, int WP, int wk0,
                                       int n_rows, int k_n, int g, int t4) {
#pragma unroll 2
    for (int ks = 0; ks < (k_n >> 3); ++ks) {
        const int kk = ks * 8 + t4;
        uint32_t ah[2][4], al[2][4];
        es_load_a<RB>(ah, al, xs, P, xk0 + kk, g);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            const int n = nt * 8 + g;
            const float* wb = ws + (n < n_rows ? n : n % n_rows) * WP + wk0 + kk;
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
template <int NT>
__device__ __forceinline__ void ds_zero(float (&dacc)[NT][2][4]) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) dacc[nt][mt][u] = 0.f;
}
// partial sums of this warp → part[warp][row][PW]
template <int NT>
__device__ __forceinline__ void ds_store_part(float* part, int PW, const float (&dacc)[NT][2][4], int wid, int g, int t4) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) part[(wid * 32 + mt * 16 + g + (u >> 1) * 8) * PW + nt * 8 + 2 * t4 + (u & 1)] = dacc[nt][mt][u];
}
__device__ __forceinline__ float ds_sum_part(const float* part, int PW, int row, int col) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < DS_WARPS; ++w) v += part[(w * 32 + row) * PW + col];
    return v;
}
template <bool RB>
__device__ __forceinline__ float ds_round(float x) { return RB ? es_rbf16(x) : x; }

// weight rows → shared memory (k-fast, pitch WP, bf16-rounded in bf16 mode); row r of the slice = src row map(r)
template <bool RB, typename F>
__device__ __forceinline__ void ds_load_rows(float* dst, int WP, const float* src, int64_t ld, int K, int n_rows, F map, int tid) {
    const int qk = K >> 2;
    for (int i = tid; i < n_rows * qk; i += DS_WARPS * 32) {
        const int r = i / qk, q = i - r * qk;
        float4 v = __ldg(reinterpret_cast<const float4*>(src + (int64_t)map(r) * ld) + q);
        if (RB) { v.x = es_rbf16(v.x); v.y = es_rbf16(v.y); v.z = es_rbf16(v.z); v.w = es_rbf16(v.w); }
        reinterpret_cast<float4*>(dst + r * WP)[q] = v;
    }
    for (int r = tid; r < n_rows; r += DS_WARPS * 32) *reinterpret_cast<float4*>(dst + r * WP + K) = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <bool RB>
__global__ void __launch_bounds__(DS_WARPS * 32, 1) dec_seq_fwd_kernel(const DecSeqFwd a) {
    extern __shared__ __align__(16) float ds_smem[];
    const int H = a.H, C = a.C, B = a.B, T = a.T, Tt = a.Tt;
    const int cta = blockIdx.x, n_cta = gridDim.x, u0 = cta * DS_UNITS;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, g = lane >> 2, t4 = lane & 3;
    const int P = H + 4, PC = C + 4;                       // pitches of H-wide and C-wide rows
    // ---- this CTA's weight slices
    float* w_hh1 = ds_smem;                                 // [12][P]  row = gate·4 + unit
    float* w_q = w_hh1 + 12 * P;                            // [8][P]   q columns 8·cta …
    float* w_hh2 = w_q + 8 * P;                             // [12][P]  (w_q and w_hh2 are contiguous: ONE 20-row slice for phase B)
    float* w_ih2 = w_hh2 + 12 * P;                          // [12][P]
    float* w_c2h = w_ih2 + 12 * P;                          // [4][PC]
    float* xs = w_c2h + 4 * PC;                             // [32][P] state tile; the attention CTA's scratch in phase C
    float* part = xs + 32 * P;                              // [8 warps][32][25]
    auto gate_row = [&](int r) { return (r >> 2) * H + u0 + (r & 3); };
    ds_load_rows<RB>(w_hh1, P, a.gru1_w_hh, H, H, 12, gate_row, tid);
    ds_load_rows<RB>(w_q, P, a.attn_h_w, H, H, 8, [&](int r) { return cta * 8 + r; }, tid);
    ds_load_rows<RB>(w_hh2, P, a.gru2_w_hh, H, H, 12, gate_row, tid);
    ds_load_rows<RB>(w_ih2, P, a.gru2_w_ih, H, H, 12, gate_row, tid);
    ds_load_rows<RB>(w_c2h, PC, a.c2h_w, C, C, 4, [&](int r) { return u0 + r; }, tid);
    // ---- epilogue roles: (row, unit) for the cells and x2 — 128 threads —, (row, column) for q — 256 threads
    const int erow = tid >> 2, eu = tid & 3;
    const bool ework = tid < 128 && erow < B;
    const int qrow_ = tid >> 3, qcol = tid & 7;
    const bool qwork = qrow_ < B;
    float b_hh1[3] = {0.f, 0.f, 0.f}, b_hh2[3] = {0.f, 0.f, 0.f}, b_ih2[3] = {0.f, 0.f, 0.f};
    if (ework) {
#pragma unroll
        for (int gt = 0; gt < 3; ++gt) {
            b_hh1[gt] = a.gru1_b_hh ? a.gru1_b_hh[gt * H + u0 + eu] : 0.f;
            b_hh2[gt] = a.gru2_b_hh ? a.gru2_b_hh[gt * H + u0 + eu] : 0.f;
            b_ih2[gt] = a.gru2_b_ih ? a.gru2_b_ih[gt * H + u0 + eu] : 0.f;
        }
    }
    __shared__ int gave_up;
    if (tid == 0) gave_up = 0;
    __syncthreads();
    const int KWH = a.kw_h, k_lo = wid * KWH, k_n = max(0, min(KWH, H - k_lo));     // warps' ranges over an H-wide contraction
    float h_state = ework ? a.h0[(int64_t)erow * H + u0 + eu] : 0.f;                  // h2 of the previous step (h0 at t = 0)
    float h1 = 0.f, gh2[3] = {0.f, 0.f, 0.f};
    float* X_h = a.xch;                                     // exchange buffers, all zero-initialised [32][pitch]
    float* X_h1 = X_h + 32 * P;
    float* X_x2 = X_h1 + 32 * P;
    float* X_q = X_x2 + 32 * P;                             // [32][PC]
    float* X_c = X_q + 32 * PC;                             // [2 halves][32][P]  (C = 2H: each half is one H-wide tile)
    int seq = 0;                                            // barrier sequence number
#ifdef DS_PROFILE
    const bool prof = tid == 0 && cta == 0;
    long long tacc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, tlast = clock64();
#endif
    for (int t = 0; t < Tt; ++t) {
        const int64_t o3 = ((int64_t)t * B + erow) * 3 * H + u0 + eu, oh = ((int64_t)t * B + erow) * H + u0 + eu;
        // ================= A: gru_1 =================
        float gi1[3] = {0.f, 0.f, 0.f};
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gi1[gt] = __ldg(a.gi1_all + o3 + gt * H);
        }
        if (t == 0) {
            ds_copy_tile(xs, P, a.h0, H, B, H, wid, lane);
            if (RB) {       // h0 enters a bf16-mode contraction: round the staged copy (later states are rounded by their producer)
                for (int i = tid; i < 32 * P; i += DS_WARPS * 32) xs[i] = es_rbf16(xs[i]);
                __syncthreads();
            }
        } else {
            es_wait(a.bar, n_cta, seq, &gave_up);
            DS_T(0);
            ds_copy_tile(xs, P, X_h, P, 32, P, wid, lane);
        }
        DS_T(1);
        {
            float dacc[2][2][4];
            ds_zero<2>(dacc);
            ds_mma<2, RB>(dacc, xs, P, k_lo, w_hh1, P, k_lo, 12, k_n, g, t4);
            ds_store_part<2>(part, 25, dacc, wid, g, t4);
        }
        __syncthreads();
        if (ework) {
            float pre[3];
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) pre[gt] = ds_sum_part(part, 25, erow, gt * 4 + eu) + b_hh1[gt];
            if (gave_up) pre[0] = __int_as_float(0x7fc00000);
            const float r = sigmoidf_precise(gi1[0] + pre[0]);
            const float z = sigmoidf_precise(gi1[1] + pre[1]);
            const float n = tanhf(gi1[2] + r * pre[2]);
            h1 = (1.0f - z) * n + z * h_state;
            X_h1[erow * P + u0 + eu] = ds_round<RB>(h1);
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gi1[gt] = pre[gt];          // kept for the deferred store
        }
        DS_T(2);
        es_arrive(a.bar + cta, ++seq);
        DS_T(3);
        if (ework) {
            a.h1_all[oh] = h1;
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) a.gh1_all[o3 + gt * H] = gi1[gt];
        }
        // ================= B: q and gh2 from h1 =================
        es_wait(a.bar, n_cta, seq, &gave_up);
        DS_T(4);
        ds_copy_tile(xs, P, X_h1, P, 32, P, wid, lane);
        DS_T(5);
        {
            float dacc[3][2][4];
            ds_zero<3>(dacc);
            ds_mma<3, RB>(dacc, xs, P, k_lo, w_q, P, k_lo, 20, k_n, g, t4);     // rows 0-7 = q, 8-19 = gh2
            ds_store_part<3>(part, 25, dacc, wid, g, t4);
        }
        __syncthreads();
        float qv = 0.f;
        if (qwork) {
            qv = ds_sum_part(part, 25, qrow_, qcol);
            X_q[qrow_ * PC + cta * 8 + qcol] = qv;
        }
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gh2[gt] = ds_sum_part(part, 25, erow, 8 + gt * 4 + eu) + b_hh2[gt];
        }
        DS_T(6);
        es_arrive(a.bar + cta, ++seq);
        if (qwork) a.q_all[((int64_t)t * B + qrow_) * C + cta * 8 + qcol] = qv;
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) a.gh2_all[o3 + gt * H] = gh2[gt];
        }
        // ================= C: attention of sentence `cta` =================
        es_wait(a.bar, n_cta, seq, &gave_up);
        DS_T(7);
        if (cta < B) {
            float* q_s = xs;                                // [C] 2·log2e·q
            float* v_s = xs + C;                            // [C]
            float* sc_s = xs + 2 * C;                       // [T] scores → α
            float* red = xs + 2 * C + 256;                  // [8]
            const int b = cta;
            float pv = 0.f;
            for (int i = tid; i < C; i += DS_WARPS * 32) {
                q_s[i] = __ldcg(X_q + b * PC + i) * kDsTwoLog2e;
                const float vv = __ldg(a.attn_v + i);
                v_s[i] = vv;
                pv += vv;
            }
            pv = warp_sum(pv);
            if (lane == 0) red[wid] = pv;
            __syncthreads();
            float vsum = 0.f;
#pragma unroll
            for (int w = 0; w < DS_WARPS; ++w) vsum += red[w];
            const float* key_b = a.keys + (int64_t)b * T * C;
            const float* ctx_b = a.enc + (int64_t)b * T * C;
            const float* mask_b = a.mask + (int64_t)b * T;
            // scores: Σ_c v_c·tanh(q_c + k_c) = Σ v − 2 Σ v_c / (1 + exp(2(q_c + k_c))), one warp per source position
            for (int tp = wid; tp < T; tp += DS_WARPS) {
                float sc = -INFINITY;
                if (__ldg(mask_b + tp) != 0.f) {            // warp-uniform
                    const float* kr = key_b + (int64_t)tp * C;
                    float acc = 0.f;
                    for (int cb = 0; cb < C; cb += 1024) {       // eight independent 16-byte loads in flight per lane
                        float4 kv[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c0 = cb + j * 128 + lane * 4;
                            kv[j] = c0 < C ? __ldg(reinterpret_cast<const float4*>(kr + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const int c0 = cb + j * 128 + lane * 4;
                            if (c0 < C) {
                                const float4 qq = *reinterpret_cast<const float4*>(q_s + c0);
                                const float4 vv = *reinterpret_cast<const float4*>(v_s + c0);
                                acc += ds_sum4(vv, fmaf(kv[j].x, kDsTwoLog2e, qq.x), fmaf(kv[j].y, kDsTwoLog2e, qq.y),
                                               fmaf(kv[j].z, kDsTwoLog2e, qq.z), fmaf(kv[j].w, kDsTwoLog2e, qq.w));
                            }
                        }
                    }
                    acc = warp_sum(acc);
                    sc = fmaf(-2.0f, acc, vsum);
                }
                if (lane == 0) sc_s[tp] = sc;
            }
            __syncthreads();
            DS_T(8);
            if (wid == 0) {                                 // masked soft-max over T (T <= 256)
                float m = -INFINITY;
                for (int tp = lane; tp < T; tp += 32) m = fmaxf(m, sc_s[tp]);
                m = warp_max(m);
                float sum = 0.f;
                for (int tp = lane; tp < T; tp += 32) {
                    const float e = expf(sc_s[tp] - m);
                    sc_s[tp] = e;
                    sum += e;
                }
                sum = warp_sum(sum);
                for (int tp = lane; tp < T; tp += 32) {
                    const float al = sc_s[tp] / sum;
                    sc_s[tp] = al;
                    a.alpha_all[((int64_t)t * B + b) * T + tp] = al;
                }
            }
            __syncthreads();
            // context: thread owns 4 channels; positions with α = 0 (masked) are skipped
            for (int c0 = tid * 4; c0 < C; c0 += DS_WARPS * 32 * 4) {
                float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int tb = 0; tb < T; tb += 8) {          // eight positions in flight (masked ones, α = 0, are not loaded)
                    float4 cv[8];
                    float al[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        al[j] = tb + j < T ? sc_s[tb + j] : 0.f;
                        cv[j] = al[j] != 0.f ? __ldg(reinterpret_cast<const float4*>(ctx_b + (int64_t)(tb + j) * C + c0)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        acc.x = fmaf(al[j], cv[j].x, acc.x); acc.y = fmaf(al[j], cv[j].y, acc.y);
                        acc.z = fmaf(al[j], cv[j].z, acc.z); acc.w = fmaf(al[j], cv[j].w, acc.w);
                    }
                }
                if (gave_up) acc.x = __int_as_float(0x7fc00000);
                *reinterpret_cast<float4*>(a.c_all + ((int64_t)t * B + b) * C + c0) = acc;
                const int half = c0 / H, cc = c0 - half * H;
                *reinterpret_cast<float4*>(X_c + ((size_t)half * 32 + b) * P + cc) =
                    make_float4(ds_round<RB>(acc.x), ds_round<RB>(acc.y), ds_round<RB>(acc.z), ds_round<RB>(acc.w));
            }
        }
        DS_T(9);
        es_arrive(a.bar + cta, ++seq);
        // ================= D: x2 = c·W_c2hᵀ (two H-wide halves) =================
        es_wait(a.bar, n_cta, seq, &gave_up);
        DS_T(10);
        {
            float dacc[1][2][4];
            ds_zero<1>(dacc);
            for (int half = 0; half < 2; ++half) {
                if (half) __syncthreads();                  // everybody is done with the first half's tile
                ds_copy_tile(xs, P, X_c + (size_t)half * 32 * P, P, 32, P, wid, lane);
                ds_mma<1, RB>(dacc, xs, P, k_lo, w_c2h, PC, half * H + k_lo, 4, k_n, g, t4);
            }
            ds_store_part<1>(part, 25, dacc, wid, g, t4);
        }
        __syncthreads();
        float x2 = 0.f;
        if (ework) {
            x2 = ds_sum_part(part, 25, erow, eu);
            X_x2[erow * P + u0 + eu] = ds_round<RB>(x2);
        }
        DS_T(11);
        es_arrive(a.bar + cta, ++seq);
        if (ework) a.x2_all[oh] = x2;
        // ================= E: gru_2 =================
        es_wait(a.bar, n_cta, seq, &gave_up);
        DS_T(12);
        ds_copy_tile(xs, P, X_x2, P, 32, P, wid, lane);
        {
            float dacc[2][2][4];
            ds_zero<2>(dacc);
            ds_mma<2, RB>(dacc, xs, P, k_lo, w_ih2, P, k_lo, 12, k_n, g, t4);
            ds_store_part<2>(part, 25, dacc, wid, g, t4);
        }
        __syncthreads();
        float gi2[3] = {0.f, 0.f, 0.f};
        if (ework) {
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) gi2[gt] = ds_sum_part(part, 25, erow, gt * 4 + eu) + b_ih2[gt];
            if (gave_up) gi2[0] = __int_as_float(0x7fc00000);
            const float r = sigmoidf_precise(gi2[0] + gh2[0]);
            const float z = sigmoidf_precise(gi2[1] + gh2[1]);
            const float n = tanhf(gi2[2] + r * gh2[2]);
            h_state = (1.0f - z) * n + z * h1;
            X_h[erow * P + u0 + eu] = ds_round<RB>(h_state);
        }
        DS_T(13);
        if (t + 1 < Tt) es_arrive(a.bar + cta, ++seq);
        else __syncthreads();
        DS_T(14);
        if (ework) {
            a.h2_all[oh] = h_state;
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) a.gi2_all[o3 + gt * H] = gi2[gt];
        }
    }
#ifdef DS_PROFILE
    if (prof) printf("dec fwd Tt%d: A wait %lld copy %lld mma+cell %lld arrive %lld | B wait %lld copy %lld mma+q %lld | C arr+wait %lld scores %lld softmax+ctx %lld | D arr+wait %lld copy+mma %lld | E arr+wait %lld copy+mma+cell %lld arrive %lld\n",
                     Tt, tacc[0], tacc[1], tacc[2], tacc[3], tacc[4], tacc[5], tacc[6], tacc[7], tacc[8], tacc[9], tacc[10], tacc[11], tacc[12], tacc[13], tacc[14]);
#endif
}

int ds_num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    }
    return n;
}
// Opt-in (VAG_DEC_SEQ=1, read at every call).  Measured on the replayed training step (batch 32, Ts 26, Tt 11, bf16 mode): 1.75 ms
// with this kernel against 1.70 ms with the per-step kernels it replaces — 28 us per step against 23.  Per-step cycles of CTA 0
// (-DDS_PROFILE): A 6.8 k (wait 1.3, copy 2.2, product + cell 2.2, release 1.1) · B 7.5 k · C 19.7 k (wait 4.3, scores 10.9 — one CTA
// per sentence, 96 of 128 CTAs idle —, soft-max + context 4.5) · D 8.4 k · E 12 k: five flag barriers (≈ 2.7 k cycles each with the
// release) and five 66 KB tile copies that all 128 CTAs pull from the same L2 lines cost as much as the five launches they replace.
// What it would take to win: the attention split over four CTAs per sentence, bf16 exchange tiles, and fewer barriers per step.
bool ds_enabled() {
    const char* e = getenv("VAG_DEC_SEQ");
    return e && e[0] == '1';
}
int ds_kw(int H) { return ((H + DS_WARPS - 1) / DS_WARPS + 7) / 8 * 8; }
size_t ds_smem_bytes(int H, int C) { return sizeof(float) * ((size_t)44 * (H + 4) + 4 * (size_t)(C + 4) + 32 * (size_t)(H + 4) + (size_t)DS_WARPS * 32 * 25); }

}  // namespace

size_t dec_seq_scratch_bytes(int H, int C) {   // flags + X_h, X_h1, X_x2 [32][H+4], X_q [32][C+4], X_c [2][32][H+4]
    return ES_MAX_CTAS * sizeof(int) + sizeof(float) * ((size_t)5 * 32 * (H + 4) + (size_t)32 * (C + 4)) + 256;
}
bool dec_seq_fwd_ok(int B, int T, int Tt, int H, int C) {
    return ds_enabled() && B >= 1 && B <= 32 && T >= 1 && T <= 256 && Tt >= 1 && H >= 32 && (H % 8) == 0 && C == 2 * H &&
           H / DS_UNITS <= ds_num_sms() && H / DS_UNITS <= ES_MAX_CTAS && H / DS_UNITS >= B && (H + 4) / 4 <= 160 &&
           2 * C + 256 + 8 <= 32 * (H + 4) && ds_smem_bytes(H, C) <= 220 * 1024;
}

int dec_seq_fwd(DecSeqFwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st) {
    if (!dec_seq_fwd_ok(a.B, a.T, a.Tt, a.H, a.C)) {
        set_error("dec_seq_fwd: unsupported shape");
        return VAG_ERR_UNSUPPORTED;
    }
    if (!scratch || scratch_bytes < dec_seq_scratch_bytes(a.H, a.C)) {
        set_error("dec_seq_fwd: scratch too small");
        return VAG_ERR_WORKSPACE;
    }
    VAG_CUDA(cudaMemsetAsync(scratch, 0, dec_seq_scratch_bytes(a.H, a.C), st));     // flags, pad columns, rows >= B
    a.bar = reinterpret_cast<int*>(scratch);
    a.xch = reinterpret_cast<float*>(reinterpret_cast<char*>(scratch) + ES_MAX_CTAS * sizeof(int));
    a.kw_h = ds_kw(a.H);
    const size_t smem = ds_smem_bytes(a.H, a.C);
    static size_t configured[2] = {0, 0};
    if (smem > configured[round_bf16]) {
        if (round_bf16) VAG_CUDA(cudaFuncSetAttribute(dec_seq_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else VAG_CUDA(cudaFuncSetAttribute(dec_seq_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured[round_bf16] = smem;
    }
    const dim3 grid(a.H / DS_UNITS);
    if (round_bf16) dec_seq_fwd_kernel<true><<<grid, DS_WARPS * 32, smem, st>>>(a);
    else dec_seq_fwd_kernel<false><<<grid, DS_WARPS * 32, smem, st>>>(a);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

}  // namespace vag
