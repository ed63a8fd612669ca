// Fused attention: score → masked softmax over the source positions → context vector.
//
// One CTA per (sentence, group of ≤16 rows).  The K beams of a sentence sit in consecutive rows and share
// the sentence's keys/ctx [T, C], which are therefore read from HBM/L2 once per CTA instead of K times (the
// reference tiles them K times, V11:253-254).  Three phases inside the CTA:
//   1. scores: each warp owns source positions t = warp, warp+nw, …; a lane holds 4-channel slices of the key
//      row in registers and loops over the rows (q staged in shared memory), warp-shuffle reducing over C;
//   2. masked softmax over T, one warp per row;
//   3. context: each thread owns 4 channels, streams ctx[t] once and accumulates all rows.
// Bandwidth bound: algorithmic bytes per sentence = 2·T·C·4 (keys + ctx) + R·C·4·2 (q in, c out).
#include "common.cuh"
#include "split.cuh"
#include <math.h>
#include <stdlib.h>
#include <type_traits>

namespace vag {

constexpr int kMaxRowsPerCta = 16;

// tanh(x) = 1 - 2 / (exp(2x) + 1) on the SFU (ex2 + rcp): ABSOLUTE error ≤ ~1.5e-7 (relative error is poor only
// where |tanh| is tiny).  The attention score sums v_c·tanh(·) over C channels, so absolute accuracy is what
// matters; it sits at the FP32 rounding level of the sum itself while costing 6 instructions instead of ~25.
__device__ __forceinline__ float tanh_abs(float x) {
    const float t = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, t + 1.0f);
}

template <int MODE, bool VEC>
__global__ void __launch_bounds__(256)
attention_kernel(float* __restrict__ c_out, int64_t ld_c, float* __restrict__ alpha_out, const float* __restrict__ q,
                 int64_t ld_q, const float* __restrict__ keys, const float* __restrict__ ctx,
                 const float* __restrict__ v, const float* __restrict__ mask, int rows, int rows_per_sent, int T, int C) {
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x;
    const int r_base = blockIdx.y * kMaxRowsPerCta;  // first row of this group inside the sentence
    const int R = min(kMaxRowsPerCta, rows_per_sent - r_base);
    const int row0 = b * rows_per_sent + r_base;
    if (row0 >= rows) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;

    const int r_cap = min(kMaxRowsPerCta, rows_per_sent);
    float* q_s = smem;                     // [r_cap][C]
    float* v_s = q_s + (size_t)r_cap * C;  // [C]
    float* sc_s = v_s + C;                 // [r_cap][T]  scores then α
    const float* key_b = keys + (int64_t)b * T * C;
    const float* ctx_b = ctx + (int64_t)b * T * C;
    const float* mask_b = mask ? mask + (int64_t)b * T : nullptr;

    for (int i = tid; i < R * C; i += blockDim.x) {
        const int r = i / C, c = i % C;
        q_s[r * C + c] = (row0 + r < rows) ? q[(int64_t)(row0 + r) * ld_q + c] : 0.f;
    }
    if (MODE == VAG_ATTN_MLP)
        for (int i = tid; i < C; i += blockDim.x) v_s[i] = v[i];
    __syncthreads();

    // ---- phase 1: scores
    for (int t = wid; t < T; t += nw) {
        const bool live = mask_b ? (mask_b[t] != 0.f) : true;
        float part[kMaxRowsPerCta];
#pragma unroll
        for (int r = 0; r < kMaxRowsPerCta; ++r) part[r] = 0.f;
        if (live) {
            const float* kr = key_b + (int64_t)t * C;
            if (VEC) {
                for (int c = lane * 4; c < C; c += 128) {
                    const float4 kv = *reinterpret_cast<const float4*>(kr + c);
                    float4 vv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (MODE == VAG_ATTN_MLP) vv = *reinterpret_cast<const float4*>(v_s + c);
#pragma unroll
                    for (int r = 0; r < kMaxRowsPerCta; ++r) {
                        if (r < R) {
                            const float4 qv = *reinterpret_cast<const float4*>(q_s + r * C + c);
                            if (MODE == VAG_ATTN_MLP) {
                                part[r] = fmaf(vv.x, tanh_abs(qv.x + kv.x), part[r]);
                                part[r] = fmaf(vv.y, tanh_abs(qv.y + kv.y), part[r]);
                                part[r] = fmaf(vv.z, tanh_abs(qv.z + kv.z), part[r]);
                                part[r] = fmaf(vv.w, tanh_abs(qv.w + kv.w), part[r]);
                            } else {
                                part[r] = fmaf(qv.x, kv.x, part[r]);
                                part[r] = fmaf(qv.y, kv.y, part[r]);
                                part[r] = fmaf(qv.z, kv.z, part[r]);
                                part[r] = fmaf(qv.w, kv.w, part[r]);
                            }
                        }
                    }
                }
            } else {
                for (int c = lane; c < C; c += 32) {
                    const float kv = kr[c];
#pragma unroll
                    for (int r = 0; r < kMaxRowsPerCta; ++r) {
                        if (r < R) {
                            if (MODE == VAG_ATTN_MLP)
                                part[r] = fmaf(v_s[c], tanh_abs(q_s[r * C + c] + kv), part[r]);
                            else
                                part[r] = fmaf(q_s[r * C + c], kv, part[r]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kMaxRowsPerCta; ++r) {
            if (r < R) {
                const float s = warp_sum(part[r]);
                if (lane == 0) sc_s[r * T + t] = live ? s : -INFINITY;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: softmax over T, one warp per row
    for (int r = wid; r < R; r += nw) {
        float* s = sc_s + r * T;
        float m = -INFINITY;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
        m = warp_max(m);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float e = expf(s[t] - m);
            s[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        for (int t = lane; t < T; t += 32) {
            const float a = s[t] / sum;
            s[t] = a;
            if (alpha_out && row0 + r < rows) alpha_out[(int64_t)(row0 + r) * T + t] = a;
        }
    }
    __syncthreads();

    // ---- phase 3: context
    if (VEC) {
        for (int c = tid * 4; c < C; c += blockDim.x * 4) {
            float4 acc[kMaxRowsPerCta];
#pragma unroll
            for (int r = 0; r < kMaxRowsPerCta; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int t = 0; t < T; ++t) {
                if (mask_b && mask_b[t] == 0.f) continue;  // α is exactly 0 there
                const float4 x = *reinterpret_cast<const float4*>(ctx_b + (int64_t)t * C + c);
#pragma unroll
                for (int r = 0; r < kMaxRowsPerCta; ++r) {
                    if (r < R) {
                        const float a = sc_s[r * T + t];
                        acc[r].x = fmaf(a, x.x, acc[r].x);
                        acc[r].y = fmaf(a, x.y, acc[r].y);
                        acc[r].z = fmaf(a, x.z, acc[r].z);
                        acc[r].w = fmaf(a, x.w, acc[r].w);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kMaxRowsPerCta; ++r)
                if (r < R && row0 + r < rows) *reinterpret_cast<float4*>(c_out + (int64_t)(row0 + r) * ld_c + c) = acc[r];
        }
    } else {
        for (int c = tid; c < C; c += blockDim.x) {
            float acc[kMaxRowsPerCta];
#pragma unroll
            for (int r = 0; r < kMaxRowsPerCta; ++r) acc[r] = 0.f;
            for (int t = 0; t < T; ++t) {
                if (mask_b && mask_b[t] == 0.f) continue;
                const float x = ctx_b[(int64_t)t * C + c];
#pragma unroll
                for (int r = 0; r < kMaxRowsPerCta; ++r)
                    if (r < R) acc[r] = fmaf(sc_s[r * T + t], x, acc[r]);
            }
#pragma unroll
            for (int r = 0; r < kMaxRowsPerCta; ++r)
                if (r < R && row0 + r < rows) c_out[(int64_t)(row0 + r) * ld_c + c] = acc[r];
        }
    }
}


// 1 / (1 + exp(2x)) from a pre-scaled argument u = 2·log2(e)·x: two SFU ops (ex2, rcp) and one add.  tanh(x) = 1 − 2·this.
// Saturates cleanly: u → +inf gives 0 (tanh = 1), u → −inf gives 1 (tanh = −1).
__device__ __forceinline__ float inv_one_plus_exp2(float u) {
    float t, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(u));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t + 1.0f));
    return r;
}
constexpr float kTwoLog2e = 2.8853900817779268f;

// Σ_{i<4} v_i / (1 + 2^{u_i}) with FIVE SFU operations instead of eight: the four reciprocals share one rcp through the
// common denominator  Π(1 + 2^{u_i}).  The kernel is SFU-bound (16 MUFU lanes per SM against 128 FMA lanes), so trading
// three rcp for ~14 FMA-pipe operations shortens the critical pipe by 3/8.  u is clamped to 30: beyond that
// 1/(1+2^u) < 1e-9 — below half an ulp of the O(1) sum it is added to — and the product of four terms stays < 2^124.
__device__ __forceinline__ float sum4_v_over_one_plus_exp2(const float4& v, float u0, float u1, float u2, float u3) {
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

// Factored form for the beam loop: exp(2(q + k)) = exp(2q)·exp(2k).  exp(2k) depends on (sentence, position, channel) only and is
// computed ONCE per decode call (attn_exp_keys), exp(2q) once per (row, channel) when a CTA stages its q rows — the inner loop
// over (row, position, channel) then needs no exponential at all:  Σ_{i<4} v_i / (1 + Eq_i·Ek_i)  costs ONE SFU operation (the
// shared reciprocal) per four elements instead of five, i.e. 0.25 instead of 1.25 MUFU per element, and the kernel moves from the
// SFU roof (16 lanes per SM) to the FMA pipe (128 lanes).  Both exponentials are the PRECISE expf (≤ 2 ulp), so the product
// carries ≤ 3e-7 relative error — an absolute error ≤ 1.5e-7 in tanh, the same as the single-exponential formula above.
// Each 1 + Eq·Ek is clamped to 2^30 (beyond that 1/(1+E) < 1e-9, below half an ulp of the O(1) sum) so that the product of four
// stays finite.  Valid while |q|, |k| ≤ 40 (both exponentials finite and normal); a CTA that sees anything larger takes the
// single-exponential path — checked per sentence for k (attn_exp_keys) and per CTA for q.
constexpr float kFactoredMaxAbs = 40.0f;

// CLAMP = false: the caller has checked max|q| + max|k| ≤ kNoClampMaxSum, i.e. every 1 + Eq·Ek < 2^31 and the product of four
// stays finite without the four min operations (a fifth of the loop's instructions); the results are bit-identical, the clamp
// never engages in that range.
constexpr float kNoClampMaxSum = 10.5f;
template <bool CLAMP>
__device__ __forceinline__ float sum4_v_over_one_plus_prod(const float4& v, const float4& eq, const float4& ek) {
    float a0 = fmaf(eq.x, ek.x, 1.0f), a1 = fmaf(eq.y, ek.y, 1.0f), a2 = fmaf(eq.z, ek.z, 1.0f), a3 = fmaf(eq.w, ek.w, 1.0f);
    if (CLAMP) { a0 = fminf(a0, 1073741824.0f); a1 = fminf(a1, 1073741824.0f); a2 = fminf(a2, 1073741824.0f); a3 = fminf(a3, 1073741824.0f); }
    const float p01 = a0 * a1, p23 = a2 * a3;
    const float n01 = fmaf(v.x, a1, v.y * a0), n23 = fmaf(v.z, a3, v.w * a2);
    const float num = fmaf(n01, p23, n23 * p01);
    return num * rcp_approx(p01 * p23);
}

// ekeys = exp(2·keys) (precise), kflag[b] = bits of max |key| of sentence b (+inf when a key is NaN or outside ±kFactoredMaxAbs:
// non-negative floats order like their bit patterns, so the blocks of a sentence combine with an integer atomicMax).
// One block per (sentence, chunk).
__global__ void __launch_bounds__(256)
attn_exp_keys_kernel(float* __restrict__ ekeys, int* __restrict__ kflag, const float* __restrict__ keys, int TC) {
    const int b = blockIdx.y;
    const float4* src = reinterpret_cast<const float4*>(keys + (int64_t)b * TC);
    float4* dst = reinterpret_cast<float4*>(ekeys + (int64_t)b * TC);
    bool bad = false;
    float kmax = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < TC / 4; i += gridDim.x * blockDim.x) {
        const float4 k = src[i];
        bad |= !(fabsf(k.x) <= kFactoredMaxAbs) | !(fabsf(k.y) <= kFactoredMaxAbs) | !(fabsf(k.z) <= kFactoredMaxAbs) | !(fabsf(k.w) <= kFactoredMaxAbs);
        kmax = fmaxf(fmaxf(kmax, fmaxf(fabsf(k.x), fabsf(k.y))), fmaxf(fabsf(k.z), fabsf(k.w)));
        dst[i] = make_float4(expf(2.0f * k.x), expf(2.0f * k.y), expf(2.0f * k.z), expf(2.0f * k.w));
    }
    kmax = warp_max(kmax);
    const bool any_bad = __syncthreads_or(bad);
    if ((threadIdx.x & 31) == 0) atomicMax(kflag + b, any_bad ? 0x7f800000 : __float_as_int(kmax));
}
int attn_exp_keys(float* ekeys, int* kflag, const float* keys, int B, int T, int C, cudaStream_t st) {
    if ((C & 3) || B <= 0) {
        set_error("attn_exp_keys: C must be a multiple of 4");
        return VAG_ERR_UNSUPPORTED;
    }
    VAG_CUDA(cudaMemsetAsync(kflag, 0, sizeof(int) * (size_t)B, st));
    const int TC = T * C;
    dim3 grid(ceil_div(TC / 4, 256 * 4) > 0 ? ceil_div(TC / 4, 256 * 4) : 1, B);
    attn_exp_keys_kernel<<<grid, 256, 0, st>>>(ekeys, kflag, keys, TC);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Tuned variant for C % 4 == 0 (every real configuration): rows-per-CTA is a template parameter so that the
// accumulator arrays are exactly as large as the beam, a warp loads a whole key row (up to 1024 channels) with
// independent 128-bit loads BEFORE the SFU-heavy row loop (memory-level parallelism), and the context phase
// streams ctx with four positions in flight.  MLP-mode cost is 1.25 SFU ops per (row, position, channel) in the single-exponential
// form, 0.25 in the factored form the beam loop uses (exp(2·keys) precomputed per call, exp(2q) per staged row).
// ---------------------------------------------------------------------------------------------------------
// FULLC: C is a multiple of 1024, so the per-chunk bounds checks (and the branches that fence the SFU chains apart)
// disappear from the inner loops.
template <int MODE, int RCAP, bool FULLC>
__global__ void __launch_bounds__(256, RCAP <= 6 ? 4 : (RCAP <= 12 ? 3 : 2))
attention_tuned_kernel(float* __restrict__ c_out, int64_t ld_c, float* __restrict__ alpha_out, const float* __restrict__ q,
                       int64_t ld_q, const float* __restrict__ keys, const float* __restrict__ ctx,
                       const float* __restrict__ v, const float* __restrict__ mask, int rows, int rows_per_sent, int T, int C,
                       SplitDst sd, const int* __restrict__ done, const float* __restrict__ ekeys, const int* __restrict__ kflag,
                       int force_clamp) {
    pdl_trigger();   // the contraction that follows may start its prologue while this kernel drains
    pdl_wait();      // launched programmatically itself: the query projection before it must have landed
    if (done && *reinterpret_cast<const volatile int*>(done)) return;   // beam search over (block-uniform)
    extern __shared__ __align__(16) float smem[];
    const int b = blockIdx.x;
    const int r_base = blockIdx.y * RCAP;
    const int R = min(RCAP, rows_per_sent - r_base);
    const int row0 = b * rows_per_sent + r_base;
    if (row0 >= rows) return;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    constexpr int NW = 8;
    float* q_s = smem;                    // [RCAP][C]
    float* v_s = q_s + (size_t)RCAP * C;  // [C]
    float* sc_s = v_s + C;                // [RCAP][T]
    constexpr int RPAD = (RCAP + 3) / 4 * 4;
    float* al_s = sc_s + (((size_t)RCAP * T + 3) & ~(size_t)3);  // [T][RPAD], 16-byte aligned: α transposed, so the context phase reads a position's weights as float4
    const float* key_b = keys + (int64_t)b * T * C;
    const float* ctx_b = ctx + (int64_t)b * T * C;
    const float* mask_b = mask ? mask + (int64_t)b * T : nullptr;

    // Stage the q rows.  Factored exponentials (see sum4_v_over_one_plus_prod) are usable when the sentence's keys and this CTA's q
    // rows are in range: the rows are staged as exp(2q) in ONE pass that also checks the range; only a CTA that fails the check
    // (never with real models) stages them again in the single-exponential form.  All RCAP rows are staged — rows the CTA does
    // not own hold zeros (their scores are never read) — so that the score loop carries no per-row guard.
    bool fast = false, noclamp = false;
    {
        const bool try_fast = MODE == VAG_ATTN_MLP && ekeys != nullptr;
        const float kmax = try_fast ? __int_as_float(kflag[b]) : INFINITY;   // +inf: a key out of range (attn_exp_keys)
        bool bad = try_fast && !(kmax <= kFactoredMaxAbs);
        float qmax = 0.f;
        int sr = (tid * 4) / C, scol = tid * 4 - sr * C;     // (row, channel) of element i, advanced without a division per pass
        const int adv_r = 1024 / C, adv_c = 1024 - adv_r * C;
        for (int i = tid * 4; i < RCAP * C; i += 1024) {
            const int r = sr, c = scol;
            sr += adv_r; scol += adv_c;
            if (scol >= C) { scol -= C; ++sr; }
            float4 val = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < R && row0 + r < rows) {
                val = *reinterpret_cast<const float4*>(q + (int64_t)(row0 + r) * ld_q + c);
                if (try_fast) {
                    bad |= !(fabsf(val.x) <= kFactoredMaxAbs) | !(fabsf(val.y) <= kFactoredMaxAbs) | !(fabsf(val.z) <= kFactoredMaxAbs) | !(fabsf(val.w) <= kFactoredMaxAbs);
                    qmax = fmaxf(fmaxf(qmax, fmaxf(fabsf(val.x), fabsf(val.y))), fmaxf(fabsf(val.z), fabsf(val.w)));
                    val.x = exp2x_comp(val.x); val.y = exp2x_comp(val.y); val.z = exp2x_comp(val.z); val.w = exp2x_comp(val.w);
                } else if (MODE == VAG_ATTN_MLP) {
                    val.x *= kTwoLog2e; val.y *= kTwoLog2e; val.z *= kTwoLog2e; val.w *= kTwoLog2e;
                }
            }
            *reinterpret_cast<float4*>(q_s + r * C + c) = val;
        }
        if (try_fast) {
            fast = !__syncthreads_or(bad);
            noclamp = !__syncthreads_or(bad || !(qmax + kmax <= kNoClampMaxSum)) && !force_clamp;
            if (!fast) {
                for (int i = tid * 4; i < R * C; i += 1024) {
                    const int r = i / C, c = i - r * C;
                    if (row0 + r < rows) {
                        float4 val = *reinterpret_cast<const float4*>(q + (int64_t)(row0 + r) * ld_q + c);
                        val.x *= kTwoLog2e; val.y *= kTwoLog2e; val.z *= kTwoLog2e; val.w *= kTwoLog2e;
                        *reinterpret_cast<float4*>(q_s + r * C + c) = val;
                    }
                }
            }
        }
    }
    // MLP mode: Σ_c v_c·tanh(x_c) = Σ_c v_c − 2·Σ_c v_c / (1 + exp(2 x_c)); the first sum is a per-CTA constant
    __shared__ float vsum_s[8];
    float vsum = 0.f;
    if (MODE == VAG_ATTN_MLP) {
        float part_v = 0.f;
        for (int i = tid; i < C; i += 256) { const float vv = v[i]; v_s[i] = vv; part_v += vv; }
        part_v = warp_sum(part_v);
        if (lane == 0) vsum_s[wid] = part_v;
    }
    __syncthreads();
    if (MODE == VAG_ATTN_MLP) {
#pragma unroll
        for (int w = 0; w < 8; ++w) vsum += vsum_s[w];
    }

    // ---- phase 1: scores.  Work items = (live source position, group of RG rows), dealt round-robin to the 8 warps so
    // that short sentences still keep every warp busy (a whole position per warp quantises badly when T_live ~ 8-16).
    constexpr int RG = (RCAP + 3) / 4;
    __shared__ int live_t[256];
    __shared__ int n_live_s;
    if (wid == 0) {  // compact the live positions (any mask pattern, not only prefixes)
        int n = 0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const int t = t0 + lane;
            const bool on = t < T && (mask_b ? mask_b[t] != 0.f : true);
            const unsigned bal = __ballot_sync(0xffffffffu, on);
            if (on && n + __popc(bal & ((1u << lane) - 1)) < 256) live_t[n + __popc(bal & ((1u << lane) - 1))] = t;
            n += __popc(bal);
        }
        if (lane == 0) n_live_s = min(n, 256);
    }
    for (int i = tid; i < R * T; i += 256) sc_s[i] = -INFINITY;
    for (int i = tid; i < T * RPAD; i += 256) al_s[i] = 0.f;
    __syncthreads();
    const int n_live = n_live_s;
    const int n_groups = (R + RG - 1) / RG;
    if (MODE == VAG_ATTN_MLP && fast) {
        const float* ekey_b = ekeys + (int64_t)b * T * C;
        auto score_loop = [&](auto clamp_tag) {
            constexpr bool CLAMP = decltype(clamp_tag)::value;
            // item = (live position, row group), dealt round-robin; the pair is advanced incrementally (no division per item)
            const int step_t = NW / n_groups, step_g = NW - step_t * n_groups;
            int it_t = wid / n_groups, it_g = wid - it_t * n_groups;
            for (int item = wid; item < n_live * n_groups; item += NW) {
                const int t = live_t[it_t];
                const int r_lo = it_g * RG;
                it_t += step_t; it_g += step_g;
                if (it_g >= n_groups) { it_g -= n_groups; ++it_t; }
                float part[RG];
#pragma unroll
                for (int g = 0; g < RG; ++g) part[g] = 0.f;
                const float* kr = ekey_b + (int64_t)t * C;
                for (int cb = 0; cb < C; cb += 1024) {
                    float4 kv[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = cb + j * 128 + lane * 4;
                        kv[j] = (FULLC || c < C) ? *reinterpret_cast<const float4*>(kr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = cb + j * 128 + lane * 4;
                        if (FULLC || c < C) {
                            const float4 vv = *reinterpret_cast<const float4*>(v_s + c);
#pragma unroll
                            for (int g = 0; g < RG; ++g)   // r_lo + g < RCAP always (RG·n_groups covers RCAP rows at most): no guard
                                part[g] += sum4_v_over_one_plus_prod<CLAMP>(vv, *reinterpret_cast<const float4*>(q_s + min(r_lo + g, RCAP - 1) * C + c), kv[j]);
                        }
                    }
                }
                if (RG == 2) {
                    // two reductions in five shuffles: the halves of the warp swap the sum they do not keep, then both butterflies
                    // run in the same four instructions (the same additions in the same order as two separate warp sums)
                    const bool lo = lane < 16;
                    float keep = (lo ? part[0] : part[RG - 1]) + __shfl_xor_sync(0xffffffffu, lo ? part[RG - 1] : part[0], 16);
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) keep += __shfl_xor_sync(0xffffffffu, keep, o);
                    const int r = r_lo + (lo ? 0 : 1);
                    if ((lane & 15) == 0 && r < R) sc_s[r * T + t] = fmaf(-2.0f, keep, vsum);
                } else {
#pragma unroll
                    for (int g = 0; g < RG; ++g) {
                        const int r = r_lo + g;
                        if (r < R) {
                            const float sum = warp_sum(part[g]);
                            if (lane == 0) sc_s[r * T + t] = fmaf(-2.0f, sum, vsum);
                        }
                    }
                }
            }
        };
        if (noclamp) score_loop(std::false_type{});
        else score_loop(std::true_type{});
    } else
    for (int item = wid; item < n_live * n_groups; item += NW) {
        const int t = live_t[item / n_groups];
        const int r_lo = (item % n_groups) * RG;
        float part[RG];
#pragma unroll
        for (int g = 0; g < RG; ++g) part[g] = 0.f;
        const float* kr = key_b + (int64_t)t * C;
        for (int cb = 0; cb < C; cb += 1024) {
            float4 kv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cb + j * 128 + lane * 4;
                kv[j] = (FULLC || c < C) ? *reinterpret_cast<const float4*>(kr + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int g = 0; g < RG; ++g) {
                const int r = r_lo + g;
                if (r < R) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int c = cb + j * 128 + lane * 4;
                        if (FULLC || c < C) {
                            const float4 qv = *reinterpret_cast<const float4*>(q_s + r * C + c);
                            if (MODE == VAG_ATTN_MLP) {
                                const float4 vv = *reinterpret_cast<const float4*>(v_s + c);
                                part[g] += sum4_v_over_one_plus_exp2(vv, fmaf(kv[j].x, kTwoLog2e, qv.x), fmaf(kv[j].y, kTwoLog2e, qv.y),
                                                                     fmaf(kv[j].z, kTwoLog2e, qv.z), fmaf(kv[j].w, kTwoLog2e, qv.w));
                            } else {
                                part[g] = fmaf(qv.x, kv[j].x, part[g]);
                                part[g] = fmaf(qv.y, kv[j].y, part[g]);
                                part[g] = fmaf(qv.z, kv[j].z, part[g]);
                                part[g] = fmaf(qv.w, kv[j].w, part[g]);
                            }
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < RG; ++g) {
            const int r = r_lo + g;
            if (r < R) {
                float s = warp_sum(part[g]);
                if (MODE == VAG_ATTN_MLP) s = fmaf(-2.0f, s, vsum);
                if (lane == 0) sc_s[r * T + t] = s;
            }
        }
    }
    __syncthreads();

    // ---- phase 2: masked softmax over T, one warp per row
    for (int r = wid; r < R; r += NW) {
        float* s = sc_s + r * T;
        float m = -INFINITY;
        for (int t = lane; t < T; t += 32) m = fmaxf(m, s[t]);
        m = warp_max(m);
        float sum = 0.f;
        for (int t = lane; t < T; t += 32) {
            const float e = expf(s[t] - m);
            s[t] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        for (int t = lane; t < T; t += 32) {
            const float a = s[t] / sum;
            s[t] = a;
            al_s[t * RPAD + r] = a;
            if (alpha_out && row0 + r < rows) alpha_out[(int64_t)(row0 + r) * T + t] = a;
        }
    }
    __syncthreads();

    // ---- phase 3: context, each thread owns 4 channels and keeps 4 LIVE positions in flight (α is exactly 0 at masked positions:
    //      they are never loaded and cost no arithmetic)
    for (int c = tid * 4; c < C; c += 1024) {
        float4 acc[RCAP];
#pragma unroll
        for (int r = 0; r < RCAP; ++r) acc[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int i0 = 0; i0 < n_live; i0 += 4) {
            float4 x[4];
            int tt[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const bool on = i0 + u < n_live;
                tt[u] = live_t[on ? i0 + u : 0];
                x[u] = on ? *reinterpret_cast<const float4*>(ctx_b + (int64_t)tt[u] * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u < n_live) {
                    float al[RPAD];
#pragma unroll
                    for (int r4 = 0; r4 < RPAD / 4; ++r4)   // rows ≥ R hold zeros: no per-row guard in the FMA loop
                        *reinterpret_cast<float4*>(al + 4 * r4) = *reinterpret_cast<const float4*>(al_s + tt[u] * RPAD + 4 * r4);
#pragma unroll
                    for (int r = 0; r < RCAP; ++r) {
                        acc[r].x = fmaf(al[r], x[u].x, acc[r].x);
                        acc[r].y = fmaf(al[r], x[u].y, acc[r].y);
                        acc[r].z = fmaf(al[r], x[u].z, acc[r].z);
                        acc[r].w = fmaf(al[r], x[u].w, acc[r].w);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RCAP; ++r)
            if (r < R && row0 + r < rows) {
                if (c_out) *reinterpret_cast<float4*>(c_out + (int64_t)(row0 + r) * ld_c + c) = acc[r];
                if (sd.hi) split_store4(sd, row0 + r, c, acc[r]);   // operand planes for context2hid / the read-out
            }
    }
}

template <int MODE, int RCAP, bool FULLC>
static int launch_attention_tuned(float* c_out, int64_t ld_c, float* alpha, const float* q, int64_t ld_q, const float* keys,
                                  const float* ctx, const float* v, const float* mask, int rows, int rows_per_sent, int T, int C,
                                  cudaStream_t st, SplitDst sd = SplitDst(), const int* done = nullptr, const float* ekeys = nullptr,
                                  const int* kflag = nullptr) {
    const size_t smem = ((size_t)RCAP * C + C + (size_t)RCAP * T + 4 + (size_t)T * ((RCAP + 3) / 4 * 4)) * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("vag_attention_f32: C=%d T=%d needs %zu B of shared memory", C, T, smem);
        return VAG_ERR_UNSUPPORTED;
    }
    static size_t configured = 0;
    if (smem > configured) {
        VAG_CUDA(cudaFuncSetAttribute(attention_tuned_kernel<MODE, RCAP, FULLC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    dim3 grid(rows / rows_per_sent, ceil_div(rows_per_sent, RCAP));
    static const int force_clamp = getenv("VAG_ATTN_CLAMP") && getenv("VAG_ATTN_CLAMP")[0] == '1';   // A/B runs: keep the clamped loop
    VAG_CUDA(launch_pdl(PDL_ATTN, attention_tuned_kernel<MODE, RCAP, FULLC>, grid, dim3(256), smem, st, c_out, ld_c, alpha, q, ld_q, keys, ctx, v,
                        mask, rows, rows_per_sent, T, C, sd, done, ekeys, kflag, force_clamp));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

template <int MODE>
static int dispatch_attention_tuned(float* c_out, int64_t ld_c, float* alpha, const float* q, int64_t ld_q, const float* keys,
                                    const float* ctx, const float* v, const float* mask, int rows, int rows_per_sent, int T,
                                    int C, cudaStream_t st, SplitDst sd = SplitDst(), const int* done = nullptr,
                                    const float* ekeys = nullptr, const int* kflag = nullptr) {
#define VAG_ATT(RC)                                                                                                        \
    do {                                                                                                                   \
        if (C % 1024 == 0)                                                                                                 \
            return launch_attention_tuned<MODE, RC, true>(c_out, ld_c, alpha, q, ld_q, keys, ctx, v, mask, rows, rows_per_sent, T, C, st, sd, done, ekeys, kflag); \
        return launch_attention_tuned<MODE, RC, false>(c_out, ld_c, alpha, q, ld_q, keys, ctx, v, mask, rows, rows_per_sent, T, C, st, sd, done, ekeys, kflag); \
    } while (0)
    if (rows_per_sent == 1) VAG_ATT(1);
    // One CTA per sentence re-reads nothing, but with few sentences (the reference's eval batch of 16; 125 sentences per GPU when
    // a 1000-sentence test set is sharded over 8 GPUs) it leaves most SMs idle and a single CTA's latency IS the kernel's
    // duration: split a sentence's rows over several CTAs (each re-reads the sentence's keys / context from L2) until the grid
    // holds about three CTAs per SM.
    const int B = rows / rows_per_sent, slots = 3 * num_sms();
    static const int env_rcap = getenv("VAG_ATTN_RCAP") ? atoi(getenv("VAG_ATTN_RCAP")) : 0;   // A/B runs: rows per CTA
    if (env_rcap == 4) VAG_ATT(4);
    if (env_rcap == 6) VAG_ATT(6);
    if (env_rcap == 12 && rows_per_sent <= 12) VAG_ATT(12);
    if (rows_per_sent <= 4 || B * ceil_div(rows_per_sent, 8) <= slots) {
        if (rows_per_sent > 4 && B * ceil_div(rows_per_sent, 4) > slots) VAG_ATT(8);
        VAG_ATT(4);
    }
    if (rows_per_sent <= 8) VAG_ATT(8);
    // 9-12 rows per sentence (beam 12): two CTAs of six rows each — 64 registers and 30 KB of shared memory per CTA, four CTAs
    // (32 warps) per SM instead of three CTAs of twelve rows (24 warps); the sentence's exp(2·keys) rows are read twice from L2.
    // Measured on the 1000-sentence decode: bf16 mode 36.2 -> 35.0 ms, FP32 mode (power-capped) unchanged.
    if (rows_per_sent <= 12) VAG_ATT(6);
    VAG_ATT(16);
#undef VAG_ATT
}

// Decoder-step attention of the fused step: the context leaves the kernel only as tensor-core operand planes.
// Requirements (the caller's workspace guarantees them): C % 4 == 0, 16-byte aligned q / keys / ctx, rows_per_sent <= 16.
int attention_mlp_split(SplitDst sd, const float* q, int64_t ld_q, const float* keys, const float* ctx, const float* v,
                        const float* mask, int rows, int rows_per_sent, int T, int C, cudaStream_t st, const int* done, const float* ekeys,
                        const int* kflag) {
    return dispatch_attention_tuned<VAG_ATTN_MLP>(nullptr, 0, nullptr, q, ld_q, keys, ctx, v, mask, rows, rows_per_sent, T, C, st, sd, done, ekeys,
                                                  kflag);
}

}  // namespace vag

using namespace vag;

extern "C" int vag_attention_f32(float* c_out, int64_t ld_c, float* alpha, const float* q, int64_t ld_q, const float* keys,
                                 const float* ctx, const float* v, const float* mask, int rows, int rows_per_sent, int T,
                                 int C, int mode, vag_stream_t stream) {
    VAG_REQUIRE(c_out && q && keys && ctx, "vag_attention_f32: null pointer");
    VAG_REQUIRE(mode == VAG_ATTN_MLP || mode == VAG_ATTN_DOT, "vag_attention_f32: bad mode %d", mode);
    VAG_REQUIRE(mode == VAG_ATTN_DOT || v, "vag_attention_f32: MLP mode needs v");
    VAG_REQUIRE(rows >= 0 && rows_per_sent > 0 && T > 0 && C > 0, "vag_attention_f32: bad shape");
    VAG_REQUIRE(rows % rows_per_sent == 0, "vag_attention_f32: rows (%d) not a multiple of rows_per_sent (%d)", rows, rows_per_sent);
    if (rows == 0) return VAG_OK;
    const int B = rows / rows_per_sent;
    const int r_cap = rows_per_sent < kMaxRowsPerCta ? rows_per_sent : kMaxRowsPerCta;
    const size_t smem = ((size_t)r_cap * C + C + (size_t)r_cap * T) * sizeof(float);
    if (smem > 227 * 1024) {
        set_error("vag_attention_f32: C=%d T=%d needs %zu B of shared memory", C, T, smem);
        return VAG_ERR_UNSUPPORTED;
    }
    auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
    const bool vec = (C % 4 == 0) && (ld_c % 4 == 0) && al(keys) && al(ctx) && al(c_out);
    if (vec && (ld_q % 4 == 0) && al(q)) {
        if (mode == VAG_ATTN_MLP)
            return dispatch_attention_tuned<VAG_ATTN_MLP>(c_out, ld_c, alpha, q, ld_q, keys, ctx, v, mask, rows, rows_per_sent, T, C,
                                                          (cudaStream_t)stream);
        return dispatch_attention_tuned<VAG_ATTN_DOT>(c_out, ld_c, alpha, q, ld_q, keys, ctx, v, mask, rows, rows_per_sent, T, C,
                                                      (cudaStream_t)stream);
    }
    dim3 grid(B, ceil_div(rows_per_sent, kMaxRowsPerCta));
    cudaStream_t st = (cudaStream_t)stream;
#define VAG_ATTN_LAUNCH(MODE, VEC)                                                                                  \
    do {                                                                                                            \
        VAG_CUDA(cudaFuncSetAttribute(attention_kernel<MODE, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize,     \
                                      (int)smem));                                                                  \
        attention_kernel<MODE, VEC><<<grid, 256, smem, st>>>(c_out, ld_c, alpha, q, ld_q, keys, ctx, v, mask, rows, \
                                                             rows_per_sent, T, C);                                  \
    } while (0)
    if (mode == VAG_ATTN_MLP) {
        if (vec) VAG_ATTN_LAUNCH(VAG_ATTN_MLP, true); else VAG_ATTN_LAUNCH(VAG_ATTN_MLP, false);
    } else {
        if (vec) VAG_ATTN_LAUNCH(VAG_ATTN_DOT, true); else VAG_ATTN_LAUNCH(VAG_ATTN_DOT, false);
    }
#undef VAG_ATTN_LAUNCH
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
