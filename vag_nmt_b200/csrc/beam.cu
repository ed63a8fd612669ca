// Beam-search kernels: candidate selection (masking + top-K per sentence), parent reorder, finalisation.
// Reference: V11.beamsearch, models/NMT_AttentionImagine_Seq2Seq_Beam_V11.py:233-337.
//
// Selection = one CTA per sentence scanning its K·V candidates once, coalesced; every thread keeps a sorted
// private top-K in registers (static indexing only), then K rounds of a block arg-max pop the winners in
// canonical order (score descending, flat index k·V+v ascending on ties).  HBM bound: the logits are read once.
#include "common.cuh"
#include "split.cuh"
#include <stdlib.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <math.h>

namespace vag {

constexpr int kMaxBeam = 16;
constexpr int kEOS = 3;
constexpr float kBeamInf = -1e5f;  // V11:257

__device__ __forceinline__ bool cand_better(float av, int ai, float bv, int bi) {
    return av > bv || (av == bv && ai < bi);
}

template <int KMAX>
struct TopList {
    float v[KMAX];
    int i[KMAX];
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < KMAX; ++j) { v[j] = -INFINITY; i[j] = 0x7fffffff; }
    }
    // keep the best KMAX seen so far, sorted best-first (a superset of the top-K for any K <= KMAX; every
    // index is static so the list stays in registers)
    __device__ __forceinline__ void push(float c, int ci) {
        if (!cand_better(c, ci, v[KMAX - 1], i[KMAX - 1])) return;  // quick reject against the current worst
        bool prev_better = true;  // "old[j-1] is better than c"; true for j == 0 by convention
        float carry_v = 0.f; int carry_i = 0;
#pragma unroll
        for (int j = 0; j < KMAX; ++j) {
            const float ov = v[j]; const int oi = i[j];
            const bool old_better = cand_better(ov, oi, c, ci);
            if (!old_better) {
                if (prev_better) { v[j] = c; i[j] = ci; }       // insertion point
                else { v[j] = carry_v; i[j] = carry_i; }       // shifted down
            }
            carry_v = ov; carry_i = oi;
            prev_better = old_better;
        }
    }
    __device__ __forceinline__ void pop() {
#pragma unroll
        for (int j = 0; j + 1 < KMAX; ++j) { v[j] = v[j + 1]; i[j] = i[j + 1]; }
        v[KMAX - 1] = -INFINITY; i[KMAX - 1] = 0x7fffffff;
    }
};

// logits [rows, V] (ld), lse [rows] or NULL (then `logits` already holds log-probabilities).
// done (optional): device flag; when *done != 0 the kernel leaves all state untouched.
template <int KMAX>
__global__ void __launch_bounds__(256)
beam_select_kernel(const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
                   const int64_t* __restrict__ prev_tokens, float* __restrict__ nll, int64_t* __restrict__ tokens_out,
                   int32_t* __restrict__ parents_out, int K, int64_t V, int step, int avoid_double,
                   const int* __restrict__ done, int* __restrict__ fin_counter) {
    if (done && *done) return;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ float nll_s[kMaxBeam];
    __shared__ float lse_s[kMaxBeam];
    __shared__ int64_t cur_s[kMaxBeam];
    __shared__ float red_v[8];
    __shared__ int red_i[8];
    __shared__ int red_t[8];
    __shared__ int win_t;

    const int Kin = step == 0 ? 1 : K;  // rows of this sentence feeding the selection
    if (tid < Kin) {
        const int row = b * Kin + tid;
        nll_s[tid] = step == 0 ? 0.f : nll[(int64_t)b * K + tid];
        lse_s[tid] = lse ? lse[row] : 0.f;
        cur_s[tid] = step == 0 ? -1 : prev_tokens[(int64_t)b * K + tid];
    }
    __syncthreads();

    TopList<KMAX> top;
    top.init();
    const int64_t total = (int64_t)Kin * V;
    for (int k = 0; k < Kin; ++k) {
        const float* row = logits + (int64_t)(b * Kin + k) * ld;
        const float base = nll_s[k];
        const float l = lse_s[k];
        const int64_t cur = cur_s[k];
        const bool fin = (step > 0) && (cur == kEOS);
        for (int64_t vtok = tid; vtok < V; vtok += blockDim.x) {
            float lp;
            if (fin) {
                lp = (vtok == kEOS) ? 0.f : kBeamInf;  // V11:291-294
            } else {
                lp = row[vtok] - l;
                if (avoid_double && step > 0 && vtok == cur) lp = kBeamInf;  // V11:279-280
            }
            const float c = step == 0 ? lp : base + lp;  // V11:297
            top.push(c, (int)(k * V + vtok));
        }
    }
    (void)total;

    int n_eos = 0;
    for (int round = 0; round < K; ++round) {
        float bv = top.v[0];
        int bi = top.i[0];
        int bt = tid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
        }
        if (lane == 0) { red_v[wid] = bv; red_i[wid] = bi; red_t[wid] = bt; }
        __syncthreads();
        if (tid == 0) {
            float fv = red_v[0]; int fi = red_i[0]; int ft = red_t[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
                if (cand_better(red_v[w], red_i[w], fv, fi)) { fv = red_v[w]; fi = red_i[w]; ft = red_t[w]; }
            win_t = ft;
            const int64_t tok = (int64_t)fi % V;
            const int par = (int)((int64_t)fi / V);
            nll[(int64_t)b * K + round] = fv;
            tokens_out[(int64_t)b * K + round] = tok;
            parents_out[(int64_t)b * K + round] = par;
            n_eos += (tok == kEOS);
        }
        __syncthreads();
        if (tid == win_t) top.pop();
    }
    if (tid == 0 && fin_counter) atomicAdd(fin_counter, n_eos);
}

// ---------------------------------------------------------------------------------------------------------
// Fast selection for real vocabularies (V >= kFastMinV): one CTA per sentence, ONE pass over the logits.
//
// Every thread owns a strided slice of each of the sentence's K rows.  During the single pass it keeps, per row,
// an online (max, Σexp) pair and the BEST element of its slice (shared memory, [row][thread]).  After the rows'
// log-sum-exps are combined, the best-of-slice values become scores  nll_k + (logit − lse_k)  and the block pops
// the K winners in canonical order (score desc, flat index k·V+v asc) with K rounds of a block arg-max; only the
// thread that owned a winner rescans that one slice (L2-resident) for its next-best element.  Exact for any data.
// The repeated token is skipped (V11:279-280); a finished hypothesis offers exactly one candidate, <eos> at +0
// (V11:291-294) — their −1e5 siblings can never be selected when V − 1 >= K, which every accepted V satisfies.
// Instruction budget ≈ 6 per logit (ex2, max, compare) so the kernel is bound by the one read of the logits.
// ---------------------------------------------------------------------------------------------------------
constexpr int kFastMinV = 512;
constexpr float kLog2e = 1.4426950408889634f;

// best element of thread `tid`'s slice of `row` that is strictly worse than (pv, pi) in the canonical order
__device__ __forceinline__ void slice_next_best(const float* __restrict__ row, int V, int tid, bool vec4, int skip, float pv,
                                                int pi, float& bv, int& bi) {
    bv = -INFINITY;
    bi = 0x7fffffff;
    if (vec4) {
        for (int v0 = tid * 4; v0 < V; v0 += 1024) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int v = v0 + j;
                if (v < V) {
                    const float x = row[v];
                    if (v != skip && cand_better(pv, pi, x, v) && cand_better(x, v, bv, bi)) { bv = x; bi = v; }
                }
            }
        }
    } else {
        for (int v = tid; v < V; v += 256) {
            const float x = row[v];
            if (v != skip && cand_better(pv, pi, x, v) && cand_better(x, v, bv, bi)) { bv = x; bi = v; }
        }
    }
}

template <int KMAX>
__global__ void __launch_bounds__(256)
beam_select_fast_kernel(const float* __restrict__ logits, int64_t ld, int subtract_lse,
                        const int64_t* __restrict__ prev_tokens, float* __restrict__ nll, int64_t* __restrict__ tokens_out,
                        int32_t* __restrict__ parents_out, int K, int V, int step, int avoid_double,
                        const int* __restrict__ done, int* __restrict__ fin_counter) {
    if (done && *done) return;
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ float best_v[KMAX][256];   // best remaining element of slice (row, thread): raw logit, later the score
    __shared__ int best_i[KMAX][256];     // its token id (0x7fffffff: slice exhausted)
    __shared__ float red_m[KMAX][8], red_s[KMAX][8];
    __shared__ float nll_s[KMAX], lse_s[KMAX];
    __shared__ int cur_s[KMAX];
    __shared__ float red_a[8];
    __shared__ int red_i[8], red_t[8];
    __shared__ int win_t, win_k, win_tok;
    __shared__ float win_v;

    const int Kin = step == 0 ? 1 : K;
    const bool vec4 = ((ld & 3) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);  // thread owns 4 consecutive tokens
    if (tid < Kin) {
        nll_s[tid] = step == 0 ? 0.f : nll[(int64_t)b * K + tid];
        cur_s[tid] = step == 0 ? -1 : (int)prev_tokens[(int64_t)b * K + tid];
    }
    __syncthreads();

    // ---- single pass over the sentence's rows
    for (int k = 0; k < Kin; ++k) {
        const int cur = cur_s[k];
        if (step > 0 && cur == kEOS) {  // finished hypothesis: the only candidate is <eos> at +0, parked in slice 0
            best_v[k][tid] = tid == 0 ? 0.f : -INFINITY;
            best_i[k][tid] = tid == 0 ? kEOS : 0x7fffffff;
            if (lane == 0) { red_m[k][wid] = 0.f; red_s[k][wid] = wid == 0 ? 1.f : 0.f; }  // lse = 0 ⇒ score = nll_k + 0
            continue;
        }
        const float* row = logits + (int64_t)(b * Kin + k) * ld;
        const int skip = (avoid_double && step > 0) ? cur : -1;
        float m = -INFINITY, s = 0.f, bv = -INFINITY;
        int bi = 0x7fffffff;
        if (vec4) {
            const int v_full = V & ~3;  // float4 groups entirely below V
            for (int v0 = tid * 4; v0 < V; v0 += 2048) {
                const int v1 = v0 + 1024;
                float x[8];
                const float4 a = *reinterpret_cast<const float4*>(row + v0);  // rows are padded to a multiple of 4
                const float4 c = v1 < V ? *reinterpret_cast<const float4*>(row + v1) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = c.x; x[5] = c.y; x[6] = c.z; x[7] = c.w;
                if (v0 >= v_full) {  // ragged tail group: mask the padding
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (v0 + j >= V) x[j] = -INFINITY;
                }
                if (v1 < V && v1 >= v_full) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (v1 + j >= V) x[4 + j] = -INFINITY;
                }
                float cm = x[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) cm = fmaxf(cm, x[j]);
                if (cm > m) { s *= exp2f((m - cm) * kLog2e); m = cm; }
                const float m2 = m * kLog2e;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += exp2f(fmaf(x[j], kLog2e, -m2));
                if (cm >= bv) {  // rare after the first groups: someone may beat the slice's best
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int v = (j < 4 ? v0 : v1 - 4) + j;
                        if (x[j] != -INFINITY && v != skip && cand_better(x[j], v, bv, bi)) { bv = x[j]; bi = v; }
                    }
                }
            }
        } else {
            for (int v0 = tid; v0 < V; v0 += 2048) {
                float x[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) x[j] = v0 + j * 256 < V ? row[v0 + j * 256] : -INFINITY;
                float cm = x[0];
#pragma unroll
                for (int j = 1; j < 8; ++j) cm = fmaxf(cm, x[j]);
                if (cm > m) { s *= exp2f((m - cm) * kLog2e); m = cm; }
                const float m2 = m * kLog2e;
#pragma unroll
                for (int j = 0; j < 8; ++j) s += exp2f(fmaf(x[j], kLog2e, -m2));
                if (cm >= bv) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int v = v0 + j * 256;
                        if (x[j] != -INFINITY && v != skip && cand_better(x[j], v, bv, bi)) { bv = x[j]; bi = v; }
                    }
                }
            }
        }
        best_v[k][tid] = bv;
        best_i[k][tid] = bi;
        // warp-level (max, Σexp); the skipped token still belongs to the softmax denominator
        const float wm = warp_max(m);
        const float ws = warp_sum(m == -INFINITY ? 0.f : s * exp2f((m - wm) * kLog2e));
        if (lane == 0) { red_m[k][wid] = wm; red_s[k][wid] = ws; }
    }
    __syncthreads();
    if (tid < Kin) {
        float fm = red_m[tid][0];
        for (int w = 1; w < 8; ++w) fm = fmaxf(fm, red_m[tid][w]);
        float fs = 0.f;
        for (int w = 0; w < 8; ++w) fs += red_s[tid][w] * exp2f((red_m[tid][w] - fm) * kLog2e);
        const bool fin = step > 0 && cur_s[tid] == kEOS;
        lse_s[tid] = (subtract_lse && !fin) ? fm + logf(fs) : 0.f;
    }
    __syncthreads();
    // ---- raw logits → scores; every thread caches the head of its K slices
    float hv = -INFINITY;
    int hi = 0x7fffffff;  // flat index k·V + token
    for (int k = 0; k < Kin; ++k) {
        const int ti = best_i[k][tid];
        if (ti != 0x7fffffff) {
            const float lp = best_v[k][tid] - lse_s[k];
            const float sc = step == 0 ? lp : nll_s[k] + lp;  // V11:297
            best_v[k][tid] = sc;
            if (cand_better(sc, k * V + ti, hv, hi)) { hv = sc; hi = k * V + ti; }
        }
    }
    // ---- K rounds: block arg-max, then only the winner's owner refills that one slice
    int n_eos = 0;
    for (int round = 0; round < K; ++round) {
        float bv = hv;
        int bi = hi, bt = tid;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int ot = __shfl_xor_sync(0xffffffffu, bt, o);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bt = ot; }
        }
        if (lane == 0) { red_a[wid] = bv; red_i[wid] = bi; red_t[wid] = bt; }
        __syncthreads();
        if (tid == 0) {
            float fv = red_a[0];
            int fi = red_i[0], ft = red_t[0];
            for (int w = 1; w < 8; ++w)
                if (cand_better(red_a[w], red_i[w], fv, fi)) { fv = red_a[w]; fi = red_i[w]; ft = red_t[w]; }
            const int par = fi / V, tok = fi - par * V;
            win_t = ft; win_k = par; win_tok = tok; win_v = fv;
            nll[(int64_t)b * K + round] = fv;
            tokens_out[(int64_t)b * K + round] = tok;
            parents_out[(int64_t)b * K + round] = par;
            n_eos += (tok == kEOS);
        }
        __syncthreads();
        if (tid == win_t) {
            const int k = win_k;
            const int cur = cur_s[k];
            float nv = -INFINITY;
            int ni = 0x7fffffff;
            if (!(step > 0 && cur == kEOS)) {
                const float* row = logits + (int64_t)(b * Kin + k) * ld;
                const int skip = (avoid_double && step > 0) ? cur : -1;
                // the popped element's RAW logit is needed as the bound; recover it from the row itself
                slice_next_best(row, V, tid, vec4, skip, row[win_tok], win_tok, nv, ni);
                if (ni != 0x7fffffff) {
                    const float lp = nv - lse_s[k];
                    nv = step == 0 ? lp : nll_s[k] + lp;
                }
            }
            best_v[k][tid] = nv;
            best_i[k][tid] = ni;
            hv = -INFINITY;
            hi = 0x7fffffff;
            for (int kk = 0; kk < Kin; ++kk) {
                const int ti = best_i[kk][tid];
                if (ti != 0x7fffffff && cand_better(best_v[kk][tid], kk * V + ti, hv, hi)) { hv = best_v[kk][tid]; hi = kk * V + ti; }
            }
        }
    }
    if (tid == 0 && fin_counter) atomicAdd(fin_counter, n_eos);
}

// ---------------------------------------------------------------------------------------------------------
// Selection from the projection epilogue's summaries: the tensor-core kernel that produced the logits also left,
// per (row, column tile of width tile_w), the tile's running max, Σexp(x − max) and arg-max (lowest column on
// ties).  One CTA per sentence combines the tiles' (max, Σexp) into the row log-sum-exps, turns every tile's
// arg-max into a score and pops the K winners with K block arg-max rounds; after each pop ONE warp rescans the
// winner's tile in the logits (tile_w floats, L2-resident) for its next-best element, so the result is exactly
// what a scan of all K·V logits gives while only ~K tiles per sentence are ever re-read.
// ---------------------------------------------------------------------------------------------------------
template <int KMAX>
__global__ void __launch_bounds__(256)
beam_select_summary_kernel(const float4* __restrict__ summ, int n_tiles, int tile_w, const float* __restrict__ logits, int64_t ld,
                           const int64_t* __restrict__ prev_tokens, float* __restrict__ nll, int64_t* __restrict__ tokens_out,
                           int32_t* __restrict__ parents_out, int K, int V, int step, int avoid_double,
                           const int* __restrict__ done, int* __restrict__ fin_counter) {
    if (done && *done) return;
    extern __shared__ float dyn[];
    const int b = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int Kin = step == 0 ? 1 : K;
    const int n_sl = Kin * n_tiles;
    float* sl_v = dyn;                                   // [Kin·n_tiles] score of the slice's best remaining element
    int* sl_i = reinterpret_cast<int*>(dyn + n_sl);      // its token (0x7fffffff: exhausted)
    __shared__ float nll_s[KMAX], lse_s[KMAX];
    __shared__ int cur_s[KMAX];
    __shared__ float red_a[8];
    __shared__ int red_i[8], red_t[8];
    __shared__ int win_sl, win_tok;
    constexpr float kL2e = 1.4426950408889634f;

    if (tid < Kin) {
        nll_s[tid] = step == 0 ? 0.f : nll[(int64_t)b * K + tid];
        cur_s[tid] = step == 0 ? -1 : (int)prev_tokens[(int64_t)b * K + tid];
    }
    __syncthreads();
    // ---- row log-sum-exps from the tile summaries (one warp per row)
    for (int k = wid; k < Kin; k += 8) {
        const float4* sr = summ + (int64_t)(b * Kin + k) * n_tiles;
        float m = -INFINITY;
        for (int t = lane; t < n_tiles; t += 32) m = fmaxf(m, sr[t].x);
        m = warp_max(m);
        float s = 0.f;
        for (int t = lane; t < n_tiles; t += 32) s += sr[t].y * exp2f((sr[t].x - m) * kL2e);
        s = warp_sum(s);
        if (lane == 0) lse_s[k] = (step > 0 && cur_s[k] == kEOS) ? 0.f : m + logf(s);
    }
    __syncthreads();

    // best element of tile `tile` of row k strictly after (pv, pi) in canonical order, skipping `skip`; warp-cooperative
    auto tile_next = [&](int k, int tile, int skip, float pv, int pi, float& bv, int& bi) {
        const float* row = logits + (int64_t)(b * Kin + k) * ld;
        const int c_lo = tile * tile_w, c_hi = min(V, c_lo + tile_w);
        bv = -INFINITY;
        bi = 0x7fffffff;
        for (int c = c_lo + lane; c < c_hi; c += 32) {
            const float x = row[c];
            if (c != skip && cand_better(pv, pi, x, c) && cand_better(x, c, bv, bi)) { bv = x; bi = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
        }
    };

    // ---- slice scores
    for (int idx = tid; idx < n_sl; idx += 256) {
        const int k = idx / n_tiles, tile = idx - k * n_tiles;
        float sc = -INFINITY;
        int tok = 0x7fffffff;
        if (step > 0 && cur_s[k] == kEOS) {              // finished hypothesis: one candidate, <eos> at +0 (V11:291-294)
            if (tile == 0) { sc = nll_s[k] + 0.f; tok = kEOS; }
        } else {
            const float4 e = summ[(int64_t)(b * Kin + k) * n_tiles + tile];
            tok = __float_as_int(e.w);
            const float lp = e.z - lse_s[k];
            sc = step == 0 ? lp : nll_s[k] + lp;        // V11:297
        }
        sl_v[idx] = sc;
        sl_i[idx] = tok;
    }
    __syncthreads();
    // the repeated token may not be chosen (V11:279-280): where it is a tile's arg-max, replace it by the runner-up
    if (avoid_double && step > 0) {
        for (int k = wid; k < Kin; k += 8) {
            const int cur = cur_s[k];
            if (cur == kEOS || cur < 0 || cur >= V) continue;
            const int tile = cur / tile_w;
            const int idx = k * n_tiles + tile;
            if (sl_i[idx] == cur) {
                float nv; int ni;
                tile_next(k, tile, cur, INFINITY, -1, nv, ni);
                if (lane == 0) {
                    sl_i[idx] = ni;
                    sl_v[idx] = ni == 0x7fffffff ? -INFINITY : nll_s[k] + (nv - lse_s[k]);
                }
            }
        }
        __syncthreads();
    }

    // ---- K rounds of block arg-max + one-tile refill
    int n_eos = 0;
    for (int round = 0; round < K; ++round) {
        float bv = -INFINITY;
        int bi = 0x7fffffff, bs = -1;                    // bi = flat index k·V + token
        for (int idx = tid; idx < n_sl; idx += 256) {
            const int tok = sl_i[idx];
            if (tok != 0x7fffffff) {
                const int flat = (idx / n_tiles) * V + tok;
                if (cand_better(sl_v[idx], flat, bv, bi)) { bv = sl_v[idx]; bi = flat; bs = idx; }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            const int os = __shfl_xor_sync(0xffffffffu, bs, o);
            if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; bs = os; }
        }
        if (lane == 0) { red_a[wid] = bv; red_i[wid] = bi; red_t[wid] = bs; }
        __syncthreads();
        if (tid == 0) {
            float fv = red_a[0];
            int fi = red_i[0], fs = red_t[0];
            for (int w = 1; w < 8; ++w)
                if (cand_better(red_a[w], red_i[w], fv, fi)) { fv = red_a[w]; fi = red_i[w]; fs = red_t[w]; }
            const int par = fi / V, tok = fi - par * V;
            win_sl = fs; win_tok = tok;
            nll[(int64_t)b * K + round] = fv;
            tokens_out[(int64_t)b * K + round] = tok;
            parents_out[(int64_t)b * K + round] = par;
            n_eos += (tok == kEOS);
        }
        __syncthreads();
        if (wid == 0 && round + 1 < K) {                 // refill the winner's tile
            const int idx = win_sl, k = idx / n_tiles, tile = idx - k * n_tiles;
            const int cur = cur_s[k];
            float nv = -INFINITY;
            int ni = 0x7fffffff;
            if (!(step > 0 && cur == kEOS)) {
                const float* row = logits + (int64_t)(b * Kin + k) * ld;
                tile_next(k, tile, (avoid_double && step > 0) ? cur : -1, row[win_tok], win_tok, nv, ni);
            }
            if (lane == 0) {
                sl_i[idx] = ni;
                sl_v[idx] = ni == 0x7fffffff ? -INFINITY : (step == 0 ? nv - lse_s[k] : nll_s[k] + (nv - lse_s[k]));
            }
        }
        __syncthreads();
    }
    if (tid == 0 && fin_counter) atomicAdd(fin_counter, n_eos);
}

// ---------------------------------------------------------------------------------------------------------
// Selection WITHOUT logits (fused decode step): the vocabulary contraction leaves only, per (row, 128-column tile),
// the best two elements in canonical order and Σexp relative to the best (tc_gemm_top2).  The K winners are popped as
// in the kernel above; a tile that has to supply a THIRD candidate (rare: ≈ C(K,3)/tiles² of the step-0 rows) has its
// 128 logits recomputed by one warp from the operand planes of the read-out and of the projection matrix with the
// tensor core's own three-product arithmetic, excluding the columns already taken.  451 MB of logits per step (1000
// sentences, beam 12) are neither written nor read.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float plane_f(int mode, uint16_t u) {
    return mode == 2 ? __bfloat162float(__ushort_as_bfloat16(u)) : __half2float(__ushort_as_half(u));
}

constexpr int kSel2Threads = 128;   // 4 warps per sentence; 28 KB of entries ⇒ 7 CTAs per SM (1000 sentences in one wave)

// Candidate of one slice held in registers by the thread that owns the slice's entry.
struct SelCand {
    float sc;      // score (V11:297) of the slice's next candidate
    int flat;      // k·V + token  (0x7fffffff: empty list slot)
    int idx;       // entry index = slice·Kin + k
    int depth;     // 0: the slice's best, 1: its second, 2: recomputed
    float sc2;     // depth 0 only: score of the slice's second element ...
    int tok2;      // ... and its column (0xFFFF: the slice has no second; -2: not loaded, read the summary again)
};

// Optional tail of the selection kernel: the work of beam_advance_fused_kernel for the CTA's own sentence — state rows gathered
// by parent (fp32 + operand planes), operand planes of the chosen tokens' embeddings — and, by the LAST CTA to finish (ticket
// counter), the stop test and the host progress word.  One launch less per decoder step; the copies (≈ 100 MB per step at 12 000
// rows) overlap the latency-bound pops of the other sentences on the SM.
struct SelAdvance {
    float* h_next = nullptr;          // nullptr: disabled (the caller launches beam_advance_fused itself)
    const float* h_cur = nullptr;
    int H = 0, E = 0;
    SplitDst h_sd, e_sd;
    const uint16_t* emb_hi = nullptr;
    const uint16_t* emb_lo = nullptr;
    int64_t ld_emb = 0;
    int* done = nullptr;
    int* steps_run = nullptr;
    int* ticket = nullptr;            // zeroed per decode call, one word per step
    volatile int32_t* host_progress = nullptr;
    int nonce = 0;
};

// NT = threads per sentence: 128 when there are enough sentences to fill the GPU (7 CTAs per SM), 512 for small batches (the
// reference's eval batch of 16; a test set sharded over 8 GPUs), where the kernel's duration is ONE CTA's latency and that latency
// is the summary scan (294 x K float4 loads per sentence, four in flight per thread): more workers per row shorten it fourfold.
template <int KMAX, int NT, int UNR>
__global__ void __launch_bounds__(NT)
beam_select_top2_kernel(const float4* __restrict__ summ /*[n_slices][n_rows]*/, int n_slices, int slice_w, int n_rows,
                        const uint16_t* __restrict__ t_hi, const uint16_t* __restrict__ t_lo, int64_t ld_t,
                        const uint16_t* __restrict__ w_hi, const uint16_t* __restrict__ w_lo, int64_t ld_w,
                        const float* __restrict__ bias, int E, int mode, const int64_t* __restrict__ prev_tokens,
                        float* __restrict__ nll, int64_t* __restrict__ tokens_out, int32_t* __restrict__ parents_out, int K,
                        int V, int step, int avoid_double, const int* __restrict__ done, int* __restrict__ fin_counter,
                        int force_recompute, long long* __restrict__ dbg, const SelAdvance adv) {
    pdl_trigger();   // the kernel that follows may become resident while this one drains
    pdl_wait();      // launched programmatically itself (PDL_BEAM): the vocabulary summaries must have landed
    if (done && *reinterpret_cast<const volatile int*>(done)) return;
    extern __shared__ float dyn[];
    const int b = blockIdx.x;
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int Kin = step == 0 ? 1 : K;
    const int n_sl = Kin * n_slices;                     // entry idx = slice·Kin + k (k fastest: coalesced summary reads)
    // Thread t works for ONE row: k = t mod Kin, as worker j = t / Kin of the row's J = NT / Kin workers (slices j, j+J, …;
    // threads ≥ J·Kin idle).  One pass over the summaries then yields, in registers, the row's soft-max partials AND the
    // thread's three best candidates (ordering inside a row does not depend on the row's log-sum-exp).
    const int J = NT / Kin, my_k = tid % Kin, my_j = tid / Kin;
    const bool worker = my_j < J;
    // 8 bytes per entry, only read again when a thread's list runs empty: (raw logit of the entry's next candidate — or its
    // SCORE once it has been advanced, flagged by bit 20 —, token | depth << 16), 0x7fffffff = exhausted
    float* sl_v = dyn;
    int* sl_i = reinterpret_cast<int*>(dyn + n_sl);
    __shared__ float nll_s[KMAX], lse_s[KMAX];
    __shared__ int cur_s[KMAX];
    __shared__ float part_m[NT], part_s[NT];
    __shared__ float red_a[2][NW], rc_a[NW];
    __shared__ int red_k[2][NW], rc_i[NW];
    __shared__ int win_par[KMAX], win_tok[KMAX];
    __shared__ float win_v[KMAX];
    __shared__ int req_flag[2], req_idx, req_tid;        // pending block-wide recomputation of one slice (flag per iteration parity)
    __shared__ int n_eos_s;
    constexpr float kL2e = 1.4426950408889634f;
    const int64_t row0 = (int64_t)b * Kin;

    if (tid < Kin) {
        nll_s[tid] = step == 0 ? 0.f : nll[(int64_t)b * K + tid];
        cur_s[tid] = step == 0 ? -1 : (int)prev_tokens[(int64_t)b * K + tid];
    }
    if (tid == 0) { req_flag[0] = req_flag[1] = 0; n_eos_s = 0; }
    __syncthreads();

    // ---- the thread's sorted list of its three best entries (sc = RAW logit until the log-sum-exps are known)
    SelCand L0, L1, L2;
    L0.flat = L1.flat = L2.flat = 0x7fffffff;
    L0.sc = L1.sc = L2.sc = -INFINITY;
    int n_l = 0;
    auto insert = [&](const SelCand& c) {   // keeps the best three; the caller guarantees c may legally enter (see apply_next)
        if (!cand_better(c.sc, c.flat, L2.sc, L2.flat)) return;
        if (cand_better(c.sc, c.flat, L0.sc, L0.flat)) { L2 = L1; L1 = L0; L0 = c; }
        else if (cand_better(c.sc, c.flat, L1.sc, L1.flat)) { L2 = L1; L1 = c; }
        else { L2 = c; }
        n_l = min(n_l + 1, 3);
    };
    const int my_cur = cur_s[my_k];
    const bool row_done = step > 0 && my_cur == kEOS;
    const int my_skip = (avoid_double && step > 0) ? my_cur : -1;   // the repeated token may not be chosen (V11:279-280)
    float pm = -INFINITY, ps = 0.f;
    if (worker) {
        if (row_done) {                                  // finished hypothesis: one candidate, <eos> at +0 (V11:291-294)
            for (int sl = my_j; sl < n_slices; sl += J) sl_i[sl * Kin + my_k] = 0x7fffffff;
            if (my_j == 0) {
                SelCand c;
                c.sc = 0.f; c.flat = my_k * V + kEOS; c.idx = my_k; c.depth = 2; c.sc2 = -INFINITY; c.tok2 = 0xFFFF;
                sl_v[my_k] = 0.f;
                sl_i[my_k] = kEOS | (2 << 16);
                insert(c);
            }
        } else {
            // Straight-line scan (round 2: the branchy version — early-out insert with three struct shuffles, two-way soft-max update
            // — ran every path for nearly every entry because SOME lane of the warp took it): the list is three (raw logit, token,
            // entry) triples updated with selects, the soft-max partial is one FMA whose operands are selected (bit-identical to
            // the two-branch form), and the full candidates are built once after the loop.
            const float4* sp = summ + row0 + my_k;
            float lv0 = -INFINITY, lv1 = -INFINITY, lv2 = -INFINITY;
            int lt0 = 0x7fffffff, lt1 = 0x7fffffff, lt2 = 0x7fffffff, li0 = 0, li1 = 0, li2 = 0;
            for (int sl0 = my_j; sl0 < n_slices; sl0 += UNR * J) {
                float4 e[UNR];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {          // UNR independent loads in flight
                    const int sl = sl0 + u * J;
                    e[u] = sl < n_slices ? sp[(int64_t)sl * n_rows] : make_float4(-INFINITY, 0.f, -INFINITY, __int_as_float((int)0xFFFFFFFFu));
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int sl = sl0 + u * J;
                    if (sl >= n_slices) break;
                    const int idx = sl * Kin + my_k;
                    {   // soft-max partial: ps·2^(pm − x) + y when x is a new maximum, else ps + y·2^(x − pm) — one FMA either way
                        const bool up = e[u].x > pm;
                        const float ex = exp2f((up ? pm - e[u].x : e[u].x - pm) * kL2e);
                        const float upd = fmaf(up ? ps : e[u].y, ex, up ? e[u].y : ps);
                        ps = (up || e[u].x != -INFINITY) ? upd : ps;
                        pm = up ? e[u].x : pm;
                    }
                    const int bits = __float_as_int(e[u].w), i1 = bits & 0xFFFF, i2 = (bits >> 16) & 0xFFFF;
                    const bool first = i1 != 0xFFFF && i1 != my_skip;                  // the slice's best may be chosen
                    const bool second = !first && i1 != 0xFFFF && i2 != 0xFFFF;          // i1 == skip ⇒ i2 != skip
                    const float cv = first ? e[u].x : (second ? e[u].z : -INFINITY);
                    const int ct = first ? i1 : (second ? i2 : 0x7fffffff);
                    sl_v[idx] = cv;
                    sl_i[idx] = first ? i1 : (second ? (i2 | (1 << 16)) : 0x7fffffff);
                    // keep the best three (logit descending, token ascending): an empty candidate (−inf, INT_MAX) beats nothing
                    const bool b0 = cand_better(cv, ct, lv0, lt0), b1 = cand_better(cv, ct, lv1, lt1), b2 = cand_better(cv, ct, lv2, lt2);
                    lv2 = b1 ? lv1 : (b2 ? cv : lv2); lt2 = b1 ? lt1 : (b2 ? ct : lt2); li2 = b1 ? li1 : (b2 ? idx : li2);
                    lv1 = b0 ? lv0 : (b1 ? cv : lv1); lt1 = b0 ? lt0 : (b1 ? ct : lt1); li1 = b0 ? li0 : (b1 ? idx : li1);
                    lv0 = b0 ? cv : lv0;              lt0 = b0 ? ct : lt0;              li0 = b0 ? idx : li0;
                }
            }
            // full candidates of the three survivors: depth from the entry just written, the slice's runner-up (depth 0 only) from
            // the summary again (three independent loads, L2-resident)
            auto build = [&](SelCand& L, float v, int t, int idx) {
                if (t == 0x7fffffff) return;
                L.sc = v; L.flat = my_k * V + t; L.idx = idx; L.depth = (sl_i[idx] >> 16) & 0xF; L.sc2 = -INFINITY; L.tok2 = 0xFFFF;
                if (L.depth == 0) {
                    const float4 e2 = sp[(int64_t)(idx / Kin) * n_rows];
                    L.sc2 = e2.z;
                    L.tok2 = (__float_as_int(e2.w) >> 16) & 0xFFFF;
                }
                ++n_l;
            };
            build(L0, lv0, lt0, li0);
            build(L1, lv1, lt1, li1);
            build(L2, lv2, lt2, li2);
        }
    }
    part_m[tid] = pm;
    part_s[tid] = ps;
    __syncthreads();
    if (tid < Kin) {                                     // row log-sum-exp from its J workers' partials
        float m = -INFINITY;
        for (int j = 0; j < J; ++j) m = fmaxf(m, part_m[j * Kin + tid]);
        float sum = 0.f;
        for (int j = 0; j < J; ++j) {
            const float pmj = part_m[j * Kin + tid];
            if (pmj != -INFINITY) sum += part_s[j * Kin + tid] * exp2f((pmj - m) * kL2e);
        }
        lse_s[tid] = (step > 0 && cur_s[tid] == kEOS) ? 0.f : m + logf(sum);
    }
    __syncthreads();
    auto score_of = [&](int k, float logit) { const float lp = logit - lse_s[k]; return step == 0 ? lp : nll_s[k] + lp; };
    // raw logits → scores (a finished row's <eos> candidate is nll + 0)
    if (L0.flat != 0x7fffffff) { L0.sc = score_of(my_k, L0.sc); L0.sc2 = score_of(my_k, L0.sc2); }
    if (L1.flat != 0x7fffffff) { L1.sc = score_of(my_k, L1.sc); L1.sc2 = score_of(my_k, L1.sc2); }
    if (L2.flat != 0x7fffffff) { L2.sc = score_of(my_k, L2.sc); L2.sc2 = score_of(my_k, L2.sc2); }
    // rebuild the list from the entry arrays (only when it ran empty: the thread took three winners in a row); depth-0
    // entries still hold raw logits, advanced entries hold scores
    auto rescan = [&]() {
        L0.flat = L1.flat = L2.flat = 0x7fffffff;
        L0.sc = L1.sc = L2.sc = -INFINITY;
        n_l = 0;
        if (!worker) return;
        for (int sl = my_j; sl < n_slices; sl += J) {
            const int idx = sl * Kin + my_k;
            const int pk = sl_i[idx];
            if (pk != 0x7fffffff) {
                SelCand c;
                c.depth = (pk >> 16) & 0xF;
                c.sc = (pk & (1 << 20)) ? sl_v[idx] : score_of(my_k, sl_v[idx]);   // bit 20: the value is already a score
                c.flat = my_k * V + (pk & 0xFFFF); c.idx = idx; c.sc2 = -INFINITY; c.tok2 = -2;
                insert(c);
            }
        }
    };
    // the owner replaces the popped head by the slice's next candidate (nt == 0x7fffffff: the slice is exhausted)
    auto apply_next = [&](int idx, int k, float nsc, int nt, int nd) {
        sl_i[idx] = nt == 0x7fffffff ? nt : (nt | (nd << 16) | (1 << 20));
        sl_v[idx] = nsc;
        L0 = L1; L1 = L2; L2.flat = 0x7fffffff; L2.sc = -INFINITY;   // pop the head
        --n_l;
        if (n_l == 0) { rescan(); return; }                          // includes the entry just written
        if (nt == 0x7fffffff) return;
        SelCand c;
        c.sc = nsc; c.flat = k * V + nt; c.idx = idx; c.depth = nd; c.sc2 = -INFINITY; c.tok2 = 0xFFFF;
        // the list holds the TRUE best n_l entries of this thread; a candidate below its tail may be beaten by entries the
        // list never saw, so it only enters when it beats the tail (or fills a slot the pop just freed AND beats the tail)
        const SelCand& tail = n_l == 2 ? L1 : L0;
        if (cand_better(c.sc, c.flat, tail.sc, tail.flat)) insert(c);
    };
    __syncthreads();

    // K rounds of a block arg-max over the list heads.  A head travels as (score, flat << 9 | holder's thread): the flat index
    // sits in the high bits, so the packed word orders ties exactly like the flat index alone, and the holder — which IS the
    // owner of the winner's slice — is known without the two integer divisions every thread used to execute per round.  Every
    // thread finishes the cross-warp step itself (double-buffered partials), so a round has ONE block barrier and no
    // single-thread section: the owner records the winner and advances its slice while the others already reduce the next round.
    int it = 0;
    for (int round = 0; round < K;) {
        float bv = L0.sc;
        int bk = L0.flat == 0x7fffffff ? 0x7fffffff : ((L0.flat << 9) | tid);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
            if (cand_better(ov, ok, bv, bk)) { bv = ov; bk = ok; }
        }
        const int pb = it & 1;
        ++it;
        if (lane == 0) { red_a[pb][wid] = bv; red_k[pb][wid] = bk; }
        __syncthreads();
        // The flag is double-buffered like the partials: this iteration reads word pb, an owner that asks for a recomputation
        // writes word pb ^ 1 (read after the NEXT barrier) — with one barrier per round a single word would be written by a fast
        // owner while slower warps still test it.
        if (req_flag[pb]) {
            // block-uniform: the previous winner's slice needs a third (or later) candidate.  The CTA recomputes the slice's
            // logits from the operand planes with the tensor core's own three products (hi·hi + (lo·hi + hi·lo)·2^-11);
            // warp = every NW-th column, lane = 8 consecutive k; columns already taken are excluded.
            const int idx = req_idx, sl = idx / Kin, k = idx - sl * Kin;
            const int skip = (avoid_double && step > 0) ? cur_s[k] : -1;
            float nv = -INFINITY;
            int ni = 0x7fffffff;
            if (force_recompute != 2) {
                const uint16_t* th = t_hi + (row0 + k) * ld_t;
                const uint16_t* tl = t_lo + (row0 + k) * ld_t;
                const int c_lo = sl * slice_w, c_hi = min(V, c_lo + slice_w);
                for (int c = c_lo + wid; c < c_hi; c += NW) {
                    bool taken = c == skip;
                    for (int w = 0; w < round; ++w) taken |= (win_par[w] == k && win_tok[w] == c);
                    if (taken) continue;   // warp-uniform
                    const uint16_t* wh = w_hi + (int64_t)c * ld_w;
                    const uint16_t* wl = w_lo + (int64_t)c * ld_w;
                    float main_acc = 0.f, cross_acc = 0.f;
                    for (int e0 = lane * 8; e0 < E; e0 += 256) {
                        const uint4 a_h = *reinterpret_cast<const uint4*>(th + e0), b_h = *reinterpret_cast<const uint4*>(wh + e0);
                        const uint32_t ah[4] = {a_h.x, a_h.y, a_h.z, a_h.w}, bh[4] = {b_h.x, b_h.y, b_h.z, b_h.w};
                        uint32_t al[4] = {0u, 0u, 0u, 0u}, bl[4] = {0u, 0u, 0u, 0u};
                        if (mode != 2) {
                            const uint4 a_l = *reinterpret_cast<const uint4*>(tl + e0), b_l = *reinterpret_cast<const uint4*>(wl + e0);
                            al[0] = a_l.x; al[1] = a_l.y; al[2] = a_l.z; al[3] = a_l.w;
                            bl[0] = b_l.x; bl[1] = b_l.y; bl[2] = b_l.z; bl[3] = b_l.w;
                        }
#pragma unroll
                        for (int u = 0; u < 8; ++u) {
                            const int sh = (u & 1) * 16;
                            const float xh = plane_f(mode, (uint16_t)(ah[u >> 1] >> sh)), yh = plane_f(mode, (uint16_t)(bh[u >> 1] >> sh));
                            main_acc = fmaf(xh, yh, main_acc);
                            if (mode != 2) {
                                const float xl = plane_f(mode, (uint16_t)(al[u >> 1] >> sh)), yl = plane_f(mode, (uint16_t)(bl[u >> 1] >> sh));
                                cross_acc = fmaf(xl, yh, cross_acc);
                                cross_acc = fmaf(xh, yl, cross_acc);
                            }
                        }
                    }
                    main_acc = warp_sum(main_acc);
                    cross_acc = warp_sum(cross_acc);
                    const float x = (main_acc + cross_acc * (1.0f / 2048.0f)) + (bias ? bias[c] : 0.f);
                    if (cand_better(x, c, nv, ni)) { nv = x; ni = c; }
                }
            }
            if (lane == 0) { rc_a[wid] = nv; rc_i[wid] = ni; }
            __syncthreads();
            if (tid == req_tid) {                         // the owner installs the recomputed candidate
                nv = rc_a[0]; ni = rc_i[0];
                for (int w = 1; w < NW; ++w)
                    if (cand_better(rc_a[w], rc_i[w], nv, ni)) { nv = rc_a[w]; ni = rc_i[w]; }
                apply_next(idx, k, ni == 0x7fffffff ? -INFINITY : score_of(k, nv), ni, 2);
                req_flag[pb] = 0;
            }
            __syncthreads();
            continue;                                     // redo the arg-max of this round with the owner's updated list
        }
        // cross-warp step, by every warp: lane l takes partial l mod NW, log2(NW) butterfly levels leave the winner in all lanes
        float fv = red_a[pb][lane & (NW - 1)];
        int fk = red_k[pb][lane & (NW - 1)];
#pragma unroll
        for (int o = NW / 2; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, fv, o);
            const int ok = __shfl_xor_sync(0xffffffffu, fk, o);
            if (cand_better(ov, ok, fv, fk)) { fv = ov; fk = ok; }
        }
        const bool none = fk == 0x7fffffff;               // no candidate left (needs V <= K + 1: excluded by the launcher)
        if (tid == (none ? 0 : (fk & 511))) {             // the holder of the winning head = the owner of its slice
            const int k = my_k, idx = L0.idx;
            const int tok = none ? 0 : L0.flat - k * V;
            win_par[round] = none ? 0 : k;
            win_tok[round] = tok;
            win_v[round] = fv;
            if (tok == kEOS) atomicAdd(&n_eos_s, 1);
            if (round + 1 < K && !none) {                 // next candidate of the winner's slice
                const int sl = idx / Kin;
                const int cur = cur_s[k];
                const int skip = (avoid_double && step > 0) ? cur : -1;
                bool recompute = false;
                float nsc = -INFINITY;
                int nt = 0x7fffffff;
                if (!(step > 0 && cur == kEOS)) {
                    if (L0.depth == 0 && force_recompute != 1) {
                        int tok2 = L0.tok2;
                        float sc2 = L0.sc2;
                        if (tok2 == -2) {                 // list entry rebuilt by a rescan: fetch the slice's second again
                            const float4 e = summ[(int64_t)sl * n_rows + row0 + k];
                            tok2 = (__float_as_int(e.w) >> 16) & 0xFFFF;
                            sc2 = tok2 == 0xFFFF ? -INFINITY : score_of(k, e.z);
                        }
                        if (tok2 == 0xFFFF) { /* single-column slice: exhausted */ }
                        else if (tok2 != skip) { nsc = sc2; nt = tok2; }
                        else recompute = true;
                    } else {
                        recompute = true;
                    }
                }
                if (dbg) { atomicAdd((unsigned long long*)dbg + 20, (unsigned long long)recompute); atomicAdd((unsigned long long*)dbg + 21, 1ull); }
                if (recompute) { req_idx = idx; req_tid = tid; req_flag[pb ^ 1] = 1; }   // seen by everybody after the next barrier
                else apply_next(idx, k, nsc, nt, 1);
            }
        }
        ++round;
    }
    __syncthreads();
    if (tid < K) {
        nll[(int64_t)b * K + tid] = win_v[tid];
        tokens_out[(int64_t)b * K + tid] = win_tok[tid];
        parents_out[(int64_t)b * K + tid] = win_par[tid];
    }
    if (adv.h_next) {
        // ---- the sentence's share of the reorder (see SelAdvance): new row n = b·K + k takes the state of row b·Kin + parent
        const int H4 = adv.H >> 2, E8 = adv.E >> 3;
        for (int k0 = 0; k0 < K; k0 += 4) {               // four rows' loads in flight per thread
            for (int c = tid; c < H4; c += NT) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u < K) v[u] = reinterpret_cast<const float4*>(adv.h_cur + (row0 + (Kin == 1 ? 0 : win_par[k0 + u])) * adv.H)[c];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (k0 + u < K) {
                        const int64_t n = (int64_t)b * K + k0 + u;
                        reinterpret_cast<float4*>(adv.h_next + n * adv.H)[c] = v[u];
                        split_store4(adv.h_sd, n, c * 4, v[u]);
                    }
            }
        }
        if (adv.e_sd.hi) {
            for (int i = tid; i < K * E8; i += NT) {
                const int k = i / E8, c = i - k * E8;
                int64_t id = win_tok[k];
                if (id < 0 || id >= V) id = 0;
                const int64_t n = (int64_t)b * K + k;
                *reinterpret_cast<uint4*>(adv.e_sd.hi + n * adv.e_sd.ld + c * 8) = *reinterpret_cast<const uint4*>(adv.emb_hi + id * adv.ld_emb + c * 8);
                if (adv.e_sd.mode != 2)
                    *reinterpret_cast<uint4*>(adv.e_sd.lo + n * adv.e_sd.ld + c * 8) = *reinterpret_cast<const uint4*>(adv.emb_lo + id * adv.ld_emb + c * 8);
            }
        }
        if (tid == 0) {
            // stop test (V11:265-269) by the last CTA to get here: every CTA adds its <eos> count BEFORE it takes a ticket
            atomicAdd(fin_counter, n_eos_s);
            __threadfence();
            if (atomicAdd(adv.ticket, 1) == (int)gridDim.x - 1) {
                __threadfence();
                const int fin = atomicAdd(fin_counter, 0) == (int)gridDim.x * K;
                adv.steps_run[0] = step + 1;
                if (fin) *adv.done = 1;
                if (adv.host_progress) {   // mapped host memory: the host throttles / stops its launch loop on this word
                    *adv.host_progress = (adv.nonce << 16) | (fin << 15) | (step + 1);
                    __threadfence_system();
                }
            }
        }
        return;
    }
    if (tid == 0 && fin_counter) atomicAdd(fin_counter, n_eos_s);
}

// Row gather by parent + early-stop bookkeeping.
//   h_next[b*K + k, :] = h_cur[b*Kin + parents[b,k], :]
// Block (0,0) thread 0 also turns the per-step EOS counter into the `done` flag / steps_run the way the
// reference's host-side test does (V11:265-269): all B·K tokens chosen at this step are EOS ⇒ stop.
__global__ void __launch_bounds__(128)
beam_advance_kernel(float* __restrict__ h_next, const float* __restrict__ h_cur, const int32_t* __restrict__ parents, int B,
                    int K, int Kin, int H, int step, int* __restrict__ done, const int* __restrict__ fin_counter,
                    int* __restrict__ steps_run) {
    if (*done) return;
    const int n = blockIdx.x;  // new row
    const int b = n / K;
    const int p = b * Kin + (Kin == 1 ? 0 : parents[n]);
    const float4* src = reinterpret_cast<const float4*>(h_cur + (int64_t)p * H);
    float4* dst = reinterpret_cast<float4*>(h_next + (int64_t)n * H);
    if (H % 4 == 0) {
        for (int c = threadIdx.x; c < H / 4; c += blockDim.x) dst[c] = src[c];
    } else {
        for (int c = threadIdx.x; c < H; c += blockDim.x) h_next[(int64_t)n * H + c] = h_cur[(int64_t)p * H + c];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) steps_run[0] = step + 1;
}
// Fused-step flavour: one launch reorders the hidden state by parent (fp32 + tensor-core operand planes), gathers the
// operand planes of the next step's input embeddings from the pre-split table, and evaluates the stop test (the
// separate one-thread kernel of the plain path).  A block that sees the flag flip mid-launch may skip its copy: the
// state of a finished search is never read again.
__global__ void __launch_bounds__(128)
beam_advance_fused_kernel(float* __restrict__ h_next, const float* __restrict__ h_cur, const int32_t* __restrict__ parents,
                          const int64_t* __restrict__ tokens, int B, int K, int Kin, int H, int step, int* __restrict__ done,
                          const int* __restrict__ fin_counter, int* __restrict__ steps_run, SplitDst h_sd, SplitDst e_sd,
                          const uint16_t* __restrict__ t_hi, const uint16_t* __restrict__ t_lo, int64_t ld_t, int E, int64_t V,
                          volatile int32_t* host_progress, int nonce) {
    pdl_trigger();   // the contraction that follows may start its prologue while this kernel drains
    pdl_wait();      // launched programmatically itself (PDL_BEAM): the selection's parents / tokens must have landed
    if (*reinterpret_cast<volatile int*>(done)) return;
    constexpr int RPB = 4;   // rows per block: 12000 one-row blocks were launch-overhead bound
    for (int n = blockIdx.x * RPB; n < min((int)(blockIdx.x + 1) * RPB, B * K); ++n) {   // new row
        const int b = n / K;
        const int p = b * Kin + (Kin == 1 ? 0 : parents[n]);
        const float4* src = reinterpret_cast<const float4*>(h_cur + (int64_t)p * H);
        float4* dst = reinterpret_cast<float4*>(h_next + (int64_t)n * H);
        for (int c = threadIdx.x; c < H / 4; c += blockDim.x) {
            const float4 v = src[c];
            dst[c] = v;
            split_store4(h_sd, n, c * 4, v);
        }
        int64_t id = e_sd.hi ? tokens[n] : 0;
        if (id < 0 || id >= V) id = 0;
        for (int c = threadIdx.x; e_sd.hi && c < E / 8; c += blockDim.x) {
            *reinterpret_cast<uint4*>(e_sd.hi + (int64_t)n * e_sd.ld + c * 8) = *reinterpret_cast<const uint4*>(t_hi + id * ld_t + c * 8);
            if (e_sd.mode != 2)
                *reinterpret_cast<uint4*>(e_sd.lo + (int64_t)n * e_sd.ld + c * 8) = *reinterpret_cast<const uint4*>(t_lo + id * ld_t + c * 8);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        steps_run[0] = step + 1;
        const int fin = *fin_counter == B * K;
        if (fin) *done = 1;
        if (host_progress) {   // mapped host memory: the host throttles / stops its launch loop on this word (vag_beam_decode_f32)
            *host_progress = (nonce << 16) | (fin << 15) | (step + 1);
            __threadfence_system();
        }
    }
}
__global__ void beam_done_kernel(int* __restrict__ done, const int* __restrict__ fin_counter, int total, const int* __restrict__ steps_run,
                                 volatile int32_t* host_progress, int nonce) {
    const int already = *done;
    const int fin = already || *fin_counter == total;
    if (fin) *done = 1;
    if (host_progress) {
        *host_progress = (nonce << 16) | (fin << 15) | (*steps_run & 0x7FFF);
        __threadfence_system();
    }
}

// Finalisation (V11:315-337): backtrace every final hypothesis through the parent pointers, force EOS in the
// last row, normalise by #tokens>3 (clamped to 1), take the best hypothesis per sentence.
// One warp per sentence, lane k = hypothesis k.
__global__ void __launch_bounds__(32)
beam_finalize_kernel(const int64_t* __restrict__ tok_hist /*[L,B,K]*/, const int32_t* __restrict__ par_hist /*[L,B,K]*/,
                     const float* __restrict__ nll, const int* __restrict__ steps_run, int B, int K, int L,
                     int64_t* __restrict__ hyp_out /*[B,L]*/, int32_t* __restrict__ hyp_len, int64_t* __restrict__ beam_out /*[L,B,K] or NULL*/) {
    const int b = blockIdx.x;
    const int k = threadIdx.x;
    const int S = *steps_run;  // rows of the beam that were filled
    int count = 0;
    if (k < K) {
        int p = k;
        for (int di = L - 1; di >= 0; --di) {
            int64_t tok;
            if (di >= S) {
                tok = 0;
            } else {
                tok = tok_hist[((int64_t)di * B + b) * K + p];
                p = par_hist[((int64_t)di * B + b) * K + p];
            }
            if (di == L - 1) tok = kEOS;  // V11:315
            if (beam_out) beam_out[((int64_t)di * B + b) * K + k] = tok;
            count += (tok > 3);
        }
    }
    float score = -INFINITY;
    if (k < K) score = nll[(int64_t)b * K + k] / fmaxf((float)count, 1.0f);  // V11:318-321
    float bv = score;
    int bk = k;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int ok = __shfl_xor_sync(0xffffffffu, bk, o);
        if (ov > bv || (ov == bv && ok < bk)) { bv = ov; bk = ok; }
    }
    if (k == bk) {
        int p = k;
        int first_eos = L;
        for (int di = L - 1; di >= 0; --di) {
            int64_t tok;
            if (di >= S) {
                tok = 0;
            } else {
                tok = tok_hist[((int64_t)di * B + b) * K + p];
                p = par_hist[((int64_t)di * B + b) * K + p];
            }
            if (di == L - 1) tok = kEOS;
            hyp_out[(int64_t)b * L + di] = tok;
            if (tok == kEOS) first_eos = di;
        }
        hyp_len[b] = first_eos;
    }
}

// Row arg-max over V (greedy decoding, V11:207-216); ties → lowest index.  Writes the token twice:
// into the [B, L] result (column `step`) and into the next-step input vector.
__global__ void __launch_bounds__(256)
row_argmax_kernel(const float* __restrict__ logits, int64_t ld, int64_t V, int64_t* __restrict__ out, int64_t out_stride,
                  int64_t* __restrict__ next_in) {
    const int r = blockIdx.x;
    const float* row = logits + (int64_t)r * ld;
    float bv = -INFINITY;
    int bi = 0x7fffffff;
    for (int64_t i = threadIdx.x; i < V; i += blockDim.x) {
        const float x = row[i];
        if (cand_better(x, (int)i, bv, bi)) { bv = x; bi = (int)i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (cand_better(ov, oi, bv, bi)) { bv = ov; bi = oi; }
    }
    __shared__ float sv[8];
    __shared__ int si[8];
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
            if (cand_better(sv[w], si[w], bv, bi)) { bv = sv[w]; bi = si[w]; }
        out[(int64_t)r * out_stride] = bi;
        if (next_in) next_in[r] = bi;
    }
}

int beam_select(const float* logits, int64_t ld, const float* lse, const int64_t* prev_tokens, float* nll,
                int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step, int avoid_double,
                const int* done, int* fin_counter, cudaStream_t st) {
    if (K > kMaxBeam) {
        set_error("beam size %d > %d is not supported", K, kMaxBeam);
        return VAG_ERR_UNSUPPORTED;
    }
    if ((int64_t)K * V >= 0x7fffffff) {
        set_error("K*V overflows the 31-bit candidate index");
        return VAG_ERR_UNSUPPORTED;
    }
    const bool fast = V >= kFastMinV && V > K + 1;
    const int subtract = lse ? 1 : 0;  // the fast kernel computes the row log-sum-exp itself
#define VAG_SELECT(KM)                                                                                                    \
    do {                                                                                                                  \
        if (fast)                                                                                                         \
            beam_select_fast_kernel<KM><<<B, 256, 0, st>>>(logits, ld, subtract, prev_tokens, nll, tokens_out, parents_out, \
                                                           K, (int)V, step, avoid_double, done, fin_counter);              \
        else                                                                                                              \
            beam_select_kernel<KM><<<B, 256, 0, st>>>(logits, ld, lse, prev_tokens, nll, tokens_out, parents_out, K, V, step, \
                                                      avoid_double, done, fin_counter);                                   \
    } while (0)
    if (K <= 4) VAG_SELECT(4);
    else if (K <= 8) VAG_SELECT(8);
    else if (K <= 12) VAG_SELECT(12);
    else VAG_SELECT(16);
#undef VAG_SELECT
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int beam_select_summary(const float4* summ, int tile_w, const float* logits, int64_t ld, const int64_t* prev_tokens, float* nll,
                        int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step, int avoid_double,
                        const int* done, int* fin_counter, cudaStream_t st) {
    if (K > kMaxBeam || (int64_t)K * V >= 0x7fffffff || V <= K + 1) {
        set_error("beam_select_summary: unsupported K=%d V=%lld", K, (long long)V);
        return VAG_ERR_UNSUPPORTED;
    }
    const int n_tiles = (int)((V + tile_w - 1) / tile_w);
    const int Kin = step == 0 ? 1 : K;
    const size_t smem = (size_t)Kin * n_tiles * 8;
#define VAG_SELS(KM)                                                                                                          \
    beam_select_summary_kernel<KM><<<B, 256, smem, st>>>(summ, n_tiles, tile_w, logits, ld, prev_tokens, nll, tokens_out,      \
                                                         parents_out, K, (int)V, step, avoid_double, done, fin_counter)
    if (K <= 4) VAG_SELS(4);
    else if (K <= 8) VAG_SELS(8);
    else if (K <= 12) VAG_SELS(12);
    else VAG_SELS(16);
#undef VAG_SELS
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

long long* tc_debug();
// summ: [ceil(V / slice_w)][n_rows] (tc_gemm_top2), n_rows = B (step 0) or B·K
int beam_select_top2(const float4* summ, int slice_w, SplitDst t, const uint16_t* w_hi, const uint16_t* w_lo, int64_t ld_w,
                     const float* bias, int E, const int64_t* prev_tokens, float* nll, int64_t* tokens_out, int32_t* parents_out,
                     int B, int K, int64_t V, int step, int avoid_double, const int* done, int* fin_counter, cudaStream_t st,
                     const SelAdvance* adv_in) {
    const SelAdvance adv = adv_in ? *adv_in : SelAdvance();
    if (K > kMaxBeam || (int64_t)K * V >= 0x7fffffff || V <= K + 1 || V >= 0xFFFF || (E % 8)) {
        set_error("beam_select_top2: unsupported K=%d V=%lld E=%d", K, (long long)V, E);
        return VAG_ERR_UNSUPPORTED;
    }
    const int n_slices = (int)((V + slice_w - 1) / slice_w);
    const int Kin = step == 0 ? 1 : K;
    const size_t smem = (size_t)Kin * n_slices * 8 + 16;
    if (smem > 200 * 1024) {
        set_error("beam_select_top2: K=%d V=%lld needs %zu B of shared memory", K, (long long)V, smem);
        return VAG_ERR_UNSUPPORTED;
    }
    // VAG_SELECT_RECOMPUTE=1 (tests): never use a slice's stored runner-up, always recompute — exercises the rare path
    const char* fe = getenv("VAG_SELECT_RECOMPUTE");
    const int force = fe ? (fe[0] == '1' ? 1 : (fe[0] == '2' ? 2 : 0)) : 0;   // 2: timing experiments only (never recompute)
    // threads per sentence / summary loads in flight: VAG_SELECT_NT = 128 | 256 | 512, VAG_SELECT_UNR = 4 | 8 override the choice (A/B runs)
    static const int env_nt = getenv("VAG_SELECT_NT") ? atoi(getenv("VAG_SELECT_NT")) : 0;
    static const int env_unr = getenv("VAG_SELECT_UNR") ? atoi(getenv("VAG_SELECT_UNR")) : 0;
    const int nt = env_nt ? env_nt : (B <= 2 * num_sms() ? 512 : kSel2Threads);   // few sentences: 512 threads each (see the kernel)
    const int unr = env_unr ? env_unr : 4;
#define VAG_SEL2_NT(KM, NT_, UNR_)                                                                                            \
    do {                                                                                                                      \
        static size_t configured = 0;                                                                                         \
        if (smem > 48 * 1024 && smem > configured) {                                                                          \
            VAG_CUDA(cudaFuncSetAttribute(beam_select_top2_kernel<KM, NT_, UNR_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            configured = smem;                                                                                                \
        }                                                                                                                     \
        VAG_CUDA(launch_pdl(PDL_BEAM, beam_select_top2_kernel<KM, NT_, UNR_>, dim3(B), dim3(NT_), smem, st, summ, n_slices, slice_w, B * Kin, \
                            (const uint16_t*)t.hi, (const uint16_t*)t.lo, (int64_t)t.ld, w_hi, w_lo, ld_w, bias, E, (int)t.mode, prev_tokens, nll, tokens_out, \
                            parents_out, K, (int)V, step, avoid_double, done, fin_counter, force, tc_debug(), adv));           \
    } while (0)
#define VAG_SEL2(KM)                                                                                                          \
    do {                                                                                                                      \
        if (nt >= 512) VAG_SEL2_NT(KM, 512, 4);                                                                               \
        else if (nt >= 256 && unr >= 8) VAG_SEL2_NT(KM, 256, 8);                                                              \
        else if (nt >= 256) VAG_SEL2_NT(KM, 256, 4);                                                                          \
        else if (unr >= 8) VAG_SEL2_NT(KM, 128, 8);                                                                           \
        else VAG_SEL2_NT(KM, 128, 4);                                                                                         \
    } while (0)
    if (K <= 4) VAG_SEL2(4);
    else if (K <= 8) VAG_SEL2(8);
    else if (K <= 12) VAG_SEL2(12);
    else VAG_SEL2(16);
#undef VAG_SEL2
#undef VAG_SEL2_NT
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int beam_advance(float* h_next, const float* h_cur, const int32_t* parents, int B, int K, int Kin, int H, int step,
                 int* done, int* fin_counter, int* steps_run, cudaStream_t st, volatile int32_t* host_progress, int nonce) {
    beam_advance_kernel<<<B * K, 128, 0, st>>>(h_next, h_cur, parents, B, K, Kin, H, step, done, fin_counter, steps_run);
    VAG_LAUNCH_CHECK();
    beam_done_kernel<<<1, 1, 0, st>>>(done, fin_counter, B * K, steps_run, host_progress, nonce);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int beam_advance_fused(float* h_next, const float* h_cur, const int32_t* parents, const int64_t* tokens, int B, int K, int Kin,
                       int H, int step, int* done, int* fin_counter, int* steps_run, SplitDst h_sd, SplitDst e_sd,
                       const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, int64_t V, cudaStream_t st,
                       volatile int32_t* host_progress, int nonce) {
    VAG_CUDA(launch_pdl(PDL_BEAM, beam_advance_fused_kernel, dim3((B * K + 3) / 4), dim3(128), 0, st, h_next, h_cur, parents, tokens, B, K, Kin, H,
                        step, done, (const int*)fin_counter, steps_run, h_sd, e_sd, t_hi, t_lo, ld_t, E, V, host_progress, nonce));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int beam_finalize(const int64_t* tok_hist, const int32_t* par_hist, const float* nll, const int* steps_run, int B, int K,
                  int L, int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, cudaStream_t st) {
    beam_finalize_kernel<<<B, 32, 0, st>>>(tok_hist, par_hist, nll, steps_run, B, K, L, hyp_out, hyp_len, beam_out);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int row_argmax(const float* logits, int64_t ld, int rows, int64_t V, int64_t* out, int64_t out_stride, int64_t* next_in,
               cudaStream_t st) {
    row_argmax_kernel<<<rows, 256, 0, st>>>(logits, ld, V, out, out_stride, next_in);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

}  // namespace vag

using namespace vag;

extern "C" int vag_beam_select_f32(const float* logp, int64_t ld_logp, const int64_t* prev_tokens, float* nll,
                                   int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step,
                                   int avoid_double, vag_stream_t stream) {
    VAG_REQUIRE(logp && nll && tokens_out && parents_out, "vag_beam_select_f32: null pointer");
    VAG_REQUIRE(step == 0 || prev_tokens, "vag_beam_select_f32: prev_tokens required for step > 0");
    VAG_REQUIRE(B > 0 && K > 0 && V > 0 && ld_logp >= V && step >= 0, "vag_beam_select_f32: bad shape");
    VAG_REQUIRE(K <= V, "vag_beam_select_f32: beam %d larger than the vocabulary %lld", K, (long long)V);
    return beam_select(logp, ld_logp, nullptr, prev_tokens, nll, tokens_out, parents_out, B, K, V, step, avoid_double,
                       nullptr, nullptr, (cudaStream_t)stream);
}

extern "C" int vag_row_argmax_f32(const float* logits, int64_t ld, int rows, int64_t V, int64_t* out, vag_stream_t stream) {
    VAG_REQUIRE(logits && out, "vag_row_argmax_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && V > 0 && ld >= V, "vag_row_argmax_f32: bad shape");
    if (rows == 0) return VAG_OK;
    return row_argmax(logits, ld, rows, V, out, 1, nullptr, (cudaStream_t)stream);
}
