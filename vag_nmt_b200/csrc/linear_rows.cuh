// ≤ 32-row contraction kernel (linear_rows.cu): problem descriptors for the multi-problem / multi-segment launches.
#pragma once
#include "common.cuh"

namespace vag {

// One K-segment of a contraction: x [rows, K] (pitch ldx) against W(c, k) = w[c·ldw + k] (wk) or w[k·ldw + c] (!wk).
struct Rows32Seg {
    const float* x;
    const float* w;
    int64_t ldx, ldw;
    int K;
};
// y[rows, N] (+)= Σ_seg x_seg · W_seg (+ bias) (tanh).  Up to two segments accumulate into one output (a K-concatenated
// contraction, e.g. dh1 = dgh2·W_hh2 + dq·W_attn_h); up to two problems with their own N / pitches share a launch
// (both GRU directions; the two contractions that read h1).
struct Rows32Problem {
    float* y;
    const float* bias;
    int64_t ldy;
    int N;
    int nseg;
    Rows32Seg seg[2];
};

bool rows32_ok(const float* x, int64_t ldx, const float* w, int64_t ldw, int rows, int K, int N, bool wk);
bool rows32_problem_ok(const Rows32Problem& p, int rows, bool wk);
int linear_rows32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
                  int K, int N, int flags, bool wk, bool round_bf16, cudaStream_t st);
int linear_rows32_multi(const Rows32Problem* probs, int nprob, int rows, int flags, bool wk, bool round_bf16, cudaStream_t st);

}  // namespace vag
