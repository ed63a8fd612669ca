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

// GRU flavour (linear_rows32_gru): the contraction produces one of the two pre-activation matrices of a GRU cell,
// pre[rows, 3H] = x · Wᵀ + bias (gate order r | z | n like torch.nn.GRU), and the epilogue finishes the cell —
//   r = σ(gi_r + gh_r), z = σ(gi_z + gh_z), n = tanh(gi_n + r · gh_n), h' = (1 − z) · n + z · h
// — with the other pre-activation matrix read from memory.  A CTA owns 4 hidden units = 12 columns (4 per gate), so the gate
// kernel that used to follow the contraction (one more node of ~4 µs in the captured step) disappears.
struct Rows32Gru {
    float* pre;              // [rows, 3H] the pre-activations this contraction produces (kept: back-propagation reads them)
    const float* bias;       // [3H]
    const float* other;      // [rows, 3H] the other pre-activation matrix
    int produces_gi;         // 1: pre = gi (input side), other = gh;  0: pre = gh, other = gi
    const float* h_prev;     // [rows, H]
    float* h_out;            // [rows, H]  (must not alias h_prev or x: other CTAs are still reading them)
    float* out2;             // optional second copy of h' (the encoder's context slot), row pitch ld_out2
    int64_t ld_out2;
    const int32_t* lengths;  // optional [rows]: rows with lengths[row] <= t are masked (h' = h, out2 untouched, pre = 0)
    int t;
    int H;
    Rows32Seg seg;           // x [rows, K] against w [3H, K] (k-fast)
};
// Backward of a GRU cell fused behind the contraction that finishes its incoming gradient:
//   g[row, u] = Σ_seg x_seg · W_seg (columns = hidden units, W read transposed)  (+ base[row, u])  (+ add2[row, u])
//   → dgi / dgh [rows, 3H] of the cell (gate order r | z | n) and dh_out[row, u] = g · z, the part of the gradient that flows
//   straight to the previous state.  Rows with lengths[row] <= t get zeros.  hprev_store (optional) keeps the previous state the
//   cell saw (the encoder reads it back from its context output).
struct Rows32GruBwd {
    Rows32Seg seg[2];
    int nseg;
    const float* base;       // optional, pitch ld_base
    int64_t ld_base;
    const float* add2;       // optional, pitch ld_add2
    int64_t ld_add2;
    const float* gi;         // [rows, 3H] saved pre-activations
    const float* gh;
    const float* h_prev;     // optional (null: 0), pitch ld_hprev
    int64_t ld_hprev;
    float* hprev_store;      // optional [rows, H]
    float* dgi;              // [rows, 3H]
    float* dgh;
    float* dh_out;           // [rows, H]
    const int32_t* lengths;  // optional
    int t;
    int H;
};
bool rows32_gru_bwd_ok(const Rows32GruBwd& p, int rows);
int linear_rows32_gru_bwd(const Rows32GruBwd* probs, int nprob, int rows, bool round_bf16, cudaStream_t st);
bool rows32_gru_ok(const Rows32Gru& p, int rows);
int linear_rows32_gru(const Rows32Gru* probs, int nprob, int rows, bool round_bf16, cudaStream_t st);

bool rows32_ok(const float* x, int64_t ldx, const float* w, int64_t ldw, int rows, int K, int N, bool wk);
bool rows32_problem_ok(const Rows32Problem& p, int rows, bool wk);
int linear_rows32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
                  int K, int N, int flags, bool wk, bool round_bf16, cudaStream_t st);
int linear_rows32_multi(const Rows32Problem* probs, int nprob, int rows, int flags, bool wk, bool round_bf16, cudaStream_t st);

}  // namespace vag
