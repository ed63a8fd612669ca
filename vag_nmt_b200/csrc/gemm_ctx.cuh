// Contraction context of a composite operator: routes every y = act(x·Wᵀ + b) to the tcgen05 split-precision
// kernel when the shape allows it (else to the FP32 FFMA kernel) and owns the operand-split bookkeeping:
//   * weights are split into hi/lo planes ONCE per composite call and reused by every step that follows;
//   * activation splits live in a per-step bump region and are shared by the contractions that read the same
//     buffer inside one step (h1 feeds attn_h and gru_2, the context feeds context2hid and W2, …);
//   * several (x_i, W_i) pairs can be laid side by side along K so that  Σ_i x_i·W_iᵀ  runs as ONE contraction
//     (the three-matrix read-out  tanh(W1 h2 + W3 e + W2 c)  of NMT_Decoder.py:137).
#pragma once
#include "common.cuh"
#include "split.cuh"

namespace vag {

int linear_simt(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                int rows, int K, int N, int flags, cudaStream_t st);
int tc_elem_bytes();
int tc_split(const float* x, int64_t ldx, int rows, int K, void* hi, void* lo, int64_t ld_out, int64_t col_off, cudaStream_t st);
int tc_split_pair(const float* x, int64_t ldx, bool xt, int M, void* xh, void* xl, const float* w, int64_t ldw, bool wt, int N,
                  void* wh, void* wl, int K, int Kp, cudaStream_t st);
int tc_gemm(float* y, int64_t ldy, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
            const float* bias, int rows, int K, int N, int flags, cudaStream_t st, float4* summ, int* summ_tile_w);
bool tc_enabled();
int gemm_mode();
int tc_gemm_top2(float4* summ, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
                 const float* bias, int rows, int K, int N, cudaStream_t st);
int tc_gemm_split_out(SplitDst out, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
                      const float* bias, int rows, int K, int N, int flags, cudaStream_t st);
bool tc_call_supported(const float* y, int64_t ldy, int flags);
// While alive, every tcgen05 contraction launched from this thread receives `done` (device flag, may be null): its CTAs return
// right after their prologue when *done != 0 (beam search: every hypothesis has ended, the remaining steps are dead).
struct TcDoneScope {
    const int* prev;
    explicit TcDoneScope(const int* done);
    ~TcDoneScope();
};
int bias_sum3(float* out, const float* a, const float* b, const float* c, int n, cudaStream_t st);

// GRU cell fused into its contractions (gru_pair.cuh): h' = GRU(x, h_prev) from operand planes of x (phase X; absent when the
// input pre-activations come from a per-token table g1) and of h_prev (phase H) against PERMUTED weight planes
// (tc_gru_prepare_weight); writes h' as fp32 and as operand planes.
struct GruCall {
    const void *xh = nullptr, *xl = nullptr;   // x planes [rows, Kx], pitch ldx
    int64_t ldx = 0;
    int Kx = 0;
    const void *hh = nullptr, *hl = nullptr;   // h_prev planes [rows, Kh], pitch ldh
    int64_t ldh = 0;
    int Kh = 0;
    const void *wih_h = nullptr, *wih_l = nullptr, *whh_h = nullptr, *whh_l = nullptr;   // permuted weight planes [3H, K]
    const float* bias4 = nullptr;              // [4][H] b_r, b_z, b_in, b_hn (tc_gru_prepare_bias)
    const float* g1 = nullptr;                 // [V, 3H] per-token input pre-activations (Kx == 0)
    const int64_t* tokens = nullptr;           // nullptr with g1: table row = row index
    int64_t V = 0;
    const float* h_prev = nullptr;             // [rows, H] fp32
    float* h_out = nullptr;                    // [rows, H] fp32 (must not alias h_prev)
    SplitDst out;                              // operand planes of h'
    int rows = 0, H = 0;
    float* y2 = nullptr;                       // optional second fp32 copy of h' (the encoder's context slab), row pitch ld_y2
    int64_t ld_y2 = 0;
};
bool tc_gru_supported(int rows, int H, int Kx, int Kh);
int tc_gru(const GruCall& c, cudaStream_t st);
int tc_gru_prepare_weight(void* hi, void* lo, const float* w, int H, int K, bool hidden, float* scratch, cudaStream_t st);
int tc_gru_prepare_bias(float* out4h, const float* b_ih, const float* b_hh, int H, bool with_ih, cudaStream_t st);

struct GemmCtx {
    struct Ent {
        const void* src;
        int rows, K;
        void *hi, *lo;
        int64_t ld;
    };
    cudaStream_t st = nullptr;
    char* wbase = nullptr;
    size_t wcap = 0, woff = 0;
    char* abase = nullptr;
    size_t acap = 0, aoff = 0;
    Ent wc[32];
    int nw = 0;
    Ent ac[16];
    int na = 0;
    bool tc = false;

    GemmCtx() {}
    GemmCtx(cudaStream_t s, void* wregion, size_t wbytes, void* aregion, size_t abytes)
        : st(s), wbase((char*)wregion), wcap(wbytes), abase((char*)aregion), acap(abytes) {
        tc = tc_enabled() && wregion && aregion;
    }
    // bytes of split storage for a [rows, K] operand (both planes, 4 B/element upper bound, 256 B aligned)
    static size_t split_bytes(int64_t rows, int64_t K) { return align_up((size_t)rows * K * 4, 256) * 2; }
    void new_step() { aoff = 0; na = 0; }

    static bool shape_ok(int rows, int K, int N) { return rows >= 64 && N >= 64 && K >= 32 && (K % 8 == 0); }
    static bool ptr_ok(const float* p, int64_t ld) { return (((uintptr_t)p & 15) == 0) && (ld % 4 == 0); }

    char* take(char* base, size_t cap, size_t& off, size_t bytes) {
        const size_t start = align_up(off, 256);
        if (!base || start + bytes > cap) return nullptr;
        off = start + bytes;
        return base + start;
    }
    // Find or create the split of a [rows, Ktot]-wide operand whose first segment is `src`.
    Ent* lookup(Ent* tab, int n, const void* src, int rows, int K) {
        for (int i = 0; i < n; ++i)
            if (tab[i].src == src && tab[i].rows == rows && tab[i].K == K) return &tab[i];
        return nullptr;
    }
    Ent* make(bool weight, const void* key, int rows, int Ktot) {
        Ent* tab = weight ? wc : ac;
        int& n = weight ? nw : na;
        if (n >= (weight ? 32 : 16)) return nullptr;
        const size_t plane = align_up((size_t)rows * Ktot * tc_elem_bytes(), 256);
        char* p = weight ? take(wbase, wcap, woff, 2 * plane) : take(abase, acap, aoff, 2 * plane);
        if (!p) return nullptr;
        tab[n] = Ent{key, rows, Ktot, p, p + plane, Ktot};
        return &tab[n++];
    }

    // Planes that already exist (vag_decoder_prepare_f32): later lookups of `w` hit them and no split is launched.
    void preset(const float* w, int N, int K, void* hi, void* lo) {
        if (nw < 32 && !lookup(wc, nw, w, N, K)) wc[nw++] = Ent{w, N, K, hi, lo, K};
    }
    void preset3(const float* w0, int N, int Kt, void* hi, void* lo, float* bsum) {
        if (nw + 1 < 32 && !lookup(wc, nw, w0, N, Kt)) {
            wc[nw++] = Ent{w0, N, Kt, hi, lo, Kt};
            wc[nw++] = Ent{(const void*)((uintptr_t)w0 + 1), N, Kt, bsum, nullptr, 0};
        }
    }
    // Split planes of a weight matrix, made once per composite call (nullptr: cache full / no region → FFMA path).
    int weight(Ent** out, const float* w, int64_t ldw, int N, int K) {
        Ent* we = lookup(wc, nw, w, N, K);
        if (!we && (we = make(true, w, N, K))) VAG_TRY(tc_split(w, ldw, N, K, we->hi, we->lo, K, 0, st));
        *out = we;
        return VAG_OK;
    }
    // Three weight matrices laid side by side along K plus the sum of their biases (the read-out of NMT_Decoder.py:137).
    int weight3(Ent** out, float** bsum_out, const float* const w[3], const int64_t ldw[3], const int K[3],
                const float* const bias[3], int N) {
        const int Kt = K[0] + K[1] + K[2];
        *out = nullptr;
        *bsum_out = nullptr;
        Ent* we = lookup(wc, nw, w[0], N, Kt);
        if (!we) {
            we = make(true, w[0], N, Kt);
            char* bp = we ? take(wbase, wcap, woff, align_up((size_t)N * 4, 256)) : nullptr;
            if (!we || !bp) return VAG_OK;
            int off = 0;
            for (int i = 0; i < 3; ++i) {
                VAG_TRY(tc_split(w[i], ldw[i], N, K[i], we->hi, we->lo, Kt, off, st));
                off += K[i];
            }
            VAG_TRY(bias_sum3((float*)bp, bias[0], bias[1], bias[2], N, st));
            // remember the bias vector right behind the entry (second cache slot keyed by the bias pointer)
            if (nw < 32) wc[nw++] = Ent{(const void*)((uintptr_t)w[0] + 1), N, Kt, bp, nullptr, 0};
        }
        Ent* be = lookup(wc, nw, (const void*)((uintptr_t)w[0] + 1), N, Kt);
        if (be) { *out = we; *bsum_out = (float*)be->hi; }
        return VAG_OK;
    }

    // summ / summ_tile_w: optional per-(row, column-tile) soft-max / arg-max summary written by the tensor-core
    // epilogue (see tc_gemm); *summ_tile_w stays 0 when the FP32 FFMA kernel ran and no summary exists.
    int linear(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
               int K, int N, int flags, float4* summ = nullptr, int* summ_tile_w = nullptr) {
        if (summ_tile_w) *summ_tile_w = 0;
        if (rows == 0 || N == 0) return VAG_OK;
        if (tc && shape_ok(rows, K, N) && ptr_ok(x, ldx) && ptr_ok(w, ldw) && tc_call_supported(y, ldy, flags)) {
            Ent* we = nullptr;
            VAG_TRY(weight(&we, w, ldw, N, K));
            Ent* xe = we ? lookup(ac, na, x, rows, K) : nullptr;
            if (we && !xe && (xe = make(false, x, rows, K))) VAG_TRY(tc_split(x, ldx, rows, K, xe->hi, xe->lo, K, 0, st));
            if (we && xe)
                return tc_gemm(y, ldy, xe->hi, xe->lo, xe->ld, we->hi, we->lo, we->ld, bias, rows, K, N, flags, st, summ, summ_tile_w);
        }
        return linear_simt(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, st);
    }

    // y = act( Σ_i x_i·W_iᵀ + (b_0 + b_1 + b_2) ), three segments concatenated along K.
    int linear3(float* y, int64_t ldy, const float* const x[3], const int64_t ldx[3], const int K[3], const float* const w[3],
                const int64_t ldw[3], const float* const bias[3], int rows, int N, int flags) {
        const int Kt = K[0] + K[1] + K[2];
        bool ok = tc && shape_ok(rows, Kt, N) && tc_call_supported(y, ldy, flags);
        for (int i = 0; i < 3; ++i) ok = ok && (K[i] % 8 == 0) && ptr_ok(x[i], ldx[i]) && ptr_ok(w[i], ldw[i]);
        if (ok) {
            Ent* we = nullptr;
            float* bsum = nullptr;
            VAG_TRY(weight3(&we, &bsum, w, ldw, K, bias, N));
            Ent* xe = (we && bsum) ? make(false, x[0], rows, Kt) : nullptr;
            if (xe) {
                int off = 0;
                for (int i = 0; i < 3; ++i) {
                    VAG_TRY(tc_split(x[i], ldx[i], rows, K[i], xe->hi, xe->lo, Kt, off, st));
                    off += K[i];
                }
                return tc_gemm(y, ldy, xe->hi, xe->lo, Kt, we->hi, we->lo, Kt, bsum, rows, Kt, N, flags, st, nullptr, nullptr);
            }
        }
        // FP32 FFMA path: three accumulating contractions, summed left to right like the reference
        VAG_TRY(linear_simt(y, ldy, x[0], ldx[0], w[0], ldw[0], bias[0], rows, K[0], N, flags & VAG_LIN_ACCUMULATE, st));
        VAG_TRY(linear_simt(y, ldy, x[1], ldx[1], w[1], ldw[1], bias[1], rows, K[1], N, VAG_LIN_ACCUMULATE, st));
        return linear_simt(y, ldy, x[2], ldx[2], w[2], ldw[2], bias[2], rows, K[2], N, VAG_LIN_ACCUMULATE | (flags & VAG_LIN_TANH), st);
    }
};

}  // namespace vag
