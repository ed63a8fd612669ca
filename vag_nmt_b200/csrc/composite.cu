// Host-side composition of the kernels into the reference's operators: encoder, visual-attention pooling,
// decoder step, beam / greedy decoding.  Everything is enqueued on the caller's stream; no allocation, no
// host synchronisation.
#include "common.cuh"
#include "gemm_ctx.cuh"
#include <algorithm>
#include <stdlib.h>
#include <vector>

namespace vag {

static inline size_t max_sz(size_t a, size_t b) { return a > b ? a : b; }
int row_lse(float* lse_out, const float* logits, int64_t ld, int rows, int64_t V, cudaStream_t st);
int beam_select(const float* logits, int64_t ld, const float* lse, const int64_t* prev_tokens, float* nll,
                int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step, int avoid_double,
                const int* done, int* fin_counter, cudaStream_t st);
int beam_select_summary(const float4* summ, int tile_w, const float* logits, int64_t ld, const int64_t* prev_tokens, float* nll,
                        int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step, int avoid_double,
                        const int* done, int* fin_counter, cudaStream_t st);
int beam_advance(float* h_next, const float* h_cur, const int32_t* parents, int B, int K, int Kin, int H, int step,
                 int* done, int* fin_counter, int* steps_run, cudaStream_t st);
int beam_finalize(const int64_t* tok_hist, const int32_t* par_hist, const float* nll, const int* steps_run, int B, int K,
                  int L, int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, cudaStream_t st);
int row_argmax(const float* logits, int64_t ld, int rows, int64_t V, int64_t* out, int64_t out_stride, int64_t* next_in,
               cudaStream_t st);

// ------------------------------------------------------------------ encoder
// out[(t*B + b), :] = table[src[b, t], :]  (time-major rows so that a timestep is a contiguous row block),
// mask[b, t] = src[b, t] != 0   (Encoder.py:47,50)
__global__ void encoder_embed_kernel(float* __restrict__ out, float* __restrict__ mask, const float* __restrict__ table,
                                     const int64_t* __restrict__ src, int B, int T, int E, int64_t vocab) {
    const int row = blockIdx.x * blockDim.y + threadIdx.y;  // t*B + b
    if (row >= B * T) return;
    const int t = row / B, b = row % B;
    int64_t id = src[(int64_t)b * T + t];
    if (threadIdx.x == 0) mask[(int64_t)b * T + t] = id != 0 ? 1.0f : 0.0f;
    if (id < 0 || id >= vocab) id = 0;
    const float* s = table + id * E;
    float* d = out + (int64_t)row * E;
    for (int c = threadIdx.x; c < E; c += blockDim.x) d[c] = s[c];
}

__global__ void fill_i64_kernel(int64_t* p, int64_t v, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

struct EncoderWs {
    float *x, *gi[2], *gh[2], *h[2];
    void *wreg, *areg;
    size_t wbytes, abytes;
};
template <typename A>
static void encoder_layout(A& a, int B, int T, int E, int H, EncoderWs* ws) {
    float* x = (float*)a.template take<float>((size_t)T * B * E);
    float* gi0 = (float*)a.template take<float>((size_t)T * B * 3 * H);
    float* gi1 = (float*)a.template take<float>((size_t)T * B * 3 * H);
    float* gh0 = (float*)a.template take<float>((size_t)B * 3 * H);
    float* gh1 = (float*)a.template take<float>((size_t)B * 3 * H);
    float* h0 = (float*)a.template take<float>((size_t)B * H);
    float* h1 = (float*)a.template take<float>((size_t)B * H);
    const size_t wb = 2 * (GemmCtx::split_bytes(3 * H, E) + GemmCtx::split_bytes(3 * H, H)) + 4096;
    const size_t ab = max_sz(GemmCtx::split_bytes((int64_t)T * B, E), GemmCtx::split_bytes(B, H)) + 4096;
    void* wr = a.template take<char>(wb);
    void* ar = a.template take<char>(ab);
    if (ws) {
        ws->x = x; ws->gi[0] = gi0; ws->gi[1] = gi1; ws->gh[0] = gh0; ws->gh[1] = gh1; ws->h[0] = h0; ws->h[1] = h1;
        ws->wreg = wr; ws->wbytes = wb; ws->areg = ar; ws->abytes = ab;
    }
}
struct SizerAdapter {
    ArenaSizer s;
    template <typename T> void* take(size_t n) { s.take<T>(n); return nullptr; }
};
struct ArenaAdapter {
    Arena a;
    ArenaAdapter(void* p, size_t n) : a(p, n) {}
    template <typename T> void* take(size_t n) { return a.take<T>(n); }
};

}  // namespace vag

using namespace vag;

extern "C" size_t vag_encoder_workspace_bytes(int B, int T, int E, int H) {
    SizerAdapter s;
    encoder_layout(s, B, T, E, H, nullptr);
    return s.s.total();
}

extern "C" int vag_encoder_fwd_f32(const vag_encoder_weights* w, const int64_t* src, const int32_t* lengths_host, int B, int T,
                                   float* ctx_out, float* mask_out, void* workspace, size_t workspace_bytes,
                                   vag_stream_t stream) {
    VAG_REQUIRE(w && src && lengths_host && ctx_out && mask_out, "vag_encoder_fwd_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0 && w->E > 0 && w->H > 0, "vag_encoder_fwd_f32: bad shape");
    for (int b = 0; b < B; ++b) {
        VAG_REQUIRE(lengths_host[b] >= 1 && lengths_host[b] <= T, "vag_encoder_fwd_f32: length[%d]=%d outside [1,%d]", b, lengths_host[b], T);
        VAG_REQUIRE(b == 0 || lengths_host[b] <= lengths_host[b - 1],
                    "vag_encoder_fwd_f32: lengths must be sorted in decreasing order (pack_padded_sequence, Encoder.py:55)");
    }
    VAG_REQUIRE(lengths_host[0] == T, "vag_encoder_fwd_f32: longest sentence (%d) must span the padded width (%d)", lengths_host[0], T);
    const int E = w->E, H = w->H;
    cudaStream_t st = (cudaStream_t)stream;
    ArenaAdapter ar(workspace, workspace_bytes);
    EncoderWs ws;
    encoder_layout(ar, B, T, E, H, &ws);
    if (ar.a.overflow) {
        set_error("vag_encoder_fwd_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    {
        dim3 block(64, 4);
        encoder_embed_kernel<<<ceil_div(B * T, 4), block, 0, st>>>(ws.x, mask_out, w->emb, src, B, T, E, w->vocab);
        VAG_LAUNCH_CHECK();
    }
    VAG_CUDA(cudaMemsetAsync(ctx_out, 0, (size_t)B * T * 2 * H * sizeof(float), st));
    // active-row count per timestep (rows sorted by length ⇒ a prefix)
    std::vector<int> n_act(T);
    for (int t = 0; t < T; ++t) {
        int n = 0;
        while (n < B && lengths_host[n] > t) ++n;
        n_act[t] = n;
    }
    GemmCtx gemm(st, ws.wreg, ws.wbytes, ws.areg, ws.abytes);
    for (int d = 0; d < 2; ++d) {
        VAG_TRY(gemm.linear(ws.gi[d], 3 * H, ws.x, E, w->w_ih[d], E, w->b_ih[d], T * B, E, 3 * H, 0));
        VAG_CUDA(cudaMemsetAsync(ws.h[d], 0, (size_t)B * H * sizeof(float), st));
    }
    // the two directions are independent chains; interleave them so neighbouring launches can overlap their tails
    for (int s = 0; s < T; ++s) {
        for (int d = 0; d < 2; ++d) {
            const int t = d == 0 ? s : T - 1 - s;
            const int n = n_act[t];
            if (n == 0) continue;
            gemm.new_step();  // h changed: its split is stale
            VAG_TRY(gemm.linear(ws.gh[d], 3 * H, ws.h[d], H, w->w_hh[d], H, w->b_hh[d], n, H, 3 * H, 0));
            VAG_TRY(vag_gru_gates_f32(ws.h[d], H, ctx_out + (int64_t)t * 2 * H + d * H, (int64_t)T * 2 * H,
                                      ws.gi[d] + (int64_t)t * B * 3 * H, 3 * H, ws.gh[d], 3 * H, ws.h[d], H, n, H, stream));
        }
    }
    return VAG_OK;
}

// ------------------------------------------------------------------ visual-attention pooling
static size_t vse_wbytes(int I, int C, int S) {
    return GemmCtx::split_bytes(S, I) + GemmCtx::split_bytes(C, S) + GemmCtx::split_bytes(C, C) + GemmCtx::split_bytes(S, C) + 4096;
}
static size_t vse_abytes(int B, int T, int I, int C, int S) {
    return GemmCtx::split_bytes(B, I) + GemmCtx::split_bytes(B, S) + GemmCtx::split_bytes((int64_t)B * T, C) + GemmCtx::split_bytes(B, C) + 4096;
}
extern "C" size_t vag_vse_workspace_bytes(int B, int T, int I, int C, int S) {
    ArenaSizer s;
    s.take<float>((size_t)B * T * C);
    s.take<float>((size_t)B * C);
    s.take<char>(vse_wbytes(I, C, S));
    s.take<char>(vse_abytes(B, T, I, C, S));
    return s.total();
}

extern "C" int vag_vse_pool_fwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                                    float* im_emb, float* txt_emb, float* ctx_vec, float* beta, void* workspace,
                                    size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(w && im && ctx && im_emb && txt_emb && ctx_vec, "vag_vse_pool_fwd_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0, "vag_vse_pool_fwd_f32: bad shape");
    VAG_REQUIRE(w->method == VAG_ATTN_DOT || (w->method == VAG_ATTN_MLP && w->mlp_w), "vag_vse_pool_fwd_f32: bad attention method");
    const int I = w->I, C = w->C, S = w->S;
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    float* pk = ar.take<float>((size_t)B * T * C);
    float* iq = ar.take<float>((size_t)B * C);
    const size_t wb = vse_wbytes(I, C, S), ab = vse_abytes(B, T, I, C, S);
    void* wr = ar.take<char>(wb);
    void* areg = ar.take<char>(ab);
    if (ar.overflow) {
        set_error("vag_vse_pool_fwd_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    GemmCtx gemm(st, wr, wb, areg, ab);
    const int act = w->activation ? VAG_LIN_TANH : 0;
    VAG_TRY(gemm.linear(im_emb, S, im, I, w->im_w, I, w->im_b, B, I, S, act));            // VSE_Imagine_Enc.py:123-127
    VAG_TRY(vag_l2norm_rows_f32(im_emb, S, B, S, stream));                                        // :132
    VAG_TRY(gemm.linear(iq, C, im_emb, S, w->emb2ctx_w, S, nullptr, B, S, C, 0));         // :58
    VAG_TRY(gemm.linear(pk, C, ctx, C, w->ctx2ctx_w, C, nullptr, B * T, C, C, 0));        // :57
    VAG_TRY(vag_attention_f32(ctx_vec, C, beta, iq, C, pk, ctx, w->mlp_w, mask, B, 1, T, C, w->method, stream));  // :135-137
    VAG_TRY(gemm.linear(txt_emb, S, ctx_vec, C, w->txt_w, C, w->txt_b, B, C, S, act));    // :138-140
    VAG_TRY(vag_l2norm_rows_f32(txt_emb, S, B, S, stream));                                       // :145
    return VAG_OK;
}

// ------------------------------------------------------------------ decoder
extern "C" size_t vag_attn_keys_workspace_bytes(int B, int T, int C) {
    return GemmCtx::split_bytes(C, C) + GemmCtx::split_bytes((int64_t)B * T, C) + 8192;
}

extern "C" int vag_attn_keys_f32(const vag_decoder_weights* w, const float* ctx, int B, int T, float* keys, void* workspace,
                                 size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(w && ctx && keys && B > 0 && T > 0, "vag_attn_keys_f32: bad argument");
    const size_t wb = GemmCtx::split_bytes(w->C, w->C) + 4096;
    Arena ar(workspace, workspace_bytes);
    void* wr = ar.take<char>(wb);
    const size_t ab = ar.overflow || workspace_bytes < ar.off + 512 ? 0 : workspace_bytes - align_up(ar.off, 256) - 256;
    void* areg = ab ? ar.take<char>(ab) : nullptr;
    GemmCtx gemm((cudaStream_t)stream, ar.overflow ? nullptr : wr, wb, ar.overflow ? nullptr : areg, ab);
    return gemm.linear(keys, w->C, ctx, w->C, w->attn_e_w, w->C, nullptr, B * T, w->C, w->C, 0);
}

extern "C" size_t vag_decoder_init_workspace_bytes(int B, int C, int H) {
    return align_up((size_t)B * C * 4, 256) + GemmCtx::split_bytes(H, C) + GemmCtx::split_bytes(B, C) + 16384;
}

extern "C" int vag_decoder_init_f32(const vag_decoder_weights* w, const float* ctx_vec, const float* ctx, const float* mask,
                                    float split, int B, int T, float* h0, void* workspace, size_t workspace_bytes,
                                    vag_stream_t stream) {
    VAG_REQUIRE(w && ctx && mask && h0 && B > 0 && T > 0, "vag_decoder_init_f32: bad argument");
    VAG_REQUIRE(w->ini_w && w->ini_b, "vag_decoder_init_f32: decoderini weights missing");
    Arena ar(workspace, workspace_bytes);
    float* z = ar.take<float>((size_t)B * w->C);
    if (ar.overflow) {
        set_error("vag_decoder_init_f32: workspace %zu B too small (need B*C floats)", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    VAG_TRY(vag_init_mix_f32(z, ctx_vec, ctx, mask, split, B, T, w->C, stream));
    const size_t wb = GemmCtx::split_bytes(w->H, w->C) + 4096, ab = GemmCtx::split_bytes(B, w->C) + 4096;
    void* wr = ar.take<char>(wb);   // optional: without them the FP32 FFMA kernel runs
    void* areg = ar.take<char>(ab);
    GemmCtx gemm((cudaStream_t)stream, ar.overflow ? nullptr : wr, wb, ar.overflow ? nullptr : areg, ab);
    return gemm.linear(h0, w->H, z, w->C, w->ini_w, w->C, w->ini_b, B, w->C, w->H, VAG_LIN_TANH);
}

namespace vag {
struct StepWs {
    float *e, *gi, *gh, *h1, *q, *c, *x2, *t;
    void *wreg, *areg;
    size_t wbytes, abytes;
};
static size_t step_wbytes(int E, int H, int C, int64_t V) {
    return GemmCtx::split_bytes(3 * H, E) + 3 * GemmCtx::split_bytes(3 * H, H) + GemmCtx::split_bytes(C, H) +
           GemmCtx::split_bytes(H, C) + GemmCtx::split_bytes(E, H + E + C) + GemmCtx::split_bytes(V, E) + 16384;
}
static size_t step_abytes(int rows, int E, int H, int C) {
    return 2 * GemmCtx::split_bytes(rows, E) + 4 * GemmCtx::split_bytes(rows, H) + GemmCtx::split_bytes(rows, C) +
           GemmCtx::split_bytes(rows, H + E + C) + 16384;
}
template <typename A>
static void step_layout(A& a, int rows, int E, int H, int C, int64_t V, StepWs* ws) {
    float* e = (float*)a.template take<float>((size_t)rows * E);
    float* gi = (float*)a.template take<float>((size_t)rows * 3 * H);
    float* gh = (float*)a.template take<float>((size_t)rows * 3 * H);
    float* h1 = (float*)a.template take<float>((size_t)rows * H);
    float* q = (float*)a.template take<float>((size_t)rows * C);
    float* c = (float*)a.template take<float>((size_t)rows * C);
    float* x2 = (float*)a.template take<float>((size_t)rows * H);
    float* t = (float*)a.template take<float>((size_t)rows * E);
    const size_t wb = step_wbytes(E, H, C, V), ab = step_abytes(rows, E, H, C);
    void* wr = a.template take<char>(wb);
    void* ar = a.template take<char>(ab);
    if (ws) {
        ws->e = e; ws->gi = gi; ws->gh = gh; ws->h1 = h1; ws->q = q; ws->c = c; ws->x2 = x2; ws->t = t;
        ws->wreg = wr; ws->wbytes = wb; ws->areg = ar; ws->abytes = ab;
    }
}

// One conditional-GRU step up to (and including) the vocabulary logits.  NMT_Decoder.py:109-143.
static int decoder_step_core(GemmCtx& gemm, const vag_decoder_weights* w, const StepWs& ws, const int64_t* tokens, const float* h_prev,
                             const float* keys, const float* ctx, const float* mask, int rows, int rows_per_sent, int T,
                             float* h_out, float* logits, int64_t ld_logits, float* alpha_out, cudaStream_t st,
                             float4* summ = nullptr, int* summ_tile_w = nullptr) {
    const int E = w->E, H = w->H, C = w->C;
    const int64_t V = w->V;
    vag_stream_t vs = (vag_stream_t)st;
    gemm.new_step();
    VAG_TRY(vag_embed_rows_f32(ws.e, E, w->emb, E, tokens, rows, V, vs));                                              // :118
    VAG_TRY(gemm.linear(ws.gi, 3 * H, ws.e, E, w->gru1_w_ih, E, w->gru1_b_ih, rows, E, 3 * H, 0));              // :121
    VAG_TRY(gemm.linear(ws.gh, 3 * H, h_prev, H, w->gru1_w_hh, H, w->gru1_b_hh, rows, H, 3 * H, 0));
    VAG_TRY(vag_gru_gates_f32(ws.h1, H, nullptr, 0, ws.gi, 3 * H, ws.gh, 3 * H, h_prev, H, rows, H, vs));
    VAG_TRY(gemm.linear(ws.q, C, ws.h1, H, w->attn_h_w, H, nullptr, rows, H, C, 0));                            // :47
    VAG_TRY(vag_attention_f32(ws.c, C, alpha_out, ws.q, C, keys, ctx, w->attn_v, mask, rows, rows_per_sent, T, C,
                              VAG_ATTN_MLP, vs));                                                                       // :124-126
    VAG_TRY(gemm.linear(ws.x2, H, ws.c, C, w->c2h_w, C, nullptr, rows, C, H, 0));                               // :127
    VAG_TRY(gemm.linear(ws.gi, 3 * H, ws.x2, H, w->gru2_w_ih, H, w->gru2_b_ih, rows, H, 3 * H, 0));             // :129
    VAG_TRY(gemm.linear(ws.gh, 3 * H, ws.h1, H, w->gru2_w_hh, H, w->gru2_b_hh, rows, H, 3 * H, 0));
    VAG_TRY(vag_gru_gates_f32(h_out, H, nullptr, 0, ws.gi, 3 * H, ws.gh, 3 * H, ws.h1, H, rows, H, vs));
    // t = tanh((W1 h2 + b1) + (W3 e + b3) + (W2 c + b2)), summed left to right like :137
    {
        const float* const xs[3] = {h_out, ws.e, ws.c};
        const int64_t lds[3] = {H, E, C};
        const int Ks[3] = {H, E, C};
        const float* const wts[3] = {w->w1_w, w->w3_w, w->w2_w};
        const float* const bs[3] = {w->w1_b, w->w3_b, w->w2_b};
        VAG_TRY(gemm.linear3(ws.t, E, xs, lds, Ks, wts, lds, bs, rows, E, VAG_LIN_TANH));
    }
    if (logits) VAG_TRY(gemm.linear(logits, ld_logits, ws.t, E, w->out_w, E, w->out_b, rows, E, (int)V, 0, summ, summ_tile_w));            // :143
    return VAG_OK;
}

// ------------------------------------------------------------------ fused decode step (tensor-core path only)
// Same arithmetic as decoder_step_core, but no activation ever takes the detour "fp32 in HBM → split kernel → planes":
// every producer writes the operand planes of what it produced (split.cuh), and activations that only feed
// contractions (embedding row, context, context2hid output, read-out) exist ONLY as planes.  Nine split launches, the
// embedding gather and four fp32 round trips per step disappear.  The read-out input [h2 | e | c] is one plane pair of
// pitch H+E+C; the embedding and the context are column windows of it (TMA takes any 16-byte aligned pitch).
int gru_gates_split(float* h_out, int64_t ld_ho, const float* gi, int64_t ld_gi, const float* gh, int64_t ld_gh,
                    const float* h_prev, int64_t ld_hp, int rows, int H, SplitDst sd, cudaStream_t st,
                    const int64_t* gi_rows = nullptr, int64_t gi_n_rows = 0);
int embed_split_rows(SplitDst dst, const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, const int64_t* tokens, int rows,
                     int64_t V, cudaStream_t st);
int attention_mlp_split(SplitDst sd, const float* q, int64_t ld_q, const float* keys, const float* ctx, const float* v,
                        const float* mask, int rows, int rows_per_sent, int T, int C, cudaStream_t st);
int beam_select_top2(const float4* summ, int tile_w, SplitDst t, const uint16_t* w_hi, const uint16_t* w_lo, int64_t ld_w,
                     const float* bias, int E, const int64_t* prev_tokens, float* nll, int64_t* tokens_out, int32_t* parents_out,
                     int B, int K, int64_t V, int step, int avoid_double, const int* done, int* fin_counter, cudaStream_t st);
int beam_advance_fused(float* h_next, const float* h_cur, const int32_t* parents, const int64_t* tokens, int B, int K, int Kin,
                       int H, int step, int* done, int* fin_counter, int* steps_run, SplitDst h_sd, SplitDst e_sd,
                       const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, int64_t V, cudaStream_t st);

struct FusedStep {
    bool ok = false;
    SplitDst hprev, h1, x2, t, cat;                 // cat = [h2 | e | c], pitch H + E + C
    GemmCtx::Ent *w_g1i = nullptr, *w_g1h = nullptr, *w_ah = nullptr, *w_c2h = nullptr, *w_g2i = nullptr, *w_g2h = nullptr,
                 *w_ro = nullptr, *w_out = nullptr, *w_emb = nullptr;
    float* b_ro = nullptr;
    float* g1 = nullptr;   // [V, 3H] table  Emb·W_ihᵀ + b_ih  of gru_1: its input pre-activations depend on the token only
    SplitDst cat_e(int H) const { SplitDst d = cat; d.hi += H; if (d.lo) d.lo += H; return d; }
    SplitDst cat_c(int H, int E) const { SplitDst d = cat; d.hi += H + E; if (d.lo) d.lo += H + E; return d; }
};

static int fused_setup(GemmCtx& gemm, const vag_decoder_weights* w, const StepWs& ws, int n_rows, int rows_per_sent, float* g1,
                       FusedStep* f) {
    const int E = w->E, H = w->H, C = w->C, Kt = H + E + C;
    const int64_t V = w->V;
    const int mode = gemm_mode();
    f->ok = false;
    if (!gemm.tc || (mode != 1 && mode != 2) || n_rows <= 128 || rows_per_sent > 16 || V < 64) return VAG_OK;
    if ((E % 8) || (H % 8) || (C % 8)) return VAG_OK;
    const float* ws_[] = {w->gru1_w_ih, w->gru1_w_hh, w->attn_h_w, w->c2h_w, w->gru2_w_ih, w->gru2_w_hh, w->w1_w, w->w2_w, w->w3_w,
                          w->out_w, w->emb};
    for (const float* p : ws_)
        if (!p || ((uintptr_t)p & 15)) return VAG_OK;
    VAG_TRY(gemm.weight(&f->w_g1i, w->gru1_w_ih, E, 3 * H, E));
    VAG_TRY(gemm.weight(&f->w_g1h, w->gru1_w_hh, H, 3 * H, H));
    VAG_TRY(gemm.weight(&f->w_ah, w->attn_h_w, H, C, H));
    VAG_TRY(gemm.weight(&f->w_c2h, w->c2h_w, C, H, C));
    VAG_TRY(gemm.weight(&f->w_g2i, w->gru2_w_ih, H, 3 * H, H));
    VAG_TRY(gemm.weight(&f->w_g2h, w->gru2_w_hh, H, 3 * H, H));
    VAG_TRY(gemm.weight(&f->w_out, w->out_w, E, (int)V, E));
    {
        const float* const wts[3] = {w->w1_w, w->w3_w, w->w2_w};
        const float* const bs[3] = {w->w1_b, w->w3_b, w->w2_b};
        const int64_t lds[3] = {H, E, C};
        const int Ks[3] = {H, E, C};
        VAG_TRY(gemm.weight3(&f->w_ro, &f->b_ro, wts, lds, Ks, bs, E));
    }
    if (w->emb == w->out_w) f->w_emb = f->w_out;   // tied: the gather reads the projection's planes
    else VAG_TRY(gemm.weight(&f->w_emb, w->emb, E, (int)V, E));
    if (!f->w_g1i || !f->w_g1h || !f->w_ah || !f->w_c2h || !f->w_g2i || !f->w_g2h || !f->w_out || !f->w_ro || !f->b_ro || !f->w_emb)
        return VAG_OK;
    // activation planes, carved once from the step's activation region (the plain path's per-step splits are not used)
    Arena ar(ws.areg, ws.abytes);
    auto planes = [&](SplitDst* d, int K) {
        d->hi = (uint16_t*)ar.take<uint16_t>((size_t)n_rows * K);
        d->lo = (uint16_t*)ar.take<uint16_t>((size_t)n_rows * K);
        d->ld = K;
        d->mode = mode;
    };
    planes(&f->hprev, H);
    planes(&f->h1, H);
    planes(&f->x2, H);
    planes(&f->t, E);
    planes(&f->cat, Kt);
    if (ar.overflow) return VAG_OK;
    // gru_1's input contraction once per call for EVERY token instead of once per step for every row: the same kernel on
    // the same operand rows, so each table row is bit-identical to what the per-step contraction produced
    if (g1 && V > 128) {
        VAG_TRY(tc_gemm(g1, 3 * H, f->w_emb->hi, f->w_emb->lo, f->w_emb->ld, f->w_g1i->hi, f->w_g1i->lo, f->w_g1i->ld, w->gru1_b_ih,
                        (int)V, E, 3 * H, 0, gemm.st, nullptr, nullptr));
        f->g1 = g1;
    }
    f->ok = true;
    return VAG_OK;
}

// One fused step: the embedding planes (cat_e) and the previous state's planes (hprev) are already in place.
static int decoder_step_fused(const FusedStep& f, const vag_decoder_weights* w, const StepWs& ws, const int64_t* tokens, const float* h_prev, const float* keys,
                              const float* ctx, const float* mask, int rows, int rows_per_sent, int T, float* h_out, float* logits,
                              int64_t ld_logits, cudaStream_t st, float4* summ, int* summ_tile_w) {
    const int E = w->E, H = w->H, C = w->C, Kt = H + E + C;
    const int64_t V = w->V;
    const SplitDst ce = f.cat_e(H), cc = f.cat_c(H, E);
    auto gemm = [&](float* y, int64_t ldy, const SplitDst& x, const GemmCtx::Ent* we, const float* bias, int K, int N, float4* sm,
                    int* tw) {
        return tc_gemm(y, ldy, x.hi, x.lo, x.ld, we->hi, we->lo, we->ld, bias, rows, K, N, 0, st, sm, tw);
    };
    if (!f.g1) VAG_TRY(gemm(ws.gi, 3 * H, ce, f.w_g1i, w->gru1_b_ih, E, 3 * H, nullptr, nullptr));           // NMT_Decoder.py:121
    VAG_TRY(gemm(ws.gh, 3 * H, f.hprev, f.w_g1h, w->gru1_b_hh, H, 3 * H, nullptr, nullptr));
    if (f.g1) VAG_TRY(gru_gates_split(ws.h1, H, f.g1, 3 * H, ws.gh, 3 * H, h_prev, H, rows, H, f.h1, st, tokens, V));
    else VAG_TRY(gru_gates_split(ws.h1, H, ws.gi, 3 * H, ws.gh, 3 * H, h_prev, H, rows, H, f.h1, st));
    VAG_TRY(gemm(ws.q, C, f.h1, f.w_ah, nullptr, H, C, nullptr, nullptr));                                    // :47
    VAG_TRY(attention_mlp_split(cc, ws.q, C, keys, ctx, w->attn_v, mask, rows, rows_per_sent, T, C, st));     // :124-126
    VAG_TRY(tc_gemm_split_out(f.x2, cc.hi, cc.lo, Kt, f.w_c2h->hi, f.w_c2h->lo, f.w_c2h->ld, nullptr, rows, C, H, 0, st));   // :127
    VAG_TRY(gemm(ws.gi, 3 * H, f.x2, f.w_g2i, w->gru2_b_ih, H, 3 * H, nullptr, nullptr));                     // :129
    VAG_TRY(gemm(ws.gh, 3 * H, f.h1, f.w_g2h, w->gru2_b_hh, H, 3 * H, nullptr, nullptr));
    VAG_TRY(gru_gates_split(h_out, H, ws.gi, 3 * H, ws.gh, 3 * H, ws.h1, H, rows, H, f.cat, st));
    VAG_TRY(tc_gemm_split_out(f.t, f.cat.hi, f.cat.lo, Kt, f.w_ro->hi, f.w_ro->lo, f.w_ro->ld, f.b_ro, rows, Kt, E, VAG_LIN_TANH, st));  // :137
    if (!logits) {   // beam loop: only the top-2 / Σexp summaries of every 128-column tile leave the projection
        if (summ_tile_w) *summ_tile_w = 128;
        return tc_gemm_top2(summ, f.t.hi, f.t.lo, f.t.ld, f.w_out->hi, f.w_out->lo, f.w_out->ld, w->out_b, rows, E, (int)V, st);
    }
    return gemm(logits, ld_logits, f.t, f.w_out, w->out_b, E, (int)V, summ, summ_tile_w);                     // :143
}
}  // namespace vag

extern "C" size_t vag_decoder_step_workspace_bytes(int rows, int E, int H, int C, int64_t V) {
    SizerAdapter s;
    step_layout(s, rows, E, H, C, V, nullptr);
    return s.s.total();
}

extern "C" int vag_decoder_step_f32(const vag_decoder_weights* w, const int64_t* tokens, const float* h_prev, const float* keys,
                                    const float* ctx, const float* mask, int rows, int rows_per_sent, int T, float* h_out,
                                    float* logits_or_logp, int want_logp, float* alpha_out, void* workspace,
                                    size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(w && tokens && h_prev && keys && ctx && h_out && logits_or_logp, "vag_decoder_step_f32: null pointer");
    VAG_REQUIRE(rows > 0 && rows_per_sent > 0 && rows % rows_per_sent == 0 && T > 0, "vag_decoder_step_f32: bad shape");
    VAG_REQUIRE(h_out != h_prev, "vag_decoder_step_f32: h_out must not alias h_prev");
    ArenaAdapter ar(workspace, workspace_bytes);
    StepWs ws;
    step_layout(ar, rows, w->E, w->H, w->C, w->V, &ws);
    if (ar.a.overflow) {
        set_error("vag_decoder_step_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    GemmCtx gemm((cudaStream_t)stream, ws.wreg, ws.wbytes, ws.areg, ws.abytes);
    VAG_TRY(decoder_step_core(gemm, w, ws, tokens, h_prev, keys, ctx, mask, rows, rows_per_sent, T, h_out, logits_or_logp, w->V,
                              alpha_out, (cudaStream_t)stream));
    if (want_logp) VAG_TRY(vag_log_softmax_f32(logits_or_logp, logits_or_logp, rows, (int)w->V, stream));
    return VAG_OK;
}

// ------------------------------------------------------------------ beam search
namespace vag {
struct BeamWs {
    StepWs step;
    float *logits, *lse, *h_a, *h_b, *nll, *g1;
    float4* summ;
    int64_t *tok_hist, *sos;
    int32_t* par_hist;
    int* flags;  // [0] done, [1] steps_run, [2..2+L) per-step EOS counters
};
template <typename A>
static void beam_layout(A& a, int B, int K, int L, int E, int H, int C, int64_t V, BeamWs* ws) {
    const int N = B * K;
    StepWs sw;
    step_layout(a, N, E, H, C, V, &sw);
    float* logits = (float*)a.template take<float>((size_t)N * ((V + 3) / 4 * 4));  // rows padded to 16 B
    float* lse = (float*)a.template take<float>((size_t)N);
    float* g1 = (float*)a.template take<float>((size_t)V * 3 * H);   // per-token input pre-activations of gru_1 (fused path)
    float4* summ = (float4*)a.template take<float4>((size_t)N * ((V + 31) / 32));   // per 32-column slice (top-2 kernel); the per-128 summaries need a quarter
    float* h_a = (float*)a.template take<float>((size_t)N * H);
    float* h_b = (float*)a.template take<float>((size_t)N * H);
    float* nll = (float*)a.template take<float>((size_t)N);
    int64_t* tok_hist = (int64_t*)a.template take<int64_t>((size_t)L * N);
    int64_t* sos = (int64_t*)a.template take<int64_t>((size_t)B);
    int32_t* par_hist = (int32_t*)a.template take<int32_t>((size_t)L * N);
    int* flags = (int*)a.template take<int>((size_t)L + 2);
    if (ws) {
        ws->step = sw; ws->logits = logits; ws->lse = lse; ws->g1 = g1; ws->summ = summ; ws->h_a = h_a; ws->h_b = h_b; ws->nll = nll;
        ws->tok_hist = tok_hist; ws->sos = sos; ws->par_hist = par_hist; ws->flags = flags;
    }
}
}  // namespace vag

extern "C" size_t vag_beam_decode_workspace_bytes(int B, int K, int T, int L, int E, int H, int C, int64_t V) {
    SizerAdapter s;
    beam_layout(s, B, K, L, E, H, C, V, nullptr);
    (void)T;
    return s.s.total();
}

extern "C" int vag_beam_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                                   const float* mask, int B, int K, int T, int L, int avoid_double, int64_t* hyp_out,
                                   int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out, void* workspace,
                                   size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(w && h0 && keys && ctx && mask && hyp_out && hyp_len, "vag_beam_decode_f32: null pointer");
    VAG_REQUIRE(B > 0 && K > 1 && T > 0 && L > 0, "vag_beam_decode_f32: bad shape B=%d K=%d T=%d L=%d", B, K, T, L);
    VAG_REQUIRE(K <= w->V, "vag_beam_decode_f32: beam larger than the vocabulary");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = w->E, H = w->H, C = w->C;
    const int64_t V = w->V;
    const int N = B * K;
    ArenaAdapter ar(workspace, workspace_bytes);
    BeamWs ws;
    beam_layout(ar, B, K, L, E, H, C, V, &ws);
    if (ar.a.overflow) {
        set_error("vag_beam_decode_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    int* done = ws.flags;
    int* steps_run = ws.flags + 1;
    int* fin = ws.flags + 2;
    VAG_CUDA(cudaMemsetAsync(ws.flags, 0, sizeof(int) * (size_t)(L + 2), st));
    fill_i64_kernel<<<ceil_div(B, 256), 256, 0, st>>>(ws.sos, 2 /*SOS*/, B);
    VAG_LAUNCH_CHECK();
    const int64_t ldl = (V + 3) / 4 * 4;
    GemmCtx gemm(st, ws.step.wreg, ws.step.wbytes, ws.step.areg, ws.step.abytes);
    FusedStep fused;
    VAG_TRY(fused_setup(gemm, w, ws.step, N, K, ws.g1, &fused));
    for (int di = 0; di < L; ++di) {
        const int rows = di == 0 ? B : N;
        const int rps = di == 0 ? 1 : K;
        const int64_t* tokens = di == 0 ? ws.sos : ws.tok_hist + (size_t)(di - 1) * N;
        const float* h_prev = di == 0 ? h0 : ws.h_a;
        int tile_w = 0;  // > 0 when the tensor-core projection also produced the per-tile soft-max / arg-max summaries
        if (fused.ok && rows > 128) {
            if (di == 0) {   // planes of the initial state and of the <sos> embedding; later steps get them from the reorder
                VAG_TRY(tc_split(h0, H, B, H, fused.hprev.hi, fused.hprev.lo, H, 0, st));
                VAG_TRY(embed_split_rows(fused.cat_e(H), (const uint16_t*)fused.w_emb->hi, (const uint16_t*)fused.w_emb->lo,
                                         fused.w_emb->ld, E, ws.sos, B, V, st));
            }
            const bool no_logits = V >= 512 && V < 0xFFFF && !getenv("VAG_KEEP_LOGITS");
            VAG_TRY(decoder_step_fused(fused, w, ws.step, tokens, h_prev, keys, ctx, mask, rows, rps, T, ws.h_b, no_logits ? nullptr : ws.logits,
                                       ldl, st, V >= 512 ? ws.summ : nullptr, &tile_w));
            if (no_logits) {
                VAG_TRY(beam_select_top2(ws.summ, 32, fused.t, (const uint16_t*)fused.w_out->hi, (const uint16_t*)fused.w_out->lo,
                                         fused.w_out->ld, w->out_b, E, di == 0 ? nullptr : tokens, ws.nll, ws.tok_hist + (size_t)di * N,
                                         ws.par_hist + (size_t)di * N, B, K, V, di, avoid_double, done, fin + di, st));
                tile_w = -1;   // selection done
            }
        } else
        VAG_TRY(decoder_step_core(gemm, w, ws.step, tokens, h_prev, keys, ctx, mask, rows, rps, T, ws.h_b, ws.logits, ldl, nullptr, st,
                                  V >= 512 ? ws.summ : nullptr, &tile_w));
        if (tile_w < 0) {
        } else if (tile_w > 0) {
            VAG_TRY(beam_select_summary(ws.summ, tile_w, ws.logits, ldl, di == 0 ? nullptr : tokens, ws.nll,
                                        ws.tok_hist + (size_t)di * N, ws.par_hist + (size_t)di * N, B, K, V, di, avoid_double, done,
                                        fin + di, st));
        } else {
            const bool fused_lse = V >= 512;  // the fast scan kernel folds the log-sum-exp into its single pass
            if (!fused_lse) VAG_TRY(row_lse(ws.lse, ws.logits, ldl, rows, V, st));
            VAG_TRY(beam_select(ws.logits, ldl, ws.lse, di == 0 ? nullptr : tokens, ws.nll, ws.tok_hist + (size_t)di * N,
                                ws.par_hist + (size_t)di * N, B, K, V, di, avoid_double, done, fin + di, st));
        }
        if (fused.ok)
            VAG_TRY(beam_advance_fused(ws.h_a, ws.h_b, ws.par_hist + (size_t)di * N, ws.tok_hist + (size_t)di * N, B, K, rps, H, di, done,
                                       fin + di, steps_run, fused.hprev, fused.cat_e(H), (const uint16_t*)fused.w_emb->hi,
                                       (const uint16_t*)fused.w_emb->lo, fused.w_emb->ld, E, V, st));
        else
            VAG_TRY(beam_advance(ws.h_a, ws.h_b, ws.par_hist + (size_t)di * N, B, K, rps, H, di, done, fin + di, steps_run, st));
    }
    VAG_TRY(beam_finalize(ws.tok_hist, ws.par_hist, ws.nll, steps_run, B, K, L, hyp_out, hyp_len, beam_out, st));
    if (nll_out) VAG_CUDA(cudaMemcpyAsync(nll_out, ws.nll, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, st));
    if (steps_out) VAG_CUDA(cudaMemcpyAsync(steps_out, steps_run, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return VAG_OK;
}

extern "C" int vag_greedy_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                                     const float* mask, int B, int T, int L, int64_t* tokens_out, void* workspace,
                                     size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(w && h0 && keys && ctx && mask && tokens_out, "vag_greedy_decode_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0 && L > 0, "vag_greedy_decode_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    ArenaAdapter ar(workspace, workspace_bytes);
    BeamWs ws;
    beam_layout(ar, B, 1, L, w->E, w->H, w->C, w->V, &ws);
    if (ar.a.overflow) {
        set_error("vag_greedy_decode_f32: workspace %zu B too small (use vag_beam_decode_workspace_bytes with K=1)", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    fill_i64_kernel<<<ceil_div(B, 256), 256, 0, st>>>(ws.sos, 2, B);
    VAG_LAUNCH_CHECK();
    VAG_CUDA(cudaMemcpyAsync(ws.h_a, h0, sizeof(float) * (size_t)B * w->H, cudaMemcpyDeviceToDevice, st));
    float* h_cur = ws.h_a;
    float* h_nxt = ws.h_b;
    const int64_t ldl = (w->V + 3) / 4 * 4;
    GemmCtx gemm(st, ws.step.wreg, ws.step.wbytes, ws.step.areg, ws.step.abytes);
    for (int di = 0; di < L; ++di) {
        VAG_TRY(decoder_step_core(gemm, w, ws.step, ws.sos, h_cur, keys, ctx, mask, B, 1, T, h_nxt, ws.logits, ldl, nullptr, st));
        VAG_TRY(row_argmax(ws.logits, ldl, B, w->V, tokens_out + di, L, ws.sos, st));
        std::swap(h_cur, h_nxt);
    }
    return VAG_OK;
}
