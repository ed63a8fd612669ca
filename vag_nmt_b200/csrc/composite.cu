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
                 int* done, int* fin_counter, int* steps_run, cudaStream_t st, volatile int32_t* host_progress = nullptr, int nonce = 0);
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
    // GRU cell fused into its hidden contraction (gru_pair_kernel, > 128 active rows): permuted W_hh planes and unit-major
    // biases per direction, a second state buffer (the fused kernel must not write the rows other tiles still read) and the
    // operand planes of both state buffers
    void *whp[2], *wlp[2];
    float* bias4[2];
    float* h_alt[2];
    void *php[2][2], *plp[2][2];
    float* permute_scratch;
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
    void *whp[2], *wlp[2], *php[2][2], *plp[2][2];
    float *bias4[2], *h_alt[2];
    for (int d = 0; d < 2; ++d) {
        whp[d] = a.template take<uint16_t>((size_t)3 * H * H);
        wlp[d] = a.template take<uint16_t>((size_t)3 * H * H);
        bias4[d] = (float*)a.template take<float>((size_t)4 * H);
        h_alt[d] = (float*)a.template take<float>((size_t)B * H);
        for (int q = 0; q < 2; ++q) {
            php[d][q] = a.template take<uint16_t>((size_t)B * H);
            plp[d][q] = a.template take<uint16_t>((size_t)B * H);
        }
    }
    float* permute_scratch = (float*)a.template take<float>((size_t)3 * H * H);
    if (ws) {
        ws->x = x; ws->gi[0] = gi0; ws->gi[1] = gi1; ws->gh[0] = gh0; ws->gh[1] = gh1; ws->h[0] = h0; ws->h[1] = h1;
        ws->wreg = wr; ws->wbytes = wb; ws->areg = ar; ws->abytes = ab;
        for (int d = 0; d < 2; ++d) {
            ws->whp[d] = whp[d]; ws->wlp[d] = wlp[d]; ws->bias4[d] = bias4[d]; ws->h_alt[d] = h_alt[d];
            for (int q = 0; q < 2; ++q) { ws->php[d][q] = php[d][q]; ws->plp[d][q] = plp[d][q]; }
        }
        ws->permute_scratch = permute_scratch;
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

// Operand planes of one weight matrix [rows, K] (16-bit modes: 2 bytes per element, lo unused in bf16 mode).
struct Planes {
    void *hi = nullptr, *lo = nullptr;
    int64_t ld = 0;
};
// Decode-call invariants of a weight set (vag_decoder_prepare_f32): every decoder matrix as operand planes, the K-concatenated
// read-out matrix [W1 | W3 | W2] with its summed bias, the per-token gru_1 table.  The layout is a pure function of
// (E, H, C, V): the same code sizes the buffer, fills it and finds the pieces again in a later call.
struct Prepared {
    Planes g1i, g1h, ah, ae, c2h, g2i, g2h, ro, out, emb, ini;
    Planes g1h_p, g2i_p, g2h_p;   // row-permuted copies for the GRU-fused contraction (gru_pair.cuh)
    float* gb1 = nullptr;         // [4][H] unit-major biases of gru_1 (input side in the per-token table) and gru_2
    float* gb2 = nullptr;
    float* scratch = nullptr;     // [3H, max(E, H)] fp32 staging of a permuted matrix
    float* b_ro = nullptr;
    float* g1 = nullptr;   // [V, 3H] table  Emb·W_ihᵀ + b_ih  of gru_1: its input pre-activations depend on the token only
};
template <typename A>
static void prepared_layout(A& a, int E, int H, int C, int64_t V, Prepared* p) {
    auto planes = [&](Planes* d, int64_t rows, int64_t K) {
        void* hi = a.template take<uint16_t>((size_t)rows * K);
        void* lo = a.template take<uint16_t>((size_t)rows * K);
        if (d) { d->hi = hi; d->lo = lo; d->ld = K; }
    };
    planes(p ? &p->g1i : nullptr, 3 * H, E);
    planes(p ? &p->g1h : nullptr, 3 * H, H);
    planes(p ? &p->ah : nullptr, C, H);
    planes(p ? &p->ae : nullptr, C, C);
    planes(p ? &p->c2h : nullptr, H, C);
    planes(p ? &p->g2i : nullptr, 3 * H, H);
    planes(p ? &p->g2h : nullptr, 3 * H, H);
    planes(p ? &p->ro : nullptr, E, H + E + C);
    planes(p ? &p->out : nullptr, V, E);
    planes(p ? &p->emb : nullptr, V, E);
    planes(p ? &p->ini : nullptr, H, C);
    planes(p ? &p->g1h_p : nullptr, 3 * H, H);
    planes(p ? &p->g2i_p : nullptr, 3 * H, H);
    planes(p ? &p->g2h_p : nullptr, 3 * H, H);
    float* gb1 = (float*)a.template take<float>((size_t)4 * H);
    float* gb2 = (float*)a.template take<float>((size_t)4 * H);
    float* scratch = (float*)a.template take<float>((size_t)3 * H * (E > H ? E : H));
    float* b_ro = (float*)a.template take<float>((size_t)E);
    float* g1 = (float*)a.template take<float>((size_t)V * 3 * H);
    if (p) { p->b_ro = b_ro; p->g1 = g1; p->gb1 = gb1; p->gb2 = gb2; p->scratch = scratch; }
}
static bool prepared_supported(const vag_decoder_weights* w, int mode) {
    const int E = w->E, H = w->H, C = w->C;
    if (!tc_enabled() || (mode != 1 && mode != 2) || (E % 8) || (H % 8) || (C % 8) || w->V < 64) return false;
    const float* ws_[] = {w->gru1_w_ih, w->gru1_w_hh, w->attn_h_w, w->attn_e_w, w->c2h_w, w->gru2_w_ih, w->gru2_w_hh, w->w1_w, w->w2_w, w->w3_w,
                          w->out_w, w->emb};
    for (const float* q : ws_)
        if (!q || ((uintptr_t)q & 15)) return false;
    return true;
}
// Fill a prepared region (mode = the current gemm mode, 1 or 2).
static int prepared_fill(const vag_decoder_weights* w, Prepared& pr, cudaStream_t st) {
    const int E = w->E, H = w->H, C = w->C, Kt = H + E + C;
    const int V = (int)w->V;
    auto split = [&](const Planes& d, const float* src, int rows, int K, int64_t col_off = 0, int64_t ld = 0) {
        return tc_split(src, K, rows, K, d.hi, d.lo, ld ? ld : d.ld, col_off, st);
    };
    VAG_TRY(split(pr.g1i, w->gru1_w_ih, 3 * H, E));
    VAG_TRY(split(pr.g1h, w->gru1_w_hh, 3 * H, H));
    VAG_TRY(split(pr.ah, w->attn_h_w, C, H));
    VAG_TRY(split(pr.ae, w->attn_e_w, C, C));
    VAG_TRY(split(pr.c2h, w->c2h_w, H, C));
    VAG_TRY(split(pr.g2i, w->gru2_w_ih, 3 * H, H));
    VAG_TRY(split(pr.g2h, w->gru2_w_hh, 3 * H, H));
    VAG_TRY(split(pr.out, w->out_w, V, E));
    if (w->emb != w->out_w) VAG_TRY(split(pr.emb, w->emb, V, E));
    if (w->ini_w && !((uintptr_t)w->ini_w & 15)) VAG_TRY(split(pr.ini, w->ini_w, H, C));
    VAG_TRY(split(pr.ro, w->w1_w, E, H, 0, Kt));        // read-out [W1 | W3 | W2] laid side by side along K (NMT_Decoder.py:137)
    VAG_TRY(split(pr.ro, w->w3_w, E, E, H, Kt));
    VAG_TRY(split(pr.ro, w->w2_w, E, C, H + E, Kt));
    VAG_TRY(bias_sum3(pr.b_ro, w->w1_b, w->w3_b, w->w2_b, E, st));
    if (H % 32 == 0) {   // GRU-fused contraction: permuted matrices + unit-major biases
        VAG_TRY(tc_gru_prepare_weight(pr.g1h_p.hi, pr.g1h_p.lo, w->gru1_w_hh, H, H, true, pr.scratch, st));
        VAG_TRY(tc_gru_prepare_weight(pr.g2i_p.hi, pr.g2i_p.lo, w->gru2_w_ih, H, H, false, pr.scratch, st));
        VAG_TRY(tc_gru_prepare_weight(pr.g2h_p.hi, pr.g2h_p.lo, w->gru2_w_hh, H, H, true, pr.scratch, st));
        VAG_TRY(tc_gru_prepare_bias(pr.gb1, nullptr, w->gru1_b_hh, H, false, st));
        VAG_TRY(tc_gru_prepare_bias(pr.gb2, w->gru2_b_ih, w->gru2_b_hh, H, true, st));
    }
    // gru_1's input contraction once for EVERY token instead of once per step for every row: the same kernel on the same operand
    // rows, so each table row is bit-identical to what the per-step contraction produces
    if (V > 128) {
        const Planes& em = (w->emb == w->out_w) ? pr.out : pr.emb;
        VAG_TRY(tc_gemm(pr.g1, 3 * H, em.hi, em.lo, em.ld, pr.g1i.hi, pr.g1i.lo, pr.g1i.ld, w->gru1_b_ih, V, E, 3 * H, 0, st, nullptr, nullptr));
    }
    return VAG_OK;
}
// The prepared data of this call: the caller's buffer (w->prepared) when it is large enough, else `fallback` (a region of the
// call's own workspace) filled now.  *have = false when the tensor-core step cannot run in this mode / for these shapes.
static int prepared_get(const vag_decoder_weights* w, void* fallback, size_t fallback_bytes, cudaStream_t st, Prepared* pr, bool* have) {
    *have = false;
    const int mode = gemm_mode();
    if (!prepared_supported(w, mode)) return VAG_OK;
    SizerAdapter sz;
    prepared_layout(sz, w->E, w->H, w->C, w->V, nullptr);
    const size_t need = sz.s.total();
    if (w->prepared && w->prepared_bytes >= need) {
        ArenaAdapter ar(const_cast<void*>(w->prepared), w->prepared_bytes);
        prepared_layout(ar, w->E, w->H, w->C, w->V, pr);
        *have = !ar.a.overflow;
        return VAG_OK;
    }
    if (!fallback || fallback_bytes < need) return VAG_OK;
    ArenaAdapter ar(fallback, fallback_bytes);
    prepared_layout(ar, w->E, w->H, w->C, w->V, pr);
    if (ar.a.overflow) return VAG_OK;
    VAG_TRY(prepared_fill(w, *pr, st));
    *have = true;
    return VAG_OK;
}
// Make the prepared planes visible to the generic contraction context (its weight cache is keyed by the fp32 pointer).
static void prepared_register(GemmCtx& gemm, const vag_decoder_weights* w, const Prepared& pr) {
    const int E = w->E, H = w->H, C = w->C;
    gemm.preset(w->gru1_w_ih, 3 * H, E, pr.g1i.hi, pr.g1i.lo);
    gemm.preset(w->gru1_w_hh, 3 * H, H, pr.g1h.hi, pr.g1h.lo);
    gemm.preset(w->attn_h_w, C, H, pr.ah.hi, pr.ah.lo);
    gemm.preset(w->attn_e_w, C, C, pr.ae.hi, pr.ae.lo);
    gemm.preset(w->c2h_w, H, C, pr.c2h.hi, pr.c2h.lo);
    gemm.preset(w->gru2_w_ih, 3 * H, H, pr.g2i.hi, pr.g2i.lo);
    gemm.preset(w->gru2_w_hh, 3 * H, H, pr.g2h.hi, pr.g2h.lo);
    gemm.preset(w->out_w, (int)w->V, E, pr.out.hi, pr.out.lo);
    if (w->ini_w && !((uintptr_t)w->ini_w & 15)) gemm.preset(w->ini_w, H, C, pr.ini.hi, pr.ini.lo);
    gemm.preset3(w->w1_w, E, H + E + C, pr.ro.hi, pr.ro.lo, pr.b_ro);
}


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
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
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
    // Steps with more than 128 active rows run the whole cell in ONE launch (gru_pair_kernel: hidden contraction + gates on the TMEM
    // drain, the input pre-activations gi[t] gathered like the decoder's per-token table with table row = sentence, h' written
    // into the state buffer, its operand planes and the context slab) instead of contraction + operand split + gate kernel.  The
    // fused kernel ping-pongs between two state buffers; rows beyond the active prefix are never written, so a sentence that
    // starts later (backward direction) finds zeros in both.  Smaller steps keep the three-kernel path on the current buffer.
    static const bool enc_fused_off = getenv("VAG_ENC_FUSED") && getenv("VAG_ENC_FUSED")[0] == '0';
    const int mode = gemm_mode();
    bool fused_ok = !enc_fused_off && (mode == 1 || mode == 2) && tc_gru_supported(n_act[0], H, 0, H) && (H % 16) == 0 &&
                    ((reinterpret_cast<uintptr_t>(ctx_out) & 15) == 0);
    float* hbuf[2][2] = {{ws.h[0], ws.h_alt[0]}, {ws.h[1], ws.h_alt[1]}};
    int cur[2] = {0, 0};
    bool planes_valid[2] = {false, false};
    if (fused_ok) {
        for (int d = 0; d < 2; ++d) {
            VAG_TRY(tc_gru_prepare_weight(ws.whp[d], ws.wlp[d], w->w_hh[d], H, H, true, ws.permute_scratch, st));
            VAG_TRY(tc_gru_prepare_bias(ws.bias4[d], nullptr, w->b_hh[d], H, false, st));
            VAG_CUDA(cudaMemsetAsync(ws.h_alt[d], 0, (size_t)B * H * sizeof(float), st));
        }
    }
    // the two directions are independent chains; interleave them so neighbouring launches can overlap their tails
    for (int s = 0; s < T; ++s) {
        for (int d = 0; d < 2; ++d) {
            const int t = d == 0 ? s : T - 1 - s;
            const int n = n_act[t];
            if (n == 0) continue;
            float* hc = hbuf[d][cur[d]];
            if (fused_ok && tc_gru_supported(n, H, 0, H)) {
                float* hn = hbuf[d][cur[d] ^ 1];
                if (!planes_valid[d]) {   // first fused step of this direction (or after three-kernel steps): planes of ALL rows of the current state
                    VAG_TRY(tc_split(hc, H, B, H, ws.php[d][cur[d]], ws.plp[d][cur[d]], H, 0, st));
                    VAG_CUDA(cudaMemsetAsync(ws.php[d][cur[d] ^ 1], 0, (size_t)B * H * 2, st));
                    VAG_CUDA(cudaMemsetAsync(ws.plp[d][cur[d] ^ 1], 0, (size_t)B * H * 2, st));
                    planes_valid[d] = true;
                }
                GruCall g;
                g.hh = ws.php[d][cur[d]]; g.hl = ws.plp[d][cur[d]]; g.ldh = H; g.Kh = H;
                g.whh_h = ws.whp[d]; g.whh_l = ws.wlp[d];
                g.bias4 = ws.bias4[d]; g.g1 = ws.gi[d] + (int64_t)t * B * 3 * H; g.tokens = nullptr; g.V = B;
                g.h_prev = hc; g.h_out = hn;
                g.out.hi = (uint16_t*)ws.php[d][cur[d] ^ 1]; g.out.lo = (uint16_t*)ws.plp[d][cur[d] ^ 1]; g.out.ld = H; g.out.mode = mode;
                g.rows = n; g.H = H;
                g.y2 = ctx_out + (int64_t)t * 2 * H + d * H; g.ld_y2 = (int64_t)T * 2 * H;
                VAG_TRY(tc_gru(g, st));
                cur[d] ^= 1;
                continue;
            }
            planes_valid[d] = false;
            gemm.new_step();  // h changed: its split is stale
            VAG_TRY(gemm.linear(ws.gh[d], 3 * H, hc, H, w->w_hh[d], H, w->b_hh[d], n, H, 3 * H, 0));
            VAG_TRY(vag_gru_gates_f32(hc, H, ctx_out + (int64_t)t * 2 * H + d * H, (int64_t)T * 2 * H,
                                      ws.gi[d] + (int64_t)t * B * 3 * H, 3 * H, ws.gh[d], 3 * H, hc, H, n, H, stream));
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

namespace vag {
// Shared body of the inference forward and the training forward (which keeps what the backward needs: `sv`).
static int vse_pool_core(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T, float* im_emb,
                         float* txt_emb, float* ctx_vec, float* beta, const vag_vse_saved* sv, void* workspace, size_t workspace_bytes,
                         vag_stream_t stream, const char* who) {
    VAG_REQUIRE(w && im && ctx && im_emb && txt_emb && ctx_vec, "%s: null pointer", who);
    VAG_REQUIRE(B > 0 && T > 0, "%s: bad shape", who);
    VAG_REQUIRE(w->method == VAG_ATTN_DOT || (w->method == VAG_ATTN_MLP && w->mlp_w), "%s: bad attention method", who);
    const int I = w->I, C = w->C, S = w->S;
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    float* pk = ar.take<float>((size_t)B * T * C);
    float* iq = ar.take<float>((size_t)B * C);
    const size_t wb = vse_wbytes(I, C, S), ab = vse_abytes(B, T, I, C, S);
    void* wr = ar.take<char>(wb);
    void* areg = ar.take<char>(ab);
    if (ar.overflow) {
        set_error("%s: workspace %zu B too small", who, workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    if (sv) {   // training: the projected keys / query are outputs, not scratch
        VAG_REQUIRE(sv->a_im && sv->iq && sv->pk && sv->a_txt, "%s: saved-activation pointers missing", who);
        pk = sv->pk;
        iq = sv->iq;
    }
    float* a_im = sv ? sv->a_im : im_emb;
    float* a_txt = sv ? sv->a_txt : txt_emb;
    GemmCtx gemm(st, wr, wb, areg, ab);
    const int act = w->activation ? VAG_LIN_TANH : 0;
    VAG_TRY(gemm.linear(a_im, S, im, I, w->im_w, I, w->im_b, B, I, S, act));              // VSE_Imagine_Enc.py:123-127
    if (sv) VAG_CUDA(cudaMemcpyAsync(im_emb, a_im, sizeof(float) * (size_t)B * S, cudaMemcpyDeviceToDevice, st));
    VAG_TRY(vag_l2norm_rows_f32(im_emb, S, B, S, stream));                                        // :132
    VAG_TRY(gemm.linear(iq, C, im_emb, S, w->emb2ctx_w, S, nullptr, B, S, C, 0));         // :58
    VAG_TRY(gemm.linear(pk, C, ctx, C, w->ctx2ctx_w, C, nullptr, B * T, C, C, 0));        // :57
    VAG_TRY(vag_attention_f32(ctx_vec, C, beta, iq, C, pk, ctx, w->mlp_w, mask, B, 1, T, C, w->method, stream));  // :135-137
    VAG_TRY(gemm.linear(a_txt, S, ctx_vec, C, w->txt_w, C, w->txt_b, B, C, S, act));      // :138-140
    if (sv) VAG_CUDA(cudaMemcpyAsync(txt_emb, a_txt, sizeof(float) * (size_t)B * S, cudaMemcpyDeviceToDevice, st));
    VAG_TRY(vag_l2norm_rows_f32(txt_emb, S, B, S, stream));                                       // :145
    return VAG_OK;
}
}  // namespace vag

extern "C" int vag_vse_pool_fwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                                    float* im_emb, float* txt_emb, float* ctx_vec, float* beta, void* workspace,
                                    size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    return vse_pool_core(w, im, ctx, mask, B, T, im_emb, txt_emb, ctx_vec, beta, nullptr, workspace, workspace_bytes, stream, "vag_vse_pool_fwd_f32");
}

extern "C" int vag_vse_pool_train_fwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                                          float* im_emb, float* txt_emb, float* ctx_vec, float* beta, const vag_vse_saved* saved,
                                          void* workspace, size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(saved && beta, "vag_vse_pool_train_fwd_f32: saved activations and beta are required");
    return vse_pool_core(w, im, ctx, mask, B, T, im_emb, txt_emb, ctx_vec, beta, saved, workspace, workspace_bytes, stream, "vag_vse_pool_train_fwd_f32");
}

// ------------------------------------------------------------------ decoder
extern "C" size_t vag_attn_keys_workspace_bytes(int B, int T, int C) {
    return GemmCtx::split_bytes(C, C) + GemmCtx::split_bytes((int64_t)B * T, C) + 8192;
}

extern "C" int vag_attn_keys_f32(const vag_decoder_weights* w, const float* ctx, int B, int T, float* keys, void* workspace,
                                 size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && ctx && keys && B > 0 && T > 0, "vag_attn_keys_f32: bad argument");
    const size_t wb = GemmCtx::split_bytes(w->C, w->C) + 4096;
    Arena ar(workspace, workspace_bytes);
    void* wr = ar.take<char>(wb);
    const size_t ab = ar.overflow || workspace_bytes < ar.off + 512 ? 0 : workspace_bytes - align_up(ar.off, 256) - 256;
    void* areg = ab ? ar.take<char>(ab) : nullptr;
    GemmCtx gemm((cudaStream_t)stream, ar.overflow ? nullptr : wr, wb, ar.overflow ? nullptr : areg, ab);
    Prepared pr;
    bool have_pr = false;
    VAG_TRY(prepared_get(w, nullptr, 0, (cudaStream_t)stream, &pr, &have_pr));
    if (have_pr) gemm.preset(w->attn_e_w, w->C, w->C, pr.ae.hi, pr.ae.lo);
    return gemm.linear(keys, w->C, ctx, w->C, w->attn_e_w, w->C, nullptr, B * T, w->C, w->C, 0);
}

extern "C" size_t vag_decoder_init_workspace_bytes(int B, int C, int H) {
    return align_up((size_t)B * C * 4, 256) + GemmCtx::split_bytes(H, C) + GemmCtx::split_bytes(B, C) + 16384;
}

extern "C" int vag_decoder_init_f32(const vag_decoder_weights* w, const float* ctx_vec, const float* ctx, const float* mask,
                                    float split, int B, int T, float* h0, void* workspace, size_t workspace_bytes,
                                    vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
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
    Prepared pr;
    bool have_pr = false;
    VAG_TRY(prepared_get(w, nullptr, 0, (cudaStream_t)stream, &pr, &have_pr));
    if (have_pr && !((uintptr_t)w->ini_w & 15)) gemm.preset(w->ini_w, w->H, w->C, pr.ini.hi, pr.ini.lo);
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
                    const int64_t* gi_rows = nullptr, int64_t gi_n_rows = 0, const int* done = nullptr);
int embed_split_rows(SplitDst dst, const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, const int64_t* tokens, int rows,
                     int64_t V, cudaStream_t st);
int attention_mlp_split(SplitDst sd, const float* q, int64_t ld_q, const float* keys, const float* ctx, const float* v,
                        const float* mask, int rows, int rows_per_sent, int T, int C, cudaStream_t st, const int* done = nullptr,
                        const float* ekeys = nullptr, const int* kflag = nullptr);
int attn_exp_keys(float* ekeys, int* kflag, const float* keys, int B, int T, int C, cudaStream_t st);
struct SelAdvance {   // mirrors beam.cu (tail of the selection kernel that replaces beam_advance_fused)
    float* h_next = nullptr;
    const float* h_cur = nullptr;
    int H = 0, E = 0;
    SplitDst h_sd, e_sd;
    const uint16_t* emb_hi = nullptr;
    const uint16_t* emb_lo = nullptr;
    int64_t ld_emb = 0;
    int* done = nullptr;
    int* steps_run = nullptr;
    int* ticket = nullptr;
    volatile int32_t* host_progress = nullptr;
    int nonce = 0;
};
int beam_select_top2(const float4* summ, int tile_w, SplitDst t, const uint16_t* w_hi, const uint16_t* w_lo, int64_t ld_w,
                     const float* bias, int E, const int64_t* prev_tokens, float* nll, int64_t* tokens_out, int32_t* parents_out,
                     int B, int K, int64_t V, int step, int avoid_double, const int* done, int* fin_counter, cudaStream_t st,
                     const SelAdvance* adv = nullptr);
int beam_advance_fused(float* h_next, const float* h_cur, const int32_t* parents, const int64_t* tokens, int B, int K, int Kin,
                       int H, int step, int* done, int* fin_counter, int* steps_run, SplitDst h_sd, SplitDst e_sd,
                       const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, int64_t V, cudaStream_t st,
                       volatile int32_t* host_progress = nullptr, int nonce = 0);

struct FusedStep {
    bool ok = false;
    SplitDst hprev, h1, x2, t, cat;                 // cat = [h2 | e | c], pitch H + E + C
    Prepared pr;
    const Planes* w_emb = nullptr;                  // tied: the gather reads the projection's planes
    float* g1 = nullptr;
    const int* done = nullptr;                      // device flag: every hypothesis has ended, later steps return at once
    const float* ekeys = nullptr;                   // exp(2·keys) [B, T, C] and its per-sentence range flags (attention.cu, factored form)
    const int* kflag = nullptr;
    bool gru_fused = false;                         // the two GRU cells run as gru_pair_kernel (contractions + gates in one launch)
    SplitDst cat_e(int H) const { SplitDst d = cat; d.hi += H; if (d.lo) d.lo += H; return d; }
    SplitDst cat_c(int H, int E) const { SplitDst d = cat; d.hi += H + E; if (d.lo) d.lo += H + E; return d; }
};

static int fused_setup(const vag_decoder_weights* w, const StepWs& ws, int n_rows, int rows_per_sent, const Prepared& pr, FusedStep* f) {
    const int E = w->E, H = w->H, Kt = H + E + w->C;
    const int mode = gemm_mode();
    f->ok = false;
    if (n_rows <= 128 || rows_per_sent > 16) return VAG_OK;
    f->pr = pr;
    f->w_emb = (w->emb == w->out_w) ? &f->pr.out : &f->pr.emb;
    f->g1 = w->V > 128 ? pr.g1 : nullptr;
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
    f->ok = true;
    f->gru_fused = tc_gru_supported(n_rows, H, H, H) && !getenv("VAG_GRU_UNFUSED");
    return VAG_OK;
}

// One fused step: the embedding planes (cat_e) and the previous state's planes (hprev) are already in place.
static int decoder_step_fused(const FusedStep& f, const vag_decoder_weights* w, const StepWs& ws, const int64_t* tokens, const float* h_prev, const float* keys,
                              const float* ctx, const float* mask, int rows, int rows_per_sent, int T, float* h_out, float* logits,
                              int64_t ld_logits, cudaStream_t st, float4* summ, int* summ_tile_w) {
    const int E = w->E, H = w->H, C = w->C, Kt = H + E + C;
    const int64_t V = w->V;
    const SplitDst ce = f.cat_e(H), cc = f.cat_c(H, E);
    const Prepared& pr = f.pr;
    TcDoneScope ds(f.done);   // the contractions launched below return at once when *done is set
    auto gemm = [&](float* y, int64_t ldy, const SplitDst& x, const Planes& we, const float* bias, int K, int N, float4* sm,
                    int* tw) {
        return tc_gemm(y, ldy, x.hi, x.lo, x.ld, we.hi, we.lo, we.ld, bias, rows, K, N, 0, st, sm, tw);
    };
    const bool fused_cells = f.gru_fused && rows > 128;
    if (fused_cells && f.g1) {   // gru_1: hidden contraction + per-token input pre-activations + gates in one launch  (NMT_Decoder.py:121)
        GruCall g;
        g.hh = f.hprev.hi; g.hl = f.hprev.lo; g.ldh = f.hprev.ld; g.Kh = H;
        g.whh_h = pr.g1h_p.hi; g.whh_l = pr.g1h_p.lo;
        g.bias4 = pr.gb1; g.g1 = f.g1; g.tokens = tokens; g.V = V;
        g.h_prev = h_prev; g.h_out = ws.h1; g.out = f.h1; g.rows = rows; g.H = H;
        VAG_TRY(tc_gru(g, st));
    } else {
        if (!f.g1) VAG_TRY(gemm(ws.gi, 3 * H, ce, pr.g1i, w->gru1_b_ih, E, 3 * H, nullptr, nullptr));
        VAG_TRY(gemm(ws.gh, 3 * H, f.hprev, pr.g1h, w->gru1_b_hh, H, 3 * H, nullptr, nullptr));
        if (f.g1) VAG_TRY(gru_gates_split(ws.h1, H, f.g1, 3 * H, ws.gh, 3 * H, h_prev, H, rows, H, f.h1, st, tokens, V, f.done));
        else VAG_TRY(gru_gates_split(ws.h1, H, ws.gi, 3 * H, ws.gh, 3 * H, h_prev, H, rows, H, f.h1, st, nullptr, 0, f.done));
    }
    VAG_TRY(gemm(ws.q, C, f.h1, pr.ah, nullptr, H, C, nullptr, nullptr));                                    // :47
    VAG_TRY(attention_mlp_split(cc, ws.q, C, keys, ctx, w->attn_v, mask, rows, rows_per_sent, T, C, st, f.done, f.ekeys, f.kflag));     // :124-126
    VAG_TRY(tc_gemm_split_out(f.x2, cc.hi, cc.lo, Kt, pr.c2h.hi, pr.c2h.lo, pr.c2h.ld, nullptr, rows, C, H, 0, st));   // :127
    if (fused_cells) {           // gru_2: both contractions + gates in one launch                                     (:129)
        GruCall g;
        g.xh = f.x2.hi; g.xl = f.x2.lo; g.ldx = f.x2.ld; g.Kx = H;
        g.hh = f.h1.hi; g.hl = f.h1.lo; g.ldh = f.h1.ld; g.Kh = H;
        g.wih_h = pr.g2i_p.hi; g.wih_l = pr.g2i_p.lo; g.whh_h = pr.g2h_p.hi; g.whh_l = pr.g2h_p.lo;
        g.bias4 = pr.gb2;
        g.h_prev = ws.h1; g.h_out = h_out; g.out = f.cat; g.rows = rows; g.H = H;
        VAG_TRY(tc_gru(g, st));
    } else {
        VAG_TRY(gemm(ws.gi, 3 * H, f.x2, pr.g2i, w->gru2_b_ih, H, 3 * H, nullptr, nullptr));
        VAG_TRY(gemm(ws.gh, 3 * H, f.h1, pr.g2h, w->gru2_b_hh, H, 3 * H, nullptr, nullptr));
        VAG_TRY(gru_gates_split(h_out, H, ws.gi, 3 * H, ws.gh, 3 * H, ws.h1, H, rows, H, f.cat, st, nullptr, 0, f.done));
    }
    VAG_TRY(tc_gemm_split_out(f.t, f.cat.hi, f.cat.lo, Kt, pr.ro.hi, pr.ro.lo, pr.ro.ld, pr.b_ro, rows, Kt, E, VAG_LIN_TANH, st));  // :137
    if (!logits) {   // beam loop: only the top-2 / Σexp summaries of every 128-column tile leave the projection
        if (summ_tile_w) *summ_tile_w = 128;
        return tc_gemm_top2(summ, f.t.hi, f.t.lo, f.t.ld, pr.out.hi, pr.out.lo, pr.out.ld, w->out_b, rows, E, (int)V, st);
    }
    return gemm(logits, ld_logits, f.t, pr.out, w->out_b, E, (int)V, summ, summ_tile_w);                     // :143
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
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
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
    {
        Prepared pr;
        bool have_pr = false;
        VAG_TRY(prepared_get(w, nullptr, 0, (cudaStream_t)stream, &pr, &have_pr));
        if (have_pr) prepared_register(gemm, w, pr);
    }
    VAG_TRY(decoder_step_core(gemm, w, ws, tokens, h_prev, keys, ctx, mask, rows, rows_per_sent, T, h_out, logits_or_logp, w->V,
                              alpha_out, (cudaStream_t)stream));
    if (want_logp) VAG_TRY(vag_log_softmax_f32(logits_or_logp, logits_or_logp, rows, (int)w->V, stream));
    return VAG_OK;
}

// ------------------------------------------------------------------ beam search
namespace vag {
struct BeamWs {
    StepWs step;
    float *logits, *lse, *h_a, *h_b, *nll;
    void* prep;          // the call's own copy of the decode invariants (unused when the caller passes w->prepared)
    size_t prep_bytes;
    float4* summ;
    int64_t *tok_hist, *sos;
    int32_t* par_hist;
    int* flags;  // [0] done, [1] steps_run, [2..2+L) per-step EOS counters, [2+L..2+2L) per-step CTA tickets (fused reorder)
    float* ekeys;
    int* kflag;
};
template <typename A>
static void beam_layout(A& a, int B, int K, int T, int L, int E, int H, int C, int64_t V, BeamWs* ws) {
    const int N = B * K;
    StepWs sw;
    step_layout(a, N, E, H, C, V, &sw);
    float* logits = (float*)a.template take<float>((size_t)N * ((V + 3) / 4 * 4));  // rows padded to 16 B
    float* lse = (float*)a.template take<float>((size_t)N);
    SizerAdapter psz;
    prepared_layout(psz, E, H, C, V, nullptr);
    const size_t prep_bytes = psz.s.total();
    void* prep = a.template take<char>(prep_bytes);
    float4* summ = (float4*)a.template take<float4>((size_t)N * ((V + 31) / 32));   // per 32-column slice (top-2 kernel); the per-128 summaries need a quarter
    float* h_a = (float*)a.template take<float>((size_t)N * H);
    float* h_b = (float*)a.template take<float>((size_t)N * H);
    float* nll = (float*)a.template take<float>((size_t)N);
    int64_t* tok_hist = (int64_t*)a.template take<int64_t>((size_t)L * N);
    int64_t* sos = (int64_t*)a.template take<int64_t>((size_t)B);
    int32_t* par_hist = (int32_t*)a.template take<int32_t>((size_t)L * N);
    int* flags = (int*)a.template take<int>(2 * (size_t)L + 2);
    float* ekeys = (float*)a.template take<float>((size_t)B * T * C);   // exp(2·keys) for the factored attention scores
    int* kflag = (int*)a.template take<int>((size_t)B);
    if (ws) {
        ws->ekeys = ekeys; ws->kflag = kflag;
        ws->step = sw; ws->logits = logits; ws->lse = lse; ws->prep = prep; ws->prep_bytes = prep_bytes; ws->summ = summ; ws->h_a = h_a; ws->h_b = h_b; ws->nll = nll;
        ws->tok_hist = tok_hist; ws->sos = sos; ws->par_hist = par_hist; ws->flags = flags;
    }
}
}  // namespace vag

extern "C" size_t vag_beam_decode_workspace_bytes(int B, int K, int T, int L, int E, int H, int C, int64_t V) {
    SizerAdapter s;
    beam_layout(s, B, K, T, L, E, H, C, V, nullptr);
    return s.s.total();
}

extern "C" size_t vag_decoder_prepared_bytes(int E, int H, int C, int64_t V) {
    SizerAdapter s;
    prepared_layout(s, E, H, C, V, nullptr);
    return s.s.total();
}

extern "C" int vag_decoder_prepare_f32(const vag_decoder_weights* w, void* prepared, size_t prepared_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && prepared, "vag_decoder_prepare_f32: null pointer");
    if (!prepared_supported(w, gemm_mode())) {
        set_error("vag_decoder_prepare_f32: needs 16-byte aligned weights, E/H/C multiples of 8, V >= 64 and a 16-bit tensor-core mode");
        return VAG_ERR_UNSUPPORTED;
    }
    ArenaAdapter ar(prepared, prepared_bytes);
    Prepared pr;
    prepared_layout(ar, w->E, w->H, w->C, w->V, &pr);
    if (ar.a.overflow) {
        set_error("vag_decoder_prepare_f32: buffer %zu B too small (vag_decoder_prepared_bytes)", prepared_bytes);
        return VAG_ERR_WORKSPACE;
    }
    return prepared_fill(w, pr, (cudaStream_t)stream);
}

// host_progress word: (call nonce & 0x7FFF) << 16 | done << 15 | steps finished (L < 32768)
static inline int progress_nonce() {
    static thread_local unsigned n = 0;
    n = (n + 1) & 0x7FFF;
    if (n == 0) n = 1;
    return (int)n;
}

// Steps [step_begin, step_end) of the search on the state the workspace holds; step_begin == 0 initialises that state, step_end >= L
// appends the epilogue.  vag_beam_decode_f32 is the range [0, L); vag_beam_decode_steps_f32 exposes the ranges so that a caller can
// capture the loop as a few CUDA graphs (one per chunk of steps) and look at `done` between chunks.
static int beam_decode_range(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                             const float* mask, int B, int K, int T, int L, int avoid_double, int step_begin, int step_end,
                             int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out,
                             volatile int32_t* host_progress, int32_t* done_host, void* workspace, size_t workspace_bytes,
                             vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && h0 && keys && ctx && mask && hyp_out && hyp_len, "vag_beam_decode_f32: null pointer");
    VAG_REQUIRE(B > 0 && K > 1 && T > 0 && L > 0 && L < 32768, "vag_beam_decode_f32: bad shape B=%d K=%d T=%d L=%d", B, K, T, L);
    VAG_REQUIRE(K <= w->V, "vag_beam_decode_f32: beam larger than the vocabulary");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = w->E, H = w->H, C = w->C;
    const int64_t V = w->V;
    const int N = B * K;
    ArenaAdapter ar(workspace, workspace_bytes);
    BeamWs ws;
    beam_layout(ar, B, K, T, L, E, H, C, V, &ws);
    if (ar.a.overflow) {
        set_error("vag_beam_decode_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    int* done = ws.flags;
    int* steps_run = ws.flags + 1;
    int* fin = ws.flags + 2;
    const bool first = step_begin == 0, last = step_end >= L;
    if (first) {
        VAG_CUDA(cudaMemsetAsync(ws.flags, 0, sizeof(int) * (2 * (size_t)L + 2), st));
        fill_i64_kernel<<<ceil_div(B, 256), 256, 0, st>>>(ws.sos, 2 /*SOS*/, B);
        VAG_LAUNCH_CHECK();
    }
    const int64_t ldl = (V + 3) / 4 * 4;
    GemmCtx gemm(st, ws.step.wreg, ws.step.wbytes, ws.step.areg, ws.step.abytes);
    Prepared pr;
    bool have_pr = false;
    VAG_TRY(prepared_get(w, ws.prep, ws.prep_bytes, st, &pr, &have_pr));
    FusedStep fused;
    if (have_pr) {
        prepared_register(gemm, w, pr);
        VAG_TRY(fused_setup(w, ws.step, N, K, pr, &fused));
        fused.done = done;
        if (fused.ok && (C % 4 == 0) && !getenv("VAG_ATTN_SINGLE_EXP")) {   // factored attention scores: exp(2·keys) once per call
            if (first) VAG_TRY(attn_exp_keys(ws.ekeys, ws.kflag, keys, B, T, C, st));
            fused.ekeys = ws.ekeys;
            fused.kflag = ws.kflag;
        }
    }
    const int nonce = host_progress ? progress_nonce() : 0;
    // VAG_SEL_ADVANCE=1: the reorder runs as the tail of the selection kernel (one launch less per step).  Opt-in: token-identical
    // but measured SLOWER than the separate bandwidth-shaped launch in every regime (1000 sentences bf16 34.86 vs 34.32 ms, FP32
    // 49.2 vs 48.9; 16 sentences 9.85 vs 9.76 ms) — a sentence's 84 KB of copies serialise behind its own pops.
    static const bool sel_advance = getenv("VAG_SEL_ADVANCE") && getenv("VAG_SEL_ADVANCE")[0] == '1';
    constexpr int kLookahead = 4;     // steps the host may enqueue ahead of the device when it polls host_progress
    for (int di = step_begin; di < (last ? L : step_end); ++di) {
        if (host_progress && di > kLookahead) {
            // The reference tests `n_fini == B·K` on the host after EVERY step (V11:265-269).  Here the device publishes its
            // progress in mapped host memory and the host only stays kLookahead steps ahead: when `done` appears it stops
            // enqueueing, so a search that ends after s steps costs s (+ at most kLookahead returned-at-once) steps, not L.
            bool stop = false;
            for (unsigned spin = 0;; ++spin) {
                const int v = *host_progress;
                const bool mine = ((v >> 16) & 0x7FFF) == nonce;
                if (mine && (v & 0x8000)) { stop = true; break; }
                if (mine && (v & 0x7FFF) >= di - kLookahead) break;
                if ((spin & 0x3FF) == 0x3FF) {   // every ~1000 polls: has the stream died or drained without progress?
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q != cudaErrorNotReady) {
                        if (q != cudaSuccess) { set_error("vag_beam_decode_f32: stream error while polling: %s", cudaGetErrorString(q)); return VAG_ERR_CUDA; }
                        const int v2 = *host_progress;   // drained: the last store is visible now
                        if (((v2 >> 16) & 0x7FFF) == nonce && (v2 & 0x8000)) stop = true;
                        break;
                    }
                }
            }
            if (stop) break;
        }
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
                // the reorder (state rows by parent, embedding planes, stop test) runs as the tail of the selection kernel
                SelAdvance adv;
                if (sel_advance && (H % 4) == 0 && (E % 8) == 0) {
                    adv.h_next = ws.h_a; adv.h_cur = ws.h_b; adv.H = H; adv.E = E; adv.h_sd = fused.hprev; adv.e_sd = fused.cat_e(H);
                    adv.emb_hi = (const uint16_t*)fused.w_emb->hi; adv.emb_lo = (const uint16_t*)fused.w_emb->lo; adv.ld_emb = fused.w_emb->ld;
                    adv.done = done; adv.steps_run = steps_run; adv.ticket = fin + L + di; adv.host_progress = host_progress; adv.nonce = nonce;
                }
                VAG_TRY(beam_select_top2(ws.summ, 32, fused.t, (const uint16_t*)pr.out.hi, (const uint16_t*)pr.out.lo,
                                         pr.out.ld, w->out_b, E, di == 0 ? nullptr : tokens, ws.nll, ws.tok_hist + (size_t)di * N,
                                         ws.par_hist + (size_t)di * N, B, K, V, di, avoid_double, done, fin + di, st, &adv));
                tile_w = adv.h_next ? -2 : -1;   // selection done (-2: and the reorder with it)
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
        if (tile_w == -2) {
        } else if (fused.ok)
            VAG_TRY(beam_advance_fused(ws.h_a, ws.h_b, ws.par_hist + (size_t)di * N, ws.tok_hist + (size_t)di * N, B, K, rps, H, di, done,
                                       fin + di, steps_run, fused.hprev, fused.cat_e(H), (const uint16_t*)fused.w_emb->hi,
                                       (const uint16_t*)fused.w_emb->lo, fused.w_emb->ld, E, V, st, host_progress, nonce));
        else
            VAG_TRY(beam_advance(ws.h_a, ws.h_b, ws.par_hist + (size_t)di * N, B, K, rps, H, di, done, fin + di, steps_run, st,
                                 host_progress, nonce));
    }
    if (done_host) VAG_CUDA(cudaMemcpyAsync(done_host, done, sizeof(int), cudaMemcpyDeviceToHost, st));   // pinned host word: `done` after this range
    if (!last) return VAG_OK;
    VAG_TRY(beam_finalize(ws.tok_hist, ws.par_hist, ws.nll, steps_run, B, K, L, hyp_out, hyp_len, beam_out, st));
    if (nll_out) VAG_CUDA(cudaMemcpyAsync(nll_out, ws.nll, sizeof(float) * (size_t)N, cudaMemcpyDeviceToDevice, st));
    if (steps_out) VAG_CUDA(cudaMemcpyAsync(steps_out, steps_run, sizeof(int), cudaMemcpyDeviceToDevice, st));
    return VAG_OK;
}

extern "C" int vag_beam_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                                   const float* mask, int B, int K, int T, int L, int avoid_double, int64_t* hyp_out,
                                   int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out,
                                   volatile int32_t* host_progress, void* workspace, size_t workspace_bytes, vag_stream_t stream) {
    return beam_decode_range(w, h0, keys, ctx, mask, B, K, T, L, avoid_double, 0, L, hyp_out, hyp_len, beam_out, nll_out, steps_out,
                             host_progress, nullptr, workspace, workspace_bytes, stream);
}

/* Steps [step_begin, step_end) of the same search, on the state a previous range left in `workspace` (same buffer, same shape
 * arguments): step_begin == 0 initialises it, step_end >= L appends the epilogue and writes the outputs (step_begin == step_end == L:
 * the epilogue alone).  done_host (optional): one int32 of PINNED host memory that receives the `done` flag as it stands after the
 * range — asynchronously, in stream order.  No host polling inside: every range can be captured in a CUDA graph, and a caller that
 * replays chunk graphs one ahead of an event can stop launching once `done` arrives (V11:265-269). */
extern "C" int vag_beam_decode_steps_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                                         const float* mask, int B, int K, int T, int L, int avoid_double, int step_begin, int step_end,
                                         int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out,
                                         int32_t* done_host, void* workspace, size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(step_begin >= 0 && step_begin <= step_end && step_begin <= L, "vag_beam_decode_steps_f32: bad step range [%d, %d) of %d",
                step_begin, step_end, L);
    return beam_decode_range(w, h0, keys, ctx, mask, B, K, T, L, avoid_double, step_begin, step_end, hyp_out, hyp_len, beam_out, nll_out,
                             steps_out, nullptr, done_host, workspace, workspace_bytes, stream);
}

/* The epilogue of the search alone (V11:315-337) for callers that drive the steps themselves with vag_decoder_step_f32 +
 * vag_beam_select_f32: tok_hist / par_hist [L, B, K] (token and parent-beam index chosen at every step), nll [B, K],
 * steps_run[1] on the device (rows >= steps_run count as 0, the last row is forced to <eos>). */
extern "C" int vag_beam_finalize_f32(const int64_t* tok_hist, const int32_t* par_hist, const float* nll, const int32_t* steps_run, int B,
                                     int K, int L, int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, vag_stream_t stream) {
    VAG_REQUIRE(tok_hist && par_hist && nll && steps_run && hyp_out && hyp_len, "vag_beam_finalize_f32: null pointer");
    VAG_REQUIRE(B > 0 && K > 0 && K <= 32 && L > 0, "vag_beam_finalize_f32: bad shape");
    return beam_finalize(tok_hist, par_hist, nll, (const int*)steps_run, B, K, L, hyp_out, hyp_len, beam_out, (cudaStream_t)stream);
}

extern "C" int vag_greedy_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                                     const float* mask, int B, int T, int L, int64_t* tokens_out, void* workspace,
                                     size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && h0 && keys && ctx && mask && tokens_out, "vag_greedy_decode_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0 && L > 0, "vag_greedy_decode_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    ArenaAdapter ar(workspace, workspace_bytes);
    BeamWs ws;
    beam_layout(ar, B, 1, T, L, w->E, w->H, w->C, w->V, &ws);
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
    {
        Prepared pr;
        bool have_pr = false;
        VAG_TRY(prepared_get(w, nullptr, 0, st, &pr, &have_pr));   // only the caller's buffer: greedy has no table to build
        if (have_pr) prepared_register(gemm, w, pr);
    }
    for (int di = 0; di < L; ++di) {
        VAG_TRY(decoder_step_core(gemm, w, ws.step, ws.sos, h_cur, keys, ctx, mask, B, 1, T, h_nxt, ws.logits, ldl, nullptr, st));
        VAG_TRY(row_argmax(ws.logits, ldl, B, w->V, tokens_out + di, L, ws.sos, st));
        std::swap(h_cur, h_nxt);
    }
    return VAG_OK;
}
