/*
 * vag_nmt.h — C ABI of libvagnmt.so: the B200 (sm_100a) kernels behind the VAG-NMT
 * per-timestep translation hot path.
 *
 * The reference (Eurus-Holmes/VAG-NMT) has no FFI: its boundary is the Python
 * nn.Module API.  The Python package vag_nmt_b200 mirrors that
 * API and binds this library with ctypes; every entry point below names the reference
 * code (file:line, relative to the upstream repository root) whose arithmetic it
 * replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - row-major, strides (ld*) in ELEMENTS; weights keep PyTorch's [out, in] layout;
 *   - no allocation and no global state inside a call: scratch is caller-owned (…_workspace_bytes() says how much), and the
 *     arithmetic mode travels WITH the call (the `precision` field of the weight structs, the VAG_LIN_BF16 flag or a
 *     `precision` argument) — there is no library-wide or thread-wide mode switch;
 *   - no host synchronisation, with one opt-in exception: vag_beam_decode_f32's `host_progress` early stop;
 *   - every call enqueues on `stream` (a cudaStream_t) and returns a vag_status;
 *   - on failure vag_last_error() (thread-local) describes the reason.
 */
#ifndef VAG_NMT_H
#define VAG_NMT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* vag_stream_t; /* cudaStream_t */

typedef enum {
    VAG_OK = 0,
    VAG_ERR_INVALID = -1,     /* bad argument (null pointer, non-positive size, unsorted lengths …) */
    VAG_ERR_CUDA = -2,        /* a CUDA runtime call / launch failed */
    VAG_ERR_WORKSPACE = -3,   /* caller workspace too small */
    VAG_ERR_UNSUPPORTED = -4  /* shape outside what the kernels implement */
} vag_status;

/* Arithmetic of the contractions of one call.
 *   VAG_PREC_FP32  FP32-exact (default): every operand is split into FP16 hi/lo planes, three tcgen05 products, FP32-level
 *                  accuracy — the token-exact mode.  (VAG_GEMM=tf32x3 in the environment selects a TF32 hi/lo split instead:
 *                  any FP32 range at half the rate; A/B runs only.)
 *   VAG_PREC_BF16  north_star's bf16 mode: operands rounded to bfloat16, ONE product, FP32 accumulation; the small-shape FFMA
 *                  kernels round their operands the same way.  State, soft-max, attention scores, losses and Adam stay FP32. */
typedef enum { VAG_PREC_FP32 = 0, VAG_PREC_BF16 = 1 } vag_precision;

const char* vag_last_error(void);
int vag_abi_version(void);
/* 1 when the current device is sm_100 (B200); kernels are compiled for sm_100a only. */
int vag_device_supported(void);
/* Number of kernels this library has launched in the process so far (diagnostic; bench.py reports the delta). */
long long vag_launch_count(void);

/* ------------------------------------------------------------------------------------
 * Dense contraction  y = act(x · Wᵀ + bias [+ y])
 * Replaces every nn.Linear / F.linear on the path (layers/NMT_Decoder.py:38-40,127,137,143;
 * layers/VSE_Imagine_Enc.py:57-58,123,138; decoderini V11:118) and the gate GEMMs inside
 * nn.GRU (layers/Encoder.py:58, layers/NMT_Decoder.py:121,129).
 * ---------------------------------------------------------------------------------- */
enum {
    VAG_LIN_TANH = 1,       /* y = tanh(...) */
    VAG_LIN_ACCUMULATE = 2, /* add the previous content of y before the activation */
    VAG_LIN_FORCE_SIMT = 4, /* (internal) never take the tensor-core path */
    VAG_LIN_FORCE_TC = 8,   /* (internal) fail with VAG_ERR_UNSUPPORTED instead of falling to SIMT FP32 */
    VAG_LIN_BF16 = 16       /* this contraction runs in VAG_PREC_BF16 (default VAG_PREC_FP32) */
};
/* x [rows, in_dim] (ld ldx), w [out_dim, in_dim] (ld ldw), bias [out_dim] or NULL,
 * y [rows, out_dim] (ld ldy).  Any shape / alignment; FP32 FFMA kernel. */
int vag_linear_f32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                   const float* bias, int rows, int in_dim, int out_dim, int flags, vag_stream_t stream);

/* The same contraction on the 5th-generation tensor cores (tcgen05.mma kind::f16, TMEM accumulators, TMA-fed):
 * every operand is split into hi = rn_f16(v) and lo = rn_f16((v - hi)·2^11) and hi·hi + (hi·lo + lo·hi)·2^-11 is
 * accumulated in FP32 (two TMEM accumulators), which keeps FP32-level accuracy so that decoded tokens stay exact; with
 * VAG_LIN_BF16 one bfloat16 product instead.  Needs rows >= 64, out_dim >= 64, in_dim >= 32 and a multiple of 8,
 * 16-byte aligned x / w (else VAG_ERR_UNSUPPORTED).  The composites below take this path automatically for eligible
 * shapes; VAG_GEMM=simt in the environment forces the FFMA kernel. */
/* Vocabulary projection of the beam loop WITHOUT an output matrix (replaces NMT_Decoder.py:143 `out` + log_softmax + the
 * topk of V11:300 for the tensor-core path): summ [ceil(out_dim / 32), rows] x 4 floats =
 *   (best logit of the 32-column slice, sum of exp(logit - best), second-best logit, best column | second column << 16)
 * in canonical order (value descending, column ascending; 0xFFFF = none).  Needs rows > 128 and a 16-bit gemm mode. */
int vag_tc_gemm_top2_f32(float* summ, const void* x_hi, const void* x_lo, int64_t ldx, const void* w_hi, const void* w_lo,
                         int64_t ldw, const float* bias, int rows, int in_dim, int out_dim, int precision, vag_stream_t stream);
size_t vag_linear_tc_workspace_bytes(int rows, int in_dim, int out_dim);
int vag_linear_tc_f32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                      const float* bias, int rows, int in_dim, int out_dim, int flags, void* workspace,
                      size_t workspace_bytes, vag_stream_t stream);

/* The two halves of vag_linear_tc_f32 for callers that keep operands split across calls (weights are split once
 * per decode; an activation read by several contractions is split once per step).
 *   vag_tc_elem_bytes(precision): bytes per element of a split plane (2; 4 under VAG_GEMM=tf32x3 in VAG_PREC_FP32).
 *   vag_tc_split_f32 : x [rows, K] → hi / lo planes [rows, ld_out] (ld_out in elements, multiple of 8; planes 128-B aligned;
 *                      VAG_PREC_BF16 writes the hi plane only).
 *   vag_tc_gemm_f32  : y = act(x·Wᵀ + bias [+ y]) from split operands (tcgen05.mma, TMEM accumulators, TMA loads);
 *                      VAG_LIN_BF16 in flags when the planes were made with VAG_PREC_BF16. */
int vag_tc_elem_bytes(int precision);
/* Tools only: device buffer of 32 int64 that the CTA-pair contraction fills with per-role cycle counters (NULL = off). */
int vag_tc_set_debug(void* device_i64x32);
int vag_tc_split_f32(const float* x, int64_t ldx, int rows, int K, void* hi, void* lo, int64_t ld_out, int precision,
                     vag_stream_t stream);
int vag_tc_gemm_f32(float* y, int64_t ldy, const void* x_hi, const void* x_lo, int64_t ldx, const void* w_hi,
                    const void* w_lo, int64_t ldw, const float* bias, int rows, int in_dim, int out_dim, int flags,
                    vag_stream_t stream);

/* out[r, :] = table[ids[r], :]   (nn.Embedding lookups: Encoder.py:50, NMT_Decoder.py:118) */
int vag_embed_rows_f32(float* out, int64_t ldo, const float* table, int dim, const int64_t* ids, int rows,
                       int64_t table_rows, vag_stream_t stream);

/* PyTorch GRU cell gates (order r, z, n) given the two pre-activations
 *   gi = W_ih x + b_ih   [rows, 3H]      gh = W_hh h + b_hh   [rows, 3H]
 *   h' = (1-z) n + z h    with r = σ(gi_r+gh_r), z = σ(gi_z+gh_z), n = tanh(gi_n + r·gh_n)
 * h_out2 (optional) receives a second copy of h' (the encoder writes its context rows with it).
 * Replaces the cuDNN/ATen GRU cell behind Encoder.py:58 and NMT_Decoder.py:121,129. */
int vag_gru_gates_f32(float* h_out, int64_t ld_ho, float* h_out2, int64_t ld_ho2, const float* gi, int64_t ld_gi,
                      const float* gh, int64_t ld_gh, const float* h_prev, int64_t ld_hp, int rows, int H,
                      vag_stream_t stream);

/* Fused attention: score → masked softmax over T → context.
 *   mode VAG_ATTN_MLP : s[n,t] = Σ_c v[c]·tanh(q[n,c] + keys[b,t,c])   BahdanauAttn, NMT_Decoder.py:27-51
 *   mode VAG_ATTN_DOT : s[n,t] = Σ_c q[n,c]·keys[b,t,c]                 ImagineAttn.score_dot, VSE_Imagine_Enc.py:48-65
 *   α[n,:] = softmax_t(s[n,t] where mask[b,t] != 0 else -inf);   c[n,:] = Σ_t α[n,t]·ctx[b,t,:]   (NMT_Decoder.py:126)
 * Row n belongs to sentence b = n / rows_per_sent: the K beams of a sentence share keys/ctx, which are
 * never tiled K times (the reference does, V11:253-254).  keys/ctx are SENTENCE-MAJOR [B, T, C].
 * alpha may be NULL. */
enum { VAG_ATTN_MLP = 0, VAG_ATTN_DOT = 1 };
int vag_attention_f32(float* c_out, int64_t ld_c, float* alpha, const float* q, int64_t ld_q, const float* keys,
                      const float* ctx, const float* v, const float* mask, int rows, int rows_per_sent, int T,
                      int C, int mode, vag_stream_t stream);

/* z[b,:] = split·ctx_vec[b,:] + (1-split)·Σ_t ctx[b,t,:] / Σ_t mask[b,t]     (V11:118,201; V2:85,142)
 * ctx_vec may be NULL (text-only model: z = mean). */
int vag_init_mix_f32(float* z, const float* ctx_vec, const float* ctx, const float* mask, float split, int B,
                     int T, int C, vag_stream_t stream);

/* x[r,:] /= max(‖x[r,:]‖₂, 1e-12)      utils/utils.py:6-10 */
int vag_l2norm_rows_f32(float* x, int64_t ldx, int rows, int dim, vag_stream_t stream);

/* logp = log_softmax(logits) row-wise (NMT_Decoder.py:143).  In place when logp == logits. */
int vag_log_softmax_f32(float* logp, const float* logits, int rows, int V, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Encoder  (LIUMCVC_Encoder.forward, layers/Encoder.py:36-65, eval-mode dropouts)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int E, H;                 /* embedding size, hidden size per direction */
    int precision;            /* vag_precision of every contraction issued on behalf of these weights */
    int64_t vocab;            /* rows of emb */
    const float* emb;         /* encoder.embedding.weight [vocab, E] */
    const float* w_ih[2];     /* encoder.gru.weight_ih_l0{,_reverse} [3H, E] */
    const float* w_hh[2];     /* [3H, H] */
    const float* b_ih[2];     /* [3H] */
    const float* b_hh[2];     /* [3H] */
} vag_encoder_weights;

size_t vag_encoder_workspace_bytes(int B, int T, int E, int H);
/* src int64 [B, T] 0-padded, rows sorted by length descending; lengths_host[B] (HOST, like the reference's
 * python list).  ctx_out [B, T, 2H] sentence-major (exact zeros at pads), mask_out [B, T] (1.0 where src != 0). */
int vag_encoder_fwd_f32(const vag_encoder_weights* w, const int64_t* src, const int32_t* lengths_host, int B, int T,
                        float* ctx_out, float* mask_out, void* workspace, size_t workspace_bytes,
                        vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Visual-attention text pooling  (VSE_Imagine_Enc.forward/get_emb_vec, layers/VSE_Imagine_Enc.py:110-172,
 * ImagineAttn :29-79, l2norm utils/utils.py:6-10)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int I, C, S;              /* image feature size, context size (2H), shared embedding size */
    int method;               /* VAG_ATTN_DOT or VAG_ATTN_MLP (imagine_attn 'dot' / 'mlp') */
    int activation;           /* activation_vse */
    int precision;            /* vag_precision */
    const float* im_w;        /* vse_imagine.im_embedding.weight [S, I] */
    const float* im_b;        /* [S] */
    const float* txt_w;       /* vse_imagine.text_embedding.weight [S, C] */
    const float* txt_b;       /* [S] */
    const float* ctx2ctx_w;   /* vse_imagine.imagine_attn.ctx2ctx.weight [C, C] */
    const float* emb2ctx_w;   /* vse_imagine.imagine_attn.emb2ctx.weight [C, S] */
    const float* mlp_w;       /* vse_imagine.imagine_attn.mlp.weight [C] (method mlp only) */
} vag_vse_weights;

size_t vag_vse_workspace_bytes(int B, int T, int I, int C, int S);
/* im [B, I]; ctx [B, T, C]; mask [B, T] →
 * im_emb [B, S], txt_emb [B, S], ctx_vec [B, C], beta [B, T] (beta may be NULL). */
int vag_vse_pool_fwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                         float* im_emb, float* txt_emb, float* ctx_vec, float* beta, void* workspace,
                         size_t workspace_bytes, vag_stream_t stream);

/* Training pair of the pooling: the forward that keeps what back-propagation needs, and the whole backward chain in one call
 * (what loss.backward() does through VSE_Imagine_Enc.forward, layers/VSE_Imagine_Enc.py:110-152, and ImagineAttn, :29-79).
 *   saved: a_im [B, S] = act(im_embedding(im)) before l2norm, iq [B, C] = emb2ctx(im_emb), pk [B, T, C] = ctx2ctx(ctx),
 *          a_txt [B, S] = act(text_embedding(ctx_vec)) before l2norm — caller-owned, filled by the forward.
 *   backward: d_im_emb / d_txt_emb / d_ctx_vec may be NULL (no gradient arrives there);  g: gradients like the weights (mlp_w only
 *   for the mlp attention), overwritten;  d_ctx [B, T, C] overwritten with the gradient w.r.t. the encoder context. */
typedef struct {
    float *a_im, *iq, *pk, *a_txt;
} vag_vse_saved;
typedef struct {
    float *im_w, *im_b, *txt_w, *txt_b, *ctx2ctx_w, *emb2ctx_w, *mlp_w;
} vag_vse_grads;
int vag_vse_pool_train_fwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                               float* im_emb, float* txt_emb, float* ctx_vec, float* beta, const vag_vse_saved* saved,
                               void* workspace, size_t workspace_bytes, vag_stream_t stream);
size_t vag_vse_pool_bwd_workspace_bytes(int B, int T, int I, int C, int S);
int vag_vse_pool_bwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                         const vag_vse_saved* saved, const float* im_emb, const float* beta, const float* ctx_vec,
                         const float* d_im_emb, const float* d_txt_emb, const float* d_ctx_vec, const vag_vse_grads* g,
                         float* d_ctx, void* workspace, size_t workspace_bytes, vag_stream_t stream);

/* Bidirectional hinge ranking loss, SUM over pairs (losses/PairwiseRankingLoss.py:9-24); with
 * one_direction != 0 only cost_s (losses/ImageRetrievalRankingLoss.py:9-21).
 * im [B, S], s [B, S] → loss_out[1]; grad_im/grad_s ([B, S], may both be NULL) receive dLoss/d·. */
size_t vag_rank_loss_workspace_bytes(int B, int S);
int vag_rank_loss_f32(const float* im, const float* s, int B, int S, float margin, int one_direction,
                      float* loss_out, float* grad_im, float* grad_s, void* workspace, size_t workspace_bytes,
                      vag_stream_t stream);

/* ranks[i] = #{j : q_i·g_j > q_i·g_i} + #{j < i : q_i·g_j == q_i·g_i}   (utils/im_retrieval_eval.py:15-22) */
size_t vag_recall_ranks_workspace_bytes(int n, int S);
int vag_recall_ranks_f32(const float* queries, const float* gallery, int n, int S, int32_t* ranks, void* workspace,
                         size_t workspace_bytes, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Conditional-GRU attention decoder  (NMT_Decoder.forward, layers/NMT_Decoder.py:109-145)
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int E, H, C;              /* embedding, hidden, context (2·encoder H) */
    int precision;            /* vag_precision */
    int64_t V;                /* target vocabulary */
    const float* emb;         /* decoder.embedding.weight [V, E] */
    const float* gru1_w_ih;   /* [3H, E] */
    const float* gru1_w_hh;   /* [3H, H] */
    const float* gru1_b_ih;
    const float* gru1_b_hh;
    const float* attn_h_w;    /* decoder.attn.attn_h.weight [C, H] */
    const float* attn_e_w;    /* decoder.attn.attn_e.weight [C, C] */
    const float* attn_v;      /* decoder.attn.v [C] */
    const float* c2h_w;       /* decoder.context2hid.weight [H, C] */
    const float* gru2_w_ih;   /* [3H, H] */
    const float* gru2_w_hh;   /* [3H, H] */
    const float* gru2_b_ih;
    const float* gru2_b_hh;
    const float* w1_w;        /* [E, H] */
    const float* w1_b;
    const float* w2_w;        /* [E, C] */
    const float* w2_b;
    const float* w3_w;        /* [E, E] */
    const float* w3_b;
    const float* out_w;       /* decoder.out.weight [V, E] (== emb when tied, NMT_Decoder.py:105-106) */
    const float* out_b;       /* [V] */
    const float* ini_w;       /* decoderini.weight [H, C]  (V11:75) */
    const float* ini_b;       /* [H] */
    const void* prepared;     /* optional: buffer filled by vag_decoder_prepare_f32 for exactly these weights and this */
    size_t prepared_bytes;    /* precision (NULL / 0: the decode calls derive the same data into their workspace) */
} vag_decoder_weights;

/* Decode-call invariants of a weight set, computed ONCE instead of once per vag_beam_decode_f32 call: the tensor-core operand
 * planes of every decoder matrix, the K-concatenated read-out matrix [W1 | W3 | W2] with its summed bias, and the per-token
 * table  Emb·W_ihᵀ + b_ih  of gru_1 ([V, 3H]: its input pre-activations depend on the token only).  7.4 GFLOP + 58 MB for
 * EN→DE — negligible against a 1000-sentence call, dominant at the reference's eval batch of 16
 * (nmt_multimodal_beam_DE.py:542-547).  The caller keeps `prepared` alive and re-prepares after the weights change. */
size_t vag_decoder_prepared_bytes(int E, int H, int C, int64_t V);
int vag_decoder_prepare_f32(const vag_decoder_weights* w, void* prepared, size_t prepared_bytes, vag_stream_t stream);

/* Hoisted step-invariant half of the attention MLP: keys[b,t,:] = attn_e(ctx[b,t,:])
 * (the reference recomputes it on the K-times tiled context every step, NMT_Decoder.py:39-40,47). */
size_t vag_attn_keys_workspace_bytes(int B, int T, int C);
int vag_attn_keys_f32(const vag_decoder_weights* w, const float* ctx, int B, int T, float* keys, void* workspace,
                      size_t workspace_bytes, vag_stream_t stream);

/* h0 = tanh(decoderini(split·ctx_vec + (1-split)·mean_t ctx))  (V11:118,201; V2:85,142).
 * workspace: vag_decoder_init_workspace_bytes(); B·C floats suffice for the FP32 FFMA path. */
size_t vag_decoder_init_workspace_bytes(int B, int C, int H);
int vag_decoder_init_f32(const vag_decoder_weights* w, const float* ctx_vec, const float* ctx, const float* mask,
                         float split, int B, int T, float* h0, void* workspace, size_t workspace_bytes,
                         vag_stream_t stream);

size_t vag_decoder_step_workspace_bytes(int rows, int E, int H, int C, int64_t V);
/* One decoder step on `rows` rows (rows_per_sent consecutive rows share a sentence).
 * tokens int64 [rows]; h_prev [rows, H]; keys/ctx [B, T, C]; mask [B, T] →
 * h_out [rows, H]; logits_or_logp [rows, V] (log-softmax applied when want_logp != 0);
 * alpha_out [rows, T] optional (NULL to skip). */
int vag_decoder_step_f32(const vag_decoder_weights* w, const int64_t* tokens, const float* h_prev, const float* keys,
                         const float* ctx, const float* mask, int rows, int rows_per_sent, int T, float* h_out,
                         float* logits_or_logp, int want_logp, float* alpha_out, void* workspace,
                         size_t workspace_bytes, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Beam search  (V11.beamsearch, models/NMT_AttentionImagine_Seq2Seq_Beam_V11.py:233-337;
 * identical text-only copy models/NMT_Seq2Seq_Beam_V2.py:173-276)
 * ---------------------------------------------------------------------------------- */
/* One selection step on materialised log-probabilities (the op the parity tests drive with
 * hand-crafted logp): repeat suppression (V11:279-280), finished-hypothesis masking (:291-294),
 * cand = nll + logp (:297), top-K per sentence (:300) in canonical order (score desc, flat index asc).
 *   step == 0 : logp [B, V], picks top-K of each row;   step > 0 : logp [B·K, V].
 *   prev_tokens int64 [B, K] (tokens chosen at step-1; ignored at step 0)
 *   nll [B, K] in/out;  tokens_out int64 [B, K];  parents_out int32 [B, K] (index of the parent beam). */
int vag_beam_select_f32(const float* logp, int64_t ld_logp, const int64_t* prev_tokens, float* nll,
                        int64_t* tokens_out, int32_t* parents_out, int B, int K, int64_t V, int step,
                        int avoid_double, vag_stream_t stream);

size_t vag_beam_decode_workspace_bytes(int B, int K, int T, int L, int E, int H, int C, int64_t V);
/* The whole loop: L decoder steps on B·K rows + selection + parent reorder, the all-finished early stop
 * evaluated on the device (the reference synchronises with the host every step, V11:266-267), then the
 * epilogue (:315-337): forced final EOS, length normalisation by #tokens>3, best hypothesis per sentence.
 *   h0 [B, H]; ctx/keys [B, T, C]; mask [B, T]
 *   hyp_out int64 [B, L]   tokens of the best hypothesis (full row of the beam, EOS/pad included)
 *   hyp_len int32 [B]      number of tokens before the first EOS (what the reference returns)
 *   beam_out int64 [L, B, K], nll_out [B, K] (un-normalised), steps_out int32[1]: optional (NULL to skip)
 * Early stop.  Once every hypothesis has ended (`done`), every kernel of the remaining steps returns at once, so a finished
 * search costs launch latency only.  host_progress (optional): ONE int32 of pinned, device-mapped HOST memory
 * (cudaHostAlloc; with unified addressing the same pointer is valid on the device).  The device stores
 * (steps finished | done << 30) there after every step and the call — like the reference's per-step `fini_idxs` test,
 * V11:265-269 — keeps at most 4 steps enqueued ahead of it and stops enqueueing when `done` appears.  NULL: all L steps are
 * enqueued without touching the host (the form a CUDA graph can capture). */
int vag_beam_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                        const float* mask, int B, int K, int T, int L, int avoid_double, int64_t* hyp_out,
                        int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out,
                        volatile int32_t* host_progress, void* workspace, size_t workspace_bytes, vag_stream_t stream);

/* Steps [step_begin, step_end) of the same search (V11:255-313), on the state a previous range left in `workspace` (same buffer, same
 * shape arguments): step_begin == 0 initialises it, step_end >= L appends the epilogue and writes the outputs
 * (step_begin == step_end == L: the epilogue alone).  done_host (optional): one int32 of PINNED host memory that receives the `done`
 * flag as it stands after the range, asynchronously and in stream order.  No host polling inside the call, so every range can be
 * captured in a CUDA graph: a caller replays chunk graphs one ahead of an event and stops launching once `done` has arrived — the
 * reference's per-step `fini_idxs` test (V11:265-269) at chunk granularity. */
int vag_beam_decode_steps_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                              const float* mask, int B, int K, int T, int L, int avoid_double, int step_begin, int step_end,
                              int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, float* nll_out, int32_t* steps_out,
                              int32_t* done_host, void* workspace, size_t workspace_bytes, vag_stream_t stream);

/* The epilogue of the search alone (V11:315-337), for callers that drive the steps themselves (vag_decoder_step_f32 +
 * vag_beam_select_f32): tok_hist int64 / par_hist int32 [L, B, K] = token and parent-beam index chosen at every step,
 * nll [B, K], steps_run int32[1] ON THE DEVICE (rows >= steps_run count as 0; the last row is forced to <eos>, :315).
 * → hyp_out int64 [B, L], hyp_len int32 [B], beam_out int64 [L, B, K] (optional, the back-traced `beam` of the reference). */
int vag_beam_finalize_f32(const int64_t* tok_hist, const int32_t* par_hist, const float* nll, const int32_t* steps_run, int B,
                          int K, int L, int64_t* hyp_out, int32_t* hyp_len, int64_t* beam_out, vag_stream_t stream);

/* Greedy branch (beam_size == 1, V11:207-226): tokens_out int64 [B, L] = argmax at every step. */
int vag_greedy_decode_f32(const vag_decoder_weights* w, const float* h0, const float* keys, const float* ctx,
                          const float* mask, int B, int T, int L, int64_t* tokens_out, void* workspace,
                          size_t workspace_bytes, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Training-side forward pieces (V11.forward :136-164)
 * ---------------------------------------------------------------------------------- */
/* loss_rows[r] += -weight[tgt[r]] · logp[r, tgt[r]]  with logp = log_softmax(logits) computed on the fly
 * (nn.NLLLoss(weight, reduce=False), nmt_multimodal_beam_DE.py:286-291).  weight may be NULL (all ones).
 * lse_out [rows] optional: row log-sum-exp kept for the backward pass. */
int vag_nll_rows_f32(const float* logits, int64_t ld, const int64_t* tgt, const float* weight, int rows, int64_t V,
                     float* loss_rows, float* lse_out, vag_stream_t stream);

/* out[r] = argmax_v logits[r, v] (ties → lowest index): the free-running branch of forward (V11:157-160)
 * and greedy decoding (V11:211). */
int vag_row_argmax_f32(const float* logits, int64_t ld, int rows, int64_t V, int64_t* out, vag_stream_t stream);

/* Loss epilogue of forward (V11:164-166 ; NMT_Seq2Seq_Beam_V2.py:111):
 *   loss_mt = mean_b( loss_rows[b] / #{t : tgt[b,t] != 0} ),   loss = loss_w·loss_mt + (1-loss_w)·loss_vse
 * loss_vse is a device scalar or NULL (text-only model: loss = loss_mt).  out[3] = {loss, loss_mt, loss_vse}. */
int vag_translation_loss_f32(const float* loss_rows, const int64_t* tgt, int B, int Tt, const float* loss_vse,
                             float loss_w, float* out, vag_stream_t stream);
/* Its backward in one launch (what autograd derives from V11:164-166): given g[3] = d/d{loss, loss_mt, loss_vse},
 *   g_rows[b] = (g[0]·w + g[1]) / (B · #non-pad_b)   with w = loss_w (has_vse) or 1,   g_vse[0] = g[0]·(1-loss_w) + g[2]. */
int vag_translation_loss_bwd_f32(const float* g, const int64_t* tgt, int B, int Tt, float loss_w, int has_vse,
                                 float* g_rows, float* g_vse, vag_stream_t stream);

/* Source padding mask and sentence lengths, one launch (models/Encoder.py:47 `src_mask = (input_var != 0)`; the lengths are
 * what prepare_batch hands to pack_padded_sequence, Encoder.py:55).  Either output may be NULL. */
int vag_src_mask_lengths(const int64_t* src, int B, int T, float* mask, int32_t* lengths, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Backward pieces of the training step (train.py:36-51: forward, loss.backward(), clip_grad_norm_, Adam.step()).
 * The reference gets all of these from autograd + cuDNN; here each is one kernel and the Python autograd
 * Functions of vag_nmt_b200/autograd.py sequence them (BPTT over the Tt decoder / Ts encoder steps).
 * ---------------------------------------------------------------------------------- */
/* C[m,n] = alpha·Σ_k A[m·sam + k·sak]·B[k·sbk + n·sbn] + beta·C[m,n]: dX = dY·W and dW = dYᵀ·X of every nn.Linear / GRU matrix. */
int vag_gemm_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                 int64_t sbn, int M, int N, int K, float alpha, float beta, int precision, vag_stream_t stream);

/* vag_gemm_f32 with a caller-owned workspace (vag_gemm_tc_workspace_bytes): contractions with M, N >= 64 run on the tcgen05
   path — operands are split into tensor-core planes inside the workspace, transposing those that are not contraction-
   contiguous — everything else falls through to vag_gemm_f32.  Replaces the torch.mm calls autograd issues for the
   weight / input gradients of every nn.Linear and nn.GRU on the path (layers/NMT_Decoder.py:121-143 under backward()). */
size_t vag_gemm_tc_workspace_bytes(int M, int N, int K);
int vag_gemm_tc_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                    int64_t sbn, int M, int N, int K, float alpha, float beta, int precision, void* workspace,
                    size_t workspace_bytes, vag_stream_t stream);
/* GRU cell backward from the saved pre-activations: dgi, dgh [rows,3H] and dh_prev [rows,H] = dh·z (the caller adds dgh·W_hh). */
int vag_gru_gates_bwd_f32(float* dgi, float* dgh, float* dh_prev, const float* dh, int64_t ld_dh, const float* gi,
                          const float* gh, const float* h_prev, int64_t ld_hp, int rows, int H, vag_stream_t stream);
/* Backward of vag_attention_f32 with one row per sentence: dq [B,C]; dkeys, dctx [B,T,C] and dv [C] are ACCUMULATED
 * (they collect contributions from every decoder step); dctx / dv may be NULL. */
int vag_attention_bwd_f32(float* dq, int64_t ld_dq, float* dkeys, float* dctx, float* dv, const float* dc, int64_t ld_dc,
                          const float* alpha, const float* q, int64_t ld_q, const float* keys, const float* ctx,
                          const float* v, const float* mask, int B, int T, int C, int mode, vag_stream_t stream);
/* dlogits[r,v] = grad_rows[r]·weight[tgt[r]]·(softmax(logits)[r,v] − [v == tgt[r]])   (backward of vag_nll_rows_f32) */
int vag_nll_bwd_f32(float* dlogits, int64_t ldd, const float* logits, int64_t ld, const float* lse, const int64_t* tgt,
                    const float* weight, const float* grad_rows, int rows, int64_t V, vag_stream_t stream);
int vag_tanh_bwd_f32(float* dx, const float* dy, const float* y, int64_t n, vag_stream_t stream);          /* dx = dy·(1−y²) */
int vag_axpby_f32(float* y, const float* x, float a, float b, int64_t n, vag_stream_t stream);             /* y = a·x + b·y */
int vag_mul_f32(float* y, const float* m, int64_t n, vag_stream_t stream);                                /* y *= m (dropout masks) */
int vag_colsum_f32(float* out, const float* x, int64_t ldx, int rows, int cols, int accumulate, vag_stream_t stream); /* bias grads */
int vag_embed_bwd_f32(float* table_grad, const float* g, int64_t ldg, const int64_t* ids, int rows, int dim,
                      int64_t table_rows, vag_stream_t stream);                                             /* scatter-add */
int vag_l2norm_bwd_f32(float* dx, const float* dy, const float* x, int rows, int dim, vag_stream_t stream);
int vag_init_mix_bwd_f32(float* dctx_vec, float* dctx, const float* dz, const float* mask, float split, int B, int T,
                         int C, vag_stream_t stream);
/* The whole teacher-forced / free-running decoder loop of V11.forward (:136-160) with the activations its backward
 * needs, and the matching back-propagation through time — one call each.  All buffers are caller-owned. */
typedef struct {
    int64_t ld_logits;        /* row pitch of logits_all (>= V, multiple of 4) */
    float* keys;              /* [B, T, C]   attn_e(ctx) */
    float* e_all;             /* [Tt·B, E]   input embeddings            */
    float* gi1_all;           /* [Tt, B, 3H] GRU1 input pre-activations  */
    float* gh1_all;           /* [Tt, B, 3H] GRU1 hidden pre-activations */
    float* h1_all;            /* [Tt, B, H]  */
    float* q_all;             /* [Tt, B, C]  attn_h(h1) */
    float* alpha_all;         /* [Tt, B, T]  attention weights */
    float* c_all;             /* [Tt, B, C]  contexts */
    float* x2_all;            /* [Tt, B, H]  context2hid(c) */
    float* gi2_all;           /* [Tt, B, 3H] */
    float* gh2_all;           /* [Tt, B, 3H] */
    float* h2_all;            /* [Tt, B, H]  */
    float* t_all;             /* [Tt·B, E]   tanh read-out */
    float* logits_all;        /* [Tt·B, ld_logits] */
    float* lse_all;           /* [Tt, B]     row log-sum-exp */
} vag_decoder_seq_saved;

typedef struct {              /* gradient outputs, same shapes as the weights (out_w unused when tied) */
    float *emb, *gru1_w_ih, *gru1_w_hh, *gru1_b_ih, *gru1_b_hh, *attn_h_w, *attn_e_w, *attn_v, *c2h_w, *gru2_w_ih,
          *gru2_w_hh, *gru2_b_ih, *gru2_b_hh, *w1_w, *w1_b, *w2_w, *w2_b, *w3_w, *w3_b, *out_w, *out_b;
} vag_decoder_grads;

size_t vag_decoder_seq_workspace_bytes(int B, int T, int Tt, int E, int H, int C, int64_t V);
/* vag_decoder_seq_bwd_f32 `phases` — a mask, 15 = everything in one call:
 *   1  head: d logits, d read-out (all steps at once)            2  the recurrent part + the path back into the encoder context
 *   4  weight gradients of the vocabulary projection / read-out (need 1)      8  all other weight gradients (need 2)
 * All phases of one backward pass share one workspace, untouched in between.  Nothing downstream of the decoder reads a weight
 * gradient before the optimiser, so 4 and 8 may run on a second stream (they use private operand-plane scratch when called
 * without 1 / 2): 4 beside the latency-bound recurrent part, 8 beside the encoder's back-propagation.
 * tok_in int64 [Tt, B]: row 0 = <sos>; teacher != 0: rows 1.. hold tgt[:, :-1] (caller fills); else the kernel's own
 * arg-max feedback is written there.  tgt_t int64 [Tt, B].  loss_rows [B] = Σ_t NLL (nn.NLLLoss(weight, reduce=False)).
 * out_mask (optional, [Tt·B, E], values 0 or 1/(1-p)): the output dropout of NMT_Decoder.py:140-141 with a caller-drawn mask. */
int vag_decoder_seq_fwd_f32(const vag_decoder_weights* w, const float* h0, const float* enc, const float* mask,
                            int64_t* tok_in, const int64_t* tgt_t, const float* nll_weight, int B, int T, int Tt,
                            int teacher, const vag_decoder_seq_saved* s, const float* out_mask, float* loss_rows,
                            void* workspace, size_t workspace_bytes, vag_stream_t stream);
int vag_decoder_seq_bwd_f32(const vag_decoder_weights* w, const float* h0, const float* enc, const float* mask,
                            const int64_t* tok_in, const int64_t* tgt_t, const float* nll_weight, int B, int T, int Tt,
                            int tied, const vag_decoder_seq_saved* s, const float* out_mask, const float* dloss_rows,
                            const vag_decoder_grads* g, float* d_h0, float* d_enc, int phases, void* workspace,
                            size_t workspace_bytes, vag_stream_t stream);

/* Encoder training pair: forward that keeps what BPTT needs and the backward through both directions of the packed
 * bi-GRU (layers/Encoder.py:36-65 under autograd).  x [T·B, E] time-major embeddings, ids_tm int64 [T·B], gi / gh
 * [2][T, B, 3H]; gradients like the weights.  emb_mask (optional, [T·B, E] time-major, 0 or 1/(1-p)): the embedding
 * dropout of Encoder.py:51-52.
 * lengths_dev: int32 [B] ON THE DEVICE (required).  Every recurrent step runs on all B rows, both directions in one
 * launch, and rows whose sentence has ended are masked on the device, so the launch sequence depends on (B, T) only —
 * a CUDA graph captured around the training step stays valid for any batch of that shape.  lengths_host (optional, may
 * be NULL) is only validated (1 <= length <= T, sorted in decreasing order like pack_padded_sequence demands). */
size_t vag_encoder_train_workspace_bytes(int B, int T, int E, int H);
int vag_encoder_train_fwd_f32(const vag_encoder_weights* w, const int64_t* src, const int32_t* lengths_host,
                              const int32_t* lengths_dev, int B, int T, float* ctx_out, float* x, int64_t* ids_tm, float* gi,
                              float* gh, const float* emb_mask, void* workspace, size_t workspace_bytes, vag_stream_t stream);
int vag_encoder_bwd_f32(const vag_encoder_weights* w, const int32_t* lengths_host, const int32_t* lengths_dev, int B, int T,
                        const float* ctx, const float* dctx, const float* x, const int64_t* ids_tm, const float* gi,
                        const float* gh, float* d_emb, float* const* d_w_ih, float* const* d_w_hh, float* const* d_b_ih,
                        float* const* d_b_hh, const float* emb_mask, void* workspace, size_t workspace_bytes,
                        vag_stream_t stream);

/* Optimiser step of train.py:46-49 with the Adam of nmt_multimodal_beam_DE.py:303-332, fused per parameter tensor:
 *   accum[0] += Σ grad²  (vag_sumsq_f32 over every tensor, after the gradient all-reduce when data-parallel), then
 *   g ← grad·min(1, clip/(√accum + 1e-6)) [+ weight_decay·param];  Adam(m, v, step) update of param in place. */
int vag_sumsq_f32(const float* g, int64_t n, float* accum, vag_stream_t stream);
/* Multi-tensor flavours: ONE launch each over all parameter tensors (a device array of descriptors, 48 bytes each);
 * max_n = the largest tensor's element count (sizes the grid).  Same arithmetic as the per-tensor entry points. */
typedef struct vag_optim_tensor {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t n;
    float weight_decay; /* added to the gradient (torch.optim.Adam), 0 for biases (nmt_multimodal_beam_DE.py:303-312) */
    float lr;
} vag_optim_tensor;
int vag_sumsq_multi_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, float* accum, vag_stream_t stream);
/* Deterministic Σ‖g‖² (no float atomics): out[0] = the sum, overwritten.  partials: scratch of vag_sumsq_multi_partials()
 * floats.  Bit-identical from run to run and across data-parallel replicas, so the replicas stay in lockstep. */
size_t vag_sumsq_multi_partials(int n_tensors, int64_t max_n);
int vag_sumsq_multi_det_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, float* out, float* partials,
                            size_t n_partials, vag_stream_t stream);
int vag_clip_adam_multi_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, const float* grad_sumsq,
                            float clip, float beta1, float beta2, float eps, int step, vag_stream_t stream);
int vag_clip_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      const float* grad_sumsq, float clip, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, vag_stream_t stream);

/* ------------------------------------------------------------------------------------
 * Data-parallel gradient exchange over NVLink peer memory (one process per GPU on one node).  Replaces the NCCL all-reduce +
 * the norm pass of the step driver (train.py:36-51 under the commented nn.DataParallel of nmt_multimodal_beam_DE.py:277-282).
 * Every rank keeps its flat gradient buffer at the START of a peer-visible arena:
 *   vag_p2p_alloc (cudaMalloc + CUDA IPC handle; the one allocating entry point of this ABI) → exchange the 64-byte handles by any
 *   host channel → vag_p2p_open on every peer's handle → vag_dp_comm{world, rank, peers[]} (own arena at peers[rank]).
 * vag_dp_allreduce_f32 (call number `step` = 0, 1, 2 … — the same on all ranks — on every rank, once per optimisation step):
 *   barrier · reduce-scatter (rank r averages slice r over all ranks, fixed summation order) + Σ avg² block partials ·
 *   barrier · all-gather of the averaged slices + the norm (fixed order) · barrier.
 * Afterwards every rank's buffer holds the same bits (the average) and sumsq_out[0] the same Σ‖g‖² — what the replicated
 * vag_clip_adam_multi_f32 needs to keep the replicas bit-identical.  n_floats: a multiple of 4; the arena must hold
 * vag_dp_arena_bytes(n_floats).  A rank that never arrives makes the others trap after a bounded wait.
 * ---------------------------------------------------------------------------------- */
typedef struct {
    int world, rank;
    void* peers[16];          /* arena base pointer of every rank as mapped INTO THIS process (vag_p2p_open; own at [rank]) */
} vag_dp_comm;
int vag_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out_64);
int vag_p2p_free(void* dev_ptr);
int vag_p2p_open(const unsigned char* handle_64, void** dev_ptr);
int vag_p2p_close(void* dev_ptr);
size_t vag_dp_arena_bytes(int64_t n_floats);
int vag_dp_allreduce_f32(const vag_dp_comm* c, int64_t n_floats, int step, float* sumsq_out, vag_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VAG_NMT_H */
