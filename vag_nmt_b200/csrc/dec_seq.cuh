// Persistent forward time loop of the decoder's training step (dec_seq.cu): one launch walks all Tt teacher-forced steps.
#pragma once
#include "common.cuh"

namespace vag {

// Inputs: the decoder's recurrent matrices (fp32, row-major as torch.nn.GRU / nn.Linear store them), h0 [B, H], the hoisted
// attention keys [B, T, C], the encoder context enc [B, T, C], mask [B, T], gi1_all [Tt, B, 3H] = Emb(tok)·W_ih1ᵀ + b_ih1 for all
// steps.  Outputs, laid out as vag_decoder_seq_saved: gh1_all / gi2_all / gh2_all [Tt, B, 3H] (bias included), h1_all / x2_all /
// h2_all [Tt, B, H], q_all / c_all [Tt, B, C], alpha_all [Tt, B, T].
struct DecSeqFwd {
    const float *gru1_w_hh, *gru1_b_hh, *attn_h_w, *attn_v, *c2h_w, *gru2_w_ih, *gru2_w_hh, *gru2_b_ih, *gru2_b_hh;
    const float *h0, *keys, *enc, *mask, *gi1_all;
    float *gh1_all, *h1_all, *q_all, *alpha_all, *c_all, *x2_all, *gi2_all, *gh2_all, *h2_all;
    int B, T, Tt, H, C;
    int* bar;      // set by the launcher
    float* xch;
    int kw_h;
};
size_t dec_seq_scratch_bytes(int H, int C);
bool dec_seq_fwd_ok(int B, int T, int Tt, int H, int C);
int dec_seq_fwd(DecSeqFwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st);

}  // namespace vag
