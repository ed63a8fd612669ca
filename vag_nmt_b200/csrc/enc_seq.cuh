// Persistent recurrent kernels of the encoder's training step (enc_seq.cu): one launch walks all T steps of both directions.
#pragma once
#include "common.cuh"

namespace vag {

// Forward time loop.  gi / gh: [2 directions][T][B][3H] (gate order r | z | n; gi = x·W_ihᵀ + b_ih for all steps, computed before;
// gh = h·W_hhᵀ + b_hh is produced here and kept for the backward pass, zeros for masked rows).  h: [2 parities][2][B][H] scratch.
// ctx_out [B][T][2H] must be zeroed by the caller (masked positions are not written).
struct EncSeqFwd {
    const float* w_hh[2];
    const float* b_hh[2];
    const float* gi;
    float* gh;
    float* ctx_out;
    const int32_t* lengths;
    int B, T, H;
    int* bar;    // set by the launcher: flags and exchange buffer inside the scratch area, warps' contraction range
    float* xch;
    int kw;
};
// Back-propagation through time.  dgi / dgh / hprev_all as vag_encoder_bwd_f32 lays them out ([2][T][B][3H], [2][T][B][H]).
struct EncSeqBwd {
    const float* w_hh[2];
    const float* gi;
    const float* gh;
    const float* ctx;
    const float* dctx;
    float* dgi;
    float* dgh;
    float* hprev_all;
    const int32_t* lengths;
    int B, T, H;
    int* bar;    // set by the launcher
    float* xch;
    int kw, n_chunk;
};
size_t enc_seq_scratch_bytes(int H);   // device scratch both launchers need (zeroed by them)
bool enc_seq_fwd_ok(int B, int T, int H);
bool enc_seq_bwd_ok(int B, int T, int H);
int enc_seq_fwd(EncSeqFwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st);
int enc_seq_bwd(EncSeqBwd a, void* scratch, size_t scratch_bytes, bool round_bf16, cudaStream_t st);

}  // namespace vag
