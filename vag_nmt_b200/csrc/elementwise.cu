// Bandwidth-bound fused element-wise / row-wise kernels: GRU gates, decoder-init mix, l2norm,
// log-softmax, NLL rows.  All FP32, vectorised where the strides allow it.
#include "common.cuh"
#include "split.cuh"
#include <math.h>

namespace vag {

// ---------------------------------------------------------------- GRU gates
// One thread per (row, 4 hidden units).  Reads gi/gh (3 gates each) + h_prev, writes h' (twice at most).
template <bool VEC>
__global__ void __launch_bounds__(256)
gru_gates_kernel(float* h_out, int64_t ld_ho, float* h_out2, int64_t ld_ho2,
                 const float* __restrict__ gi, int64_t ld_gi, const float* __restrict__ gh, int64_t ld_gh,
                 const float* h_prev /* may alias h_out */, int64_t ld_hp, int rows, int H, SplitDst sd,
                 const int64_t* __restrict__ gi_rows = nullptr, int64_t gi_n_rows = 0, const int* __restrict__ done = nullptr) {
    pdl_trigger();   // the contraction that follows may start its prologue while this kernel drains
    if (done && *reinterpret_cast<const volatile int*>(done)) return;   // beam search over: the state is never read again
    constexpr int W = VEC ? 4 : 1;
    const int per_row = H / W;
    const int64_t total = (int64_t)rows * per_row;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(idx / per_row);
        const int j = (int)(idx % per_row) * W;
        int64_t gi_row = row;
        if (gi_rows) {   // input pre-activations looked up in a per-token table (the row's token id; out of range → 0 like the embedding)
            gi_row = gi_rows[row];
            if (gi_row < 0 || gi_row >= gi_n_rows) gi_row = 0;
        }
        const float* gir = gi + gi_row * ld_gi;
        const float* ghr = gh + (int64_t)row * ld_gh;
        float ir[W], iz[W], in_[W], hr[W], hz[W], hn[W], hp[W], out[W];
        if (VEC) {
            *reinterpret_cast<float4*>(ir) = *reinterpret_cast<const float4*>(gir + j);
            *reinterpret_cast<float4*>(iz) = *reinterpret_cast<const float4*>(gir + H + j);
            *reinterpret_cast<float4*>(in_) = *reinterpret_cast<const float4*>(gir + 2 * H + j);
            *reinterpret_cast<float4*>(hr) = *reinterpret_cast<const float4*>(ghr + j);
            *reinterpret_cast<float4*>(hz) = *reinterpret_cast<const float4*>(ghr + H + j);
            *reinterpret_cast<float4*>(hn) = *reinterpret_cast<const float4*>(ghr + 2 * H + j);
            *reinterpret_cast<float4*>(hp) = *reinterpret_cast<const float4*>(h_prev + (int64_t)row * ld_hp + j);
        } else {
            ir[0] = gir[j]; iz[0] = gir[H + j]; in_[0] = gir[2 * H + j];
            hr[0] = ghr[j]; hz[0] = ghr[H + j]; hn[0] = ghr[2 * H + j];
            hp[0] = h_prev[(int64_t)row * ld_hp + j];
        }
#pragma unroll
        for (int u = 0; u < W; ++u) {
            const float r = sigmoidf_precise(ir[u] + hr[u]);
            const float z = sigmoidf_precise(iz[u] + hz[u]);
            const float n = tanhf(in_[u] + r * hn[u]);
            out[u] = (1.0f - z) * n + z * hp[u];
        }
        if (VEC) {
            *reinterpret_cast<float4*>(h_out + (int64_t)row * ld_ho + j) = *reinterpret_cast<float4*>(out);
            if (h_out2) *reinterpret_cast<float4*>(h_out2 + (int64_t)row * ld_ho2 + j) = *reinterpret_cast<float4*>(out);
            if (sd.hi) split_store4(sd, row, j, *reinterpret_cast<float4*>(out));   // tensor-core operand planes of the new state
        } else {
            h_out[(int64_t)row * ld_ho + j] = out[0];
            if (h_out2) h_out2[(int64_t)row * ld_ho2 + j] = out[0];
        }
    }
}

// ---------------------------------------------------------------- decoder-init mix
// z[b,c] = split*ctx_vec[b,c] + (1-split) * (Σ_t ctx[b,t,c]) / (Σ_t mask[b,t])
__global__ void init_mix_kernel(float* __restrict__ z, const float* __restrict__ ctx_vec, const float* __restrict__ ctx,
                                const float* __restrict__ mask, float split, int B, int T, int C) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float msum = 0.f;
    for (int t = 0; t < T; ++t) msum += mask[(int64_t)b * T + t];
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += ctx[((int64_t)b * T + t) * C + c];
    const float mean = s / msum;
    z[(int64_t)b * C + c] = ctx_vec ? split * ctx_vec[(int64_t)b * C + c] + (1.0f - split) * mean : mean;
}

// ---------------------------------------------------------------- l2norm rows (one warp per row)
__global__ void l2norm_rows_kernel(float* __restrict__ x, int64_t ldx, int rows, int dim) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* p = x + (int64_t)row * ldx;
    float ss = 0.f;
    for (int c = lane; c < dim; c += 32) ss = fmaf(p[c], p[c], ss);
    ss = warp_sum(ss);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    for (int c = lane; c < dim; c += 32) p[c] = p[c] / nrm;
}

// ---------------------------------------------------------------- row log-sum-exp helpers
struct MaxSum {
    float m, s;
};
__device__ __forceinline__ MaxSum ms_combine(MaxSum a, MaxSum b) {
    MaxSum r;
    r.m = fmaxf(a.m, b.m);
    if (r.m == -INFINITY) { r.s = 0.f; return r; }
    r.s = a.s * expf(a.m - r.m) + b.s * expf(b.m - r.m);
    return r;
}

// Block-wide (max, Σexp(x-max)) of one row; two passes over the row (max first, then the sum) so that the
// result follows log_softmax's own arithmetic: lse = max + log Σ exp(x - max).
__device__ float block_row_lse(const float* __restrict__ row, int64_t V, float* red /* >= 33 floats */) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
    float m = -INFINITY;
    for (int64_t i = tid; i < V; i += blockDim.x) m = fmaxf(m, row[i]);
    m = warp_max(m);
    if (lane == 0) red[wid] = m;
    __syncthreads();
    if (wid == 0) {
        float v = lane < nw ? red[lane] : -INFINITY;
        v = warp_max(v);
        if (lane == 0) red[32] = v;
    }
    __syncthreads();
    m = red[32];
    __syncthreads();
    float s = 0.f;
    for (int64_t i = tid; i < V; i += blockDim.x) s += expf(row[i] - m);
    s = warp_sum(s);
    if (lane == 0) red[wid] = s;
    __syncthreads();
    if (wid == 0) {
        float v = lane < nw ? red[lane] : 0.f;
        v = warp_sum(v);
        if (lane == 0) red[32] = v;
    }
    __syncthreads();
    s = red[32];
    __syncthreads();
    return m + logf(s);
}

__global__ void __launch_bounds__(256) log_softmax_kernel(float* __restrict__ logp, const float* __restrict__ logits, int64_t V) {
    __shared__ float red[33];
    const float* src = logits + (int64_t)blockIdx.x * V;
    float* dst = logp + (int64_t)blockIdx.x * V;
    const float lse = block_row_lse(src, V, red);
    for (int64_t i = threadIdx.x; i < V; i += blockDim.x) dst[i] = src[i] - lse;
}

__global__ void __launch_bounds__(256) row_lse_kernel(float* __restrict__ lse_out, const float* __restrict__ logits, int64_t ld, int64_t V) {
    __shared__ float red[33];
    const float lse = block_row_lse(logits + (int64_t)blockIdx.x * ld, V, red);
    if (threadIdx.x == 0) lse_out[blockIdx.x] = lse;
}

__global__ void __launch_bounds__(256)
nll_rows_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ tgt, const float* __restrict__ weight,
                int64_t V, float* __restrict__ loss_rows, float* __restrict__ lse_out) {
    __shared__ float red[33];
    const int r = blockIdx.x;
    const float* row = logits + (int64_t)r * ld;
    const float lse = block_row_lse(row, V, red);
    if (threadIdx.x == 0) {
        int64_t t = tgt[r];
        if (t < 0 || t >= V) t = 0;
        const float wgt = weight ? weight[t] : 1.0f;
        loss_rows[r] += -(row[t] - lse) * wgt;
        if (lse_out) lse_out[r] = lse;
    }
}

// All Tt·B rows of a teacher-forced pass in one launch: per-row NLL into nll_out (row = t·B + b), then nll_sum_steps_kernel
// adds the steps of a sentence in time order (the same summation order as one launch per step).
__global__ void __launch_bounds__(256)
nll_rows_all_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ tgt, const float* __restrict__ weight,
                    int64_t V, float* __restrict__ nll_out, float* __restrict__ lse_out) {
    __shared__ float red[33];
    const int r = blockIdx.x;
    const float* row = logits + (int64_t)r * ld;
    const float lse = block_row_lse(row, V, red);
    if (threadIdx.x == 0) {
        int64_t t = tgt[r];
        if (t < 0 || t >= V) t = 0;
        const float wgt = weight ? weight[t] : 1.0f;
        nll_out[r] = -(row[t] - lse) * wgt;
        lse_out[r] = lse;
    }
}
__global__ void nll_sum_steps_kernel(float* __restrict__ loss_rows, const float* __restrict__ nll, int B, int Tt) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    float a = 0.f;
    for (int t = 0; t < Tt; ++t) a += nll[(int64_t)t * B + b];
    loss_rows[b] = a;
}
int nll_rows_all(const float* logits, int64_t ld, const int64_t* tgt, const float* weight, int B, int Tt, int64_t V, float* loss_rows,
                 float* lse_out, float* nll_scratch, cudaStream_t st) {
    nll_rows_all_kernel<<<B * Tt, 256, 0, st>>>(logits, ld, tgt, weight, V, nll_scratch, lse_out);
    VAG_LAUNCH_CHECK();
    nll_sum_steps_kernel<<<ceil_div(B, 128), 128, 0, st>>>(loss_rows, nll_scratch, B, Tt);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// out = {loss, loss_mt, loss_vse};  one block, deterministic order
__global__ void __launch_bounds__(256)
translation_loss_kernel(const float* __restrict__ loss_rows, const int64_t* __restrict__ tgt, int B, int Tt,
                        const float* __restrict__ loss_vse, float loss_w, float* __restrict__ out) {
    __shared__ float red[8];
    float a = 0.f;
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        float cnt = 0.f;
        for (int t = 0; t < Tt; ++t) cnt += (tgt[(int64_t)b * Tt + t] != 0) ? 1.f : 0.f;
        a += loss_rows[b] / cnt;
    }
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        const float mt = t / (float)B;
        const float vse = loss_vse ? loss_vse[0] : 0.f;
        out[0] = loss_vse ? loss_w * mt + (1.0f - loss_w) * vse : mt;
        out[1] = mt;
        out[2] = vse;
    }
}

// Backward of the mix above, one launch:  d loss_rows[b] = (g_loss·w + g_mt) / (B·count_b),  d loss_vse = g_loss·(1-w) + g_vse.
__global__ void __launch_bounds__(256)
translation_loss_bwd_kernel(const float* __restrict__ g, const int64_t* __restrict__ tgt, int B, int Tt, float loss_w,
                            int has_vse, float* __restrict__ g_rows, float* __restrict__ g_vse) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const float g0 = g[0], g1 = g[1], g2 = g[2];
    if (b == 0 && g_vse) g_vse[0] = g0 * (1.0f - loss_w) + g2;
    if (b >= B) return;
    float cnt = 0.f;
    for (int t = 0; t < Tt; ++t) cnt += (tgt[(int64_t)b * Tt + t] != 0) ? 1.f : 0.f;
    g_rows[b] = (g0 * (has_vse ? loss_w : 1.0f) + g1) / ((float)B * cnt);
}

// mask[b,t] = (src[b,t] != 0), lengths[b] = number of non-pad tokens  (Encoder.py:47); one warp per sentence
__global__ void __launch_bounds__(256)
src_mask_lengths_kernel(const int64_t* __restrict__ src, int B, int T, float* __restrict__ mask, int32_t* __restrict__ lengths) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= B) return;
    int n = 0;
    for (int t = lane; t < T; t += 32) {
        const bool live = src[(int64_t)b * T + t] != 0;
        if (mask) mask[(int64_t)b * T + t] = live ? 1.f : 0.f;
        n += live ? 1 : 0;
    }
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0 && lengths) lengths[b] = n;
}

__global__ void bias_sum3_kernel(float* __restrict__ out, const float* __restrict__ a, const float* __restrict__ b,
                                 const float* __restrict__ c, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = ((a ? a[i] : 0.f) + (b ? b[i] : 0.f)) + (c ? c[i] : 0.f);
}
int bias_sum3(float* out, const float* a, const float* b, const float* c, int n, cudaStream_t st) {
    bias_sum3_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, a, b, c, n);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int row_lse(float* lse_out, const float* logits, int64_t ld, int rows, int64_t V, cudaStream_t st) {
    if (rows == 0) return VAG_OK;
    row_lse_kernel<<<rows, 256, 0, st>>>(lse_out, logits, ld, V);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// Fused-step flavour of the gates: also writes the operand planes of the new state (all pitches multiples of 4, 16-byte
// aligned pointers — guaranteed by the caller's workspace layout).
int gru_gates_split(float* h_out, int64_t ld_ho, const float* gi, int64_t ld_gi, const float* gh, int64_t ld_gh,
                    const float* h_prev, int64_t ld_hp, int rows, int H, SplitDst sd, cudaStream_t st, const int64_t* gi_rows,
                    int64_t gi_n_rows, const int* done) {
    if (rows == 0) return VAG_OK;
    const int64_t total = (int64_t)rows * (H / 4);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
    gru_gates_kernel<true><<<blocks, 256, 0, st>>>(h_out, ld_ho, nullptr, 0, gi, ld_gi, gh, ld_gh, h_prev, ld_hp, rows, H, sd, gi_rows, gi_n_rows, done);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// dst[row, 0:E) (both planes) = split planes of table[tokens[row], :] — the embedding gather reads the ALREADY SPLIT
// table (the tied output projection keeps one for the vocabulary contraction), so no fp32 row is ever materialised.
__global__ void __launch_bounds__(256)
embed_split_rows_kernel(SplitDst dst, const uint16_t* __restrict__ t_hi, const uint16_t* __restrict__ t_lo, int64_t ld_t, int E,
                        const int64_t* __restrict__ tokens, int rows, int64_t V) {
    const int per_row = E / 8;   // 16-byte chunks per row and plane
    const int64_t total = (int64_t)rows * per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / per_row), c = (int)(i % per_row) * 8;
        int64_t id = tokens[row];
        if (id < 0 || id >= V) id = 0;
        *reinterpret_cast<uint4*>(dst.hi + (int64_t)row * dst.ld + c) = *reinterpret_cast<const uint4*>(t_hi + id * ld_t + c);
        if (dst.mode != 2) *reinterpret_cast<uint4*>(dst.lo + (int64_t)row * dst.ld + c) = *reinterpret_cast<const uint4*>(t_lo + id * ld_t + c);
    }
}
int embed_split_rows(SplitDst dst, const uint16_t* t_hi, const uint16_t* t_lo, int64_t ld_t, int E, const int64_t* tokens, int rows,
                     int64_t V, cudaStream_t st) {
    if (rows == 0) return VAG_OK;
    const int64_t total = (int64_t)rows * (E / 8);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
    embed_split_rows_kernel<<<blocks, 256, 0, st>>>(dst, t_hi, t_lo, ld_t, E, tokens, rows, V);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

}  // namespace vag

using namespace vag;


extern "C" int vag_gru_gates_f32(float* h_out, int64_t ld_ho, float* h_out2, int64_t ld_ho2, const float* gi, int64_t ld_gi,
                                 const float* gh, int64_t ld_gh, const float* h_prev, int64_t ld_hp, int rows, int H,
                                 vag_stream_t stream) {
    VAG_REQUIRE(h_out && gi && gh && h_prev, "vag_gru_gates_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && H > 0, "vag_gru_gates_f32: bad shape rows=%d H=%d", rows, H);
    if (rows == 0) return VAG_OK;
    auto al = [](const void* p) { return ((uintptr_t)p % 16) == 0; };
    const bool vec = (H % 4 == 0) && (ld_ho % 4 == 0) && (ld_gi % 4 == 0) && (ld_gh % 4 == 0) && (ld_hp % 4 == 0) &&
                     (!h_out2 || ld_ho2 % 4 == 0) && al(h_out) && al(gi) && al(gh) && al(h_prev) && (!h_out2 || al(h_out2));
    const int64_t total = (int64_t)rows * (vec ? H / 4 : H);
    const int blocks = (int)std::min<int64_t>(ceil_div64(total, 256), (int64_t)num_sms() * 16);
    if (vec)
        gru_gates_kernel<true><<<blocks, 256, 0, (cudaStream_t)stream>>>(h_out, ld_ho, h_out2, ld_ho2, gi, ld_gi, gh, ld_gh, h_prev, ld_hp, rows, H, SplitDst());
    else
        gru_gates_kernel<false><<<blocks, 256, 0, (cudaStream_t)stream>>>(h_out, ld_ho, h_out2, ld_ho2, gi, ld_gi, gh, ld_gh, h_prev, ld_hp, rows, H, SplitDst());
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_init_mix_f32(float* z, const float* ctx_vec, const float* ctx, const float* mask, float split, int B,
                                int T, int C, vag_stream_t stream) {
    VAG_REQUIRE(z && ctx && mask, "vag_init_mix_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0 && C > 0, "vag_init_mix_f32: bad shape");
    dim3 grid(ceil_div(C, 128), B);
    init_mix_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(z, ctx_vec, ctx, mask, split, B, T, C);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_l2norm_rows_f32(float* x, int64_t ldx, int rows, int dim, vag_stream_t stream) {
    VAG_REQUIRE(x, "vag_l2norm_rows_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && dim > 0, "vag_l2norm_rows_f32: bad shape");
    if (rows == 0) return VAG_OK;
    l2norm_rows_kernel<<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(x, ldx, rows, dim);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_log_softmax_f32(float* logp, const float* logits, int rows, int V, vag_stream_t stream) {
    VAG_REQUIRE(logp && logits, "vag_log_softmax_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && V > 0, "vag_log_softmax_f32: bad shape");
    if (rows == 0) return VAG_OK;
    log_softmax_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(logp, logits, V);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_nll_rows_f32(const float* logits, int64_t ld, const int64_t* tgt, const float* weight, int rows,
                                int64_t V, float* loss_rows, float* lse_out, vag_stream_t stream) {
    VAG_REQUIRE(logits && tgt && loss_rows, "vag_nll_rows_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && V > 0 && ld >= V, "vag_nll_rows_f32: bad shape");
    if (rows == 0) return VAG_OK;
    nll_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(logits, ld, tgt, weight, V, loss_rows, lse_out);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_translation_loss_bwd_f32(const float* g, const int64_t* tgt, int B, int Tt, float loss_w, int has_vse,
                                            float* g_rows, float* g_vse, vag_stream_t stream) {
    VAG_REQUIRE(g && tgt && g_rows, "vag_translation_loss_bwd_f32: null pointer");
    VAG_REQUIRE(B > 0 && Tt > 0, "vag_translation_loss_bwd_f32: bad shape");
    VAG_REQUIRE(!has_vse || g_vse, "vag_translation_loss_bwd_f32: g_vse is required when the ranking term is mixed in");
    translation_loss_bwd_kernel<<<ceil_div(B, 256), 256, 0, (cudaStream_t)stream>>>(g, tgt, B, Tt, loss_w, has_vse, g_rows,
                                                                                    has_vse ? g_vse : nullptr);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_src_mask_lengths(const int64_t* src, int B, int T, float* mask, int32_t* lengths, vag_stream_t stream) {
    VAG_REQUIRE(src && (mask || lengths), "vag_src_mask_lengths: null pointer");
    VAG_REQUIRE(B > 0 && T > 0, "vag_src_mask_lengths: bad shape");
    src_mask_lengths_kernel<<<ceil_div(B, 8), 256, 0, (cudaStream_t)stream>>>(src, B, T, mask, lengths);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_translation_loss_f32(const float* loss_rows, const int64_t* tgt, int B, int Tt, const float* loss_vse,
                                        float loss_w, float* out, vag_stream_t stream) {
    VAG_REQUIRE(loss_rows && tgt && out, "vag_translation_loss_f32: null pointer");
    VAG_REQUIRE(B > 0 && Tt > 0, "vag_translation_loss_f32: bad shape");
    translation_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(loss_rows, tgt, B, Tt, loss_vse, loss_w, out);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
