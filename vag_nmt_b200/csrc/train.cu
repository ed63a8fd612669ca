// Training-side kernels: generic strided contraction (for dX = dY·W and dW = dYᵀ·X), backward of the GRU gates,
// of the fused attention, of log-softmax + NLL, of tanh / l2norm / decoder-init mix, embedding scatter-add, column
// sums for bias gradients, and the fused clip-by-global-norm + Adam step (train.py:46-49).
// Forward counterparts live in elementwise.cu / attention.cu; the Python autograd Functions in
// vag_nmt_b200/autograd.py sequence them (back-propagation through time over the Tt decoder steps).
#include "common.cuh"
#include <math.h>
#include <algorithm>

namespace vag {

// ------------------------------------------------------------------------------------------ generic contraction
// C[m, n] = alpha · Σ_k A(m, k)·B(k, n) + beta · C[m, n],   A(m,k) = A[m·sam + k·sak],  B(k,n) = B[k·sbk + n·sbn]
// gridDim.z > 1: split-K — every z-slice handles k_chunk of the K range and atomically adds alpha·partial into C
// (the host pre-scales C by beta); skinny problems (M = batch rows) would otherwise run on a handful of CTAs.
template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256)
gemm_generic_kernel(float* __restrict__ C, int64_t ldc, const float* __restrict__ A, int64_t sam, int64_t sak,
                    const float* __restrict__ B, int64_t sbk, int64_t sbn, int M, int N, int K, float alpha, float beta,
                    int k_chunk) {
    constexpr int TM = BM / 16, TN = BN / 16;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x, tx = tid % 16, ty = tid / 16;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    const bool a_kfast = sak == 1, b_nfast = sbn == 1;
    const int k_begin = blockIdx.z * k_chunk;
    const int k_end = min(K, k_begin + k_chunk);
    const bool split = gridDim.z > 1;
    K = k_end;
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
        for (int idx = tid; idx < BM * BK; idx += 256) {
            const int m = a_kfast ? idx / BK : idx % BM;
            const int k = a_kfast ? idx % BK : idx / BM;
            const int gm = m0 + m, gk = k0 + k;
            As[k][m] = (gm < M && gk < K) ? A[(int64_t)gm * sam + (int64_t)gk * sak] : 0.f;
        }
        for (int idx = tid; idx < BN * BK; idx += 256) {
            const int n = b_nfast ? idx % BN : idx / BK;
            const int k = b_nfast ? idx / BN : idx % BK;
            const int gn = n0 + n, gk = k0 + k;
            Bs[k][n] = (gn < N && gk < K) ? B[(int64_t)gk * sbk + (int64_t)gn * sbn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int gm = m0 + ty + 16 * i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int gn = n0 + tx + 16 * j;
            if (gn >= N) continue;
            float* dst = C + (int64_t)gm * ldc + gn;
            const float v = alpha * acc[i][j];
            if (split) atomicAdd(dst, v);
            else *dst = beta == 0.f ? v : fmaf(beta, *dst, v);
        }
    }
}

// ------------------------------------------------------------------------------------------ GRU gates backward
// Given the saved pre-activations gi, gh (3H each), h_prev and dh', produce dgi, dgh and dh_prev = dh'·z.
__global__ void __launch_bounds__(256)
gru_gates_bwd_kernel(float* __restrict__ dgi, float* __restrict__ dgh, float* __restrict__ dh_prev, const float* __restrict__ dh,
                     int64_t ld_dh, const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h_prev,
                     int64_t ld_hp, int rows, int H) {
    const int64_t total = (int64_t)rows * H;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        const int r_ = (int)(idx / H), j = (int)(idx % H);
        const float* gir = gi + (int64_t)r_ * 3 * H;
        const float* ghr = gh + (int64_t)r_ * 3 * H;
        const float r = sigmoidf_precise(gir[j] + ghr[j]);
        const float z = sigmoidf_precise(gir[H + j] + ghr[H + j]);
        const float hn = ghr[2 * H + j];
        const float n = tanhf(gir[2 * H + j] + r * hn);
        const float hp = h_prev[(int64_t)r_ * ld_hp + j];
        const float g = dh[(int64_t)r_ * ld_dh + j];
        const float dn_pre = g * (1.f - z) * (1.f - n * n);
        const float dz_pre = g * (hp - n) * z * (1.f - z);
        const float dr_pre = dn_pre * hn * r * (1.f - r);
        float* a = dgi + (int64_t)r_ * 3 * H;
        float* b = dgh + (int64_t)r_ * 3 * H;
        a[j] = dr_pre; a[H + j] = dz_pre; a[2 * H + j] = dn_pre;
        b[j] = dr_pre; b[H + j] = dz_pre; b[2 * H + j] = dn_pre * r;
        dh_prev[(int64_t)r_ * H + j] = g * z;
    }
}

// ------------------------------------------------------------------------------------------ attention backward
// One CTA per sentence (rows_per_sent == 1 in training).  Inputs: dc [B, C], saved α [B, T], q [B, C] (ld_q),
// keys / ctx [B, T, C], v [C], mask.  Outputs: dq [B, C]; dkeys, dctx [B, T, C] ACCUMULATED; dv [C] accumulated with
// atomics (MLP mode).  MODE DOT: s_t = q·keys_t  →  dq = Σ_t da_t keys_t,  dkeys_t += da_t q.
template <int MODE>
__global__ void __launch_bounds__(256)
attention_bwd_kernel(float* __restrict__ dq, int64_t ld_dq, float* __restrict__ dkeys, float* __restrict__ dctx,
                     float* __restrict__ dv, const float* __restrict__ dc, int64_t ld_dc, const float* __restrict__ alpha,
                     const float* __restrict__ q, int64_t ld_q, const float* __restrict__ keys, const float* __restrict__ ctx,
                     const float* __restrict__ v, const float* __restrict__ mask, int T, int C) {
    extern __shared__ float sm[];
    float* da = sm;          // [T]
    float* red = sm + T;     // [8]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const float* key_b = keys + (int64_t)b * T * C;
    const float* ctx_b = ctx + (int64_t)b * T * C;
    float* dkey_b = dkeys + (int64_t)b * T * C;
    float* dctx_b = dctx ? dctx + (int64_t)b * T * C : nullptr;
    const float* al = alpha + (int64_t)b * T;
    const float* dcb = dc + (int64_t)b * ld_dc;
    const float* qb = q + (int64_t)b * ld_q;
    // dα_t = dc·ctx_t (one warp per t), dctx_t += α_t dc
    for (int t = wid; t < T; t += 8) {
        const bool live = mask ? mask[(int64_t)b * T + t] != 0.f : true;
        float p = 0.f;
        if (live) {
            const float a = al[t];
            for (int c = lane; c < C; c += 32) {
                const float g = dcb[c];
                p = fmaf(g, ctx_b[(int64_t)t * C + c], p);
                if (dctx_b) dctx_b[(int64_t)t * C + c] += a * g;
            }
        }
        p = warp_sum(p);
        if (lane == 0) da[t] = live ? p : 0.f;
    }
    __syncthreads();
    // softmax backward: da_t = α_t (dα_t − Σ_u α_u dα_u)
    float part = 0.f;
    for (int t = tid; t < T; t += 256) part += al[t] * da[t];
    part = warp_sum(part);
    if (lane == 0) red[wid] = part;
    __syncthreads();
    float dot = 0.f;
    for (int w = 0; w < 8; ++w) dot += red[w];
    __syncthreads();
    for (int t = tid; t < T; t += 256) da[t] = al[t] * (da[t] - dot);
    __syncthreads();
    // through the score
    for (int c = tid; c < C; c += 256) {
        float dqc = 0.f, dvc = 0.f;
        const float qc = qb[c];
        const float vc = MODE == VAG_ATTN_MLP ? v[c] : 0.f;
        for (int t = 0; t < T; ++t) {
            const float g = da[t];
            if (g == 0.f) continue;
            const float k = key_b[(int64_t)t * C + c];
            if (MODE == VAG_ATTN_MLP) {
                const float e = tanhf(qc + k);
                const float dpre = g * vc * (1.f - e * e);
                dvc = fmaf(g, e, dvc);
                dqc += dpre;
                dkey_b[(int64_t)t * C + c] += dpre;
            } else {
                dqc = fmaf(g, k, dqc);
                dkey_b[(int64_t)t * C + c] += g * qc;
            }
        }
        dq[(int64_t)b * ld_dq + c] = dqc;
        if (MODE == VAG_ATTN_MLP && dv) atomicAdd(dv + c, dvc);
    }
}

// ------------------------------------------------------------------------------------------ NLL backward
// dlogits[r, v] = g[r]·w[tgt[r]]·(softmax(logits)[r, v] − [v == tgt[r]])
__global__ void __launch_bounds__(256)
nll_bwd_kernel(float* __restrict__ dlogits, int64_t ldd, const float* __restrict__ logits, int64_t ld, const float* __restrict__ lse,
               const int64_t* __restrict__ tgt, const float* __restrict__ weight, const float* __restrict__ g, int64_t V) {
    const int r = blockIdx.x;
    int64_t t = tgt[r];
    if (t < 0 || t >= V) t = 0;
    const float scale = g[r] * (weight ? weight[t] : 1.f);
    const float l = lse[r];
    const float* src = logits + (int64_t)r * ld;
    float* dst = dlogits + (int64_t)r * ldd;
    for (int64_t i = threadIdx.x; i < V; i += blockDim.x) dst[i] = scale * (expf(src[i] - l) - (i == t ? 1.f : 0.f));
}

// ------------------------------------------------------------------------------------------ small element-wise / reductions
__global__ void tanh_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * (1.f - y[i] * y[i]);
}
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}
// out[c] (+)= Σ_r x[r, c]
__global__ void __launch_bounds__(256) colsum_kernel(float* __restrict__ out, const float* __restrict__ x, int64_t ldx, int rows, int cols, int accumulate) {
    const int c = blockIdx.x * 32 + (threadIdx.x & 31);
    const int part = threadIdx.x >> 5;
    __shared__ float sm[8][33];
    float a = 0.f;
    if (c < cols)
        for (int r = part; r < rows; r += 8) a += x[(int64_t)r * ldx + c];
    sm[part][threadIdx.x & 31] = a;
    __syncthreads();
    if (part == 0 && c < cols) {
        float t = 0.f;
        for (int p = 0; p < 8; ++p) t += sm[p][threadIdx.x & 31];
        out[c] = accumulate ? out[c] + t : t;
    }
}
// table_grad[ids[r], :] += g[r, :]
__global__ void embed_bwd_kernel(float* __restrict__ table_grad, const float* __restrict__ g, int64_t ldg, const int64_t* __restrict__ ids,
                                 int rows, int dim, int64_t table_rows) {
    const int row = blockIdx.x;
    int64_t id = ids[row];
    if (id < 0 || id >= table_rows) return;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) atomicAdd(table_grad + id * dim + c, g[(int64_t)row * ldg + c]);
}
// y = x / max(‖x‖, eps) per row;  dx = (dy − y (y·dy)) / max(‖x‖, eps)      (one warp per row)
__global__ void l2norm_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ x, int rows, int dim) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* xr = x + (int64_t)row * dim;
    const float* gr = dy + (int64_t)row * dim;
    float ss = 0.f, xg = 0.f;
    for (int c = lane; c < dim; c += 32) { ss = fmaf(xr[c], xr[c], ss); xg = fmaf(xr[c], gr[c], xg); }
    ss = warp_sum(ss);
    xg = warp_sum(xg);
    const float nrm = fmaxf(sqrtf(ss), 1e-12f);
    const float k = xg / (nrm * nrm * nrm);
    for (int c = lane; c < dim; c += 32) dx[(int64_t)row * dim + c] = gr[c] / nrm - xr[c] * k;
}
// backward of vag_init_mix_f32: dctx_vec = split·dz ; dctx[b,t,:] += (1−split)·dz / Σmask   at live positions
__global__ void init_mix_bwd_kernel(float* __restrict__ dctx_vec, float* __restrict__ dctx, const float* __restrict__ dz,
                                    const float* __restrict__ mask, float split, int B, int T, int C) {
    const int b = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float msum = 0.f;
    for (int t = 0; t < T; ++t) msum += mask[(int64_t)b * T + t];
    const float g = dz[(int64_t)b * C + c];
    if (dctx_vec) dctx_vec[(int64_t)b * C + c] = split * g;
    const float share = (dctx_vec ? (1.f - split) : 1.f) * g / msum;
    for (int t = 0; t < T; ++t)
        if (mask[(int64_t)b * T + t] != 0.f) dctx[((int64_t)b * T + t) * C + c] += share;
}

// ------------------------------------------------------------------------------------------ clip + Adam
// Σ g² over a flat buffer, accumulated into out[0] (one atomic per block)
__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
    float a = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) a = fmaf(g[i], g[i], a);
    a = warp_sum(a);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[w];
        atomicAdd(out, t);
    }
}
// torch.optim.Adam (weight decay added to the gradient) with the gradient pre-scaled by min(1, clip/(‖g‖+1e-6))
__global__ void __launch_bounds__(256)
clip_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                 const float* __restrict__ sumsq, float clip, float lr, float beta1, float beta2, float eps, float wd, float bc1,
                 float bc2_sqrt) {
    const float norm = sqrtf(sumsq[0]);
    const float coef = fminf(1.f, clip / (norm + 1e-6f));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        const float w = p[i];
        if (wd != 0.f) gi = fmaf(wd, w, gi);
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = w - (lr / bc1) * (mi / denom);
    }
}

static inline int grid_for(int64_t n) { return (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms() * 8); }

}  // namespace vag

using namespace vag;

extern "C" int vag_gemm_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                            int64_t sbn, int M, int N, int K, float alpha, float beta, vag_stream_t stream) {
    VAG_REQUIRE(C && A && B, "vag_gemm_f32: null pointer");
    VAG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "vag_gemm_f32: bad shape");
    if (M == 0 || N == 0) return VAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const bool small_m = M <= 32;
    const int tiles = small_m ? ceil_div(N, 64) * ceil_div(M, 32) : ceil_div(N, 64) * ceil_div(M, 64);
    int splits = 1;
    const int target = 2 * num_sms();
    if (tiles < target && K >= 256) splits = std::min(ceil_div(target, tiles), K / 64);
    int k_chunk = K;
    if (splits > 1) {
        k_chunk = ceil_div(ceil_div(K, splits), 16) * 16;
        splits = ceil_div(K, k_chunk);
    }
    if (splits > 1) {   // C ← beta·C first, then every slice adds its partial
        if (beta == 0.f) {
            if (ldc == N) VAG_CUDA(cudaMemsetAsync(C, 0, (size_t)M * N * sizeof(float), st));
            else VAG_CUDA(cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, st));
        } else if (beta != 1.f) {
            set_error("vag_gemm_f32: split-K path supports beta 0 or 1 only");
            return VAG_ERR_UNSUPPORTED;
        }
    }
    if (small_m) {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 32), splits);
        gemm_generic_kernel<32, 64, 16><<<grid, 256, 0, st>>>(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, alpha, beta, k_chunk);
    } else {
        dim3 grid(ceil_div(N, 64), ceil_div(M, 64), splits);
        gemm_generic_kernel<64, 64, 16><<<grid, 256, 0, st>>>(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, alpha, beta, k_chunk);
    }
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_gru_gates_bwd_f32(float* dgi, float* dgh, float* dh_prev, const float* dh, int64_t ld_dh, const float* gi,
                                     const float* gh, const float* h_prev, int64_t ld_hp, int rows, int H, vag_stream_t stream) {
    VAG_REQUIRE(dgi && dgh && dh_prev && dh && gi && gh && h_prev, "vag_gru_gates_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    gru_gates_bwd_kernel<<<grid_for((int64_t)rows * H), 256, 0, (cudaStream_t)stream>>>(dgi, dgh, dh_prev, dh, ld_dh, gi, gh, h_prev, ld_hp, rows, H);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_attention_bwd_f32(float* dq, int64_t ld_dq, float* dkeys, float* dctx, float* dv, const float* dc, int64_t ld_dc,
                                     const float* alpha, const float* q, int64_t ld_q, const float* keys, const float* ctx,
                                     const float* v, const float* mask, int B, int T, int C, int mode, vag_stream_t stream) {
    VAG_REQUIRE(dq && dkeys && dc && alpha && q && keys && ctx, "vag_attention_bwd_f32: null pointer");
    VAG_REQUIRE(mode == VAG_ATTN_DOT || v, "vag_attention_bwd_f32: MLP mode needs v");
    if (B == 0) return VAG_OK;
    const size_t smem = (size_t)(T + 8) * sizeof(float);
    if (mode == VAG_ATTN_MLP)
        attention_bwd_kernel<VAG_ATTN_MLP><<<B, 256, smem, (cudaStream_t)stream>>>(dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C);
    else
        attention_bwd_kernel<VAG_ATTN_DOT><<<B, 256, smem, (cudaStream_t)stream>>>(dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_nll_bwd_f32(float* dlogits, int64_t ldd, const float* logits, int64_t ld, const float* lse, const int64_t* tgt,
                               const float* weight, const float* grad_rows, int rows, int64_t V, vag_stream_t stream) {
    VAG_REQUIRE(dlogits && logits && lse && tgt && grad_rows, "vag_nll_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    nll_bwd_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(dlogits, ldd, logits, ld, lse, tgt, weight, grad_rows, V);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_tanh_bwd_f32(float* dx, const float* dy, const float* y, int64_t n, vag_stream_t stream) {
    VAG_REQUIRE(dx && dy && y, "vag_tanh_bwd_f32: null pointer");
    if (n == 0) return VAG_OK;
    tanh_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dx, dy, y, n);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_axpby_f32(float* y, const float* x, float a, float b, int64_t n, vag_stream_t stream) {
    VAG_REQUIRE(y && x, "vag_axpby_f32: null pointer");
    if (n == 0) return VAG_OK;
    axpby_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, x, a, b, n);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_colsum_f32(float* out, const float* x, int64_t ldx, int rows, int cols, int accumulate, vag_stream_t stream) {
    VAG_REQUIRE(out && x, "vag_colsum_f32: null pointer");
    if (cols == 0) return VAG_OK;
    colsum_kernel<<<ceil_div(cols, 32), 256, 0, (cudaStream_t)stream>>>(out, x, ldx, rows, cols, accumulate);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_embed_bwd_f32(float* table_grad, const float* g, int64_t ldg, const int64_t* ids, int rows, int dim,
                                 int64_t table_rows, vag_stream_t stream) {
    VAG_REQUIRE(table_grad && g && ids, "vag_embed_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    embed_bwd_kernel<<<rows, 128, 0, (cudaStream_t)stream>>>(table_grad, g, ldg, ids, rows, dim, table_rows);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_l2norm_bwd_f32(float* dx, const float* dy, const float* x, int rows, int dim, vag_stream_t stream) {
    VAG_REQUIRE(dx && dy && x, "vag_l2norm_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    l2norm_bwd_kernel<<<ceil_div(rows, 4), 128, 0, (cudaStream_t)stream>>>(dx, dy, x, rows, dim);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_init_mix_bwd_f32(float* dctx_vec, float* dctx, const float* dz, const float* mask, float split, int B, int T,
                                    int C, vag_stream_t stream) {
    VAG_REQUIRE(dctx && dz && mask, "vag_init_mix_bwd_f32: null pointer");
    dim3 grid(ceil_div(C, 128), B);
    init_mix_bwd_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(dctx_vec, dctx, dz, mask, split, B, T, C);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_sumsq_f32(const float* g, int64_t n, float* accum, vag_stream_t stream) {
    VAG_REQUIRE(g && accum, "vag_sumsq_f32: null pointer");
    if (n == 0) return VAG_OK;
    sumsq_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(g, n, accum);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_clip_adam_f32(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                 const float* grad_sumsq, float clip, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int step, vag_stream_t stream) {
    VAG_REQUIRE(param && grad && exp_avg && exp_avg_sq && grad_sumsq, "vag_clip_adam_f32: null pointer");
    VAG_REQUIRE(step >= 1, "vag_clip_adam_f32: step counts from 1");
    if (n == 0) return VAG_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    clip_adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, grad_sumsq, clip, lr, beta1,
                                                                    beta2, eps, weight_decay, bc1, bc2_sqrt);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
