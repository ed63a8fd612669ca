// Training-side kernels: generic strided contraction (for dX = dY·W and dW = dYᵀ·X), backward of the GRU gates,
// of the fused attention, of log-softmax + NLL, of tanh / l2norm / decoder-init mix, embedding scatter-add, column
// sums for bias gradients, and the fused clip-by-global-norm + Adam step (train.py:46-49).
// Forward counterparts live in elementwise.cu / attention.cu; the Python autograd Functions in
// vag_nmt_b200/autograd.py sequence them (back-propagation through time over the Tt decoder steps).
#include "common.cuh"
#include "linear_rows.cuh"
#include "enc_seq.cuh"
#include "dec_seq.cuh"
#include <math.h>
#include <algorithm>
#include <vector>

namespace vag {
int gemm_mode();

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
                     int64_t ld_hp, int rows, int H, const float* __restrict__ dh_add = nullptr) {
    pdl_trigger();
    pdl_wait();
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
        const float g = dh[(int64_t)r_ * ld_dh + j] + (dh_add ? dh_add[(int64_t)r_ * H + j] : 0.f);   // direct + recurrent path
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
    // blockIdx.y selects a 256-channel slab: every CTA of a sentence recomputes the (cheap) softmax backward but only
    // touches dctx / dkeys / dq / dv inside its own slab, so the expensive score backward spreads over B·C/256 CTAs
    const int c_lo = blockIdx.y * 256, c_hi = min(C, c_lo + 256);
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
                if (dctx_b && c >= c_lo && c < c_hi) dctx_b[(int64_t)t * C + c] += a * g;
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
    for (int c = c_lo + tid; c < c_hi; c += 256) {
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

// Same contract, shaped for latency (C a multiple of 128, 16-byte aligned rows): the kernel above walks its loops one
// dependent global load at a time — ~30 L2 round trips in sequence per CTA.  Here every warp issues the eight 16-byte loads
// of a context row (two rows at a time) before it touches them, dc sits in shared memory, and the score backward handles
// eight time steps per batch of loads (masked positions have dα = 0 and simply add zero).
template <int MODE>
__global__ void __launch_bounds__(256)
attention_bwd_fast_kernel(float* __restrict__ dq, int64_t ld_dq, float* __restrict__ dkeys, float* __restrict__ dctx,
                          float* __restrict__ dv, const float* __restrict__ dc, int64_t ld_dc, const float* __restrict__ alpha,
                          const float* __restrict__ q, int64_t ld_q, const float* __restrict__ keys, const float* __restrict__ ctx,
                          const float* __restrict__ v, const float* __restrict__ mask, int T, int C) {
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(16) float sm[];
    float* dcs = sm;             // [C]
    float* da = sm + C;          // [T]
    float* red = da + T;         // [8]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int c_lo = blockIdx.y * 256;
    const float* key_b = keys + (int64_t)b * T * C;
    const float* ctx_b = ctx + (int64_t)b * T * C;
    float* dkey_b = dkeys + (int64_t)b * T * C;
    float* dctx_b = dctx ? dctx + (int64_t)b * T * C : nullptr;
    const float* al = alpha + (int64_t)b * T;
    for (int c = tid * 4; c < C; c += 1024) *reinterpret_cast<float4*>(dcs + c) = *reinterpret_cast<const float4*>(dc + (int64_t)b * ld_dc + c);
    __syncthreads();
    const int nq = C >> 7;       // 16-byte pieces of a row per lane
    // dα_t = dc·ctx_t (one warp per t, two rows in flight), dctx_t += α_t dc inside this CTA's 256-channel slab
    for (int t0 = wid; t0 < T; t0 += 16) {
        const int t1 = t0 + 8;
        const bool live0 = mask ? mask[(int64_t)b * T + t0] != 0.f : true;
        const bool live1 = t1 < T && (mask ? mask[(int64_t)b * T + t1] != 0.f : true);
        float p0 = 0.f, p1 = 0.f;
        for (int i0 = 0; i0 < nq; i0 += 4) {
            float4 a0[4], a1[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int c = ((i0 + u) * 32 + lane) * 4;
                const bool in = i0 + u < nq;
                a0[u] = (live0 && in) ? *reinterpret_cast<const float4*>(ctx_b + (int64_t)t0 * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                a1[u] = (live1 && in) ? *reinterpret_cast<const float4*>(ctx_b + (int64_t)t1 * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (i0 + u >= nq) break;
                const float4 g = *reinterpret_cast<const float4*>(dcs + ((i0 + u) * 32 + lane) * 4);
                p0 += g.x * a0[u].x + g.y * a0[u].y + g.z * a0[u].z + g.w * a0[u].w;
                p1 += g.x * a1[u].x + g.y * a1[u].y + g.z * a1[u].z + g.w * a1[u].w;
            }
        }
        p0 = warp_sum(p0);
        p1 = warp_sum(p1);
        if (lane == 0) {
            da[t0] = live0 ? p0 : 0.f;
            if (t1 < T) da[t1] = live1 ? p1 : 0.f;
        }
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
    // through the score and the context: thread = channel of the slab, eight time steps per batch of loads
    const int c = c_lo + tid;
    if (c < C) {
        float dqc = 0.f, dvc = 0.f;
        const float qc = q[(int64_t)b * ld_q + c];
        const float vc = MODE == VAG_ATTN_MLP ? v[c] : 0.f;
        const float gdc = dcs[c];
        for (int t0 = 0; t0 < T; t0 += 8) {
            float k[8], dk[8], dx[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = min(t0 + u, T - 1);
                k[u] = key_b[(int64_t)t * C + c];
                dk[u] = dkey_b[(int64_t)t * C + c];
                dx[u] = dctx_b ? dctx_b[(int64_t)t * C + c] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const int t = t0 + u;
                if (t >= T) break;
                const float g = da[t];
                const bool live = mask ? mask[(int64_t)b * T + t] != 0.f : true;
                if (MODE == VAG_ATTN_MLP) {
                    const float e = tanhf(qc + k[u]);
                    const float dpre = g * vc * (1.f - e * e);
                    dvc = fmaf(g, e, dvc);
                    dqc += dpre;
                    dkey_b[(int64_t)t * C + c] = dk[u] + dpre;
                } else {
                    dqc = fmaf(g, k[u], dqc);
                    dkey_b[(int64_t)t * C + c] = dk[u] + g * qc;
                }
                if (dctx_b && live) dctx_b[(int64_t)t * C + c] = dx[u] + al[t] * gdc;
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
               const int64_t* __restrict__ tgt, const float* __restrict__ weight, const float* __restrict__ g, int64_t V, int g_mod) {
    pdl_trigger();
    pdl_wait();
    const int r = blockIdx.x;
    int64_t t = tgt[r];
    if (t < 0 || t >= V) t = 0;
    const float scale = g[g_mod > 0 ? r % g_mod : r] * (weight ? weight[t] : 1.f);   // g_mod = B: rows are t·B + b, g is per sentence
    const float l = lse[r];
    const float* src = logits + (int64_t)r * ld;
    float* dst = dlogits + (int64_t)r * ldd;
    for (int64_t i = threadIdx.x; i < V; i += blockDim.x) dst[i] = scale * (expf(src[i] - l) - (i == t ? 1.f : 0.f));
}

// ------------------------------------------------------------------------------------------ small element-wise / reductions
__global__ void tanh_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ y, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dx[i] = dy[i] * (1.f - y[i] * y[i]);
}
// y *= m (dropout mask already scaled by 1/(1-p))
__global__ void mul_kernel(float* __restrict__ y, const float* __restrict__ m, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] *= m[i];
}
// y = tanh(u)·m was stored; dx = dy·m·(1 − tanh(u)²) with tanh(u) = y/m where the unit was kept, 0 gradient elsewhere
__global__ void tanh_dropout_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ y,
                                        const float* __restrict__ m, int64_t n) {
    pdl_trigger();
    pdl_wait();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float mi = m[i];
        const float th = mi != 0.f ? y[i] / mi : 0.f;
        dx[i] = dy[i] * mi * (1.f - th * th);
    }
}
__global__ void axpby_kernel(float* __restrict__ y, const float* __restrict__ x, float a, float b, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        y[i] = a * x[i] + (b == 0.f ? 0.f : b * y[i]);
}
// out[c] (+)= Σ_r x[r, c]
__global__ void __launch_bounds__(256) colsum_kernel(float* __restrict__ out, const float* __restrict__ x, int64_t ldx, int rows, int cols, int accumulate) {
    pdl_trigger();
    pdl_wait();
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
    pdl_trigger();
    pdl_wait();
    const int row = blockIdx.x;
    int64_t id = ids[row];
    if (id < 0 || id >= table_rows) return;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) atomicAdd(table_grad + id * dim + c, g[(int64_t)row * ldg + c]);
}
// y = x / max(‖x‖, eps) per row;  dx = (dy − y (y·dy)) / max(‖x‖, eps)      (one warp per row)
__global__ void l2norm_bwd_kernel(float* __restrict__ dx, const float* __restrict__ dy, const float* __restrict__ x, int rows, int dim) {
    pdl_trigger();
    pdl_wait();
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

// Multi-tensor flavours: ONE launch covers every parameter tensor (blockIdx.y = tensor, blockIdx.x strides over its
// elements) instead of two launches per tensor — the ~37 tensors of the model cost 74 host calls per step otherwise.
__global__ void __launch_bounds__(256) sumsq_multi_kernel(const vag_optim_tensor* __restrict__ t, float* __restrict__ out) {
    const vag_optim_tensor e = t[blockIdx.y];
    float a = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * blockDim.x) a = fmaf(e.grad[i], e.grad[i], a);
    a = warp_sum(a);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        if (s != 0.f) atomicAdd(out, s);
    }
}
// Deterministic flavour: every block stores its partial (fixed reduction tree inside the block), one block then adds the
// partials in index order.  Float atomics would make Σ‖g‖² — hence the clip coefficient, hence every parameter — differ in
// the last bits from run to run and, under data parallelism, from replica to replica.
__global__ void __launch_bounds__(256) sumsq_multi_partials_kernel(const vag_optim_tensor* __restrict__ t, float* __restrict__ partials) {
    const vag_optim_tensor e = t[blockIdx.y];
    float a = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * blockDim.x) a = fmaf(e.grad[i], e.grad[i], a);
    a = warp_sum(a);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        partials[(int64_t)blockIdx.y * gridDim.x + blockIdx.x] = s;
    }
}
__global__ void __launch_bounds__(256) sumsq_finish_kernel(const float* __restrict__ partials, int n, float* __restrict__ out) {
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 256) a += partials[i];
    a = warp_sum(a);
    __shared__ float red[8];
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < 8; ++w) s += red[w];
        out[0] = s;
    }
}
__global__ void __launch_bounds__(256)
clip_adam_multi_kernel(const vag_optim_tensor* __restrict__ t, const float* __restrict__ sumsq, float clip, float beta1, float beta2,
                       float eps, float bc1, float bc2_sqrt) {
    const vag_optim_tensor e = t[blockIdx.y];
    const float norm = sqrtf(sumsq[0]);
    const float coef = fminf(1.f, clip / (norm + 1e-6f));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += (int64_t)gridDim.x * blockDim.x) {
        float gi = e.grad[i] * coef;
        const float w = e.param[i];
        if (e.weight_decay != 0.f) gi = fmaf(e.weight_decay, w, gi);
        const float mi = beta1 * e.exp_avg[i] + (1.f - beta1) * gi;
        const float vi = beta2 * e.exp_avg_sq[i] + (1.f - beta2) * gi * gi;
        e.exp_avg[i] = mi;
        e.exp_avg_sq[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        e.param[i] = w - (e.lr / bc1) * (mi / denom);
    }
}

}  // namespace vag

using namespace vag;

namespace vag {
int gemm_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, int M, int N,
             int K, float alpha, float beta, vag_stream_t stream);
}
extern "C" int vag_gemm_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                            int64_t sbn, int M, int N, int K, float alpha, float beta, int precision, vag_stream_t stream) {
    ModeScope ms(precision);
    return gemm_f32(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, alpha, beta, stream);
}
// the arithmetic mode is the enclosing C-ABI call's (ModeScope)
int vag::gemm_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, int M,
                  int N, int K, float alpha, float beta, vag_stream_t stream) {
    VAG_REQUIRE(C && A && B, "vag_gemm_f32: null pointer");
    VAG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "vag_gemm_f32: bad shape");
    if (M == 0 || N == 0) return VAG_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (alpha == 1.f && (beta == 0.f || beta == 1.f) && sak == 1 && (sbk == 1 || sbn == 1) && M <= 32) {
        // batch-sized row count: the weight-streaming kernel (linear_rows.cu); B(k, n) is the weight read k-fast or n-fast
        const bool wk = sbk == 1;
        const int64_t ldw = wk ? sbn : sbk;
        if (rows32_ok(A, sam, B, ldw, M, K, N, wk))
            return linear_rows32(C, ldc, A, sam, B, ldw, nullptr, M, K, N, beta == 1.f ? VAG_LIN_ACCUMULATE : 0, wk, gemm_mode() == 2, st);
    }
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
    VAG_CUDA(launch_pdl(PDL_SMALL, gru_gates_bwd_kernel, dim3(grid_for((int64_t)rows * H)), dim3(256), 0, (cudaStream_t)stream, dgi, dgh, dh_prev, dh, ld_dh, gi, gh, h_prev, ld_hp, rows, H, nullptr));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_attention_bwd_f32(float* dq, int64_t ld_dq, float* dkeys, float* dctx, float* dv, const float* dc, int64_t ld_dc,
                                     const float* alpha, const float* q, int64_t ld_q, const float* keys, const float* ctx,
                                     const float* v, const float* mask, int B, int T, int C, int mode, vag_stream_t stream) {
    VAG_REQUIRE(dq && dkeys && dc && alpha && q && keys && ctx, "vag_attention_bwd_f32: null pointer");
    VAG_REQUIRE(mode == VAG_ATTN_DOT || v, "vag_attention_bwd_f32: MLP mode needs v");
    if (B == 0) return VAG_OK;
    cudaStream_t st_ = (cudaStream_t)stream;
    const bool aligned = (C % 128 == 0) && (ld_dc % 4 == 0) && !(((uintptr_t)dc | (uintptr_t)ctx) & 15);
    if (aligned && (size_t)(C + T + 8) * sizeof(float) <= 48 * 1024) {
        const size_t smem_f = (size_t)(C + T + 8) * sizeof(float);
        if (mode == VAG_ATTN_MLP)
            VAG_CUDA(launch_pdl(PDL_ATTN, attention_bwd_fast_kernel<VAG_ATTN_MLP>, dim3(B, ceil_div(C, 256)), dim3(256), smem_f, st_, dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C));
        else
            VAG_CUDA(launch_pdl(PDL_ATTN, attention_bwd_fast_kernel<VAG_ATTN_DOT>, dim3(B, ceil_div(C, 256)), dim3(256), smem_f, st_, dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C));
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    const size_t smem = (size_t)(T + 8) * sizeof(float);
    if (mode == VAG_ATTN_MLP)
        attention_bwd_kernel<VAG_ATTN_MLP><<<dim3(B, ceil_div(C, 256)), 256, smem, (cudaStream_t)stream>>>(dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C);
    else
        attention_bwd_kernel<VAG_ATTN_DOT><<<dim3(B, ceil_div(C, 256)), 256, smem, (cudaStream_t)stream>>>(dq, ld_dq, dkeys, dctx, dv, dc, ld_dc, alpha, q, ld_q, keys, ctx, v, mask, T, C);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_nll_bwd_f32(float* dlogits, int64_t ldd, const float* logits, int64_t ld, const float* lse, const int64_t* tgt,
                               const float* weight, const float* grad_rows, int rows, int64_t V, vag_stream_t stream) {
    VAG_REQUIRE(dlogits && logits && lse && tgt && grad_rows, "vag_nll_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    VAG_CUDA(launch_pdl(PDL_SMALL, nll_bwd_kernel, dim3(rows), dim3(256), 0, (cudaStream_t)stream, dlogits, ldd, logits, ld, lse, tgt, weight, grad_rows, V, 0));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_tanh_bwd_f32(float* dx, const float* dy, const float* y, int64_t n, vag_stream_t stream) {
    VAG_REQUIRE(dx && dy && y, "vag_tanh_bwd_f32: null pointer");
    if (n == 0) return VAG_OK;
    VAG_CUDA(launch_pdl(PDL_SMALL, tanh_bwd_kernel, dim3(grid_for(n)), dim3(256), 0, (cudaStream_t)stream, dx, dy, y, n));
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

extern "C" int vag_mul_f32(float* y, const float* m, int64_t n, vag_stream_t stream) {
    VAG_REQUIRE(y && m, "vag_mul_f32: null pointer");
    if (n == 0) return VAG_OK;
    mul_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, m, n);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_colsum_f32(float* out, const float* x, int64_t ldx, int rows, int cols, int accumulate, vag_stream_t stream) {
    VAG_REQUIRE(out && x, "vag_colsum_f32: null pointer");
    if (cols == 0) return VAG_OK;
    VAG_CUDA(launch_pdl(PDL_SMALL, colsum_kernel, dim3(ceil_div(cols, 32)), dim3(256), 0, (cudaStream_t)stream, out, x, ldx, rows, cols, accumulate));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_embed_bwd_f32(float* table_grad, const float* g, int64_t ldg, const int64_t* ids, int rows, int dim,
                                 int64_t table_rows, vag_stream_t stream) {
    VAG_REQUIRE(table_grad && g && ids, "vag_embed_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    VAG_CUDA(launch_pdl(PDL_SMALL, embed_bwd_kernel, dim3(rows), dim3(128), 0, (cudaStream_t)stream, table_grad, g, ldg, ids, rows, dim, table_rows));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_l2norm_bwd_f32(float* dx, const float* dy, const float* x, int rows, int dim, vag_stream_t stream) {
    VAG_REQUIRE(dx && dy && x, "vag_l2norm_bwd_f32: null pointer");
    if (rows == 0) return VAG_OK;
    VAG_CUDA(launch_pdl(PDL_SMALL, l2norm_bwd_kernel, dim3(ceil_div(rows, 4)), dim3(128), 0, (cudaStream_t)stream, dx, dy, x, rows, dim));
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

extern "C" int vag_sumsq_multi_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, float* accum,
                                   vag_stream_t stream) {
    VAG_REQUIRE(tensors_device && accum && n_tensors >= 0 && max_n >= 0, "vag_sumsq_multi_f32: bad argument");
    if (n_tensors == 0 || max_n == 0) return VAG_OK;
    dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 128), (unsigned)n_tensors);
    sumsq_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tensors_device, accum);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" size_t vag_sumsq_multi_partials(int n_tensors, int64_t max_n) {
    return (size_t)std::min<int64_t>((max_n + 255) / 256, 128) * (size_t)std::max(n_tensors, 0);
}

extern "C" int vag_sumsq_multi_det_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, float* out,
                                       float* partials, size_t n_partials, vag_stream_t stream) {
    VAG_REQUIRE(tensors_device && out && partials && n_tensors >= 0 && max_n >= 0, "vag_sumsq_multi_det_f32: bad argument");
    VAG_REQUIRE(n_partials >= vag_sumsq_multi_partials(n_tensors, max_n), "vag_sumsq_multi_det_f32: partials buffer too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (n_tensors == 0 || max_n == 0) {
        VAG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
        return VAG_OK;
    }
    dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 128), (unsigned)n_tensors);
    sumsq_multi_partials_kernel<<<grid, 256, 0, st>>>(tensors_device, partials);
    VAG_LAUNCH_CHECK();
    sumsq_finish_kernel<<<1, 256, 0, st>>>(partials, (int)(grid.x * grid.y), out);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

extern "C" int vag_clip_adam_multi_f32(const vag_optim_tensor* tensors_device, int n_tensors, int64_t max_n, const float* grad_sumsq,
                                       float clip, float beta1, float beta2, float eps, int step, vag_stream_t stream) {
    VAG_REQUIRE(tensors_device && grad_sumsq && n_tensors >= 0 && max_n >= 0, "vag_clip_adam_multi_f32: bad argument");
    VAG_REQUIRE(step >= 1, "vag_clip_adam_multi_f32: step counts from 1");
    if (n_tensors == 0 || max_n == 0) return VAG_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    dim3 grid((unsigned)std::min<int64_t>((max_n + 255) / 256, 128), (unsigned)n_tensors);
    clip_adam_multi_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(tensors_device, grad_sumsq, clip, beta1, beta2, eps, bc1, bc2_sqrt);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// ==========================================================================================================
// Sequence-level composites: the whole Tt-step decoder loop (forward with saved activations, and its BPTT) in
// one C call each, so a training step costs a handful of host calls instead of one per kernel.
// Reference: V11.forward :136-160 + autograd.  Buffers are caller-owned (vag_decoder_seq_saved).
// ==========================================================================================================
#include "gemm_ctx.cuh"

namespace vag {
int row_argmax(const float* logits, int64_t ld, int rows, int64_t V, int64_t* out, int64_t out_stride, int64_t* next_in,
               cudaStream_t st);
int nll_rows_all(const float* logits, int64_t ld, const int64_t* tgt, const float* weight, int B, int Tt, int64_t V, float* loss_rows,
                 float* lse_out, float* nll_scratch, cudaStream_t st);
// ---- tensor-core route of the batched backward contractions (dW = dyᵀ·x over all batch·time rows, dx = dy·W) --------------
// tc_gemm wants both operands contraction-contiguous ([M, Kc] and [N, Kc]); an operand stored the other way round goes
// through the transposing branch of the operand split (tc_split_pair).  Scratch for the operand planes comes from the composite's workspace.
struct TcScratch {
    char* base = nullptr;
    size_t cap = 0;
};
static thread_local TcScratch g_tc_scratch;
struct TcScratchScope {
    TcScratch prev;
    TcScratchScope(void* p, size_t n) : prev(g_tc_scratch) { g_tc_scratch = TcScratch{(char*)p, n}; }
    ~TcScratchScope() { g_tc_scratch = prev; }
};
__global__ void add2d_kernel(float* __restrict__ C, int64_t ldc, const float* __restrict__ T, int M, int N) {
    pdl_trigger();
    pdl_wait();
    const int64_t total = (int64_t)M * (N >> 2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int m = (int)(i / (N >> 2)), n = (int)(i % (N >> 2)) * 4;
        float4 c = *reinterpret_cast<float4*>(C + (int64_t)m * ldc + n);
        const float4 t = *reinterpret_cast<const float4*>(T + (int64_t)m * N + n);
        c.x += t.x; c.y += t.y; c.z += t.z; c.w += t.w;
        *reinterpret_cast<float4*>(C + (int64_t)m * ldc + n) = c;
    }
}
// → 1: ran on the tensor cores, 0: not eligible (caller falls back to the FFMA kernel), < 0: error
static int gemm_tc_try(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                       int M, int N, int K, float beta, char* scratch, size_t cap, cudaStream_t st, bool a_padded = false) {
    // a_padded: A is stored [M, K] with rows zero-padded up to the next multiple of 4 past K (K itself need not be one)
    if (!scratch || !tc_enabled() || M < 64 || N < 64 || K < 32 || (N & 3)) return 0;
    if (!(beta == 0.f || beta == 1.f)) return 0;
    const bool xt = sak != 1, wt = sbk != 1;     // operand stored [Kc, rows] instead of [rows, Kc]
    if ((xt && sam != 1) || (wt && sbn != 1)) return 0;
    const int64_t ldx = xt ? sak : sam, ldw = wt ? sbk : sbn;
    if (!xt && (!GemmCtx::ptr_ok(A, ldx) || ((K & 3) && !a_padded))) return 0;
    if (a_padded && xt) return 0;
    if (!wt && (!GemmCtx::ptr_ok(B, ldw) || (K & 3))) return 0;
    if (!GemmCtx::ptr_ok(C, ldc)) return 0;
    const int Kp = (K + 7) / 8 * 8;
    const size_t esz = (size_t)tc_elem_bytes();
    Arena ar(scratch, cap);
    char* xh = ar.take<char>((size_t)M * Kp * esz);
    char* xl = ar.take<char>((size_t)M * Kp * esz);
    char* wh = ar.take<char>((size_t)N * Kp * esz);
    char* wl = ar.take<char>((size_t)N * Kp * esz);
    float* tmp = beta == 1.f ? ar.take<float>((size_t)M * N) : nullptr;
    if (ar.overflow) return 0;
    VAG_TRY(tc_split_pair(A, ldx, xt, M, xh, xl, B, ldw, wt, N, wh, wl, K, Kp, st));   // both operands, one launch, K padding zeroed
    VAG_TRY(tc_gemm(tmp ? tmp : C, tmp ? N : ldc, xh, xl, Kp, wh, wl, Kp, nullptr, M, Kp, N, 0, st, nullptr, nullptr));
    if (tmp) {
        VAG_CUDA(launch_pdl(PDL_SMALL, add2d_kernel, dim3(grid_for((int64_t)M * N / 4)), dim3(256), 0, st, C, ldc, tmp, M, N));
        VAG_LAUNCH_CHECK();
    }
    return 1;
}
static int gemm_g(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, int M,
                  int N, int K, float beta, cudaStream_t st) {
    const int r = gemm_tc_try(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, beta, g_tc_scratch.base, g_tc_scratch.cap, st);
    if (r != 0) return r < 0 ? r : VAG_OK;
    return gemm_f32(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, 1.0f, beta, (vag_stream_t)st);
}
// scratch the tensor-core route may need for one contraction of the given logical shape
static size_t gemm_tc_scratch_bytes(int64_t M, int64_t N, int64_t K) {
    const int64_t Kp = (K + 7) / 8 * 8;
    return 2 * align_up((size_t)M * Kp * 4, 256) + 2 * align_up((size_t)N * Kp * 4, 256) + align_up((size_t)M * N * 4, 256) + 4096;
}
}  // namespace vag

extern "C" size_t vag_gemm_tc_workspace_bytes(int M, int N, int K) { return gemm_tc_scratch_bytes(M, N, K); }

/* vag_gemm_f32 with a caller-owned workspace: large contractions run on the tcgen05 path (operands are split, and
   transposed where they are not contraction-contiguous, into the workspace); everything else takes the FFMA kernels. */
extern "C" int vag_gemm_tc_f32(float* C, int64_t ldc, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk,
                               int64_t sbn, int M, int N, int K, float alpha, float beta, int precision, void* workspace,
                               size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(precision);
    VAG_REQUIRE(C && A && B, "vag_gemm_tc_f32: null pointer");
    VAG_REQUIRE(M >= 0 && N >= 0 && K >= 0 && ldc >= N, "vag_gemm_tc_f32: bad shape");
    if (M == 0 || N == 0) return VAG_OK;
    if (alpha == 1.f) {
        const int r = gemm_tc_try(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, beta, (char*)workspace, workspace_bytes, (cudaStream_t)stream);
        if (r != 0) return r < 0 ? r : VAG_OK;
    }
    return gemm_f32(C, ldc, A, sam, sak, B, sbk, sbn, M, N, K, alpha, beta, stream);
}

// ----------------------------------------------------------------------------------------------------------
// Backward of the visual-attention pooling (VSE_Imagine_Enc.forward / ImagineAttn, layers/VSE_Imagine_Enc.py:29-152 under
// autograd): one call for the whole chain  l2norm ← tanh ← text_embedding ← Σβ·ctx ← softmax(score) ← ctx2ctx / emb2ctx ←
// l2norm ← tanh ← im_embedding.
// ----------------------------------------------------------------------------------------------------------
extern "C" size_t vag_vse_pool_bwd_workspace_bytes(int B, int T, int I, int C, int S) {
    const int64_t BT = (int64_t)B * T;
    return (size_t)B * S * 4 * 3 + (size_t)B * C * 4 * 2 + (size_t)BT * C * 4 + 16 * 256 +
           gemm_tc_scratch_bytes(std::max<int64_t>(BT, S), std::max<int64_t>(C, S), std::max<int64_t>({(int64_t)C, (int64_t)I, BT})) + 65536;
}

extern "C" int vag_vse_pool_bwd_f32(const vag_vse_weights* w, const float* im, const float* ctx, const float* mask, int B, int T,
                                    const vag_vse_saved* sv, const float* im_emb, const float* beta, const float* ctx_vec,
                                    const float* d_im_emb, const float* d_txt_emb, const float* d_ctx_vec, const vag_vse_grads* g,
                                    float* d_ctx, void* workspace, size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && im && ctx && sv && im_emb && beta && ctx_vec && g && d_ctx, "vag_vse_pool_bwd_f32: null pointer");
    VAG_REQUIRE(sv->a_im && sv->iq && sv->pk && sv->a_txt, "vag_vse_pool_bwd_f32: saved-activation pointers missing");
    VAG_REQUIRE(g->im_w && g->im_b && g->txt_w && g->txt_b && g->ctx2ctx_w && g->emb2ctx_w, "vag_vse_pool_bwd_f32: gradient pointers missing");
    VAG_REQUIRE(w->method == VAG_ATTN_DOT || (w->mlp_w && g->mlp_w), "vag_vse_pool_bwd_f32: the mlp attention needs its weight and gradient pointer");
    const int I = w->I, C = w->C, S = w->S, BT = B * T;
    cudaStream_t st = (cudaStream_t)stream;
    vag_stream_t vs = stream;
    Arena ar(workspace, workspace_bytes);
    float* du_t = ar.take<float>((size_t)B * S);
    float* du_i = ar.take<float>((size_t)B * S);
    float* d_ie = ar.take<float>((size_t)B * S);
    float* dcv = ar.take<float>((size_t)B * C);
    float* d_iq = ar.take<float>((size_t)B * C);
    float* dpk = ar.take<float>((size_t)BT * C);
    const size_t tcs_bytes = gemm_tc_scratch_bytes(std::max<int64_t>(BT, S), std::max<int64_t>(C, S), std::max<int64_t>({(int64_t)C, (int64_t)I, (int64_t)BT}));
    char* tcs = ar.take<char>(tcs_bytes);
    if (ar.overflow) {
        set_error("vag_vse_pool_bwd_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    TcScratchScope tc_scope(tcs, tcs_bytes);
    const bool mlp = w->method == VAG_ATTN_MLP;
    // ---- text branch: txt_emb = l2norm(a_txt), a_txt = act(text_embedding(ctx_vec))
    if (d_txt_emb) VAG_TRY(vag_l2norm_bwd_f32(du_t, d_txt_emb, sv->a_txt, B, S, vs));
    else VAG_CUDA(cudaMemsetAsync(du_t, 0, sizeof(float) * (size_t)B * S, st));
    if (w->activation) VAG_TRY(vag_tanh_bwd_f32(du_t, du_t, sv->a_txt, (int64_t)B * S, vs));
    VAG_TRY(gemm_g(g->txt_w, C, du_t, 1, S, ctx_vec, C, 1, S, C, B, 0.f, st));                 // du_tᵀ · ctx_vec
    VAG_TRY(vag_colsum_f32(g->txt_b, du_t, S, B, S, 0, vs));
    if (d_ctx_vec) VAG_CUDA(cudaMemcpyAsync(dcv, d_ctx_vec, sizeof(float) * (size_t)B * C, cudaMemcpyDeviceToDevice, st));
    else VAG_CUDA(cudaMemsetAsync(dcv, 0, sizeof(float) * (size_t)B * C, st));
    VAG_TRY(gemm_g(dcv, C, du_t, S, 1, w->txt_w, C, 1, B, C, S, 1.f, st));                      // dcv += du_t · W_txt
    // ---- pooling attention
    VAG_CUDA(cudaMemsetAsync(d_ctx, 0, sizeof(float) * (size_t)BT * C, st));
    VAG_CUDA(cudaMemsetAsync(dpk, 0, sizeof(float) * (size_t)BT * C, st));
    if (mlp) VAG_CUDA(cudaMemsetAsync(g->mlp_w, 0, sizeof(float) * (size_t)C, st));
    VAG_TRY(vag_attention_bwd_f32(d_iq, C, dpk, d_ctx, mlp ? g->mlp_w : nullptr, dcv, C, beta, sv->iq, C, sv->pk, ctx, w->mlp_w, mask, B, T, C,
                                  w->method, vs));
    VAG_TRY(gemm_g(d_ctx, C, dpk, C, 1, w->ctx2ctx_w, C, 1, BT, C, C, 1.f, st));                // d_ctx += dpk · W_ctx2ctx
    VAG_TRY(gemm_g(g->ctx2ctx_w, C, dpk, 1, C, ctx, C, 1, C, C, BT, 0.f, st));                  // dpkᵀ · ctx
    VAG_TRY(gemm_g(g->emb2ctx_w, S, d_iq, 1, C, im_emb, S, 1, C, S, B, 0.f, st));               // d_iqᵀ · im_emb
    // ---- image branch: im_emb = l2norm(a_im), a_im = act(im_embedding(im))
    if (d_im_emb) VAG_CUDA(cudaMemcpyAsync(d_ie, d_im_emb, sizeof(float) * (size_t)B * S, cudaMemcpyDeviceToDevice, st));
    else VAG_CUDA(cudaMemsetAsync(d_ie, 0, sizeof(float) * (size_t)B * S, st));
    VAG_TRY(gemm_g(d_ie, S, d_iq, C, 1, w->emb2ctx_w, S, 1, B, S, C, 1.f, st));                 // d_ie += d_iq · W_emb2ctx
    VAG_TRY(vag_l2norm_bwd_f32(du_i, d_ie, sv->a_im, B, S, vs));
    if (w->activation) VAG_TRY(vag_tanh_bwd_f32(du_i, du_i, sv->a_im, (int64_t)B * S, vs));
    VAG_TRY(gemm_g(g->im_w, I, du_i, 1, S, im, I, 1, S, I, B, 0.f, st));                        // du_iᵀ · im
    VAG_TRY(vag_colsum_f32(g->im_b, du_i, S, B, S, 0, vs));
    return VAG_OK;
}

extern "C" size_t vag_decoder_seq_workspace_bytes(int B, int T, int Tt, int E, int H, int C, int64_t V) {
    const int64_t R = (int64_t)B * Tt;
    size_t fwd = GemmCtx::split_bytes(3 * H, E) + 3 * GemmCtx::split_bytes(3 * H, H) + GemmCtx::split_bytes(C, H) +
                 GemmCtx::split_bytes(H, C) + GemmCtx::split_bytes(E, H) + GemmCtx::split_bytes(E, E) + GemmCtx::split_bytes(E, C) +
                 GemmCtx::split_bytes(V, E) + GemmCtx::split_bytes(C, C) + 65536;
    fwd += GemmCtx::split_bytes(R, E) * 2 + GemmCtx::split_bytes(R, H) + GemmCtx::split_bytes(R, C) + GemmCtx::split_bytes((int64_t)B * T, C) + 65536;
    fwd += dec_seq_scratch_bytes(H, C) + 4096;     // flags + exchange buffers of the persistent time loop
    // backward scratch: dlogits [R, V], d_t / du [R, E] x2, dH2_dir [R, H], dE [R, E], dC_dir [R, C], per-step grads
    size_t bwd = (size_t)R * (V + 4) * 4 + gemm_tc_scratch_bytes(std::max<int64_t>(V, 3 * H), std::max<int64_t>({(int64_t)C, (int64_t)3 * H, (int64_t)E}), std::max<int64_t>({(int64_t)R, (int64_t)B * T, (int64_t)3 * H})) +
                 gemm_tc_scratch_bytes(std::max<int64_t>(R, (int64_t)B * T), C, std::max<int64_t>(V + 8, 3 * H)) + (size_t)R * H * 4 + (size_t)R * E * 4 * 3 + (size_t)R * H * 4 * 2 + (size_t)R * C * 4 * 2 + (size_t)R * 3 * H * 4 * 4 +
                 (size_t)B * T * C * 4 + (size_t)B * H * 4 * 4 + (size_t)B * 3 * H * 4 * 4 + 65536;
    // … and a second operand-plane scratch for the weight-gradient phases, which may run on another stream than the recurrent part
    bwd += gemm_tc_scratch_bytes(std::max<int64_t>(V, 3 * H), std::max<int64_t>({(int64_t)C, (int64_t)3 * H, (int64_t)E}), std::max<int64_t>({(int64_t)R, (int64_t)B * T, (int64_t)3 * H})) +
           gemm_tc_scratch_bytes(std::max<int64_t>(R, (int64_t)B * T), C, std::max<int64_t>(V + 8, 3 * H)) + 4096;
    return fwd > bwd ? fwd : bwd;
}

extern "C" int vag_decoder_seq_fwd_f32(const vag_decoder_weights* w, const float* h0, const float* enc, const float* mask,
                                       int64_t* tok_in, const int64_t* tgt_t, const float* nll_weight, int B, int T, int Tt,
                                       int teacher, const vag_decoder_seq_saved* s, const float* out_mask, float* loss_rows,
                                       void* workspace, size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && h0 && enc && mask && tok_in && tgt_t && s && loss_rows, "vag_decoder_seq_fwd_f32: null pointer");
    VAG_REQUIRE(B > 0 && T > 0 && Tt > 0, "vag_decoder_seq_fwd_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = w->E, H = w->H, C = w->C;
    const int64_t V = w->V;
    const int R = B * Tt;
    const int64_t ldl = s->ld_logits;
    Arena ar(workspace, workspace_bytes);
    const size_t wb = GemmCtx::split_bytes(3 * H, E) + 3 * GemmCtx::split_bytes(3 * H, H) + GemmCtx::split_bytes(C, H) +
                      GemmCtx::split_bytes(H, C) + GemmCtx::split_bytes(E, H) + GemmCtx::split_bytes(E, E) + GemmCtx::split_bytes(E, C) +
                      GemmCtx::split_bytes(V, E) + GemmCtx::split_bytes(C, C) + 32768;
    void* wr = ar.take<char>(wb);
    const bool persistent = teacher && dec_seq_fwd_ok(B, T, Tt, H, C);     // ONE launch for the whole time loop (dec_seq.cu)
    const size_t ds_bytes = persistent ? dec_seq_scratch_bytes(H, C) : 0;
    void* ds_scratch = persistent ? ar.take<char>(ds_bytes) : nullptr;
    const size_t ab = ar.overflow ? 0 : (workspace_bytes > align_up(ar.off, 256) + 1024 ? workspace_bytes - align_up(ar.off, 256) - 512 : 0);
    void* areg = ab ? ar.take<char>(ab) : nullptr;
    GemmCtx gemm(st, ar.overflow ? nullptr : wr, wb, ar.overflow ? nullptr : areg, ab);
    vag_stream_t vs = stream;
    // hoisted keys
    VAG_TRY(gemm.linear(s->keys, C, enc, C, w->attn_e_w, C, nullptr, B * T, C, C, 0));
    if (teacher) {   // all input embeddings and their W_ih projection at once
        gemm.new_step();
        VAG_TRY(vag_embed_rows_f32(s->e_all, E, w->emb, E, tok_in, R, V, vs));
        VAG_TRY(gemm.linear(s->gi1_all, 3 * H, s->e_all, E, w->gru1_w_ih, E, w->gru1_b_ih, R, E, 3 * H, 0));
    }
    if (persistent) {
        if (ar.overflow) {
            set_error("vag_decoder_seq_fwd_f32: workspace too small");
            return VAG_ERR_WORKSPACE;
        }
        DecSeqFwd a = {w->gru1_w_hh, w->gru1_b_hh, w->attn_h_w, w->attn_v, w->c2h_w, w->gru2_w_ih, w->gru2_w_hh, w->gru2_b_ih, w->gru2_b_hh,
                       h0, s->keys, enc, mask, s->gi1_all,
                       s->gh1_all, s->h1_all, s->q_all, s->alpha_all, s->c_all, s->x2_all, s->gi2_all, s->gh2_all, s->h2_all,
                       B, T, Tt, H, C, nullptr, nullptr, 0};
        VAG_TRY(dec_seq_fwd(a, ds_scratch, ds_bytes, gemm_mode() == 2, st));
    }
    const float* h = h0;
    for (int t = 0; !persistent && t < Tt; ++t) {
        gemm.new_step();
        float* e_s = s->e_all + (size_t)t * B * E;
        float* gi1 = s->gi1_all + (size_t)t * B * 3 * H;
        float* gh1 = s->gh1_all + (size_t)t * B * 3 * H;
        float* h1 = s->h1_all + (size_t)t * B * H;
        float* q = s->q_all + (size_t)t * B * C;
        float* c = s->c_all + (size_t)t * B * C;
        float* x2 = s->x2_all + (size_t)t * B * H;
        float* gi2 = s->gi2_all + (size_t)t * B * 3 * H;
        float* gh2 = s->gh2_all + (size_t)t * B * 3 * H;
        float* h2 = s->h2_all + (size_t)t * B * H;
        if (!teacher) {
            VAG_TRY(vag_embed_rows_f32(e_s, E, w->emb, E, tok_in + (size_t)t * B, B, V, vs));
            VAG_TRY(gemm.linear(gi1, 3 * H, e_s, E, w->gru1_w_ih, E, w->gru1_b_ih, B, E, 3 * H, 0));
        }
        // gru_1: hidden-side contraction with the cell's gate math in its epilogue (one launch instead of two)
        const Rows32Gru g1 = {gh1, w->gru1_b_hh, gi1, 0, h, h1, nullptr, 0, nullptr, 0, H, {h, w->gru1_w_hh, H, H, H}};
        const bool fuse1 = rows32_gru_ok(g1, B);
        if (fuse1) {
            VAG_TRY(linear_rows32_gru(&g1, 1, B, gemm_mode() == 2, st));
        } else {
            VAG_TRY(gemm.linear(gh1, 3 * H, h, H, w->gru1_w_hh, H, w->gru1_b_hh, B, H, 3 * H, 0));
            VAG_TRY(vag_gru_gates_f32(h1, H, nullptr, 0, gi1, 3 * H, gh1, 3 * H, h, H, B, H, vs));
        }
        // the two contractions that read h1 — attention query and gru_2's hidden pre-activations — share one launch
        const Rows32Problem h1p[2] = {{q, nullptr, C, C, 1, {{h1, w->attn_h_w, H, H, H}, {}}},
                                      {gh2, w->gru2_b_hh, 3 * H, 3 * H, 1, {{h1, w->gru2_w_hh, H, H, H}, {}}}};
        const bool h1_pair = rows32_problem_ok(h1p[0], B, true) && rows32_problem_ok(h1p[1], B, true);
        if (h1_pair) VAG_TRY(linear_rows32_multi(h1p, 2, B, 0, true, gemm_mode() == 2, st));
        else VAG_TRY(gemm.linear(q, C, h1, H, w->attn_h_w, H, nullptr, B, H, C, 0));
        VAG_TRY(vag_attention_f32(c, C, s->alpha_all + (size_t)t * B * T, q, C, s->keys, enc, w->attn_v, mask, B, 1, T, C, VAG_ATTN_MLP, vs));
        VAG_TRY(gemm.linear(x2, H, c, C, w->c2h_w, C, nullptr, B, C, H, 0));
        if (!h1_pair) VAG_TRY(gemm.linear(gh2, 3 * H, h1, H, w->gru2_w_hh, H, w->gru2_b_hh, B, H, 3 * H, 0));
        // gru_2: input-side contraction (x2 → gi2) with the gate math in its epilogue; gh2 is already there
        const Rows32Gru g2 = {gi2, w->gru2_b_ih, gh2, 1, h1, h2, nullptr, 0, nullptr, 0, H, {x2, w->gru2_w_ih, H, H, H}};
        if (rows32_gru_ok(g2, B)) {
            VAG_TRY(linear_rows32_gru(&g2, 1, B, gemm_mode() == 2, st));
        } else {
            VAG_TRY(gemm.linear(gi2, 3 * H, x2, H, w->gru2_w_ih, H, w->gru2_b_ih, B, H, 3 * H, 0));
            VAG_TRY(vag_gru_gates_f32(h2, H, nullptr, 0, gi2, 3 * H, gh2, 3 * H, h1, H, B, H, vs));
        }
        h = h2;
        if (!teacher) {   // free running (V11:149-160): this step's arg-max is the next input
            float* t_s = s->t_all + (size_t)t * B * E;
            float* lg = s->logits_all + (size_t)t * B * ldl;
            VAG_TRY(gemm.linear(t_s, E, h2, H, w->w1_w, H, w->w1_b, B, H, E, 0));
            VAG_TRY(gemm.linear(t_s, E, e_s, E, w->w3_w, E, w->w3_b, B, E, E, VAG_LIN_ACCUMULATE));
            VAG_TRY(gemm.linear(t_s, E, c, C, w->w2_w, C, w->w2_b, B, C, E, VAG_LIN_ACCUMULATE | VAG_LIN_TANH));
            if (out_mask) VAG_TRY(vag_mul_f32(t_s, out_mask + (size_t)t * B * E, (int64_t)B * E, vs));   // output dropout, NMT_Decoder.py:140-141
            VAG_TRY(gemm.linear(lg, ldl, t_s, E, w->out_w, E, w->out_b, B, E, (int)V, 0));
            if (t + 1 < Tt) VAG_TRY(row_argmax(lg, ldl, B, V, tok_in + (size_t)(t + 1) * B, 1, nullptr, st));
        }
    }
    if (teacher) {   // read-out and vocabulary projection for all steps at once
        gemm.new_step();
        {   // t = tanh(W1 h2 + W3 e + W2 c + b1 + b3 + b2) as ONE contraction over the concatenated [h2 | e | c] (NMT_Decoder.py:137)
            const float* const xs[3] = {s->h2_all, s->e_all, s->c_all};
            const int64_t ldxs[3] = {H, E, C};
            const int Ks[3] = {H, E, C};
            const float* const wsp[3] = {w->w1_w, w->w3_w, w->w2_w};
            const int64_t ldws[3] = {H, E, C};
            const float* const bsp[3] = {w->w1_b, w->w3_b, w->w2_b};
            VAG_TRY(gemm.linear3(s->t_all, E, xs, ldxs, Ks, wsp, ldws, bsp, R, E, VAG_LIN_TANH));
        }
        if (out_mask) VAG_TRY(vag_mul_f32(s->t_all, out_mask, (int64_t)R * E, vs));                       // output dropout
        gemm.new_step();
        VAG_TRY(gemm.linear(s->logits_all, ldl, s->t_all, E, w->out_w, E, w->out_b, R, E, (int)V, 0));
    }
    // NLL of all Tt·B rows in one launch + a per-sentence sum over the steps in time order (V11:140-141)
    float* nll_scratch = (float*)wr;      // the weight-plane region is free again: every contraction has been enqueued
    VAG_REQUIRE(wr && !ar.overflow && wb >= sizeof(float) * (size_t)R, "vag_decoder_seq_fwd_f32: workspace too small for the NLL rows");
    VAG_TRY(nll_rows_all(s->logits_all, ldl, tgt_t, nll_weight, B, Tt, V, loss_rows, s->lse_all, nll_scratch, st));
    return VAG_OK;
}

extern "C" int vag_decoder_seq_bwd_f32(const vag_decoder_weights* w, const float* h0, const float* enc, const float* mask,
                                       const int64_t* tok_in, const int64_t* tgt_t, const float* nll_weight, int B, int T, int Tt,
                                       int tied, const vag_decoder_seq_saved* s, const float* out_mask, const float* dloss_rows,
                                       const vag_decoder_grads* g, float* d_h0, float* d_enc, int phases, void* workspace,
                                       size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && h0 && enc && mask && tok_in && tgt_t && s && dloss_rows && g && d_h0 && d_enc, "vag_decoder_seq_bwd_f32: null pointer");
    VAG_REQUIRE(phases >= 1 && phases <= 15, "vag_decoder_seq_bwd_f32: phases is a mask of 1 (head) | 2 (recurrent part) | 4 (read-out weight gradients) | 8 (other weight gradients)");
    const bool do_head = phases & 1, do_rec = phases & 2, do_wg_ro = phases & 4, do_wg = phases & 8;
    cudaStream_t st = (cudaStream_t)stream;
    vag_stream_t vs = stream;
    const int E = w->E, H = w->H, C = w->C;
    const int64_t V = w->V;
    const int R = B * Tt;
    const int64_t ldl = s->ld_logits;
    Arena ar(workspace, workspace_bytes);
    const int64_t ldd = (V + 3) / 4 * 4;                      // dlogits pitch: rows zero-padded to a multiple of 4
    float* dlogits = ar.take<float>((size_t)R * ldd);
    float* hprev_all = ar.take<float>((size_t)R * H);         // [h0; h2_0 … h2_{Tt-2}]: what gru_1 saw as its previous state
    const size_t tcs_bytes = gemm_tc_scratch_bytes(std::max<int64_t>(V, 3 * H), std::max<int64_t>({(int64_t)C, (int64_t)3 * H, (int64_t)E}),
                                                   std::max<int64_t>({(int64_t)R, (int64_t)B * T, (int64_t)3 * H})) +
                             gemm_tc_scratch_bytes(std::max<int64_t>(R, (int64_t)B * T), C, std::max<int64_t>(V + 8, 3 * H));
    char* tcs = ar.take<char>(tcs_bytes);
    char* tcs_wg = ar.take<char>(tcs_bytes);
    // the weight-gradient phases, called on their own, may overlap the recurrent part on another stream: private operand planes
    const bool wg_only = !(phases & 3);
    TcScratchScope tc_scope(wg_only ? tcs_wg : tcs, (wg_only ? tcs_wg : tcs) ? tcs_bytes : 0);
    float* d_t = ar.take<float>((size_t)R * E);
    float* du = ar.take<float>((size_t)R * E);
    float* dh2_dir = ar.take<float>((size_t)R * H);
    float* d_e = ar.take<float>((size_t)R * E);
    float* dc_dir = ar.take<float>((size_t)R * C);
    float* dgi2_all = ar.take<float>((size_t)R * 3 * H);
    float* dgh2_all = ar.take<float>((size_t)R * 3 * H);
    float* dgi1_all = ar.take<float>((size_t)R * 3 * H);
    float* dgh1_all = ar.take<float>((size_t)R * 3 * H);
    float* dx2_all = ar.take<float>((size_t)R * H);
    float* dq_all = ar.take<float>((size_t)R * C);
    float* dkeys = ar.take<float>((size_t)B * T * C);
    float* dh1 = ar.take<float>((size_t)B * H);
    float* dh_next = ar.take<float>((size_t)B * H);
    if (ar.overflow) {
        set_error("vag_decoder_seq_bwd_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    if (do_head) {
    // ---- batched over all steps: vocabulary projection and read-out
    // (the pad column of dlogits is never read: the operand split, the column sum and the transposed read all stop at V)
    VAG_CUDA(launch_pdl(PDL_SMALL, nll_bwd_kernel, dim3(R), dim3(256), 0, st, dlogits, ldd, s->logits_all, ldl, s->lse_all, tgt_t, nll_weight, dloss_rows, V, B));   // all steps
    VAG_LAUNCH_CHECK();
    {   // d_t = dlogits · out_w   (contraction over the vocabulary: rows of dlogits are padded, out_w is split transposed)
        const int r = gemm_tc_try(d_t, E, dlogits, ldd, 1, w->out_w, E, 1, R, E, (int)V, 0.f, g_tc_scratch.base, g_tc_scratch.cap, st, true);
        if (r < 0) return r;
        if (r == 0) VAG_TRY(gemm_g(d_t, E, dlogits, ldd, 1, w->out_w, E, 1, R, E, (int)V, 0.f, st));
    }
    if (out_mask) {
        VAG_CUDA(launch_pdl(PDL_SMALL, tanh_dropout_bwd_kernel, dim3(grid_for((int64_t)R * E)), dim3(256), 0, st, du, d_t, s->t_all, out_mask, (int64_t)R * E));
        VAG_LAUNCH_CHECK();
    } else {
        VAG_TRY(vag_tanh_bwd_f32(du, d_t, s->t_all, (int64_t)R * E, vs));
    }
    VAG_TRY(gemm_g(dh2_dir, H, du, E, 1, w->w1_w, H, 1, R, H, E, 0.f, st));
    VAG_TRY(gemm_g(d_e, E, du, E, 1, w->w3_w, E, 1, R, E, E, 0.f, st));
    VAG_TRY(gemm_g(dc_dir, C, du, E, 1, w->w2_w, C, 1, R, C, E, 0.f, st));
    }
    if (do_rec) {
    // ---- recurrent part, reverse time
    VAG_CUDA(cudaMemsetAsync(dkeys, 0, sizeof(float) * (size_t)B * T * C, st));
    VAG_CUDA(cudaMemsetAsync(d_enc, 0, sizeof(float) * (size_t)B * T * C, st));
    VAG_CUDA(cudaMemsetAsync(g->attn_v, 0, sizeof(float) * C, st));
    VAG_CUDA(cudaMemsetAsync(dh_next, 0, sizeof(float) * (size_t)B * H, st));
    // fused flavour of the loop below: the two contractions that finish a cell's incoming gradient carry that cell's gate
    // backward in their epilogue (dh1 → gru_1 at step t; dh_next + read-out path → gru_2 at step t-1): 5 launches per step
    Rows32GruBwd dprobe = {};
    dprobe.seg[0] = Rows32Seg{dgh2_all, w->gru2_w_hh, 3 * H, H, 3 * H};
    dprobe.seg[1] = Rows32Seg{dq_all, w->attn_h_w, C, H, C};
    dprobe.nseg = 2; dprobe.gi = s->gi1_all; dprobe.gh = s->gh1_all; dprobe.dgi = dgi1_all; dprobe.dgh = dgh1_all; dprobe.dh_out = dh_next; dprobe.H = H;
    const bool fused_dec = rows32_gru_bwd_ok(dprobe, B);
    if (fused_dec) {
        const size_t last = (size_t)(Tt - 1) * B;
        VAG_CUDA(launch_pdl(PDL_SMALL, gru_gates_bwd_kernel, dim3(grid_for((int64_t)B * H)), dim3(256), 0, st, dgi2_all + last * 3 * H, dgh2_all + last * 3 * H, dh1, dh2_dir + last * H, H,
                                                                        s->gi2_all + last * 3 * H, s->gh2_all + last * 3 * H,
                                                                        s->h1_all + last * H, H, B, H, dh_next));
        VAG_LAUNCH_CHECK();
    }
    for (int t = Tt - 1; fused_dec && t >= 0; --t) {
        const size_t o3 = (size_t)t * B * 3 * H, oh = (size_t)t * B * H, oc = (size_t)t * B * C;
        float* dx2 = dx2_all + oh;
        VAG_TRY(gemm_g(dx2, H, dgi2_all + o3, 3 * H, 1, w->gru2_w_ih, H, 1, B, H, 3 * H, 0.f, st));           // dx2 = dgi2 · W_ih2
        float* dc = dc_dir + oc;
        VAG_TRY(gemm_g(dc, C, dx2, H, 1, w->c2h_w, C, 1, B, C, H, 1.f, st));                                    // dc += dx2 · W_c2h
        float* dq = dq_all + oc;
        VAG_TRY(vag_attention_bwd_f32(dq, C, dkeys, d_enc, g->attn_v, dc, C, s->alpha_all + (size_t)t * B * T, s->q_all + oc,
                                      C, s->keys, enc, w->attn_v, mask, B, T, C, VAG_ATTN_MLP, vs));
        Rows32GruBwd a = {};      // g = dh1 + dgh2·W_hh2 + dq·W_attn_h → gru_1 backward at step t → dgi1, dgh1, dh_next = g·z
        a.seg[0] = Rows32Seg{dgh2_all + o3, w->gru2_w_hh, 3 * H, H, 3 * H};
        a.seg[1] = Rows32Seg{dq, w->attn_h_w, C, H, C};
        a.nseg = 2;
        a.base = dh1; a.ld_base = H;
        a.gi = s->gi1_all + o3; a.gh = s->gh1_all + o3;
        a.h_prev = t == 0 ? h0 : s->h2_all + (size_t)(t - 1) * B * H; a.ld_hprev = H;
        a.dgi = dgi1_all + o3; a.dgh = dgh1_all + o3; a.dh_out = dh_next; a.H = H;
        VAG_TRY(linear_rows32_gru_bwd(&a, 1, B, gemm_mode() == 2, st));
        if (t > 0) {              // g = dh_next + dgh1·W_hh1 + read-out path of step t-1 → gru_2 backward at step t-1 → dgi2, dgh2, dh1 = g·z
            const size_t p3 = (size_t)(t - 1) * B * 3 * H, ph = (size_t)(t - 1) * B * H;
            Rows32GruBwd b2 = {};
            b2.seg[0] = Rows32Seg{dgh1_all + o3, w->gru1_w_hh, 3 * H, H, 3 * H};
            b2.nseg = 1;
            b2.base = dh_next; b2.ld_base = H;
            b2.add2 = dh2_dir + ph; b2.ld_add2 = H;
            b2.gi = s->gi2_all + p3; b2.gh = s->gh2_all + p3;
            b2.h_prev = s->h1_all + ph; b2.ld_hprev = H;
            b2.dgi = dgi2_all + p3; b2.dgh = dgh2_all + p3; b2.dh_out = dh1; b2.H = H;
            VAG_TRY(linear_rows32_gru_bwd(&b2, 1, B, gemm_mode() == 2, st));
        } else {
            VAG_TRY(gemm_g(dh_next, H, dgh1_all + o3, 3 * H, 1, w->gru1_w_hh, H, 1, B, H, 3 * H, 1.f, st));   // → d h0
        }
    }
    for (int t = Tt - 1; !fused_dec && t >= 0; --t) {
        const float* dh2 = dh2_dir + (size_t)t * B * H;     // read-out path; the recurrent path dh_next is added inside the kernel
        float* dgi2 = dgi2_all + (size_t)t * B * 3 * H;
        float* dgh2 = dgh2_all + (size_t)t * B * 3 * H;
        const float* h1 = s->h1_all + (size_t)t * B * H;
        VAG_CUDA(launch_pdl(PDL_SMALL, gru_gates_bwd_kernel, dim3(grid_for((int64_t)B * H)), dim3(256), 0, st, dgi2, dgh2, dh1, dh2, H, s->gi2_all + (size_t)t * B * 3 * H,
                                                                        s->gh2_all + (size_t)t * B * 3 * H, h1, H, B, H, dh_next));
        VAG_LAUNCH_CHECK();
        float* dx2 = dx2_all + (size_t)t * B * H;
        VAG_TRY(gemm_g(dx2, H, dgi2, 3 * H, 1, w->gru2_w_ih, H, 1, B, H, 3 * H, 0.f, st));             // dx2 = dgi2 · W_ih2
        float* dc = dc_dir + (size_t)t * B * C;
        VAG_TRY(gemm_g(dc, C, dx2, H, 1, w->c2h_w, C, 1, B, C, H, 1.f, st));                            // dc += dx2 · W_c2h
        float* dq = dq_all + (size_t)t * B * C;
        VAG_TRY(vag_attention_bwd_f32(dq, C, dkeys, d_enc, g->attn_v, dc, C, s->alpha_all + (size_t)t * B * T, s->q_all + (size_t)t * B * C,
                                      C, s->keys, enc, w->attn_v, mask, B, T, C, VAG_ATTN_MLP, vs));
        {   // dh1 += dgh2 · W_hh2 + dq · W_attn_h: one launch, two K-segments accumulating into the same output
            const Rows32Problem dp = {dh1, nullptr, H, H, 2, {{dgh2, w->gru2_w_hh, 3 * H, H, 3 * H}, {dq, w->attn_h_w, C, H, C}}};
            if (rows32_problem_ok(dp, B, false)) {
                VAG_TRY(linear_rows32_multi(&dp, 1, B, VAG_LIN_ACCUMULATE, false, gemm_mode() == 2, st));
            } else {
                VAG_TRY(gemm_g(dh1, H, dgh2, 3 * H, 1, w->gru2_w_hh, H, 1, B, H, 3 * H, 1.f, st));
                VAG_TRY(gemm_g(dh1, H, dq, C, 1, w->attn_h_w, H, 1, B, H, C, 1.f, st));
            }
        }
        const float* h_prev = t == 0 ? h0 : s->h2_all + (size_t)(t - 1) * B * H;
        float* dgi1 = dgi1_all + (size_t)t * B * 3 * H;
        float* dgh1 = dgh1_all + (size_t)t * B * 3 * H;
        VAG_TRY(vag_gru_gates_bwd_f32(dgi1, dgh1, dh_next, dh1, H, s->gi1_all + (size_t)t * B * 3 * H, s->gh1_all + (size_t)t * B * 3 * H, h_prev, H, B, H, vs));
        VAG_TRY(gemm_g(dh_next, H, dgh1, 3 * H, 1, w->gru1_w_hh, H, 1, B, H, 3 * H, 1.f, st));          // dh_prev = dh1·z + dgh1 · W_hh1
    }
    VAG_CUDA(cudaMemcpyAsync(d_h0, dh_next, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    // ---- hoisted keys: the path back into the encoder context
    VAG_TRY(gemm_g(d_enc, C, dkeys, C, 1, w->attn_e_w, C, 1, B * T, C, C, 1.f, st));
    }
    if (do_wg_ro) {
    // Weight gradients: contractions over all Tt·B rows of what the recurrent part left in the workspace.  Nothing downstream of
    // the decoder reads them before the optimiser, so a caller may run this phase on a second stream while the encoder's
    // back-propagation proceeds (autograd.DecoderSeqFn).
    float* d_out_w = tied ? g->emb : g->out_w;                                                       // tied: accumulate into dEmb later
    if (!tied) VAG_CUDA(cudaMemsetAsync(g->emb, 0, sizeof(float) * (size_t)V * E, st));            // tied: the contraction below overwrites all of it
    VAG_TRY(gemm_g(d_out_w, E, dlogits, 1, ldd, s->t_all, E, 1, (int)V, E, R, 0.f, st));            // dlogitsᵀ · t_all
    VAG_TRY(vag_colsum_f32(g->out_b, dlogits, ldd, R, (int)V, 0, vs));
    VAG_TRY(gemm_g(g->w1_w, H, du, 1, E, s->h2_all, H, 1, E, H, R, 0.f, st));
    VAG_TRY(gemm_g(g->w3_w, E, du, 1, E, s->e_all, E, 1, E, E, R, 0.f, st));
    VAG_TRY(gemm_g(g->w2_w, C, du, 1, E, s->c_all, C, 1, E, C, R, 0.f, st));
    VAG_TRY(vag_colsum_f32(g->w1_b, du, E, R, E, 0, vs));
    VAG_CUDA(cudaMemcpyAsync(g->w2_b, g->w1_b, sizeof(float) * E, cudaMemcpyDeviceToDevice, st));
    VAG_CUDA(cudaMemcpyAsync(g->w3_b, g->w1_b, sizeof(float) * E, cudaMemcpyDeviceToDevice, st));
    }
    if (do_wg) {
    // ---- weight gradients, batched over steps
    VAG_TRY(gemm_g(g->gru2_w_ih, H, dgi2_all, 1, 3 * H, s->x2_all, H, 1, 3 * H, H, R, 0.f, st));
    VAG_TRY(vag_colsum_f32(g->gru2_b_ih, dgi2_all, 3 * H, R, 3 * H, 0, vs));
    VAG_TRY(gemm_g(g->gru2_w_hh, H, dgh2_all, 1, 3 * H, s->h1_all, H, 1, 3 * H, H, R, 0.f, st));
    VAG_TRY(vag_colsum_f32(g->gru2_b_hh, dgh2_all, 3 * H, R, 3 * H, 0, vs));
    VAG_TRY(gemm_g(g->c2h_w, C, dx2_all, 1, H, s->c_all, C, 1, H, C, R, 0.f, st));
    VAG_TRY(gemm_g(g->attn_h_w, H, dq_all, 1, C, s->h1_all, H, 1, C, H, R, 0.f, st));
    VAG_TRY(gemm_g(g->gru1_w_ih, E, dgi1_all, 1, 3 * H, s->e_all, E, 1, 3 * H, E, R, 0.f, st));
    VAG_TRY(vag_colsum_f32(g->gru1_b_ih, dgi1_all, 3 * H, R, 3 * H, 0, vs));
    VAG_CUDA(cudaMemcpyAsync(hprev_all, h0, sizeof(float) * (size_t)B * H, cudaMemcpyDeviceToDevice, st));
    if (Tt > 1)
        VAG_CUDA(cudaMemcpyAsync(hprev_all + (size_t)B * H, s->h2_all, sizeof(float) * (size_t)(Tt - 1) * B * H, cudaMemcpyDeviceToDevice, st));
    VAG_TRY(gemm_g(g->gru1_w_hh, H, dgh1_all, 1, 3 * H, hprev_all, H, 1, 3 * H, H, R, 0.f, st));
    VAG_TRY(vag_colsum_f32(g->gru1_b_hh, dgh1_all, 3 * H, R, 3 * H, 0, vs));
    VAG_TRY(gemm_g(d_e, E, dgi1_all, 3 * H, 1, w->gru1_w_ih, E, 1, R, E, 3 * H, 1.f, st));              // de += dgi1 · W_ih1
    VAG_TRY(vag_embed_bwd_f32(g->emb, d_e, E, tok_in, R, E, V, vs));                                     // (+ dOutW already inside when tied)
    // ---- hoisted keys: dW_attn_e
    VAG_TRY(gemm_g(g->attn_e_w, C, dkeys, 1, C, enc, C, 1, C, C, B * T, 0.f, st));
    }
    return VAG_OK;
}

// ----------------------------------------------------------------------------------------------------------
// Encoder training pair: forward that keeps what BPTT needs (time-major embeddings, gi, per-step gh), and the
// backward through both directions of the packed bi-GRU (layers/Encoder.py:36-65 under autograd).
// ----------------------------------------------------------------------------------------------------------
namespace vag {
__global__ void encoder_embed_tm_kernel(float* __restrict__ out, const float* __restrict__ table, const int64_t* __restrict__ src,
                                        int64_t* __restrict__ ids_tm, int B, int T, int E, int64_t vocab) {
    const int row = blockIdx.x * blockDim.y + threadIdx.y;  // t*B + b
    if (row >= B * T) return;
    const int t = row / B, b = row % B;
    int64_t id = src[(int64_t)b * T + t];
    if (threadIdx.x == 0) ids_tm[row] = id;
    if (id < 0 || id >= vocab) id = 0;
    const float* s = table + id * E;
    float* d = out + (int64_t)row * E;
    for (int c = threadIdx.x; c < E; c += blockDim.x) d[c] = s[c];
}
// dst[r, :] = src[r*ld_src + :]  rows × cols copy with row pitches (strided slice → contiguous and back)
__global__ void copy2d_kernel(float* __restrict__ dst, int64_t ld_dst, const float* __restrict__ src, int64_t ld_src, int rows, int cols) {
    const int64_t total = (int64_t)rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / cols), c = (int)(i % cols);
        dst[(int64_t)r * ld_dst + c] = src[(int64_t)r * ld_src + c];
    }
}
static int copy2d(float* dst, int64_t ld_dst, const float* src, int64_t ld_src, int rows, int cols, cudaStream_t st) {
    if (rows == 0) return VAG_OK;
    copy2d_kernel<<<grid_for((int64_t)rows * cols), 256, 0, st>>>(dst, ld_dst, src, ld_src, rows, cols);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
}  // namespace vag

namespace vag {

// One recurrent step of the packed bidirectional GRU for BOTH directions (blockIdx.y = direction; direction 0 is at time s,
// direction 1 at time T-1-s).  Rows whose sentence is shorter than the time index are masked on the device — the launch
// sequence does not depend on the lengths, so a captured CUDA graph of the step serves every batch of the same [B, T] shape:
// the state of a masked row is left alone (0 before a reverse chain starts), its context stays 0 (pad_packed_sequence,
// Encoder.py:60) and its saved hidden pre-activations are zeroed.
__global__ void __launch_bounds__(256)
enc_gates_fwd_kernel(float* __restrict__ h, float* __restrict__ ctx_out, const float* __restrict__ gi, float* __restrict__ gh,
                     const int32_t* __restrict__ lengths, int B, int T, int H, int s) {
    pdl_trigger();
    pdl_wait();
    const int d = blockIdx.y, t = d == 0 ? s : T - 1 - s;
    const int per_row = H >> 2;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * per_row; idx += gridDim.x * blockDim.x) {
        const int row = idx / per_row, j = (idx % per_row) << 2;
        const int64_t o3 = (((int64_t)d * T + t) * B + row) * 3 * H;
        float* ghr = gh + o3;
        if (lengths[row] <= t) {
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            *reinterpret_cast<float4*>(ghr + j) = z4;
            *reinterpret_cast<float4*>(ghr + H + j) = z4;
            *reinterpret_cast<float4*>(ghr + 2 * H + j) = z4;
            continue;
        }
        const float* gir = gi + o3;
        float ir[4], iz[4], in_[4], hr[4], hz[4], hn[4], hp[4], out[4];
        *reinterpret_cast<float4*>(ir) = *reinterpret_cast<const float4*>(gir + j);
        *reinterpret_cast<float4*>(iz) = *reinterpret_cast<const float4*>(gir + H + j);
        *reinterpret_cast<float4*>(in_) = *reinterpret_cast<const float4*>(gir + 2 * H + j);
        *reinterpret_cast<float4*>(hr) = *reinterpret_cast<const float4*>(ghr + j);
        *reinterpret_cast<float4*>(hz) = *reinterpret_cast<const float4*>(ghr + H + j);
        *reinterpret_cast<float4*>(hn) = *reinterpret_cast<const float4*>(ghr + 2 * H + j);
        float* hrow = h + ((int64_t)d * B + row) * H + j;
        *reinterpret_cast<float4*>(hp) = *reinterpret_cast<const float4*>(hrow);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float r = sigmoidf_precise(ir[u] + hr[u]);
            const float z = sigmoidf_precise(iz[u] + hz[u]);
            const float n = tanhf(in_[u] + r * hn[u]);
            out[u] = (1.0f - z) * n + z * hp[u];
        }
        *reinterpret_cast<float4*>(hrow) = *reinterpret_cast<float4*>(out);
        *reinterpret_cast<float4*>(ctx_out + ((int64_t)row * T + t) * 2 * H + (int64_t)d * H + j) = *reinterpret_cast<float4*>(out);
    }
}

// Backward of one step, both directions (direction 0 walks t = T-1 … 0, direction 1 walks t = 0 … T-1).  dh = dctx[t] + carry;
// writes this step's dgi / dgh (zeros for masked rows), the previous state the step saw (read back from ctx: exactly 0 where
// a chain starts), and carry = dh·z — the recurrent contraction then adds dgh·W_hh onto it.
__global__ void __launch_bounds__(256)
enc_gates_bwd_kernel(float* __restrict__ dgi_all, float* __restrict__ dgh_all, float* __restrict__ hprev_all, float* __restrict__ carry,
                     const float* __restrict__ dctx, const float* __restrict__ ctx, const float* __restrict__ gi,
                     const float* __restrict__ gh, const int32_t* __restrict__ lengths, int B, int T, int H, int s) {
    pdl_trigger();
    pdl_wait();
    const int d = blockIdx.y, t = d == 0 ? T - 1 - s : s, tp = d == 0 ? t - 1 : t + 1;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < B * H; idx += gridDim.x * blockDim.x) {
        const int row = idx / H, j = idx % H;
        const int64_t o3 = (((int64_t)d * T + t) * B + row) * 3 * H;
        const int64_t oh = (((int64_t)d * T + t) * B + row) * H + j;
        const float hp = (tp >= 0 && tp < T) ? ctx[((int64_t)row * T + tp) * 2 * H + (int64_t)d * H + j] : 0.f;
        hprev_all[oh] = hp;
        float* a = dgi_all + o3;
        float* b = dgh_all + o3;
        float* cr = carry + ((int64_t)d * B + row) * H + j;
        if (lengths[row] <= t) {
            a[j] = 0.f; a[H + j] = 0.f; a[2 * H + j] = 0.f;
            b[j] = 0.f; b[H + j] = 0.f; b[2 * H + j] = 0.f;
            *cr = 0.f;
            continue;
        }
        const float* gir = gi + o3;
        const float* ghr = gh + o3;
        const float r = sigmoidf_precise(gir[j] + ghr[j]);
        const float z = sigmoidf_precise(gir[H + j] + ghr[H + j]);
        const float hn = ghr[2 * H + j];
        const float n = tanhf(gir[2 * H + j] + r * hn);
        const float g = dctx[((int64_t)row * T + t) * 2 * H + (int64_t)d * H + j] + *cr;
        const float dn_pre = g * (1.f - z) * (1.f - n * n);
        const float dz_pre = g * (hp - n) * z * (1.f - z);
        const float dr_pre = dn_pre * hn * r * (1.f - r);
        a[j] = dr_pre; a[H + j] = dz_pre; a[2 * H + j] = dn_pre;
        b[j] = dr_pre; b[H + j] = dz_pre; b[2 * H + j] = dn_pre * r;
        *cr = g * z;
    }
}
}  // namespace vag

extern "C" size_t vag_encoder_train_workspace_bytes(int B, int T, int E, int H) {
    return 2 * (GemmCtx::split_bytes(3 * H, E) + GemmCtx::split_bytes(3 * H, H)) + GemmCtx::split_bytes((int64_t)T * B, E) +
           (size_t)T * B * 3 * H * 4 * 4 + (size_t)T * B * H * 4 * 2 + (size_t)B * H * 4 * 8 + (size_t)B * 3 * H * 4 * 4 + 131072 +
           (size_t)T * B * E * 4 + gemm_tc_scratch_bytes(3 * H, std::max(E, H), (int64_t)T * B) + gemm_tc_scratch_bytes((int64_t)T * B, E, 3 * H) +
           enc_seq_scratch_bytes(H) + 1024;
}

static int check_lengths_host(const int32_t* lengths_host, int B, int T, const char* who) {
    if (!lengths_host) return VAG_OK;
    for (int b = 0; b < B; ++b) {
        VAG_REQUIRE(lengths_host[b] >= 1 && lengths_host[b] <= T, "%s: bad length", who);
        VAG_REQUIRE(b == 0 || lengths_host[b] <= lengths_host[b - 1], "%s: lengths must be sorted in decreasing order", who);
    }
    return VAG_OK;
}

/* saved: x [T·B, E] time-major embeddings, ids_tm int64 [T·B], gi [2][T, B, 3H], gh [2][T, B, 3H] (zero where inactive).
   lengths_dev int32 [B] on the device drives the masking; lengths_host (optional) is only validated. */
extern "C" int vag_encoder_train_fwd_f32(const vag_encoder_weights* w, const int64_t* src, const int32_t* lengths_host,
                                         const int32_t* lengths_dev, int B, int T, float* ctx_out, float* x, int64_t* ids_tm,
                                         float* gi, float* gh, const float* emb_mask, void* workspace, size_t workspace_bytes,
                                         vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && src && lengths_dev && ctx_out && x && ids_tm && gi && gh, "vag_encoder_train_fwd_f32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int E = w->E, H = w->H;
    VAG_REQUIRE((H & 3) == 0, "vag_encoder_train_fwd_f32: H must be a multiple of 4");
    VAG_TRY(check_lengths_host(lengths_host, B, T, "vag_encoder_train_fwd_f32"));
    Arena ar(workspace, workspace_bytes);
    const size_t wb = 2 * (GemmCtx::split_bytes(3 * H, E) + GemmCtx::split_bytes(3 * H, H)) + 8192;
    const size_t ab = GemmCtx::split_bytes((int64_t)T * B, E) + 8192;
    void* wr = ar.take<char>(wb);
    void* areg = ar.take<char>(ab);
    float* h = ar.take<float>((size_t)4 * B * H);      // [parity][direction][B][H]: the fused step reads one copy and writes the other
    if (ar.overflow) {
        set_error("vag_encoder_train_fwd_f32: workspace too small");
        return VAG_ERR_WORKSPACE;
    }
    GemmCtx gemm(st, wr, wb, areg, ab);
    {
        dim3 block(64, 4);
        encoder_embed_tm_kernel<<<ceil_div(B * T, 4), block, 0, st>>>(x, w->emb, src, ids_tm, B, T, E, w->vocab);
        VAG_LAUNCH_CHECK();
        if (emb_mask) VAG_TRY(vag_mul_f32(x, emb_mask, (int64_t)T * B * E, stream));   // embedding dropout, Encoder.py:51-52
    }
    VAG_CUDA(cudaMemsetAsync(ctx_out, 0, sizeof(float) * (size_t)B * T * 2 * H, st));
    VAG_CUDA(cudaMemsetAsync(h, 0, sizeof(float) * (size_t)4 * B * H, st));
    for (int d = 0; d < 2; ++d)
        VAG_TRY(gemm.linear(gi + (size_t)d * T * B * 3 * H, 3 * H, x, E, w->w_ih[d], E, w->b_ih[d], T * B, E, 3 * H, 0));
    if (enc_seq_fwd_ok(B, T, H)) {   // persistent flavour: ONE launch for the whole time loop of both directions (enc_seq.cu)
        const size_t sb = enc_seq_scratch_bytes(H);
        void* scratch = ar.take<char>(sb);
        if (ar.overflow) {
            set_error("vag_encoder_train_fwd_f32: workspace too small");
            return VAG_ERR_WORKSPACE;
        }
        EncSeqFwd a = {{w->w_hh[0], w->w_hh[1]}, {w->b_hh[0], w->b_hh[1]}, gi, gh, ctx_out, lengths_dev, B, T, H, nullptr, nullptr, 0};
        return enc_seq_fwd(a, scratch, sb, gemm_mode() == 2, st);
    }
    {   // fused flavour: hidden-side contraction + gate math + masking of both directions in ONE launch per time step
        float* hb[2] = {h, h + (size_t)2 * B * H};
        Rows32Gru probe = {gh, w->b_hh[0], gi, 0, hb[0], hb[1], ctx_out, (int64_t)T * 2 * H, lengths_dev, 0, H, {hb[0], w->w_hh[0], H, H, H}};
        if (rows32_gru_ok(probe, B)) {
            for (int s_ = 0; s_ < T; ++s_) {
                const float* in = hb[s_ & 1];
                float* out = hb[(s_ + 1) & 1];
                Rows32Gru pp[2];
                for (int d = 0; d < 2; ++d) {
                    const int t = d == 0 ? s_ : T - 1 - s_;
                    const size_t o3 = ((size_t)d * T + t) * B * 3 * H;
                    pp[d] = Rows32Gru{gh + o3, w->b_hh[d], gi + o3, 0, in + (size_t)d * B * H, out + (size_t)d * B * H,
                                      ctx_out + (int64_t)t * 2 * H + (int64_t)d * H, (int64_t)T * 2 * H, lengths_dev, t, H,
                                      {in + (size_t)d * B * H, w->w_hh[d], H, H, H}};
                }
                VAG_TRY(linear_rows32_gru(pp, 2, B, gemm_mode() == 2, st));
            }
            return VAG_OK;
        }
    }
    const bool pair = rows32_ok(h, H, w->w_hh[0], H, B, H, 3 * H, true) && rows32_ok(h + (size_t)B * H, H, w->w_hh[1], H, B, H, 3 * H, true);
    const int gate_blocks = std::max(1, std::min(ceil_div(B * H / 4, 256), num_sms()));
    for (int s_ = 0; s_ < T; ++s_) {
        float* gh_t[2] = {gh + ((size_t)0 * T + s_) * B * 3 * H, gh + ((size_t)1 * T + (T - 1 - s_)) * B * 3 * H};
        if (pair) {
            const Rows32Problem pp[2] = {{gh_t[0], w->b_hh[0], 3 * H, 3 * H, 1, {{h, w->w_hh[0], H, H, H}, {}}},
                                         {gh_t[1], w->b_hh[1], 3 * H, 3 * H, 1, {{h + (size_t)B * H, w->w_hh[1], H, H, H}, {}}}};
            VAG_TRY(linear_rows32_multi(pp, 2, B, 0, true, gemm_mode() == 2, st));
        } else {
            for (int d = 0; d < 2; ++d) {
                gemm.new_step();
                VAG_TRY(gemm.linear(gh_t[d], 3 * H, h + (size_t)d * B * H, H, w->w_hh[d], H, w->b_hh[d], B, H, 3 * H, 0));
            }
        }
        VAG_CUDA(launch_pdl(PDL_SMALL, enc_gates_fwd_kernel, dim3(gate_blocks, 2), dim3(256), 0, st, h, ctx_out, gi, gh, lengths_dev, B, T, H, s_));
        VAG_LAUNCH_CHECK();
    }
    return VAG_OK;
}

/* grads: d_emb [vocab, E] (zeroed here), d_w_ih[2] [3H,E], d_w_hh[2] [3H,H], d_b_ih[2], d_b_hh[2] [3H] */
extern "C" int vag_encoder_bwd_f32(const vag_encoder_weights* w, const int32_t* lengths_host, const int32_t* lengths_dev, int B, int T,
                                   const float* ctx, const float* dctx, const float* x, const int64_t* ids_tm, const float* gi,
                                   const float* gh, float* d_emb, float* const* d_w_ih, float* const* d_w_hh, float* const* d_b_ih,
                                   float* const* d_b_hh, const float* emb_mask, void* workspace, size_t workspace_bytes,
                                   vag_stream_t stream) {
    ModeScope ms(w ? w->precision : VAG_PREC_FP32);
    VAG_REQUIRE(w && lengths_dev && ctx && dctx && x && ids_tm && gi && gh && d_emb, "vag_encoder_bwd_f32: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    vag_stream_t vs = stream;
    const int E = w->E, H = w->H;
    VAG_TRY(check_lengths_host(lengths_host, B, T, "vag_encoder_bwd_f32"));
    Arena ar(workspace, workspace_bytes);
    const size_t per_dir3 = (size_t)T * B * 3 * H, per_dirh = (size_t)T * B * H;
    float* dgi_all = ar.take<float>(2 * per_dir3);
    float* dgh_all = ar.take<float>(2 * per_dir3);
    float* hprev_all = ar.take<float>(2 * per_dirh);
    float* dx = ar.take<float>((size_t)T * B * E);
    float* carry = ar.take<float>((size_t)2 * B * H);
    const size_t tcs_bytes = gemm_tc_scratch_bytes(3 * H, std::max(E, H), (int64_t)T * B) + gemm_tc_scratch_bytes((int64_t)T * B, E, 3 * H);
    char* tcs = ar.take<char>(tcs_bytes);
    const size_t es_bytes = enc_seq_scratch_bytes(H);
    void* es_scratch = ar.take<char>(es_bytes);
    TcScratchScope tc_scope(tcs, tcs ? tcs_bytes : 0);
    if (ar.overflow) {
        set_error("vag_encoder_bwd_f32: workspace too small");
        return VAG_ERR_WORKSPACE;
    }
    VAG_CUDA(cudaMemsetAsync(carry, 0, sizeof(float) * (size_t)2 * B * H, st));
    const bool pair = rows32_ok(dgh_all, 3 * H, w->w_hh[0], H, B, 3 * H, H, false) && rows32_ok(dgh_all, 3 * H, w->w_hh[1], H, B, 3 * H, H, false);
    const int gate_blocks = std::max(1, std::min(ceil_div(B * H, 256), num_sms()));
    // fused flavour: step 0 runs the gate backward alone; every later step is ONE launch per time step — the recurrent
    // contraction carry + dgh(previous step)·W_hh with the gate backward of the current step in its epilogue, both directions.
    // (The contraction after the last step would only produce the gradient of the constant zero initial state.)
    Rows32GruBwd eprobe = {};
    eprobe.seg[0] = Rows32Seg{dgh_all, w->w_hh[0], 3 * H, H, 3 * H};
    eprobe.nseg = 1; eprobe.gi = gi; eprobe.gh = gh; eprobe.dgi = dgi_all; eprobe.dgh = dgh_all; eprobe.dh_out = carry; eprobe.H = H;
    const bool fused_bwd = rows32_gru_bwd_ok(eprobe, B);
    const bool persistent = enc_seq_bwd_ok(B, T, H);
    if (persistent) {   // ONE launch for the whole back-propagation through time of both directions (enc_seq.cu)
        EncSeqBwd a = {{w->w_hh[0], w->w_hh[1]}, gi, gh, ctx, dctx, dgi_all, dgh_all, hprev_all, lengths_dev, B, T, H, nullptr, nullptr, 0, 0};
        VAG_TRY(enc_seq_bwd(a, es_scratch, es_bytes, gemm_mode() == 2, st));
    }
    for (int s_ = 0; !persistent && fused_bwd && s_ < T; ++s_) {
        if (s_ == 0) {
            VAG_CUDA(launch_pdl(PDL_SMALL, enc_gates_bwd_kernel, dim3(gate_blocks, 2), dim3(256), 0, st, dgi_all, dgh_all, hprev_all, carry, dctx, ctx, gi, gh, lengths_dev, B, T, H, 0));
            VAG_LAUNCH_CHECK();
            continue;
        }
        Rows32GruBwd pp[2];
        for (int d = 0; d < 2; ++d) {
            const int t = d == 0 ? T - 1 - s_ : s_, t_before = d == 0 ? t + 1 : t - 1;      // t_before: the step handled one launch ago
            const int tp = d == 0 ? t - 1 : t + 1;                                           // where this step's previous state was produced
            const size_t o3 = ((size_t)d * T + t) * B * 3 * H, ob = ((size_t)d * T + t_before) * B * 3 * H;
            Rows32GruBwd q = {};
            q.seg[0] = Rows32Seg{dgh_all + ob, w->w_hh[d], 3 * H, H, 3 * H};
            q.nseg = 1;
            q.base = carry + (size_t)d * B * H; q.ld_base = H;
            q.add2 = dctx + (int64_t)t * 2 * H + (int64_t)d * H; q.ld_add2 = (int64_t)T * 2 * H;
            q.gi = gi + o3; q.gh = gh + o3;
            q.h_prev = (tp >= 0 && tp < T) ? ctx + (int64_t)tp * 2 * H + (int64_t)d * H : nullptr; q.ld_hprev = (int64_t)T * 2 * H;
            q.hprev_store = hprev_all + ((size_t)d * T + t) * B * H;
            q.dgi = dgi_all + o3; q.dgh = dgh_all + o3;
            q.dh_out = carry + (size_t)d * B * H;
            q.lengths = lengths_dev; q.t = t; q.H = H;
            pp[d] = q;
        }
        VAG_TRY(linear_rows32_gru_bwd(pp, 2, B, gemm_mode() == 2, st));
    }
    for (int s_ = 0; !persistent && !fused_bwd && s_ < T; ++s_) {
        VAG_CUDA(launch_pdl(PDL_SMALL, enc_gates_bwd_kernel, dim3(gate_blocks, 2), dim3(256), 0, st, dgi_all, dgh_all, hprev_all, carry, dctx, ctx, gi, gh, lengths_dev, B, T, H, s_));
        VAG_LAUNCH_CHECK();
        const int t0 = T - 1 - s_, t1 = s_;
        const float* dgh_t[2] = {dgh_all + (size_t)t0 * B * 3 * H, dgh_all + per_dir3 + (size_t)t1 * B * 3 * H};
        float* cr[2] = {carry, carry + (size_t)B * H};
        if (pair) {   // carry += dgh·W_hh, both directions in one launch
            const Rows32Problem pp[2] = {{cr[0], nullptr, H, H, 1, {{dgh_t[0], w->w_hh[0], 3 * H, H, 3 * H}, {}}},
                                         {cr[1], nullptr, H, H, 1, {{dgh_t[1], w->w_hh[1], 3 * H, H, 3 * H}, {}}}};
            VAG_TRY(linear_rows32_multi(pp, 2, B, VAG_LIN_ACCUMULATE, false, gemm_mode() == 2, st));
        } else {
            for (int d = 0; d < 2; ++d) VAG_TRY(gemm_f32(cr[d], H, dgh_t[d], 3 * H, 1, w->w_hh[d], H, 1, B, H, 3 * H, 1.f, 1.f, vs));
        }
    }
    for (int d = 0; d < 2; ++d) {
        const float* dgi_d = dgi_all + (size_t)d * per_dir3;
        const float* dgh_d = dgh_all + (size_t)d * per_dir3;
        VAG_TRY(gemm_g(dx, E, dgi_d, 3 * H, 1, w->w_ih[d], E, 1, T * B, E, 3 * H, d == 0 ? 0.f : 1.f, st));     // dx (+)= dgi·W_ih
        VAG_TRY(gemm_g(d_w_ih[d], E, dgi_d, 1, 3 * H, x, E, 1, 3 * H, E, T * B, 0.f, st));
        VAG_TRY(gemm_g(d_w_hh[d], H, dgh_d, 1, 3 * H, hprev_all + (size_t)d * per_dirh, H, 1, 3 * H, H, T * B, 0.f, st));
        VAG_TRY(vag_colsum_f32(d_b_ih[d], dgi_d, 3 * H, T * B, 3 * H, 0, vs));
        VAG_TRY(vag_colsum_f32(d_b_hh[d], dgh_d, 3 * H, T * B, 3 * H, 0, vs));
    }
    VAG_CUDA(cudaMemsetAsync(d_emb, 0, sizeof(float) * (size_t)w->vocab * E, st));
    if (emb_mask) VAG_TRY(vag_mul_f32(dx, emb_mask, (int64_t)T * B * E, vs));
    VAG_TRY(vag_embed_bwd_f32(d_emb, dx, E, ids_tm, T * B, E, w->vocab, vs));
    return VAG_OK;
}
