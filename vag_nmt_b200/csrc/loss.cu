// Shared-embedding ranking losses (forward + gradient) and retrieval ranks.
//   PairwiseRankingLoss.forward        losses/PairwiseRankingLoss.py:9-24
//   ImageRetrievalRankingLoss.forward  losses/ImageRetrievalRankingLoss.py:9-21
//   t2i / i2t rank computation         utils/im_retrieval_eval.py:15-22
// The reference zeroes the diagonal with a Python loop of 2·B tiny kernels; here the score matrix is produced
// by one contraction and a single pass computes both hinge maps, their sum and the gradient mask.
#include "common.cuh"

namespace vag {

int linear_simt(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                int rows, int K, int N, int flags, cudaStream_t st);

// One CTA per row i of scores [B, B].  Writes G[i, j] (j != i) = dLoss/dscores[i,j] and the row's loss partial.
__global__ void __launch_bounds__(256)
rank_hinge_kernel(const float* __restrict__ scores, int B, float margin, int one_direction, float* __restrict__ G,
                  float* __restrict__ row_loss) {
    const int i = blockIdx.x;
    const float d_i = scores[(int64_t)i * B + i];
    float acc = 0.f;
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
        const float s = scores[(int64_t)i * B + j];
        float g = 0.f;
        if (j != i) {
            const float d_j = scores[(int64_t)j * B + j];
            const float cs = (margin - d_j) + s;  // cost_s[i,j]  :16
            if (cs > 0.f) { acc += cs; g += 1.f; }
            if (!one_direction) {
                const float ci = (margin - d_i) + s;  // cost_im[i,j] :18
                if (ci > 0.f) { acc += ci; g += 1.f; }
            }
        }
        if (G) G[(int64_t)i * B + j] = g;
    }
    __shared__ float red[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        row_loss[i] = t;
    }
}

// Diagonal of the gradient: every hinge that is active pulls the diagonal score down.
__global__ void __launch_bounds__(256)
rank_diag_kernel(const float* __restrict__ scores, int B, float margin, int one_direction, float* __restrict__ G) {
    const int j = blockIdx.x;
    const float d = scores[(int64_t)j * B + j];
    float cnt = 0.f;
    for (int i = threadIdx.x; i < B; i += blockDim.x) {
        if (i == j) continue;
        if ((margin - d) + scores[(int64_t)i * B + j] > 0.f) cnt += 1.f;                     // column j of cost_s
        if (!one_direction && (margin - d) + scores[(int64_t)j * B + i] > 0.f) cnt += 1.f;   // row j of cost_im
    }
    __shared__ float red[8];
    cnt = warp_sum(cnt);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        G[(int64_t)j * B + j] = -t;
    }
}

__global__ void sum_to_scalar_kernel(const float* __restrict__ x, int n, float* __restrict__ out) {
    __shared__ float red[8];
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
    a = warp_sum(a);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        out[0] = t;
    }
}

// out[i, c] = Σ_j G[i, j]·m[j, c]   (TRANS == false)   or   Σ_j G[j, i]·m[j, c]   (TRANS == true)
template <bool TRANS>
__global__ void __launch_bounds__(128)
rank_grad_kernel(const float* __restrict__ G, const float* __restrict__ m, int B, int S, float* __restrict__ out) {
    const int i = blockIdx.y;
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= S) return;
    float a = 0.f;
    for (int j = 0; j < B; ++j) {
        const float g = TRANS ? G[(int64_t)j * B + i] : G[(int64_t)i * B + j];
        a = fmaf(g, m[(int64_t)j * S + c], a);
    }
    out[(int64_t)i * S + c] = a;
}

// One CTA per query row of scores [n, n].
__global__ void __launch_bounds__(256) recall_rank_kernel(const float* __restrict__ scores, int n, int32_t* __restrict__ ranks) {
    const int i = blockIdx.x;
    const float t = scores[(int64_t)i * n + i];
    int cnt = 0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        const float s = scores[(int64_t)i * n + j];
        cnt += (s > t) || (s == t && j < i);
    }
    __shared__ int red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
        ranks[i] = tot;
    }
}

}  // namespace vag

using namespace vag;

extern "C" size_t vag_rank_loss_workspace_bytes(int B, int S) {
    ArenaSizer s;
    s.take<float>((size_t)B * B);
    s.take<float>((size_t)B * B);
    s.take<float>((size_t)B);
    (void)S;
    return s.total();
}

extern "C" int vag_rank_loss_f32(const float* im, const float* s, int B, int S, float margin, int one_direction,
                                 float* loss_out, float* grad_im, float* grad_s, void* workspace, size_t workspace_bytes,
                                 vag_stream_t stream) {
    VAG_REQUIRE(im && s && loss_out, "vag_rank_loss_f32: null pointer");
    VAG_REQUIRE(B > 0 && S > 0, "vag_rank_loss_f32: bad shape");
    VAG_REQUIRE((grad_im == nullptr) == (grad_s == nullptr), "vag_rank_loss_f32: pass both gradients or neither");
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    float* scores = ar.take<float>((size_t)B * B);
    float* G = ar.take<float>((size_t)B * B);
    float* row_loss = ar.take<float>((size_t)B);
    if (ar.overflow) {
        set_error("vag_rank_loss_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    VAG_TRY(linear_simt(scores, B, im, S, s, S, nullptr, B, S, B, 0, st));  // scores = im · sᵀ   :12
    rank_hinge_kernel<<<B, 256, 0, st>>>(scores, B, margin, one_direction, grad_im ? G : nullptr, row_loss);
    VAG_LAUNCH_CHECK();
    sum_to_scalar_kernel<<<1, 256, 0, st>>>(row_loss, B, loss_out);
    VAG_LAUNCH_CHECK();
    if (grad_im) {
        rank_diag_kernel<<<B, 256, 0, st>>>(scores, B, margin, one_direction, G);
        VAG_LAUNCH_CHECK();
        dim3 grid(ceil_div(S, 128), B);
        rank_grad_kernel<false><<<grid, 128, 0, st>>>(G, s, B, S, grad_im);   // dL/dim = G · s
        VAG_LAUNCH_CHECK();
        rank_grad_kernel<true><<<grid, 128, 0, st>>>(G, im, B, S, grad_s);    // dL/ds = Gᵀ · im
        VAG_LAUNCH_CHECK();
    }
    return VAG_OK;
}

extern "C" size_t vag_recall_ranks_workspace_bytes(int n, int S) {
    ArenaSizer s;
    s.take<float>((size_t)n * n);
    (void)S;
    return s.total();
}

extern "C" int vag_recall_ranks_f32(const float* queries, const float* gallery, int n, int S, int32_t* ranks, void* workspace,
                                    size_t workspace_bytes, vag_stream_t stream) {
    VAG_REQUIRE(queries && gallery && ranks, "vag_recall_ranks_f32: null pointer");
    VAG_REQUIRE(n > 0 && S > 0, "vag_recall_ranks_f32: bad shape");
    cudaStream_t st = (cudaStream_t)stream;
    Arena ar(workspace, workspace_bytes);
    float* scores = ar.take<float>((size_t)n * n);
    if (ar.overflow) {
        set_error("vag_recall_ranks_f32: workspace %zu B too small", workspace_bytes);
        return VAG_ERR_WORKSPACE;
    }
    VAG_TRY(linear_simt(scores, n, queries, S, gallery, S, nullptr, n, S, n, 0, st));
    recall_rank_kernel<<<n, 256, 0, st>>>(scores, n, ranks);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
