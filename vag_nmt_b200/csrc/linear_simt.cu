// FP32 SIMT contraction  y = act(x · Wᵀ + bias [+ y])  and the embedding gather.
//
// This is the always-correct FP32 FFMA path: it serves shapes the tcgen05 path does not take
// (unaligned / tiny test shapes, very few rows) and is the arbiter the tensor-core path is
// checked against.  x [rows, K] and W [out, K] are both K-contiguous (PyTorch layouts), so a
// CTA stages BK-wide slabs of both through shared memory, transposed to [k][m] so that every
// thread reads its 2×4 A and 2×4 B values with conflict-free 128-bit loads.
#include "common.cuh"
#include "linear_rows.cuh"
#include <cuda_bf16.h>

namespace vag {

int gemm_mode();

template <int BM, int BN, int BK, int TM, int TN>
struct SimtCfg {
    static constexpr int kThreads = (BM / TM) * (BN / TN);
};

// bf16 mode: operands are rounded to bfloat16 on load so that the FFMA path computes what the tensor-core path does
__device__ __forceinline__ float rbf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Loads a [BROWS x BK] slab of a K-contiguous matrix into registers (float4 per thread-slot).
template <int BROWS, int BK, int THREADS, bool VEC, bool RB>
__device__ __forceinline__ void load_slab(float4 (&reg)[(BROWS * BK / 4 + THREADS - 1) / THREADS], const float* __restrict__ base,
                                          int64_t ld, int row0, int n_rows, int k0, int K, int tid) {
    constexpr int QPR = BK / 4;  // float4 per row
    constexpr int SLOTS = (BROWS * BK / 4 + THREADS - 1) / THREADS;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        int idx = tid + s * THREADS;
        int r = idx / QPR;
        int q = idx % QPR;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < BROWS * QPR) {
            int row = row0 + r;
            int k = k0 + q * 4;
            if (row < n_rows) {
                const float* p = base + (int64_t)row * ld + k;
                if (VEC && k + 3 < K) {
                    v = *reinterpret_cast<const float4*>(p);
                } else {
                    if (k + 0 < K) v.x = p[0];
                    if (k + 1 < K) v.y = p[1];
                    if (k + 2 < K) v.z = p[2];
                    if (k + 3 < K) v.w = p[3];
                }
                if (RB) { v.x = rbf(v.x); v.y = rbf(v.y); v.z = rbf(v.z); v.w = rbf(v.w); }
            }
        }
        reg[s] = v;
    }
}

template <int BROWS, int BK, int THREADS, int LDS>
__device__ __forceinline__ void store_slab(float* __restrict__ sm, const float4 (&reg)[(BROWS * BK / 4 + THREADS - 1) / THREADS],
                                           int tid) {
    constexpr int QPR = BK / 4;
    constexpr int SLOTS = (BROWS * BK / 4 + THREADS - 1) / THREADS;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
        int idx = tid + s * THREADS;
        if (idx < BROWS * QPR) {
            int r = idx / QPR;
            int q = idx % QPR;
            sm[(q * 4 + 0) * LDS + r] = reg[s].x;
            sm[(q * 4 + 1) * LDS + r] = reg[s].y;
            sm[(q * 4 + 2) * LDS + r] = reg[s].z;
            sm[(q * 4 + 3) * LDS + r] = reg[s].w;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN, bool VEC, bool RB>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
linear_simt_kernel(float* __restrict__ y, int64_t ldy, const float* __restrict__ x, int64_t ldx,
                   const float* __restrict__ w, int64_t ldw, const float* __restrict__ bias, int rows, int K, int N,
                   int flags) {
    constexpr int THREADS = (BM / TM) * (BN / TN);
    constexpr int LDA = BM + 4;
    constexpr int LDB = BN + 4;
    static_assert(TM % 4 == 0 && TN % 4 == 0, "micro tile is built from float4 pieces");
    constexpr int MQ = TM / 4;  // float4 pieces per thread along M, spaced BM/MQ apart
    constexpr int NQ = TN / 4;
    __shared__ __align__(16) float As[2][BK * LDA];
    __shared__ __align__(16) float Bs[2][BK * LDB];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);
    const int ty = tid / (BN / TN);
    const int m0 = blockIdx.y * BM;
    const int n0 = blockIdx.x * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    constexpr int ASLOTS = (BM * BK / 4 + THREADS - 1) / THREADS;
    constexpr int BSLOTS = (BN * BK / 4 + THREADS - 1) / THREADS;
    float4 ra[ASLOTS], rb[BSLOTS];

    const int n_kt = (K + BK - 1) / BK;
    load_slab<BM, BK, THREADS, VEC, RB>(ra, x, ldx, m0, rows, 0, K, tid);
    load_slab<BN, BK, THREADS, VEC, RB>(rb, w, ldw, n0, N, 0, K, tid);
    store_slab<BM, BK, THREADS, LDA>(As[0], ra, tid);
    store_slab<BN, BK, THREADS, LDB>(Bs[0], rb, tid);
    __syncthreads();

    for (int kt = 0; kt < n_kt; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < n_kt) {
            load_slab<BM, BK, THREADS, VEC, RB>(ra, x, ldx, m0, rows, (kt + 1) * BK, K, tid);
            load_slab<BN, BK, THREADS, VEC, RB>(rb, w, ldw, n0, N, (kt + 1) * BK, K, tid);
        }
        const float* a_s = As[cur];
        const float* b_s = Bs[cur];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < MQ; ++i) {
                float4 v = *reinterpret_cast<const float4*>(&a_s[k * LDA + i * (BM / MQ) + ty * 4]);
                a[i * 4 + 0] = v.x; a[i * 4 + 1] = v.y; a[i * 4 + 2] = v.z; a[i * 4 + 3] = v.w;
            }
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                float4 v = *reinterpret_cast<const float4*>(&b_s[k * LDB + j * (BN / NQ) + tx * 4]);
                b[j * 4 + 0] = v.x; b[j * 4 + 1] = v.y; b[j * 4 + 2] = v.z; b[j * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < n_kt) {
            store_slab<BM, BK, THREADS, LDA>(As[cur ^ 1], ra, tid);
            store_slab<BN, BK, THREADS, LDB>(Bs[cur ^ 1], rb, tid);
        }
        __syncthreads();
    }

    const bool do_tanh = flags & VAG_LIN_TANH;
    const bool do_acc = flags & VAG_LIN_ACCUMULATE;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + (i / 4) * (BM / MQ) + ty * 4 + (i % 4);
        if (row >= rows) continue;
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int col = n0 + (j / 4) * (BN / NQ) + tx * 4 + (j % 4);
            if (col >= N) continue;
            float v = acc[i][j];
            if (bias) v += bias[col];
            float* dst = y + (int64_t)row * ldy + col;
            if (do_acc) v = *dst + v;
            if (do_tanh) v = tanhf(v);
            *dst = v;
        }
    }
}

// Skinny problem (few rows): one warp per (row-group, output) pair, K split across lanes.
// Used when rows <= 8 so that the weight matrix is streamed exactly once.
template <int R, bool RB>
__global__ void __launch_bounds__(256)
linear_skinny_kernel(float* __restrict__ y, int64_t ldy, const float* __restrict__ x, int64_t ldx,
                     const float* __restrict__ w, int64_t ldw, const float* __restrict__ bias, int rows, int K, int N,
                     int flags) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= N) return;
    const int r0 = blockIdx.y * R;
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float* wr = w + (int64_t)warp * ldw;
    int k = lane;
    for (; k + 96 < K; k += 128) {   // four independent K slices in flight per lane (memory-level parallelism)
        float wv[4], xv[R][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) wv[u] = RB ? rbf(wr[k + 32 * u]) : wr[k + 32 * u];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float xx = (r0 + r < rows) ? x[(int64_t)(r0 + r) * ldx + k + 32 * u] : 0.f;
                xv[r][u] = RB ? rbf(xx) : xx;
            }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[r] = fmaf(xv[r][u], wv[u], acc[r]);
    }
    for (; k < K; k += 32) {
        const float wv = RB ? rbf(wr[k]) : wr[k];
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (r0 + r < rows) {
                const float xx = x[(int64_t)(r0 + r) * ldx + k];
                acc[r] = fmaf(RB ? rbf(xx) : xx, wv, acc[r]);
            }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
    if (lane == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (r0 + r >= rows) break;
            float v = acc[r];
            if (bias) v += bias[warp];
            float* dst = y + (int64_t)(r0 + r) * ldy + warp;
            if (flags & VAG_LIN_ACCUMULATE) v = *dst + v;
            if (flags & VAG_LIN_TANH) v = tanhf(v);
            *dst = v;
        }
    }
}

template <int BM, int BN, int BK, int TM, int TN>
static int launch_simt(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                       const float* bias, int rows, int K, int N, int flags, cudaStream_t st) {
    dim3 grid(ceil_div(N, BN), ceil_div(rows, BM));
    dim3 block((BM / TM) * (BN / TN));
    const bool vec = (ldx % 4 == 0) && (ldw % 4 == 0) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)w % 16 == 0);
    const bool rb = gemm_mode() == 2;
    if (vec && !rb)
        linear_simt_kernel<BM, BN, BK, TM, TN, true, false><<<grid, block, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
    else if (!rb)
        linear_simt_kernel<BM, BN, BK, TM, TN, false, false><<<grid, block, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
    else if (vec)
        linear_simt_kernel<BM, BN, BK, TM, TN, true, true><<<grid, block, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
    else
        linear_simt_kernel<BM, BN, BK, TM, TN, false, true><<<grid, block, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int linear_simt(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                int rows, int K, int N, int flags, cudaStream_t st) {
    const int sms = num_sms();
    if (rows32_ok(x, ldx, w, ldw, rows, K, N, true)) return linear_rows32(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, true, gemm_mode() == 2, st);
    if (rows <= 32 && N >= 64 && K >= 64) {   // weight-streaming regime: one warp per output feature, K across the lanes
        dim3 grid(ceil_div(N, 8), ceil_div(rows, 8));
        if (gemm_mode() == 2) linear_skinny_kernel<8, true><<<grid, 256, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
        else linear_skinny_kernel<8, false><<<grid, 256, 0, st>>>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags);
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    const int64_t big_tiles = (int64_t)ceil_div(rows, 128) * ceil_div(N, 128);
    if (big_tiles >= sms) return launch_simt<128, 128, 16, 8, 8>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, st);
    const int64_t mid_tiles = (int64_t)ceil_div(rows, 64) * ceil_div(N, 64);
    if (mid_tiles >= sms / 2 || rows > 32) return launch_simt<64, 64, 16, 4, 4>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, st);
    return launch_simt<32, 32, 16, 4, 4>(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, st);
}

__global__ void embed_rows_kernel(float* __restrict__ out, int64_t ldo, const float* __restrict__ table, int dim,
                                  const int64_t* __restrict__ ids, int rows, int64_t table_rows) {
    const int row = blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= rows) return;
    int64_t id = ids[row];
    if (id < 0 || id >= table_rows) id = 0;  // out-of-range ids read the pad row instead of faulting
    const float* src = table + id * dim;
    float* dst = out + (int64_t)row * ldo;
    for (int c = threadIdx.x; c < dim; c += blockDim.x) dst[c] = src[c];
}

}  // namespace vag

using namespace vag;

extern "C" int vag_embed_rows_f32(float* out, int64_t ldo, const float* table, int dim, const int64_t* ids, int rows,
                                  int64_t table_rows, vag_stream_t stream) {
    VAG_REQUIRE(out && table && ids, "vag_embed_rows_f32: null pointer");
    VAG_REQUIRE(dim > 0 && rows >= 0 && table_rows > 0 && ldo >= dim, "vag_embed_rows_f32: bad shape");
    if (rows == 0) return VAG_OK;
    dim3 block(64, 4);
    embed_rows_kernel<<<ceil_div(rows, 4), block, 0, (cudaStream_t)stream>>>(out, ldo, table, dim, ids, rows, table_rows);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
