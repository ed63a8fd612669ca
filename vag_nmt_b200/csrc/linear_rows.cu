// Contractions with at most 32 rows (one training batch per GPU: the recurrent steps of the encoder / decoder loops,
// forward and backward).  These are weight-streaming problems — 32 x K x N with K, N in 256..1536: 1-6 MB of weights,
// 25-100 MFLOP — whose cost is latency, not bytes or FLOPs, so the kernel is shaped for one short wave of CTAs with
// every load in flight early:
//   * lane = row (32 rows = one warp), so a weight element is a warp-uniform (broadcast) load and is read exactly
//     once per CTA;
//   * a CTA owns BN = 8 (or 4) output columns and its 16 warps split the contraction range; partial sums meet in
//     shared memory, the epilogue (bias, accumulate, tanh) runs on the first BN·32 threads;
//   * WK = true : W(c, k) = w[c·ldw + k]   (y = x·Wᵀ, the forward layout,  layers/NMT_Decoder.py:121-137)
//     WK = false: W(c, k) = w[k·ldw + c]   (dx = dy·W, the same weight read for back-propagation through time).
#include "linear_rows.cuh"
#include <cuda_bf16.h>
#include <stdlib.h>
#include <type_traits>
#include <string.h>

namespace vag {

int gemm_mode();

namespace {

__device__ __forceinline__ float rbf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
template <bool RB>
__device__ __forceinline__ float4 rnd4(float4 v) {
    if (RB) { v.x = rbf16(v.x); v.y = rbf16(v.y); v.z = rbf16(v.z); v.w = rbf16(v.w); }
    return v;
}

struct Rows32GruEpi {        // epilogue side of a Rows32Gru problem (the contraction side is mapped onto Rows32Problem)
    const float* other;
    const float* h_prev;
    float* h_out;
    float* out2;
    int64_t ld_out2;
    const int32_t* lengths;
    int t, H, produces_gi;
};
struct Rows32GruBwdEpi {     // epilogue side of a Rows32GruBwd problem
    const float* base;
    int64_t ld_base;
    const float* add2;
    int64_t ld_add2;
    const float* gi;
    const float* gh;
    const float* h_prev;
    int64_t ld_hprev;
    float* hprev_store;
    float* dgi;
    float* dgh;
    float* dh_out;
    const int32_t* lengths;
    int t, H;
};
struct Rows32Args {
    Rows32Problem p[2];
    Rows32GruEpi g[2];
};
struct Rows32BwdArgs {
    Rows32Problem p[2];
    Rows32GruBwdEpi g[2];
};

constexpr int ROWS32_WARPS = 16;
constexpr int ROWS32_KC = 32;   // contraction indices per warp per pass

__device__ __forceinline__ uint32_t to_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// One pass of a warp = a [BN columns] x [32 contraction indices] weight sub-tile: every lane fetches BN/4 DIFFERENT 16-byte
// pieces of it (coalesced; together with the lane's eight x quads that is the whole pass in flight at once), the tile
// goes through a per-warp shared-memory slot, and the FMAs read it back as broadcast LDS.128.
// MMA = true: the 32 x BN x 32 product of a pass runs on the tensor cores (mma.sync m16n8k8 TF32, two row tiles x four k steps)
// instead of the FMA loop, whose broadcast LDS.128 per 4 FMAs made shared-memory bandwidth the bound of the kernel (2.4 of 7.2 us
// at 32 x 512 x 1536).  FP32 mode: error-compensated 3xTF32 (hi/lo split of both operands, lo·hi + hi·lo + hi·hi);
// bf16 mode: the rounded operands are exact in TF32, one product.
template <int BN, bool WK, bool RB, bool MMA, bool GRU = false, typename ArgsT = Rows32Args>
__global__ void __launch_bounds__(ROWS32_WARPS * 32, 2)
linear_rows32_kernel(const __grid_constant__ ArgsT args, int rows, int flags) {
    constexpr bool GRUB = !GRU && !WK && std::is_same<ArgsT, Rows32BwdArgs>::value;   // GRU-backward epilogue (columns = hidden units)
    // blockIdx.y selects one of two independent problems (own N, pitches and segments); __grid_constant__ keeps the
    // dynamically indexed descriptors in parameter space instead of a per-thread local copy
    const Rows32Problem& pr = args.p[blockIdx.y];
    // GRU flavour: BN = 12 columns = 4 hidden units x 3 gates; tile column c is weight row (c / 4)·H + u0 + (c % 4)
    constexpr int GU = 4;
    const int N = pr.N;                                // GRU: 3H
    const int n0 = GRU ? blockIdx.x * GU : blockIdx.x * BN;
    if (n0 >= (GRU ? N / 3 : N)) return;               // the grid is sized for the wider problem
    float* __restrict__ y = pr.y;
    const float* __restrict__ bias = pr.bias;
    const int64_t ldy = pr.ldy;
    // dynamic shared memory: per-warp x tile [32 rows][36] (row pitch 36 floats: conflict-free 16-byte accesses both ways),
    // per-warp weight tile [BN·32]; the cross-warp reduction buffer [16][32][BN+1] reuses the x tiles after the main loop
    extern __shared__ __align__(16) float smem_dyn[];
    constexpr int XP = ROWS32_KC + 4;
    float* xs_all = smem_dyn;
    float* wt_all = smem_dyn + ROWS32_WARPS * 32 * XP;
    constexpr int WT = BN * XP;                        // floats per warp (the [column][36] layout is the largest)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float* xs = xs_all + wid * 32 * XP;
    float* wt = wt_all + wid * WT;
    // GRU-backward epilogue: its operands are fetched NOW, so that their L2 round trip overlaps the main loop instead of
    // following the cross-warp reduction
    float pf_gi[3] = {0.f, 0.f, 0.f}, pf_gh[3] = {0.f, 0.f, 0.f}, pf_hp = 0.f, pf_add = 0.f;
    if constexpr (GRUB) {
        if (threadIdx.x < 32 * BN) {
            const auto& ge = args.g[blockIdx.y];
            const int row = threadIdx.x / BN, u = n0 + threadIdx.x % BN, H = ge.H;
            if (row < rows && u < N) {
#pragma unroll
                for (int gt = 0; gt < 3; ++gt) {
                    pf_gi[gt] = ge.gi[(int64_t)row * 3 * H + gt * H + u];
                    pf_gh[gt] = ge.gh[(int64_t)row * 3 * H + gt * H + u];
                }
                pf_hp = ge.h_prev ? ge.h_prev[(int64_t)row * ge.ld_hprev + u] : 0.f;
                if (ge.base) pf_add += ge.base[(int64_t)row * ge.ld_base + u];
                if (ge.add2) pf_add += ge.add2[(int64_t)row * ge.ld_add2 + u];
            }
        }
    }
    float acc[BN];
#pragma unroll
    for (int c = 0; c < BN; ++c) acc[c] = 0.f;
    // tensor-core flavour: accumulator fragments of the two 16-row tiles (m16n8k8: d0/d1 = row g, cols 2t/2t+1; d2/d3 = row g+8)
    constexpr int NT = (BN + 7) / 8;            // 8-column tiles per CTA (BN = 4 uses half of one)
    float dacc[NT][2][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int u = 0; u < 4; ++u) dacc[nt][mt][u] = 0.f;
    const int g = lane >> 2, t4 = lane & 3;
    // passes of 32 contraction indices, segment 0 first, dealt round-robin to the 16 warps
    const int passes0 = (pr.seg[0].K + ROWS32_KC - 1) / ROWS32_KC;
    const int passes = passes0 + (pr.nseg > 1 ? (pr.seg[1].K + ROWS32_KC - 1) / ROWS32_KC : 0);
    for (int pass = wid; pass < passes; pass += ROWS32_WARPS) {
        const Rows32Seg& sg = pr.seg[pass < passes0 ? 0 : 1];
        const int k0 = (pass < passes0 ? pass : pass - passes0) * ROWS32_KC;
        const float* __restrict__ x = sg.x;
        const float* __restrict__ w = sg.w;
        const int64_t ldx = sg.ldx, ldw = sg.ldw;
        const int K = sg.K;
        // ---- issue every load of the pass
        float4 wv[BN / 4];
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) {
            if (WK) {   // piece = (column c, quad q) of the [BN][32] tile
                const int piece = i * 32 + lane, c = piece >> 3, q = piece & 7;
                const int col = GRU ? (c / GU) * (N / 3) + n0 + (c % GU) : min(n0 + c, N - 1);
                const int k = min(k0 + 4 * q, K - 4);
                wv[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)col * ldw + k));
            } else {    // lane = contraction index, piece i = columns 4i..4i+3 of the [32][BN] tile
                const int k = min(k0 + lane, K - 1), col = min(n0 + 4 * i, N - 4);
                wv[i] = __ldg(reinterpret_cast<const float4*>(w + (int64_t)k * ldw + col));
            }
        }
        // x tile, coalesced: instruction i covers rows 4i..4i+3, eight lanes per row (a lane-per-row read would touch 32
        // cache lines per instruction and is bound by the L1 tag rate)
        float4 xl[ROWS32_KC / 4];
#pragma unroll
        for (int i = 0; i < ROWS32_KC / 4; ++i) {
            const int row = 4 * i + (lane >> 3), k = k0 + 4 * (lane & 7);
            xl[i] = (row < rows && k < K) ? __ldg(reinterpret_cast<const float4*>(x + (int64_t)row * ldx + k)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < ROWS32_KC / 4; ++i)
            *reinterpret_cast<float4*>(xs + (4 * i + (lane >> 3)) * XP + 4 * (lane & 7)) = rnd4<RB>(xl[i]);
        // ---- weight tile through shared memory (FMA flavour: [piece] of float4 for both orientations; tensor-core flavour:
        //      [column][36] for the k-fast orientation — pitch 36 keeps the fragment reads conflict-free — else [k][BN])
#pragma unroll
        for (int i = 0; i < BN / 4; ++i) {
            if (MMA && WK) {
                const int piece = i * 32 + lane;
                *reinterpret_cast<float4*>(wt + (piece >> 3) * XP + 4 * (piece & 7)) = rnd4<RB>(wv[i]);
            } else if (MMA) {
                *reinterpret_cast<float4*>(wt + lane * BN + 4 * i) = rnd4<RB>(wv[i]);
            } else {
                reinterpret_cast<float4*>(wt)[i * 32 + lane] = rnd4<RB>(wv[i]);
            }
        }
        __syncwarp();
        if (MMA) {
#pragma unroll
            for (int ks = 0; ks < ROWS32_KC / 8; ++ks) {
                const int kk = ks * 8 + t4;
                // A fragments of both row tiles: (g, t) (g+8, t) (g, t+4) (g+8, t+4)
                uint32_t ah[2][4], alo[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    const float* xa = xs + (mt * 16 + g) * XP + kk;
                    const float af[4] = {xa[0], xa[8 * XP], xa[4], xa[8 * XP + 4]};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        ah[mt][u] = RB ? __float_as_uint(af[u]) : to_tf32(af[u]);
                        if (!RB) alo[mt][u] = to_tf32(af[u] - __uint_as_float(ah[mt][u]));
                    }
                }
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    // B fragment: b0 = W(k = kk, n), b1 = W(k = kk + 4, n) with n = 8·nt + g; columns past BN are zero
                    const int n = nt * 8 + g;
                    float bf[2] = {0.f, 0.f};
                    if (n < BN) {
                        bf[0] = WK ? wt[n * XP + kk] : wt[kk * BN + n];
                        bf[1] = WK ? wt[n * XP + kk + 4] : wt[(kk + 4) * BN + n];
                    }
                    uint32_t bh[2], bl[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        bh[u] = RB ? __float_as_uint(bf[u]) : to_tf32(bf[u]);
                        if (!RB) bl[u] = to_tf32(bf[u] - __uint_as_float(bh[u]));
                    }
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        if (!RB) {
                            mma_tf32(dacc[nt][mt], alo[mt], bh);
                            mma_tf32(dacc[nt][mt], ah[mt], bl);
                        }
                        mma_tf32(dacc[nt][mt], ah[mt], bh);
                    }
                }
            }
            __syncwarp();
            continue;
        }
        float4 xv[ROWS32_KC / 4];     // lane = row from here on
#pragma unroll
        for (int q = 0; q < ROWS32_KC / 4; ++q) xv[q] = *reinterpret_cast<const float4*>(xs + lane * XP + 4 * q);
        if (WK) {
#pragma unroll
            for (int c = 0; c < BN; ++c)
#pragma unroll
                for (int q = 0; q < ROWS32_KC / 4; ++q) {
                    const float4 wq = reinterpret_cast<const float4*>(wt)[c * 8 + q];
                    const float4 xq = xv[q];
                    acc[c] = fmaf(xq.x, wq.x, acc[c]);
                    acc[c] = fmaf(xq.y, wq.y, acc[c]);
                    acc[c] = fmaf(xq.z, wq.z, acc[c]);
                    acc[c] = fmaf(xq.w, wq.w, acc[c]);
                }
        } else {
#pragma unroll
            for (int q = 0; q < ROWS32_KC / 4; ++q) {
                const float4 xq = xv[q];
                const float xs4[4] = {xq.x, xq.y, xq.z, xq.w};
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
#pragma unroll
                    for (int i = 0; i < BN / 4; ++i) {
                        const float4 wq = reinterpret_cast<const float4*>(wt)[i * 32 + 4 * q + kk];
                        acc[4 * i + 0] = fmaf(xs4[kk], wq.x, acc[4 * i + 0]);
                        acc[4 * i + 1] = fmaf(xs4[kk], wq.y, acc[4 * i + 1]);
                        acc[4 * i + 2] = fmaf(xs4[kk], wq.z, acc[4 * i + 2]);
                        acc[4 * i + 3] = fmaf(xs4[kk], wq.w, acc[4 * i + 3]);
                    }
            }
        }
        __syncwarp();
    }
    __syncthreads();                                   // every warp is done with its x tile: reuse the space
    float (*red)[32][BN + 1] = reinterpret_cast<float (*)[32][BN + 1]>(smem_dyn);
    if (MMA) {
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int row = mt * 16 + g + (u >> 1) * 8, c = nt * 8 + 2 * t4 + (u & 1);
                    if (c < BN) red[wid][row][c] = dacc[nt][mt][u];
                }
    } else {
#pragma unroll
        for (int c = 0; c < BN; ++c) red[wid][lane][c] = acc[c];
    }
    __syncthreads();
    if constexpr (GRU) {
        if (threadIdx.x < 32 * GU) {
            const Rows32GruEpi& ge = args.g[blockIdx.y];
            const int row = threadIdx.x / GU, i = threadIdx.x % GU, u = n0 + i, H = ge.H;
            if (row < rows) {
                float pre[3], oth[3];
#pragma unroll
                for (int gt = 0; gt < 3; ++gt) {
                    float v = 0.f;
#pragma unroll
                    for (int q = 0; q < ROWS32_WARPS; ++q) v += red[q][row][gt * GU + i];
                    pre[gt] = v + (bias ? bias[gt * H + u] : 0.f);
                    oth[gt] = ge.other[(int64_t)row * 3 * H + gt * H + u];
                }
                const bool live = !ge.lengths || ge.lengths[row] > ge.t;
                const float hp = ge.h_prev[(int64_t)row * H + u];
                float hn = hp;
                if (live) {
                    const float* gi = ge.produces_gi ? pre : oth;
                    const float* gh = ge.produces_gi ? oth : pre;
                    const float r = sigmoidf_precise(gi[0] + gh[0]);
                    const float z = sigmoidf_precise(gi[1] + gh[1]);
                    const float n = tanhf(gi[2] + r * gh[2]);
                    hn = (1.0f - z) * n + z * hp;
                    if (ge.out2) ge.out2[(int64_t)row * ge.ld_out2 + u] = hn;
                }
                ge.h_out[(int64_t)row * H + u] = hn;
#pragma unroll
                for (int gt = 0; gt < 3; ++gt) y[(int64_t)row * ldy + gt * H + u] = live ? pre[gt] : 0.f;
            }
        }
        return;
    }
    if constexpr (GRUB) {
        if (threadIdx.x < 32 * BN) {
            const auto& ge = args.g[blockIdx.y];
            const int row = threadIdx.x / BN, c = threadIdx.x % BN, u = n0 + c, H = ge.H;
            if (row < rows && u < N) {
                float g = 0.f;
#pragma unroll
                for (int q = 0; q < ROWS32_WARPS; ++q) g += red[q][row][c];
                g += pf_add;
                const float hp = pf_hp;
                if (ge.hprev_store) ge.hprev_store[(int64_t)row * H + u] = hp;
                float* a = ge.dgi + (int64_t)row * 3 * H;
                float* b = ge.dgh + (int64_t)row * 3 * H;
                if (ge.lengths && ge.lengths[row] <= ge.t) {
                    a[u] = 0.f; a[H + u] = 0.f; a[2 * H + u] = 0.f;
                    b[u] = 0.f; b[H + u] = 0.f; b[2 * H + u] = 0.f;
                    ge.dh_out[(int64_t)row * H + u] = 0.f;
                } else {
                    const float r = sigmoidf_precise(pf_gi[0] + pf_gh[0]);
                    const float z = sigmoidf_precise(pf_gi[1] + pf_gh[1]);
                    const float hn = pf_gh[2];
                    const float n = tanhf(pf_gi[2] + r * hn);
                    const float dn_pre = g * (1.f - z) * (1.f - n * n);
                    const float dz_pre = g * (hp - n) * z * (1.f - z);
                    const float dr_pre = dn_pre * hn * r * (1.f - r);
                    a[u] = dr_pre; a[H + u] = dz_pre; a[2 * H + u] = dn_pre;
                    b[u] = dr_pre; b[H + u] = dz_pre; b[2 * H + u] = dn_pre * r;
                    ge.dh_out[(int64_t)row * H + u] = g * z;
                }
            }
        }
        return;
    }
    if (threadIdx.x < 32 * BN) {
        const int row = threadIdx.x / BN, c = threadIdx.x % BN;
        const int col = n0 + c;
        if (row < rows && col < N) {
            float v = 0.f;
#pragma unroll
            for (int q = 0; q < ROWS32_WARPS; ++q) v += red[q][row][c];
            if (bias) v += bias[col];
            float* dst = y + (int64_t)row * ldy + col;
            if (flags & VAG_LIN_ACCUMULATE) v = *dst + v;
            if (flags & VAG_LIN_TANH) v = tanhf(v);
            *dst = v;
        }
    }
}

template <int BN>
constexpr size_t rows32_smem_bytes() { return (size_t)(ROWS32_WARPS * 32 * (ROWS32_KC + 4) + ROWS32_WARPS * BN * (ROWS32_KC + 4)) * sizeof(float); }

static bool rows32_use_mma() {   // VAG_ROWS32=ffma keeps the FMA inner loop (A/B runs)
    static const bool on = [] { const char* e = getenv("VAG_ROWS32"); return !(e && strcmp(e, "ffma") == 0); }();
    return on;
}

template <int BN, bool WK, bool RB, bool MMA>
int launch_rows32_impl(const Rows32Problem& p0, const Rows32Problem& p1, int nprob, int rows, int flags, cudaStream_t st) {
    static bool attr_set = false;
    constexpr size_t smem = rows32_smem_bytes<BN>();
    if (!attr_set) {
        VAG_CUDA(cudaFuncSetAttribute(linear_rows32_kernel<BN, WK, RB, MMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int nmax = nprob > 1 ? (p0.N > p1.N ? p0.N : p1.N) : p0.N;
    Rows32Args args = {};
    args.p[0] = p0;
    args.p[1] = p1;
    linear_rows32_kernel<BN, WK, RB, MMA><<<dim3(ceil_div(nmax, BN), nprob), ROWS32_WARPS * 32, smem, st>>>(args, rows, flags);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

template <int BN, bool WK, bool RB>
int launch_rows32_bn(const Rows32Problem& p0, const Rows32Problem& p1, int nprob, int rows, int flags, cudaStream_t st) {
    // bf16 mode: always the tensor-core loop (one product, no conversions).  FP32 mode pays 3 products + the hi/lo split per
    // fragment: measured faster than the FMA loop only for the k-fast orientation with 8-column CTAs (6.1 vs 7.2 us at
    // 32 x 512 x 1536), slower elsewhere (8.4 vs 7.5 us transposed, 6.5 vs 5.7 us with 4-column CTAs).
    if (rows32_use_mma() && (RB || (WK && BN == 8))) return launch_rows32_impl<BN, WK, RB, true>(p0, p1, nprob, rows, flags, st);
    return launch_rows32_impl<BN, WK, RB, false>(p0, p1, nprob, rows, flags, st);
}

template <bool WK, bool RB>
int launch_rows32(const Rows32Problem& p0, const Rows32Problem& p1, int nprob, int rows, int flags, cudaStream_t st) {
    // enough CTAs for one wave on 148 SMs: 8 columns per CTA for wide outputs, 4 otherwise
    const int ntot = p0.N + (nprob > 1 ? p1.N : 0);
    // (16 columns per CTA — one wave instead of two for the two-problem launches — measured no faster: 2.852 vs 2.854 ms per step)
    if (ntot >= 960) return launch_rows32_bn<8, WK, RB>(p0, p1, nprob, rows, flags, st);
    return launch_rows32_bn<4, WK, RB>(p0, p1, nprob, rows, flags, st);
}

}  // namespace

// Eligibility: rows <= 32, 16-byte aligned operands with pitches that keep float4 loads aligned, K a multiple of 4
// (and, for the transposed weight read, N a multiple of 4).
bool rows32_ok(const float* x, int64_t ldx, const float* w, int64_t ldw, int rows, int K, int N, bool wk) {
    if (rows < 1 || rows > 32 || K < 64 || N < 32 || (K & 3)) return false;
    if (((uintptr_t)x & 15) || ((uintptr_t)w & 15) || (ldx & 3) || (ldw & 3)) return false;
    if (!wk && (N & 3)) return false;
    return true;
}
bool rows32_problem_ok(const Rows32Problem& p, int rows, bool wk) {
    if (p.nseg < 1 || p.nseg > 2) return false;
    for (int i = 0; i < p.nseg; ++i)
        if (!rows32_ok(p.seg[i].x, p.seg[i].ldx, p.seg[i].w, p.seg[i].ldw, rows, p.seg[i].K, p.N, wk)) return false;
    return true;
}

int linear_rows32_multi(const Rows32Problem* probs, int nprob, int rows, int flags, bool wk, bool round_bf16, cudaStream_t st) {
    const Rows32Problem& p0 = probs[0];
    const Rows32Problem& p1 = probs[nprob > 1 ? 1 : 0];
    if (wk) {
        if (round_bf16) return launch_rows32<true, true>(p0, p1, nprob, rows, flags, st);
        return launch_rows32<true, false>(p0, p1, nprob, rows, flags, st);
    }
    if (round_bf16) return launch_rows32<false, true>(p0, p1, nprob, rows, flags, st);
    return launch_rows32<false, false>(p0, p1, nprob, rows, flags, st);
}

bool rows32_gru_bwd_ok(const Rows32GruBwd& p, int rows) {
    if (p.nseg < 1 || p.nseg > 2 || p.H < 64 || (p.H & 3) || !p.gi || !p.gh || !p.dgi || !p.dgh || !p.dh_out) return false;
    for (int i = 0; i < p.nseg; ++i) {
        if (!rows32_ok(p.seg[i].x, p.seg[i].ldx, p.seg[i].w, p.seg[i].ldw, rows, p.seg[i].K, p.H, false)) return false;
        if (p.seg[i].x == p.dh_out) return false;
    }
    return true;
}

template <bool RB, bool MMA>
static int launch_rows32_gru_bwd(const Rows32BwdArgs& args, int nprob, int rows, cudaStream_t st) {
    static bool attr_set = false;
    constexpr size_t smem = rows32_smem_bytes<4>();
    if (!attr_set) {
        VAG_CUDA(cudaFuncSetAttribute(linear_rows32_kernel<4, false, RB, MMA, false, Rows32BwdArgs>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int H = args.g[0].H;
    linear_rows32_kernel<4, false, RB, MMA, false, Rows32BwdArgs><<<dim3(H / 4, nprob), ROWS32_WARPS * 32, smem, st>>>(args, rows, 0);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// Incoming-gradient contraction + GRU-cell backward in one launch per problem (see Rows32GruBwd).  Problems share H.
int linear_rows32_gru_bwd(const Rows32GruBwd* probs, int nprob, int rows, bool round_bf16, cudaStream_t st) {
    Rows32BwdArgs args = {};
    for (int i = 0; i < 2; ++i) {
        const Rows32GruBwd& p = probs[i < nprob ? i : 0];
        args.p[i] = Rows32Problem{nullptr, nullptr, (int64_t)p.H, p.H, p.nseg, {p.seg[0], p.seg[1]}};
        args.g[i] = Rows32GruBwdEpi{p.base, p.ld_base, p.add2, p.ld_add2, p.gi, p.gh, p.h_prev, p.ld_hprev, p.hprev_store,
                                    p.dgi, p.dgh, p.dh_out, p.lengths, p.t, p.H};
    }
    // tensor-core inner loop only in bf16 mode for this orientation (see launch_rows32_bn)
    if (round_bf16) return rows32_use_mma() ? launch_rows32_gru_bwd<true, true>(args, nprob, rows, st) : launch_rows32_gru_bwd<true, false>(args, nprob, rows, st);
    return launch_rows32_gru_bwd<false, false>(args, nprob, rows, st);
}

bool rows32_gru_ok(const Rows32Gru& p, int rows) {
    if (p.H < 64 || (p.H & 3) || !p.pre || !p.other || !p.h_prev || !p.h_out || p.h_out == p.h_prev || p.h_out == p.seg.x) return false;
    return rows32_ok(p.seg.x, p.seg.ldx, p.seg.w, p.seg.ldw, rows, p.seg.K, 3 * p.H, true);
}

template <bool RB, bool MMA>
static int launch_rows32_gru(const Rows32Args& args, int nprob, int rows, cudaStream_t st) {
    static bool attr_set = false;
    constexpr size_t smem = rows32_smem_bytes<12>();
    if (!attr_set) {
        VAG_CUDA(cudaFuncSetAttribute(linear_rows32_kernel<12, true, RB, MMA, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int H = args.g[0].H;
    linear_rows32_kernel<12, true, RB, MMA, true><<<dim3(H / 4, nprob), ROWS32_WARPS * 32, smem, st>>>(args, rows, 0);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// One GRU cell step per problem: contraction + gate epilogue in one launch (see Rows32Gru).  Problems of a launch share H.
int linear_rows32_gru(const Rows32Gru* probs, int nprob, int rows, bool round_bf16, cudaStream_t st) {
    Rows32Args args;
    for (int i = 0; i < 2; ++i) {
        const Rows32Gru& p = probs[i < nprob ? i : 0];
        args.p[i] = Rows32Problem{p.pre, p.bias, (int64_t)3 * p.H, 3 * p.H, 1, {p.seg, Rows32Seg{nullptr, nullptr, 0, 0, 0}}};
        args.g[i] = Rows32GruEpi{p.other, p.h_prev, p.h_out, p.out2, p.ld_out2, p.lengths, p.t, p.H, p.produces_gi};
    }
    const bool mma = rows32_use_mma();
    if (round_bf16) return mma ? launch_rows32_gru<true, true>(args, nprob, rows, st) : launch_rows32_gru<true, false>(args, nprob, rows, st);
    return mma ? launch_rows32_gru<false, true>(args, nprob, rows, st) : launch_rows32_gru<false, false>(args, nprob, rows, st);
}

// y[rows, N] (+)= x[rows, K] · W(c, k)  (+ bias) (tanh);  round_bf16: operands rounded to bfloat16 first (bf16 mode)
int linear_rows32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
                  int K, int N, int flags, bool wk, bool round_bf16, cudaStream_t st) {
    Rows32Problem p{y, bias, ldy, N, 1, {{x, w, ldx, ldw, K}, {nullptr, nullptr, 0, 0, 0}}};
    return linear_rows32_multi(&p, 1, rows, flags, wk, round_bf16, st);
}

}  // namespace vag
