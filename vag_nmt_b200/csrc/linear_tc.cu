// Tensor-core contraction:  y = act(x · Wᵀ + bias [+ y])  in the FP32-exact mode (error-compensated hi/lo operand split, three
// products — FP16 planes by default, TF32 planes under VAG_GEMM=tf32x3) or in the bf16 mode (one product).
//
// tcgen05 has no FP32-input MMA, and plain TF32 (10-bit mantissa) flips beams (hard part 1).
// Every operand is therefore split once into hi = rn_tf32(v) and lo = v - hi (exact in FP32) and the product is
// accumulated in TMEM (FP32) as  hi·hi + (hi·lo + lo·hi) ; the dropped lo·lo term is ≤ 2^-24 relative.
// The tensor core adds into its FP32 accumulator with truncation, a bias that grows with the number of
// accumulating instructions, so the 2^-12-times-smaller cross terms get their OWN accumulator (second half of the
// TMEM allocation): the main accumulator sees K/8 additions instead of 3K/8 and the two are summed once, in
// round-to-nearest FP32, by the epilogue.
//
// Kernel anatomy (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   TMA producer: cp.async.bulk.tensor 2-D boxes {32 fp32 = 128 B, rows} of x_hi, x_lo, w_hi, w_lo
//            into a STAGES-deep shared-memory ring (SWIZZLE_128B), completion on mbarriers (expect_tx)
//   warp 1   TMEM allocation + MMA issuer: one elected lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8)
//            from shared-memory descriptors, 3 products per k-step; tcgen05.commit releases ring slots and
//            finally signals the epilogue
//   warps 2-5  epilogue: tcgen05.ld (32 lanes x 32 columns) → bias / accumulate / tanh → global stores
#include "common.cuh"
#include "split.cuh"
#include "gemm_ctx.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>

namespace vag {

int gemm_mode();

// ------------------------------------------------------------------------------------------ operand split
// hi = round-to-nearest TF32 of v (low 13 mantissa bits zero), lo = v - hi.  Outputs are compact [rows, K].
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, float* __restrict__ hi, float* __restrict__ lo, int64_t ldo) {
    const int kq = K >> 2;
    const int64_t total = (int64_t)rows * kq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kq), c = (int)(i % kq) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)r * ldx + c);
        float4 h, l;
        uint32_t t;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.x)); h.x = __uint_as_float(t); l.x = v.x - h.x;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.y)); h.y = __uint_as_float(t); l.y = v.y - h.y;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.z)); h.z = __uint_as_float(t); l.z = v.z - h.z;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v.w)); h.w = __uint_as_float(t); l.w = v.w - h.w;
        *reinterpret_cast<float4*>(hi + (int64_t)r * ldo + c) = h;
        *reinterpret_cast<float4*>(lo + (int64_t)r * ldo + c) = l;
    }
}

// FP16 flavour of the split (default): hi = rn_f16(v), lo' = rn_f16((v - hi)·2^11).  v ≈ hi + lo'·2^-11 with a
// residual ≤ 2^-22·|v|; the cross accumulator is scaled by 2^-11 in the epilogue.  Same 11-bit pieces as TF32 but
// half the bytes per element and K = 16 per instruction: twice the tensor rate and half as many truncating
// accumulator updates.  Valid for |v| < 65504 (activations here are bounded by tanh / GRU gates, weights are O(1));
// larger magnitudes surface as inf/NaN in the output rather than as silently wrong numbers.
__global__ void __launch_bounds__(256)
split_f16_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, __half* __restrict__ hi, __half* __restrict__ lo, int64_t ldo) {
    const int kq = K >> 2;
    const int64_t total = (int64_t)rows * kq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kq), c = (int)(i % kq) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)r * ldx + c);
        const float in[4] = {v.x, v.y, v.z, v.w};
        __half h[4], l[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            h[j] = __float2half_rn(in[j]);
            l[j] = __float2half_rn((in[j] - __half2float(h[j])) * 2048.0f);
        }
        *reinterpret_cast<uint2*>(hi + (int64_t)r * ldo + c) = *reinterpret_cast<uint2*>(h);
        *reinterpret_cast<uint2*>(lo + (int64_t)r * ldo + c) = *reinterpret_cast<uint2*>(l);
    }
}

// BF16 mode (north_star "bf16 mode": bf16 operands, FP32 accumulation): one plane, no error compensation.
__global__ void __launch_bounds__(256)
round_bf16_kernel(const float* __restrict__ x, int64_t ldx, int rows, int K, __nv_bfloat16* __restrict__ hi, int64_t ldo) {
    pdl_trigger();
    pdl_wait();
    const int kq = K >> 2;
    const int64_t total = (int64_t)rows * kq;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kq), c = (int)(i % kq) * 4;
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)r * ldx + c);
        __nv_bfloat16 h[4] = {__float2bfloat16_rn(v.x), __float2bfloat16_rn(v.y), __float2bfloat16_rn(v.z), __float2bfloat16_rn(v.w)};
        *reinterpret_cast<uint2*>(hi + (int64_t)r * ldo + c) = *reinterpret_cast<uint2*>(h);
    }
}

// ------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// One elected lane of a fully converged warp (the MMA-issuing warp runs its loops warp-uniformly so that descriptors and
// barrier addresses stay in uniform registers; only the tcgen05 instructions themselves are predicated on this lane —
// issuing from inside an `if (lane == 0)` region makes the compiler wrap EVERY MMA in an elect / R2UR broadcast loop).
__device__ __forceinline__ bool elect_one_lane() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <bool F16>
__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (F16)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, see cute/arch/mma_sm100_desc.hpp):
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B)   [46,48) version = 1   [61,64) layout = 2 (SW128)
// ROWB = bytes of one K-row in shared memory = the swizzle span: 128 (SWIZZLE_128B, layout 2) or 64 (SWIZZLE_64B,
// layout 4); the stride between 8-row groups is 8·ROWB.
template <int ROWB>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((8 * ROWB) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(ROWB == 128 ? 2 : 4) << 61;
    return d;
}

// Two tile configurations:
//   ROWB = 128 ("wide"): 128 x 256 tile, 128-byte K rows, 2 stages, one CTA per SM  — least L2 traffic per flop, for long K
//   ROWB =  64 ("dual"): 128 x 128 tile,  64-byte K rows, 3 stages, TWO CTAs per SM — one CTA's epilogue overlaps the
//                        other's main loop; for short-K / epilogue-heavy problems such as the vocabulary projection
template <int BN, bool F16, int ROWB>
struct TcCfg {
    static constexpr int BM = 128;
    static constexpr int ELT = F16 ? 2 : 4;
    static constexpr int BK = ROWB / ELT;  // elements per shared-memory K row
    static constexpr int UK = 32 / ELT;    // K of one MMA instruction (32 bytes): 8 (tf32) / 16 (f16)
    static constexpr int STAGES = ROWB == 64 ? 3 : (BN == 256 ? 2 : 3);
    static constexpr int CTAS_PER_SM = ROWB == 64 ? 2 : 1;
    static constexpr int A_BYTES = BM * ROWB;
    static constexpr int B_BYTES = BN * ROWB;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr uint32_t FMT = F16 ? 0u : 2u;  // F16F32Format: F16 = 0, TF32 = 2
    static constexpr uint32_t IDESC = (1u << 4) /*D=F32*/ | (FMT << 7) /*A*/ | (FMT << 10) /*B*/ |
                                      ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
};

template <int BN, bool F16, int ROWB>
__global__ void __launch_bounds__(192, (ROWB == 64 ? 2 : 1))
linear_split3_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                     const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                     float* __restrict__ y, int64_t ldy, const float* __restrict__ bias, int rows, int K, int N, int flags,
                     float4* __restrict__ summ) {
    using Cfg = TcCfg<BN, F16, ROWB>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + Cfg::STAGES;
    uint64_t* tmem_full_bar = empty_bar + Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m0 = blockIdx.y * Cfg::BM;
    const int n0 = blockIdx.x * BN;
    const int n_kb = (K + Cfg::BK - 1) / Cfg::BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(tmem_full_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {  // TMEM: 2 x BN fp32 accumulator columns (main | cross), a power of two ≥ 32
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)(2 * BN)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % Cfg::STAGES;
                const uint32_t ph = (kb / Cfg::STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                mbar_expect_tx(&full_bar[s], Cfg::STAGE_BYTES);
                const int k0 = kb * Cfg::BK;
                tma_load_2d(st, &map_xh, &full_bar[s], k0, m0);
                tma_load_2d(st + Cfg::A_BYTES, &map_xl, &full_bar[s], k0, m0);
                tma_load_2d(st + 2 * Cfg::A_BYTES, &map_wh, &full_bar[s], k0, n0);
                tma_load_2d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &map_wl, &full_bar[s], k0, n0);
            }
        }
    } else if (warp == 1) {
        {   // warp-uniform loop, one elected lane issues (see elect_one_lane)
            for (int kb = 0; kb < n_kb; ++kb) {
                const int s = kb % Cfg::STAGES;
                const uint32_t ph = (kb / Cfg::STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                tcgen05_fence_after();
                const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t d_ah = make_smem_desc<ROWB>(st);
                const uint64_t d_al = make_smem_desc<ROWB>(st + Cfg::A_BYTES);
                const uint64_t d_bh = make_smem_desc<ROWB>(st + 2 * Cfg::A_BYTES);
                const uint64_t d_bl = make_smem_desc<ROWB>(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                if (elect_one_lane()) {
#pragma unroll
                    for (int j = 0; j < Cfg::BK / Cfg::UK; ++j) {
                        const uint64_t adv = (uint64_t)((j * 32) >> 4);  // 32 B per k-step inside the 128 B swizzle row
                        umma<F16>(tmem_base + BN, d_al + adv, d_bh + adv, Cfg::IDESC, (kb | j) != 0);  // cross terms
                        umma<F16>(tmem_base + BN, d_ah + adv, d_bl + adv, Cfg::IDESC, 1);
                        umma<F16>(tmem_base, d_ah + adv, d_bh + adv, Cfg::IDESC, (kb | j) != 0);       // main term
                    }
                    tcgen05_commit(&empty_bar[s]);  // frees the slot once the MMAs above have read it
                    if (kb == n_kb - 1) tcgen05_commit(tmem_full_bar);      // accumulator complete
                }
                __syncwarp();
            }
        }
    } else {
        // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), +32)
        const int lg = warp & 3;
        mbar_wait(tmem_full_bar, 0);
        tcgen05_fence_after();
        const bool do_tanh = flags & VAG_LIN_TANH, do_acc = flags & VAG_LIN_ACCUMULATE;
        // the bias slice of this tile, staged once (the pipeline stages are idle now); epilogue warps only
        float* bias_s = reinterpret_cast<float*>(smem) + 4 * (32 * 33);
        for (int c = threadIdx.x - 64; c < BN; c += 128) bias_s[c] = (bias && n0 + c < N) ? bias[n0 + c] : 0.f;
        asm volatile("bar.sync 1, 128;" ::: "memory");
        // optional per-(row, tile) summary for the fused beam selection: running max, Σexp(x - max), arg-max
        float sm_m = -INFINITY, sm_s = 0.f, sm_bv = -INFINITY;
        int sm_bi = 0x7fffffff;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            if (n0 + c0 >= N) break;  // warp-uniform
            uint32_t r[32], q[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                  "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
                  "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
                  "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                : "r"(taddr + (uint32_t)BN)
                : "memory");
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                  "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                  "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                  "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                : "r"(taddr)
                : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j)
                r[j] = __float_as_uint((__uint_as_float(r[j]) + (F16 ? __uint_as_float(q[j]) * (1.0f / 2048.0f) : __uint_as_float(q[j]))) +
                                       bias_s[c0 + j]);
            if (summ) {
                constexpr float kL2e = 1.4426950408889634f;
                const int n_valid = min(32, N - (n0 + c0));   // ≥ 1 here
                float cm = -INFINITY;
#pragma unroll
                for (int j = 0; j < 32; ++j) cm = fmaxf(cm, j < n_valid ? __uint_as_float(r[j]) : -INFINITY);
                if (cm > sm_m) { sm_s *= exp2f((sm_m - cm) * kL2e); sm_m = cm; }
                const float m2 = sm_m * kL2e;
#pragma unroll
                for (int j = 0; j < 32; ++j) sm_s += j < n_valid ? exp2f(fmaf(__uint_as_float(r[j]), kL2e, -m2)) : 0.f;
                if (cm > sm_bv) {   // strictly greater: on ties the lower column (earlier chunk) wins
                    sm_bv = cm;
                    int first = 31;
#pragma unroll
                    for (int j = 31; j >= 0; --j) if (j < n_valid && __uint_as_float(r[j]) == cm) first = j;
                    sm_bi = n0 + c0 + first;
                }
            }
            // transpose through shared memory (the pipeline stages are idle now) so that every store instruction
            // writes 32 consecutive floats of ONE output row instead of one float in each of 32 rows
            float* tile = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 33);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 32; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]);
            __syncwarp();
            const int col = n0 + c0 + lane;
            if (col < N) {
                const int r_lim = min(32, rows - (m0 + lg * 32));
                float* dst = y + (int64_t)(m0 + lg * 32) * ldy + col;
#pragma unroll 4
                for (int rr = 0; rr < r_lim; ++rr) {
                    float v = tile[rr * 33 + lane];
                    if (do_acc) v = dst[(int64_t)rr * ldy] + v;
                    if (do_tanh) v = tanhf(v);
                    dst[(int64_t)rr * ldy] = v;
                }
            }
        }
        if (summ) {
            const int row = m0 + lg * 32 + lane;
            if (row < rows) summ[(int64_t)row * gridDim.x + blockIdx.x] = make_float4(sm_m, sm_s, sm_bv, __int_as_float(sm_bi));
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)(2 * BN)) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ persistent kernel
// One CTA per SM walks a static round-robin list of 128 x 128 output tiles (column index fastest, so the CTAs that
// run together share an A row block through L2).  Ten warps, three roles that never join until the end:
//   warp 0      TMA producer — runs AHEAD across tile boundaries through a PSTAGES-deep ring (64-byte K rows);
//   warp 1      MMA issuer — two TMEM accumulator buffers (main | cross each) alternate between tiles;
//   warps 2-9   epilogue — drains buffer a while the MMA warp already fills buffer a^1: tcgen05.ld → (+bias, tanh,
//               optional soft-max / arg-max summary) → 128-byte-swizzled shared tile → TMA store (no per-row stores).
// Synchronisation is mbarrier-only: full/empty per ring slot, tmem_full/tmem_empty per accumulator buffer.
constexpr int PSTAGES = 5;
constexpr int P_STAGE_BYTES = 4 * 128 * 64;          // A_hi, A_lo, B_hi, B_lo: 128 rows x 64 B each = 32 KB
constexpr int P_EPI_BYTES = 8 * 4096;                // one 32 x 32 fp32 staging tile per epilogue warp
constexpr int P_SMEM_BYTES = PSTAGES * P_STAGE_BYTES + P_EPI_BYTES + 512 /*bias*/ + 2048 /*summary exchange*/ + 256 /*barriers*/ + 1024;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}

// MODE: 0 = TF32 split (3 products), 1 = FP16 split (3 products), 2 = BF16 single product (bf16 mode)
template <int MODE>
__global__ void __launch_bounds__(320, 1)
linear_split3_persistent_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                                const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                                const __grid_constant__ CUtensorMap map_y, const float* __restrict__ bias, int rows, int K, int N,
                                int flags, float4* __restrict__ summ) {
    constexpr bool F16 = MODE != 0;              // 16-bit operands (kind::f16) vs TF32
    constexpr bool SPLIT = MODE != 2;            // hi/lo error-compensated products
    constexpr int BM = 128, BN = 128, ELT = F16 ? 2 : 4, BK = 64 / ELT, UK = 32 / ELT;
    constexpr uint32_t FMT = MODE == 0 ? 2u : (MODE == 1 ? 0u : 1u);   // F16F32Format: F16 = 0, BF16 = 1, TF32 = 2
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    constexpr int T_BYTES = 128 * 64;  // one operand tile
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi = smem + PSTAGES * P_STAGE_BYTES;                       // 8 x 4 KB, 1024-B aligned
    float* bias_s = reinterpret_cast<float*>(epi + P_EPI_BYTES);         // [128]
    float4* xch = reinterpret_cast<float4*>(epi + P_EPI_BYTES + 512);    // [128] partial summaries of the right column half
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + P_EPI_BYTES + 512 + 2048);
    uint64_t* empty_bar = full_bar + PSTAGES;
    uint64_t* tfull_bar = empty_bar + PSTAGES;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (rows + BM - 1) / BM;
    const int n_tiles = tiles_n * tiles_m;
    const int n_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        for (int s = 0; s < PSTAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % PSTAGES;
                    mbar_wait(&empty_bar[s], ((g / PSTAGES) & 1) ^ 1);
                    uint8_t* st = smem + s * P_STAGE_BYTES;
                    mbar_expect_tx(&full_bar[s], SPLIT ? P_STAGE_BYTES : P_STAGE_BYTES / 2);
                    const int k0 = kb * BK;
                    tma_load_2d(st, &map_xh, &full_bar[s], k0, m0);
                    if (SPLIT) tma_load_2d(st + T_BYTES, &map_xl, &full_bar[s], k0, m0);
                    tma_load_2d(st + 2 * T_BYTES, &map_wh, &full_bar[s], k0, n0);
                    if (SPLIT) tma_load_2d(st + 3 * T_BYTES, &map_wl, &full_bar[s], k0, n0);
                }
            }
        }
    } else if (warp == 1) {
        {   // warp-uniform loops, one elected lane issues (see elect_one_lane)
            uint32_t g = 0, it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const uint32_t a = it & 1;
                mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);   // the epilogue has drained this accumulator buffer
                tcgen05_fence_after();
                const uint32_t d_main = tmem_base + a * 256, d_cross = d_main + 128;
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % PSTAGES;
                    mbar_wait(&full_bar[s], (g / PSTAGES) & 1);
                    tcgen05_fence_after();
                    const uint32_t st = smem_u32(smem + s * P_STAGE_BYTES);
                    const uint64_t d_ah = make_smem_desc<64>(st), d_al = make_smem_desc<64>(st + T_BYTES);
                    const uint64_t d_bh = make_smem_desc<64>(st + 2 * T_BYTES), d_bl = make_smem_desc<64>(st + 3 * T_BYTES);
                    if (elect_one_lane()) {
#pragma unroll
                        for (int j = 0; j < BK / UK; ++j) {
                            const uint64_t adv = (uint64_t)((j * 32) >> 4);
                            if (SPLIT) {
                                umma<F16>(d_cross, d_al + adv, d_bh + adv, IDESC, (kb | j) != 0);
                                umma<F16>(d_cross, d_ah + adv, d_bl + adv, IDESC, 1);
                            }
                            umma<F16>(d_main, d_ah + adv, d_bh + adv, IDESC, (kb | j) != 0);
                        }
                        tcgen05_commit(&empty_bar[s]);
                        if (kb == n_kb - 1) tcgen05_commit(&tfull_bar[a]);
                    }
                    __syncwarp();
                }
            }
        }
    } else {
        // ---- epilogue warps 2..9: TMEM lane group lg = warp % 4, column half ch (64 columns = two 32-wide chunks)
        const int ew = warp - 2, lg = warp & 3, ch = ew >> 2;
        uint8_t* my_tile = epi + ew * 4096;
        const bool do_tanh = flags & VAG_LIN_TANH;
        constexpr float kL2e = 1.4426950408889634f;
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
            const int tn = tile % tiles_n;
            const uint32_t a = it & 1;
            // stage this tile's bias slice: first make sure every epilogue warp has finished reading the previous one
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ew < 4) {
                const int c = ew * 32 + lane;
                bias_s[c] = (bias && n0 + c < N) ? bias[n0 + c] : 0.f;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tfull_bar[a], (it >> 1) & 1);
            tcgen05_fence_after();
            float sm_m = -INFINITY, sm_s = 0.f, sm_bv = -INFINITY;
            int sm_bi = 0x7fffffff;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = ch * 64 + cc * 32;
                if (n0 + c0 >= N) break;  // warp-uniform: this chunk lies entirely outside the matrix
                uint32_t r[32], q[32];
                const uint32_t taddr = tmem_base + a * 256 + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
                if (SPLIT)
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]),
                      "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]), "=r"(q[14]), "=r"(q[15]), "=r"(q[16]),
                      "=r"(q[17]), "=r"(q[18]), "=r"(q[19]), "=r"(q[20]), "=r"(q[21]), "=r"(q[22]), "=r"(q[23]), "=r"(q[24]),
                      "=r"(q[25]), "=r"(q[26]), "=r"(q[27]), "=r"(q[28]), "=r"(q[29]), "=r"(q[30]), "=r"(q[31])
                    : "r"(taddr + 128u)
                    : "memory");
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr)
                    : "memory");
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (cc == 1 || n0 + c0 + 32 >= N) {
                    // all TMEM reads of this warp for this tile are complete: hand the buffer back to the MMA warp early
                    tcgen05_fence_before();
                    if (lane == 0) mbar_arrive(&tempty_bar[a]);
                }
                float x[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float cross = !SPLIT ? 0.f : (F16 ? __uint_as_float(q[j]) * (1.0f / 2048.0f) : __uint_as_float(q[j]));
                    x[j] = (__uint_as_float(r[j]) + cross) + bias_s[c0 + j];
                    if (do_tanh) x[j] = tanhf(x[j]);
                }
                if (summ) {
                    const int n_valid = min(32, N - (n0 + c0));
                    float cm = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, j < n_valid ? x[j] : -INFINITY);
                    if (cm > sm_m) { sm_s *= exp2f((sm_m - cm) * kL2e); sm_m = cm; }
                    const float m2 = sm_m * kL2e;
#pragma unroll
                    for (int j = 0; j < 32; ++j) sm_s += j < n_valid ? exp2f(fmaf(x[j], kL2e, -m2)) : 0.f;
                    if (cm > sm_bv) {
                        sm_bv = cm;
                        int first = 31;
#pragma unroll
                        for (int j = 31; j >= 0; --j) if (j < n_valid && x[j] == cm) first = j;
                        sm_bi = n0 + c0 + first;
                    }
                }
                // registers → 128-byte-swizzled staging tile (row = lane, 16-byte chunk index XOR row%8) → TMA store
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // the previous store has finished reading the tile
                __syncwarp();
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    float4 v = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
                    *reinterpret_cast<float4*>(my_tile + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = v;
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_y, my_tile, n0 + c0, m0 + lg * 32);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
            if (n0 + ch * 64 >= N) {   // this warp had no chunk at all in this tile: it still owes the arrival
                tcgen05_fence_before();
                if (lane == 0) mbar_arrive(&tempty_bar[a]);
            }
            if (summ) {   // combine the two column halves of every row and write the (row, tile) summary
                const int rl = lg * 32 + lane;
                if (ch == 1) xch[rl] = make_float4(sm_m, sm_s, sm_bv, __int_as_float(sm_bi));
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (ch == 0) {
                    const float4 o = xch[rl];
                    float m = sm_m, s_ = sm_s, bv = sm_bv;
                    int bi = sm_bi;
                    if (o.x > m) { s_ = s_ * exp2f((m - o.x) * kL2e) + o.y; m = o.x; }
                    else if (o.x != -INFINITY) s_ += o.y * exp2f((o.x - m) * kL2e);
                    if (o.z > bv) { bv = o.z; bi = __float_as_int(o.w); }   // strictly greater: the left half wins ties
                    const int row = m0 + rl;
                    if (row < rows) summ[(int64_t)row * tiles_n + tn] = make_float4(m, s_, bv, __int_as_float(bi));
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------ CTA-pair kernel
// The contractions of the decode step are bound by L2 → shared-memory traffic, not by the tensor pipe (ncu:
// l1tex__m_xbar2l1tex_read_bytes ≈ 9 TB/s ≈ 85 % of the measured chip-wide cap while the tensor pipe idles half the
// time): a 128 x 128 tile with split operands pulls 32 KB per 64-byte K block for six MMAs.  This kernel pairs the two
// SMs of a TPC (cluster of 2, tcgen05 cta_group::2): the pair computes a 256 x 128 tile, every CTA loads its own 128
// activation rows but only HALF of the weight tile (64 rows) — the MMA reads the other half from the peer's shared
// memory — so the same six MMAs need 24 KB.  Roles per CTA as in the persistent kernel (warp 0 TMA producer, warp 1
// MMA issuer — leader CTA only, warps 2-9 epilogue); barriers:
//   full[s]    leader's shared memory, one arrival (leader's expect_tx of BOTH CTAs' bytes) — both producers' TMA loads
//              complete on it (cp.async.bulk.tensor ... cta_group::2 may signal the peer's barrier)
//   empty[s]   one per CTA, released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[a]   one per CTA, same multicast commit;  tempty[a]  leader's, 16 arrivals (8 epilogue warps x 2 CTAs)
// The epilogue owns TWO staging tiles per warp so that a TMA store can drain while the next chunk is being written
// (the single-buffer version was store-latency bound: 26 us for a K = 32 contraction with 73 MB of output).
constexpr int Q_STAGES = 3;
constexpr int Q_ROWB = 128;                                    // bytes of one K row in shared memory (SWIZZLE_128B)
constexpr int Q_A_BYTES = 128 * Q_ROWB, Q_B_BYTES = 64 * Q_ROWB;
constexpr int Q_STAGE_BYTES = 2 * Q_A_BYTES + 2 * Q_B_BYTES;   // A_hi, A_lo (128 rows), B_hi, B_lo (64 rows): 48 KB
constexpr int Q_EPI_BYTES = 8 * 2 * 4096;                      // two 32 x 32 fp32 staging tiles per epilogue warp
constexpr int Q_SMEM_BYTES = Q_STAGES * Q_STAGE_BYTES + Q_EPI_BYTES + 1024 /*bias per warp*/ + 4096 /*summary exchange x2*/ + 256 /*barriers*/ + 1024;

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
template <bool F16>
__device__ __forceinline__ void umma_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    if (F16)
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
    else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
            : "memory");
}
#define VAG_TMEM_LD32(v, addr)                                                                                                  \
    asm volatile(                                                                                                               \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                               \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                               \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                               \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),           \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),                \
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),               \
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                             \
        : "r"(addr)                                                                                                             \
        : "memory")

#define VAG_TMEM_LD16(v, addr)                                                                                                  \
    asm volatile(                                                                                                               \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                               \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                                        \
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),           \
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                              \
        : "r"(addr)                                                                                                             \
        : "memory")

// Role timers (tools/scratch/roles.py): compiled in with -DVAG_TC_TIMERS, otherwise every VAG_TCLK() folds to zero.
#ifdef VAG_TC_TIMERS
#define VAG_TCLK() clock64()
#else
#define VAG_TCLK() 0LL
#endif
template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(320, 1)
linear_pair_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                   const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                   const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_yh,
                   const __grid_constant__ CUtensorMap map_yl, const float* __restrict__ bias, int rows, int K, int N,
                   int flags, float4* __restrict__ summ, long long* __restrict__ dbg, const int* __restrict__ done) {
    constexpr bool F16 = MODE != 0;
    constexpr bool SPLIT = MODE != 2;
    constexpr int BMP = 256, BM = 128, BN = 128, ELT = F16 ? 2 : 4, BK = Q_ROWB / ELT, UK = 32 / ELT;
    constexpr uint32_t FMT = MODE == 0 ? 2u : (MODE == 1 ? 0u : 1u);
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BMP >> 4) << 24);
    // bf16 mode has no lo planes: its stages are half as large, so twice as many fit (it needs the depth: one product per
    // K step means a stage is consumed three times faster)
    constexpr int NST = SPLIT ? Q_STAGES : 2 * Q_STAGES;
    constexpr int A_USED = SPLIT ? 2 * Q_A_BYTES : Q_A_BYTES, B_USED = SPLIT ? 2 * Q_B_BYTES : Q_B_BYTES;
    constexpr int STG = A_USED + B_USED;
    constexpr uint32_t STAGE_TX = STG;   // bytes ONE CTA's loads deliver per stage
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* epi = smem + Q_STAGES * Q_STAGE_BYTES;                      // 16 x 4 KB, 1024-B aligned
    float* bias_w = reinterpret_cast<float*>(epi + Q_EPI_BYTES);         // [8 warps][32]
    float4* xch2 = reinterpret_cast<float4*>(epi + Q_EPI_BYTES + 1024);  // [2][128], alternating between tiles
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi + Q_EPI_BYTES + 1024 + 4096);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tfull_bar = empty_bar + NST;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (rows + BMP - 1) / BMP;
    const int n_tiles = tiles_n * tiles_m;
    const int n_kb = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_y) : "memory");
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], rank == 0 ? 9 : 8); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();   // the next kernel on the stream may start its own prologue as SMs drain
    pdl_wait();      // everything above touched no global memory; from here on the predecessor's results are needed
    // beam search: every hypothesis has ended — the flag was set by an EARLIER kernel of the stream, so all threads of all CTAs
    // read the same value and skip their role loops together (barriers untouched, TMEM released below)
    const bool dead = done && *reinterpret_cast<const volatile int*>(done) != 0;

    if (dead) {
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            long long t_wait = 0, t_begin = VAG_TCLK();
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int m0 = (tile / tiles_n) * BMP + (int)rank * BM, n0 = (tile % tiles_n) * BN + (int)rank * (BN / 2);
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % NST;
                    const long long t0 = VAG_TCLK();
#ifdef VAG_EXP_NOLOAD
                    if (g >= NST) continue;
#endif
                    mbar_wait(&empty_bar[s], ((g / NST) & 1) ^ 1);
                    t_wait += VAG_TCLK() - t0;
                    uint8_t* st = smem + s * STG;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STAGE_TX);
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
                    const int k0 = kb * BK;
                    tma_load_2d_pair(st, &map_xh, fb, k0, m0);
                    if (SPLIT) tma_load_2d_pair(st + Q_A_BYTES, &map_xl, fb, k0, m0);
                    tma_load_2d_pair(st + A_USED, &map_wh, fb, k0, n0);
                    if (SPLIT) tma_load_2d_pair(st + A_USED + Q_B_BYTES, &map_wl, fb, k0, n0);
                }
            }
            if (dbg && pair == 0) { dbg[rank * 16 + 0] = VAG_TCLK() - t_begin; dbg[rank * 16 + 1] = t_wait; dbg[rank * 16 + 2] = g; }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // the whole warp walks the loops (uniform values); one elected lane issues the tcgen05 instructions
            uint32_t g = 0, it = 0;
            long long t_we = 0, t_wf = 0, t_begin = VAG_TCLK();
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
                const uint32_t a = it & 1;
                long long t0 = VAG_TCLK();
                mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);   // both CTAs' epilogues have drained this buffer
                t_we += VAG_TCLK() - t0;
                tcgen05_fence_after();
                const uint32_t d_main = tmem_base + a * 256, d_cross = d_main + 128;
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % NST;
                    t0 = VAG_TCLK();
#ifdef VAG_EXP_NOLOAD
                    if (g < NST)
#endif
                    mbar_wait(&full_bar[s], (g / NST) & 1);
                    t_wf += VAG_TCLK() - t0;
                    tcgen05_fence_after();
                    const uint32_t st = smem_u32(smem + s * STG);
                    const uint64_t d_ah = make_smem_desc<Q_ROWB>(st), d_al = make_smem_desc<Q_ROWB>(st + Q_A_BYTES);
                    const uint64_t d_bh = make_smem_desc<Q_ROWB>(st + A_USED), d_bl = make_smem_desc<Q_ROWB>(st + A_USED + Q_B_BYTES);
                    if (elect_one_lane()) {
#pragma unroll
                        for (int j = 0; j < BK / UK; ++j) {
                            const uint64_t adv = (uint64_t)((j * 32) >> 4);
                            if (SPLIT) {
                                umma_pair<F16>(d_cross, d_al + adv, d_bh + adv, IDESC, (kb | j) != 0);
                                umma_pair<F16>(d_cross, d_ah + adv, d_bl + adv, IDESC, 1);
                            }
                            umma_pair<F16>(d_main, d_ah + adv, d_bh + adv, IDESC, (kb | j) != 0);
                        }
                        tcgen05_commit_pair(&empty_bar[s]);
                        if (kb == n_kb - 1) tcgen05_commit_pair(&tfull_bar[a]);
                    }
                    __syncwarp();
                }
            }
            if (dbg && pair == 0 && lane == 0) { dbg[4] = VAG_TCLK() - t_begin; dbg[5] = t_we; dbg[6] = t_wf; dbg[7] = it; }
        } else if (rank == 1 && lane == 0) {
            // the peer's MMA warp is idle: it forwards "my eight epilogue warps have drained buffer a" to the leader with
            // ONE remote arrival per tile, keeping the cluster-scope release off the epilogue warps' critical path
            uint32_t it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
                const uint32_t a = it & 1;
                mbar_wait(&tempty_bar[a], (it >> 1) & 1);
                mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[a]), 0));
            }
        }
    } else {
        // ---- epilogue warps 2..9: TMEM lane group lg = warp % 4, column half ch (64 columns = two 32-wide chunks)
        const int ew = warp - 2, lg = warp & 3, ch = ew >> 2;
        uint8_t* my_tiles = epi + ew * 8192;
        float* my_bias = bias_w + ew * 32;
        const bool do_tanh = flags & VAG_LIN_TANH;
        const bool split_out = flags & VAG_LIN_SPLIT_OUT;   // the output leaves as tensor-core operand planes, not fp32
        constexpr float kL2e = 1.4426950408889634f;
        uint32_t it = 0, nst = 0;
        long long t_wt = 0, t_ld = 0, t_ws = 0, t_math = 0, t_stage = 0, t_tma = 0, t_begin = VAG_TCLK();
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
            const int tn = tile % tiles_n;
            const int m0 = (tile / tiles_n) * BMP + (int)rank * BM, n0 = tn * BN;
            const uint32_t a = it & 1;
            long long t0 = VAG_TCLK();
            mbar_wait(&tfull_bar[a], (it >> 1) & 1);
            t_wt += VAG_TCLK() - t0;
            tcgen05_fence_after();
            float sm_m = -INFINITY, sm_s = 0.f, sm_bv = -INFINITY;
            int sm_bi = 0x7fffffff;
            bool released = false;
#pragma unroll 1
            for (int cc = 0; cc < 2; ++cc) {
                const int c0 = ch * 64 + cc * 32;
                if (n0 + c0 >= N) break;  // warp-uniform: this chunk lies entirely outside the matrix
                uint32_t r[32], q[32];
                const uint32_t taddr = tmem_base + a * 256 + ((uint32_t)(lg * 32) << 16) + (uint32_t)c0;
                t0 = VAG_TCLK();
                if (SPLIT) VAG_TMEM_LD32(q, taddr + 128u);
                VAG_TMEM_LD32(r, taddr);
                __syncwarp();
                my_bias[lane] = (bias && n0 + c0 + lane < N) ? bias[n0 + c0 + lane] : 0.f;
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                t_ld += VAG_TCLK() - t0;
                if (cc == 1 || n0 + c0 + 32 >= N) {
                    // all TMEM reads of this warp for this tile are complete: hand the buffer back to the MMA warp early
                    tcgen05_fence_before();
                    if (lane == 0) mbar_arrive(&tempty_bar[a]);
                    released = true;
                }
                __syncwarp();
                t0 = VAG_TCLK();
                float x[32];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(my_bias + 4 * j4);
                    const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = 4 * j4 + u;
                        const float cross = !SPLIT ? 0.f : (F16 ? __uint_as_float(q[j]) * (1.0f / 2048.0f) : __uint_as_float(q[j]));
                        x[j] = (__uint_as_float(r[j]) + cross) + bb[u];
                        if (do_tanh) x[j] = tanhf(x[j]);
                    }
                }
                if (summ) {
                    const int n_valid = min(32, N - (n0 + c0));
                    float cm = -INFINITY;
#pragma unroll
                    for (int j = 0; j < 32; ++j) cm = fmaxf(cm, j < n_valid ? x[j] : -INFINITY);
                    if (cm > sm_m) { sm_s *= exp2f((sm_m - cm) * kL2e); sm_m = cm; }
                    const float m2 = sm_m * kL2e;
#pragma unroll
                    for (int j = 0; j < 32; ++j) sm_s += j < n_valid ? exp2f(fmaf(x[j], kL2e, -m2)) : 0.f;
                    if (cm > sm_bv) {
                        sm_bv = cm;
                        int first = 31;
#pragma unroll
                        for (int j = 31; j >= 0; --j) if (j < n_valid && x[j] == cm) first = j;
                        sm_bi = n0 + c0 + first;
                    }
                }
                // registers → 128-byte-swizzled staging tile (row = lane, 16-byte chunk index XOR row%8) → TMA store;
                // two tiles alternate, so only the store issued TWO chunks ago has to have finished reading
                t_math += VAG_TCLK() - t0;
                uint8_t* my_tile = my_tiles + (nst & 1) * 4096;
                ++nst;
                t0 = VAG_TCLK();
                asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                __syncwarp();
                t_ws += VAG_TCLK() - t0;
                t0 = VAG_TCLK();
                if (!split_out) {
#pragma unroll
                    for (int c4 = 0; c4 < 8; ++c4) {
                        float4 v = make_float4(x[4 * c4], x[4 * c4 + 1], x[4 * c4 + 2], x[4 * c4 + 3]);
                        *reinterpret_cast<float4*>(my_tile + lane * 128 + ((c4 ^ (lane & 7)) << 4)) = v;
                    }
                } else {
                    // hi plane tile at +0, lo plane tile at +2048: 32 rows x 64 B, SWIZZLE_64B (chunk ^= (row >> 1) & 3)
#pragma unroll
                    for (int c8 = 0; c8 < 4; ++c8) {
                        uint32_t hw[4], lw[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            uint16_t h0, l0, h1, l1;
                            split_one(MODE, x[8 * c8 + 2 * u], h0, l0);
                            split_one(MODE, x[8 * c8 + 2 * u + 1], h1, l1);
                            hw[u] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                            lw[u] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                        }
                        const int off = lane * 64 + ((c8 ^ ((lane >> 1) & 3)) << 4);
                        *reinterpret_cast<uint4*>(my_tile + off) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                        if (SPLIT) *reinterpret_cast<uint4*>(my_tile + 2048 + off) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                t_stage += VAG_TCLK() - t0;
                t0 = VAG_TCLK();
                if (lane == 0) {
                    if (!split_out) {
                        tma_store_2d(&map_y, my_tile, n0 + c0, m0 + lg * 32);
                    } else {
                        tma_store_2d(&map_yh, my_tile, n0 + c0, m0 + lg * 32);
                        if (SPLIT) tma_store_2d(&map_yl, my_tile + 2048, n0 + c0, m0 + lg * 32);
                    }
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                __syncwarp();
                t_tma += VAG_TCLK() - t0;
            }
            if (!released) {   // this warp had no chunk at all in this tile: it still owes the arrival
                tcgen05_fence_before();
                if (lane == 0) mbar_arrive(&tempty_bar[a]);
            }
            if (summ) {   // combine the two column halves of every row and write the (row, tile) summary
                const int rl = lg * 32 + lane;
                float4* xch = xch2 + (it & 1) * 128;   // the other copy may still be read by a slower warp
                if (ch == 1) xch[rl] = make_float4(sm_m, sm_s, sm_bv, __int_as_float(sm_bi));
                asm volatile("bar.sync 2, 256;" ::: "memory");
                if (ch == 0) {
                    const float4 o = xch[rl];
                    float m = sm_m, s_ = sm_s, bv = sm_bv;
                    int bi = sm_bi;
                    if (o.x > m) { s_ = s_ * exp2f((m - o.x) * kL2e) + o.y; m = o.x; }
                    else if (o.x != -INFINITY) s_ += o.y * exp2f((o.x - m) * kL2e);
                    if (o.z > bv) { bv = o.z; bi = __float_as_int(o.w); }   // strictly greater: the left half wins ties
                    const int row = m0 + rl;
                    if (row < rows) summ[(int64_t)row * tiles_n + tn] = make_float4(m, s_, bv, __int_as_float(bi));
                }
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        if (dbg && pair == 0 && ew == 0 && lane == 0) {
            dbg[rank * 16 + 8] = VAG_TCLK() - t_begin; dbg[rank * 16 + 9] = t_wt; dbg[rank * 16 + 10] = t_ld; dbg[rank * 16 + 11] = t_ws; dbg[rank * 16 + 12] = t_math; dbg[rank * 16 + 13] = t_stage; dbg[rank * 16 + 14] = t_tma;
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();   // the peer's shared memory and the leader's barriers stay alive until both CTAs are done
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}


// ------------------------------------------------------------------------------------------ vocabulary top-2 kernel
// The vocabulary projection of the beam loop.  Same CTA-pair main loop, but NO output matrix: each of SIXTEEN epilogue
// warps owns one 32-column chunk of the CTA's 128 x 128 accumulator and reduces it, per row, to
//     (best logit, Σ exp(logit − best), second-best logit, best column | second column << 16)       [canonical order]
// written to summ[slice][row] (slice = 32-column chunk of the vocabulary, row fastest ⇒ coalesced 512-byte stores).
// beam_select_top2_kernel picks the K winners from these 294 x 16 B per row instead of 9391 x 4 B of logits.
// Why 16 warps: with K = 256 the three-product main loop lasts only ≈ 3 k cycles per tile, and the reduction (≈ 12
// instructions per element: two-level top-2 + Σexp) is latency-bound with two warps per scheduler; four per scheduler
// hide it.  No staging tiles ⇒ room for a fourth pipeline stage.
constexpr int V_STAGES = 4;
constexpr int V_SMEM_BYTES = V_STAGES * Q_STAGE_BYTES + 2048 /*bias per warp*/ + 256 /*barriers*/ + 1024;   // stationary mode: 64 KB resident weights + 4 x 32 KB

template <int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(576, 1)
vocab_top2_pair_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                       const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl,
                       const float* __restrict__ bias, int rows, int K, int N, float4* __restrict__ summ,
                       long long* __restrict__ dbg, const int* __restrict__ done) {
    constexpr bool F16 = MODE != 0;
    constexpr bool SPLIT = MODE != 2;
    constexpr int BMP = 256, BM = 128, BN = 128, ELT = F16 ? 2 : 4, BK = Q_ROWB / ELT, UK = 32 / ELT;
    constexpr uint32_t FMT = MODE == 0 ? 2u : (MODE == 1 ? 0u : 1u);
    constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BMP >> 4) << 24);
    constexpr int NST = SPLIT ? V_STAGES : 2 * V_STAGES;           // bf16 mode: half-size stages, twice as many
    constexpr int A_USED = SPLIT ? 2 * Q_A_BYTES : Q_A_BYTES, B_USED = SPLIT ? 2 * Q_B_BYTES : Q_B_BYTES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* bias_w = reinterpret_cast<float*>(smem + V_STAGES * Q_STAGE_BYTES);   // [16 warps][32]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + V_STAGES * Q_STAGE_BYTES + 2048);
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tfull_bar = empty_bar + NST;   // [2]
    uint64_t* tempty_bar = tfull_bar + 2;         // [2]
    uint64_t* bfull_bar = tempty_bar + 2;         // weight tile resident (leader's; both CTAs' loads complete on it)
    uint64_t* bempty_bar = bfull_bar + 1;         // every MMA that reads the resident weight tile has finished
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bempty_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int tiles_n = (N + BN - 1) / BN, tiles_m = (rows + BMP - 1) / BMP;
    const int n_tiles = tiles_n * tiles_m;
    const int n_kb = (K + BK - 1) / BK;
    // Every CTA pair walks a CONTIGUOUS range of tiles, row blocks fastest, so that consecutive tiles share their 128-column
    // weight tile.  Weight-stationary mode (K ≤ 4 K-blocks, i.e. the whole [64 rows x K] half tile fits in 64 KB): the weight
    // tile is loaded ONCE per column tile into a resident region and only the activations stream through the ring — a third
    // less L2 → shared-memory traffic (the kernel is L2-fed-bound: 48 KB per K block for 12 MMAs).
    const int t_begin = (int)((int64_t)pair * n_tiles / n_pairs), t_end = (int)((int64_t)(pair + 1) * n_tiles / n_pairs);
    const bool stationary = n_kb <= 4;
    constexpr int B_KB_BYTES = B_USED;                              // B_hi (+ B_lo) of one K block: 16 KB (8 KB in bf16 mode)
    const int stage_bytes = stationary ? A_USED : A_USED + B_USED;
    uint8_t* bres = smem;                                           // resident weight tile (stationary mode): 4 x 16 KB
    uint8_t* ring = stationary ? smem + 4 * B_KB_BYTES : smem;      // NST stages of A only or A + B
    constexpr uint32_t A_TX = A_USED, B_TX = B_USED;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xl) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
        mbar_init(bfull_bar, 1);
        mbar_init(bempty_bar, 1);
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], rank == 0 ? 17 : 16); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();   // the next kernel on the stream may start its own prologue as SMs drain
    pdl_wait();      // everything above touched no global memory; from here on the predecessor's results are needed
    const bool dead = done && *reinterpret_cast<const volatile int*>(done) != 0;   // see linear_pair_kernel

    if (dead) {
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0, nb = 0;
            int cur_tn = -1;
            long long t_wait = 0, t_start = VAG_TCLK();
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int tn = tile / tiles_m, tm = tile - tn * tiles_m;
                const int m0 = tm * BMP + (int)rank * BM, n0 = tn * BN + (int)rank * (BN / 2);
                if (stationary && tn != cur_tn) {
                    if (nb > 0) mbar_wait(bempty_bar, (nb - 1) & 1);   // the previous column tile's MMAs are done with the region
                    if (rank == 0) mbar_expect_tx(bfull_bar, 2 * n_kb * B_TX);
                    const uint32_t bb = mapa_u32(smem_u32(bfull_bar), 0);
                    for (int kb = 0; kb < n_kb; ++kb) {
                        tma_load_2d_pair(bres + kb * B_KB_BYTES, &map_wh, bb, kb * BK, n0);
                        if (SPLIT) tma_load_2d_pair(bres + kb * B_KB_BYTES + Q_B_BYTES, &map_wl, bb, kb * BK, n0);
                    }
                    cur_tn = tn;
                    ++nb;
                }
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % NST;
                    const long long t0 = VAG_TCLK();
                    mbar_wait(&empty_bar[s], ((g / NST) & 1) ^ 1);
                    t_wait += VAG_TCLK() - t0;
                    uint8_t* st = ring + s * stage_bytes;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * (A_TX + (stationary ? 0u : B_TX)));
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
                    const int k0 = kb * BK;
                    tma_load_2d_pair(st, &map_xh, fb, k0, m0);
                    if (SPLIT) tma_load_2d_pair(st + Q_A_BYTES, &map_xl, fb, k0, m0);
                    if (!stationary) {
                        tma_load_2d_pair(st + A_USED, &map_wh, fb, k0, n0);
                        if (SPLIT) tma_load_2d_pair(st + A_USED + Q_B_BYTES, &map_wl, fb, k0, n0);
                    }
                }
            }
            if (dbg && pair == 0) { dbg[rank * 16 + 0] = VAG_TCLK() - t_start; dbg[rank * 16 + 1] = t_wait; dbg[rank * 16 + 2] = g; }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // the whole warp walks the loops (uniform values); one elected lane issues the tcgen05 instructions
            uint32_t g = 0, it = 0, nb = 0;
            int cur_tn = -1;
            long long t_we = 0, t_wf = 0, t_start = VAG_TCLK();
            for (int tile = t_begin; tile < t_end; ++tile, ++it) {
                const int tn = tile / tiles_m;
                const uint32_t a = it & 1;
                long long t0 = VAG_TCLK();
                mbar_wait(&tempty_bar[a], ((it >> 1) & 1) ^ 1);
                t_we += VAG_TCLK() - t0;
                if (stationary && tn != cur_tn) {
                    mbar_wait(bfull_bar, nb & 1);
                    ++nb;
                    cur_tn = tn;
                }
                tcgen05_fence_after();
                const uint32_t d_main = tmem_base + a * 256, d_cross = d_main + 128;
                for (int kb = 0; kb < n_kb; ++kb, ++g) {
                    const int s = g % NST;
                    t0 = VAG_TCLK();
                    mbar_wait(&full_bar[s], (g / NST) & 1);
                    t_wf += VAG_TCLK() - t0;
                    tcgen05_fence_after();
                    const uint32_t st = smem_u32(ring + s * stage_bytes);
                    const uint32_t sb = stationary ? smem_u32(bres + kb * B_KB_BYTES) : st + A_USED;
                    const uint64_t d_ah = make_smem_desc<Q_ROWB>(st), d_al = make_smem_desc<Q_ROWB>(st + Q_A_BYTES);
                    const uint64_t d_bh = make_smem_desc<Q_ROWB>(sb), d_bl = make_smem_desc<Q_ROWB>(sb + Q_B_BYTES);
                    if (elect_one_lane()) {
#pragma unroll
                        for (int j = 0; j < BK / UK; ++j) {
                            const uint64_t adv = (uint64_t)((j * 32) >> 4);
                            if (SPLIT) {
                                umma_pair<F16>(d_cross, d_al + adv, d_bh + adv, IDESC, (kb | j) != 0);
                                umma_pair<F16>(d_cross, d_ah + adv, d_bl + adv, IDESC, 1);
                            }
                            umma_pair<F16>(d_main, d_ah + adv, d_bh + adv, IDESC, (kb | j) != 0);
                        }
                        tcgen05_commit_pair(&empty_bar[s]);
                        if (kb == n_kb - 1) {
                            tcgen05_commit_pair(&tfull_bar[a]);
                            if (stationary && (tile + 1 == t_end || (tile + 1) / tiles_m != tn)) tcgen05_commit_pair(bempty_bar);
                        }
                    }
                    __syncwarp();
                }
            }
            if (dbg && pair == 0 && lane == 0) { dbg[4] = VAG_TCLK() - t_start; dbg[5] = t_we; dbg[6] = t_wf; dbg[7] = it; }
        } else if (rank == 1 && lane == 0) {
            uint32_t it = 0;   // forward "my sixteen epilogue warps have drained buffer a" to the leader
            for (int tile = t_begin; tile < t_end; ++tile, ++it) {
                const uint32_t a = it & 1;
                mbar_wait(&tempty_bar[a], (it >> 1) & 1);
                mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[a]), 0));
            }
        }
    } else {
        // ---- epilogue warps 2..17: TMEM lane group lg = warp % 4, column quarter cq: ONE 32-column chunk per tile
        const int ew = warp - 2, lg = warp & 3, cq = ew >> 2;
        float* my_bias = bias_w + ew * 32;
        constexpr float kL2e = 1.4426950408889634f;
        const int n_slices = tiles_n * 4;
        uint32_t it = 0;
        long long t_wt = 0, t_ld = 0, t_math = 0, t_start = VAG_TCLK();
        for (int tile = t_begin; tile < t_end; ++tile, ++it) {
            const int tn = tile / tiles_m;
            const int m0 = (tile - tn * tiles_m) * BMP + (int)rank * BM, n0 = tn * BN, c0 = cq * 32;
            const uint32_t a = it & 1;
            const int n_valid = min(32, N - (n0 + c0));   // ≤ 0: the chunk lies outside the matrix (warp-uniform)
            const float bias_l = (bias && lane < n_valid) ? bias[n0 + c0 + lane] : 0.f;   // in flight while the tile is still being accumulated
            long long t0 = VAG_TCLK();
            mbar_wait(&tfull_bar[a], (it >> 1) & 1);
            t_wt += VAG_TCLK() - t0;
            tcgen05_fence_after();
            float x[32];
            t0 = VAG_TCLK();
            if (n_valid > 0) {
                __syncwarp();
                my_bias[lane] = bias_l;
                __syncwarp();
                // two 16-column halves keep the live registers at x[32] + 2 x 16 (the kernel runs 18 warps per SM)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t r[16], q[16];
                    const uint32_t taddr = tmem_base + a * 256 + ((uint32_t)(lg * 32) << 16) + (uint32_t)(c0 + 16 * h);
                    if (SPLIT) VAG_TMEM_LD16(q, taddr + 128u);
                    VAG_TMEM_LD16(r, taddr);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j4 = 0; j4 < 4; ++j4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(my_bias + 16 * h + 4 * j4);
                        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int j = 4 * j4 + u;
                            const float cross = !SPLIT ? 0.f : (F16 ? __uint_as_float(q[j]) * (1.0f / 2048.0f) : __uint_as_float(q[j]));
                            x[16 * h + j] = (__uint_as_float(r[j]) + cross) + bb[u];
                        }
                    }
                }
            }
            tcgen05_fence_before();
            if (lane == 0) mbar_arrive(&tempty_bar[a]);   // this warp's share of the accumulator is in registers
            __syncwarp();
            t_ld += VAG_TCLK() - t0;
            t0 = VAG_TCLK();
            float4 out = make_float4(-INFINITY, 0.f, -INFINITY, __int_as_float((int)0xFFFFFFFFu));
#ifdef VAG_EXP_NOMATH1
            if (n_valid > 0 && lg == 1) { float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += x[j];
                out.x = acc; }
            if (n_valid > 0 && lg != 1) {
#elif defined(VAG_EXP_NOMATH)
            if (n_valid > 0) { float acc = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) acc += x[j];
                out.x = acc; }
            if (false) {
#else
            if (n_valid > 0) {
#endif
                if (n_valid < 32) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) x[j] = j < n_valid ? x[j] : -INFINITY;
                }
                // Best two of the chunk in canonical order (value descending, column ascending), two levels — the flat version (four
                // insertion chains of 8 ALU-pipe instructions per element + three merges, ~300 per chunk) made the ALU pipe the limit of
                // this kernel in bf16 mode (63 % busy at the FP32 mode's tensor-bound pace):
                //   level 1: best (value, column) of each 8-column group            — 3 instructions per element, strict >: lowest column wins
                //   level 2: best of the four group winners                        — the chunk's best
                //   runner-up = best of { the other three group winners, the best of the winner's group WITHOUT the winner } — only the
                //   winner's group is scanned a second time (selected out of the register array with three predicated moves per element)
                float gb[4];
                int gi[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    float bv = x[8 * g];
                    int bi = 8 * g;
#pragma unroll
                    for (int t = 1; t < 8; ++t) {
                        const float xv = x[8 * g + t];
                        const bool gt = xv > bv;
                        bv = fmaxf(bv, xv);
                        bi = gt ? 8 * g + t : bi;
                    }
                    gb[g] = bv;
                    gi[g] = bi;
                }
                float c1v = gb[0];
                int k1v = gi[0];
#pragma unroll
                for (int g = 1; g < 4; ++g) {
                    const bool gt = gb[g] > c1v;
                    c1v = fmaxf(c1v, gb[g]);
                    k1v = gt ? gi[g] : k1v;
                }
                const int wg = k1v >> 3, wt = k1v & 7;                    // the winner's group and its place in it
                float sv = -INFINITY;
                int st_ = 0;
                // the winner's group out of the register array: bit selects with two lane masks (one LOP3 each) — written as a
                // conditional expression the compiler emitted a four-way divergent branch here (ncu: 18 % of the kernel's
                // instructions, 31 % of its stall samples)
                const uint32_t gm1 = (wg & 1) ? 0xFFFFFFFFu : 0u, gm2 = (wg & 2) ? 0xFFFFFFFFu : 0u;
                auto bitsel = [](uint32_t m, float a, float b) { return __uint_as_float((__float_as_uint(a) & m) | (__float_as_uint(b) & ~m)); };
#pragma unroll
                for (int t = 0; t < 8; ++t) {
                    const float xs = bitsel(gm2, bitsel(gm1, x[24 + t], x[16 + t]), bitsel(gm1, x[8 + t], x[t]));
                    const float xv = t == wt ? -INFINITY : xs;
                    const bool gt = xv > sv;
                    sv = fmaxf(sv, xv);
                    st_ = gt ? t : st_;
                }
                float c2v = -INFINITY;
                int k2v = 0xFFFF;                                         // stays "none" when nothing but -inf is left (single-column chunk)
#pragma unroll
                for (int g = 0; g < 4; ++g) {                             // candidates in ascending column order: strict > keeps the lowest column
                    const float cv = g == wg ? sv : gb[g];
                    const int ci = g == wg ? 8 * g + st_ : gi[g];
                    const bool gt = cv > c2v;
                    c2v = fmaxf(c2v, cv);
                    k2v = gt ? ci : k2v;
                }
                float c1[1] = {c1v}, c2[1] = {c2v};
                int k1[1] = {c1v == -INFINITY ? 0xFFFF : k1v}, k2[1] = {k2v};
                const float m2 = c1[0] * kL2e;
                float ps[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    float ev;
                    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ev) : "f"(fmaf(x[j], kL2e, -m2)));   // exp(-inf) = 0 for padded columns
                    ps[j & 3] += ev;
                }
                const int j1 = n0 + c0 + k1[0], j2 = k2[0] == 0xFFFF ? 0xFFFF : n0 + c0 + k2[0];
                out = make_float4(c1[0], (ps[0] + ps[1]) + (ps[2] + ps[3]), c2[0], __int_as_float(j1 | (j2 << 16)));
            }
            const int row = m0 + lg * 32 + lane;
            if (row < rows && n_valid > 0) summ[(int64_t)(tn * 4 + cq) * rows + row] = out;   // slices past ceil(N / 32) do not exist
            t_math += VAG_TCLK() - t0;
        }
        (void)n_slices;
        if (dbg && pair == 0 && ew == 0 && lane == 0) {
            dbg[rank * 16 + 8] = VAG_TCLK() - t_start; dbg[rank * 16 + 9] = t_wt; dbg[rank * 16 + 10] = t_ld; dbg[rank * 16 + 12] = t_math;
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

#include "gru_pair.cuh"

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D tensor [rows, K] (K contiguous, row pitch ld elements), box {128 B, box_rows}, SWIZZLE_128B, zero OOB fill.
static int make_map(CUtensorMap* m, const void* base, int64_t rows, int64_t K, int64_t ld, int box_rows, bool f16, int rowb) {
    const bool bf16 = gemm_mode() == 2;
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VAG_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * (f16 ? 2 : 4)};
    cuuint32_t box[2] = {(cuuint32_t)(rowb / (f16 ? 2 : 4)), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : (f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32), 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d (rows=%lld K=%lld ld=%lld)", (int)r, (long long)rows, (long long)K, (long long)ld);
        return VAG_ERR_CUDA;
    }
    return VAG_OK;
}

// fp32 output [rows, N] (pitch ldy), box {32, 32}, SWIZZLE_128B: the epilogue's staging tiles are stored with TMA
static int make_out_map(CUtensorMap* m, float* y, int64_t rows, int64_t N, int64_t ldy) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VAG_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ldy * sizeof(float)};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)y, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (output) failed with %d (rows=%lld N=%lld ldy=%lld)", (int)r, (long long)rows, (long long)N, (long long)ldy);
        return VAG_ERR_CUDA;
    }
    return VAG_OK;
}

// 16-bit operand plane [rows, N] (pitch ld elements) written by the split-output epilogue: box {32, 32}, SWIZZLE_64B
static int make_plane_out_map(CUtensorMap* m, void* base, int64_t rows, int64_t N, int64_t ld, bool bf16) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return VAG_ERR_CUDA;
    }
    cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {32u, 32u};
    cuuint32_t estr[2] = {1u, 1u};
    CUresult r = fn(m, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (operand-plane output) failed with %d (rows=%lld N=%lld ld=%lld)", (int)r, (long long)rows, (long long)N, (long long)ld);
        return VAG_ERR_CUDA;
    }
    return VAG_OK;
}

size_t linear_tc_scratch_bytes(int64_t rows, int64_t K, int64_t N) {
    return (size_t)(2 * rows * K + 2 * N * K) * sizeof(float) + 4 * 256;
}

bool linear_tc_eligible(const float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, int rows, int K, int N) {
    auto al = [](const void* p) { return ((uintptr_t)p & 15) == 0; };
    (void)y; (void)ldy;
    return rows >= 64 && N >= 64 && K >= 32 && (K % 8 == 0) && (ldx % 4 == 0) && (ldw % 4 == 0) && al(x) && al(w);
}

template <int BN, bool F16, int ROWB>
static int launch_tc(const CUtensorMap& xh, const CUtensorMap& xl, const CUtensorMap& wh, const CUtensorMap& wl, float* y,
                     int64_t ldy, const float* bias, int rows, int K, int N, int flags, float4* summ, cudaStream_t st) {
    using Cfg = TcCfg<BN, F16, ROWB>;
    static bool attr_set = false;
    if (!attr_set) {
        VAG_CUDA(cudaFuncSetAttribute(linear_split3_kernel<BN, F16, ROWB>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_set = true;
    }
    dim3 grid(ceil_div(N, BN), ceil_div(rows, Cfg::BM));
    linear_split3_kernel<BN, F16, ROWB><<<grid, 192, Cfg::SMEM_BYTES, st>>>(xh, xl, wh, wl, y, ldy, bias, rows, K, N, flags, summ);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// VAG_GEMM=tf32x3 selects the TF32 split (any FP32 range); default is the FP16 split (2x the tensor rate).
// 0 = TF32 split, 1 = FP16 split (default), 2 = BF16 single product.  Set per C-ABI call by ModeScope (common.cuh).
static thread_local int g_mode_override = -1;
static long long* g_tc_dbg = nullptr;   // optional device buffer [32] for the pair kernel's role timers (tools only)
static thread_local const int* g_tc_done = nullptr;   // TcDoneScope (gemm_ctx.cuh)
TcDoneScope::TcDoneScope(const int* done) : prev(g_tc_done) { g_tc_done = done; }
TcDoneScope::~TcDoneScope() { g_tc_done = prev; }
void set_tc_debug(long long* p) { g_tc_dbg = p; }
long long* tc_debug() { return g_tc_dbg; }
ModeScope::ModeScope(int precision) : prev(g_mode_override) { g_mode_override = precision == VAG_PREC_BF16 ? 2 : -1; }
ModeScope::~ModeScope() { g_mode_override = prev; }
int gemm_mode() {
    if (g_mode_override >= 0) return g_mode_override;
    const char* e = getenv("VAG_GEMM");
    if (e && strcmp(e, "tf32x3") == 0) return 0;
    if (e && strcmp(e, "bf16") == 0) return 2;
    return 1;
}
static bool use_f16_split() { return gemm_mode() != 0; }   // 16-bit planes

int tc_elem_bytes() { return use_f16_split() ? 2 : 4; }

// Can tc_gemm take this call in the current mode?  (bf16 mode only has the persistent TMA-store kernel.)
bool tc_call_supported(const float* y, int64_t ldy, int flags) {
    if (gemm_mode() != 2) return true;
    return ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && !(flags & VAG_LIN_ACCUMULATE);
}

// Split x [rows, K] (pitch ldx) into hi / lo planes of pitch ld_out ELEMENTS, starting at column col_off
// (so several sources can be laid side by side along K for a concatenated contraction).
int tc_split(const float* x, int64_t ldx, int rows, int K, void* hi, void* lo, int64_t ld_out, int64_t col_off, cudaStream_t st) {
    const int64_t tot = (int64_t)rows * (K / 4);
    if (tot == 0) return VAG_OK;
    const int g = (int)std::min<int64_t>(ceil_div64(tot, 256), (int64_t)num_sms() * 8);
    if (gemm_mode() == 2)
        VAG_CUDA(launch_pdl(PDL_SMALL, round_bf16_kernel, dim3(g), dim3(256), 0, st, x, ldx, rows, K, (__nv_bfloat16*)hi + col_off, ld_out));
    else if (use_f16_split())
        split_f16_kernel<<<g, 256, 0, st>>>(x, ldx, rows, K, (__half*)hi + col_off, (__half*)lo + col_off, ld_out);
    else
        split_tf32_kernel<<<g, 256, 0, st>>>(x, ldx, rows, K, (float*)hi + col_off, (float*)lo + col_off, ld_out);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// Both operands of one contraction split in ONE launch (blockIdx.z = operand): out(r, c) = split(src(r, c)) for c < cols,
// zero for cols <= c < cols_pad, where src(r, c) = src[c·ld + r] when the operand is stored contraction-major (transposed)
// and src[r·ld + c] otherwise.  Inside a captured graph every node costs ~2 µs of dependency latency whatever it does, so the
// backward contractions use this instead of two split launches (+ two memsets for the K padding) per contraction.
struct SplitJob {
    const float* src;
    int64_t ld;
    int rows, cols, cols_pad, transposed;
    void* hi;
    void* lo;
    int64_t ldo;
};
struct SplitJobs {
    SplitJob j[2];
};
template <int MODE>   // 0: TF32 planes (float), 1: FP16 hi/lo, 2: BF16 single plane
__global__ void __launch_bounds__(256) split_jobs_kernel(const __grid_constant__ SplitJobs jobs) {
    pdl_trigger();
    pdl_wait();
    const SplitJob& jb = jobs.j[blockIdx.z];
    const int tiles_c = (jb.cols_pad + 31) >> 5, tiles_r = (jb.rows + 31) >> 5;
    if ((int)blockIdx.x >= tiles_c * tiles_r) return;
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int r0 = ((int)blockIdx.x / tiles_c) << 5, c0 = ((int)blockIdx.x % tiles_c) << 5;
    if (jb.transposed) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + ty + 8 * j, r = r0 + tx;
            tile[ty + 8 * j][tx] = (c < jb.cols && r < jb.rows) ? jb.src[(int64_t)c * jb.ld + r] : 0.f;
        }
        __syncthreads();
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int r = r0 + ty + 8 * j, c = c0 + tx;
        if (r >= jb.rows || c >= jb.cols_pad) continue;
        const float v = jb.transposed ? tile[tx][ty + 8 * j] : (c < jb.cols ? jb.src[(int64_t)r * jb.ld + c] : 0.f);
        const int64_t o = (int64_t)r * jb.ldo + c;
        if (MODE == 0) {
            uint32_t t;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
            const float h = __uint_as_float(t);
            ((float*)jb.hi)[o] = h;
            ((float*)jb.lo)[o] = v - h;
        } else if (MODE == 1) {
            const __half h = __float2half_rn(v);
            ((__half*)jb.hi)[o] = h;
            ((__half*)jb.lo)[o] = __float2half_rn((v - __half2float(h)) * 2048.0f);
        } else {
            ((__nv_bfloat16*)jb.hi)[o] = __float2bfloat16_rn(v);
        }
    }
}

// x: [M, K] logical (stored [K, M] when xt), w: [N, K] logical (stored [K, N] when wt); planes of pitch Kp (zero-padded).
int tc_split_pair(const float* x, int64_t ldx, bool xt, int M, void* xh, void* xl, const float* w, int64_t ldw, bool wt, int N,
                  void* wh, void* wl, int K, int Kp, cudaStream_t st) {
    SplitJobs jobs;
    jobs.j[0] = SplitJob{x, ldx, M, K, Kp, xt ? 1 : 0, xh, xl, (int64_t)Kp};
    jobs.j[1] = SplitJob{w, ldw, N, K, Kp, wt ? 1 : 0, wh, wl, (int64_t)Kp};
    const int tiles = ceil_div(Kp, 32) * ceil_div(M > N ? M : N, 32);
    dim3 grid(tiles, 1, 2);
    if (gemm_mode() == 2) VAG_CUDA(launch_pdl(PDL_SMALL, split_jobs_kernel<2>, dim3(grid), dim3(256), 0, st, jobs));
    else if (use_f16_split()) VAG_CUDA(launch_pdl(PDL_SMALL, split_jobs_kernel<1>, dim3(grid), dim3(256), 0, st, jobs));
    else VAG_CUDA(launch_pdl(PDL_SMALL, split_jobs_kernel<0>, dim3(grid), dim3(256), 0, st, jobs));
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// The tcgen05 contraction on pre-split operands: xh/xl [rows, K] pitch ldxs, wh/wl [N, K] pitch ldws (elements).
// summ (optional): [rows, ceil(N / tile_w)] float4 (max, Σexp(x-max), best value, best column as int bits) per
// (row, column tile); *summ_tile_w receives the tile width the chosen configuration uses.
// Split-output flavour (fused decode step): y = act(x·Wᵀ + bias) leaves the kernel as operand planes `out` (no fp32
// copy).  CTA-pair kernel only: rows > 128, 16-bit modes, out.ld % 8 == 0, 16-byte aligned planes.
int tc_gemm_split_out(SplitDst out, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
                      const float* bias, int rows, int K, int N, int flags, cudaStream_t st) {
    const int mode = gemm_mode();
    if (mode == 0 || rows <= 128 || !out.hi || (out.ld & 7) || (reinterpret_cast<uintptr_t>(out.hi) & 15) ||
        (mode == 1 && (!out.lo || (reinterpret_cast<uintptr_t>(out.lo) & 15)))) {
        set_error("tc_gemm_split_out: needs a 16-bit mode, rows > 128 and 16-byte aligned output planes");
        return VAG_ERR_UNSUPPORTED;
    }
    CUtensorMap mxh, mxl, mwh, mwl, myh, myl;
    VAG_TRY(make_map(&mxh, xh, rows, K, ldxs, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mxl, xl, rows, K, ldxs, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mwh, wh, N, K, ldws, 64, true, Q_ROWB));
    VAG_TRY(make_map(&mwl, wl, N, K, ldws, 64, true, Q_ROWB));
    VAG_TRY(make_plane_out_map(&myh, out.hi, rows, N, out.ld, mode == 2));
    VAG_TRY(make_plane_out_map(&myl, mode == 2 ? out.hi : out.lo, rows, N, out.ld, mode == 2));
    const int n_tiles = ceil_div(N, 128) * ceil_div(rows, 256);
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
    static bool attr_set[3] = {false, false, false};
    if (mode == 1) {
        if (!attr_set[1]) { VAG_CUDA(cudaFuncSetAttribute(linear_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES)); attr_set[1] = true; }
        VAG_CUDA(launch_pdl(PDL_TC, linear_pair_kernel<1>, dim3(grid), dim3(320), Q_SMEM_BYTES, st, mxh, mxl, mwh, mwl, myh, myh, myl, bias, rows, K, N, flags | VAG_LIN_SPLIT_OUT, (float4*)nullptr, g_tc_dbg, g_tc_done));
    } else {
        if (!attr_set[2]) { VAG_CUDA(cudaFuncSetAttribute(linear_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES)); attr_set[2] = true; }
        VAG_CUDA(launch_pdl(PDL_TC, linear_pair_kernel<2>, dim3(grid), dim3(320), Q_SMEM_BYTES, st, mxh, mxl, mwh, mwl, myh, myh, myl, bias, rows, K, N, flags | VAG_LIN_SPLIT_OUT, (float4*)nullptr, g_tc_dbg, g_tc_done));
    }
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// Summary-only vocabulary projection (fused decode step): NO output matrix.  summ [ceil(N / 32), rows] float4 =
// (best value, Σexp(x − best), second-best value, best column | second column << 16 — 0xFFFF: none) per (32-column slice,
// row), canonical order (value descending, column ascending).  CTA-pair kernel: rows > 128, 16-bit modes, N < 65535.
int tc_gemm_top2(float4* summ, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
                 const float* bias, int rows, int K, int N, cudaStream_t st) {
    const int mode = gemm_mode();
    if (mode == 0 || rows <= 128 || !summ || N >= 0xFFFF) {
        set_error("tc_gemm_top2: needs a 16-bit mode, rows > 128 and fewer than 65535 columns");
        return VAG_ERR_UNSUPPORTED;
    }
    CUtensorMap mxh, mxl, mwh, mwl;
    VAG_TRY(make_map(&mxh, xh, rows, K, ldxs, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mxl, xl, rows, K, ldxs, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mwh, wh, N, K, ldws, 64, true, Q_ROWB));
    VAG_TRY(make_map(&mwl, wl, N, K, ldws, 64, true, Q_ROWB));
    const int n_tiles = ceil_div(N, 128) * ceil_div(rows, 256);
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
    static bool attr_set[3] = {false, false, false};
    if (mode == 1) {
        if (!attr_set[1]) { VAG_CUDA(cudaFuncSetAttribute(vocab_top2_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES)); attr_set[1] = true; }
        VAG_CUDA(launch_pdl(PDL_TC, vocab_top2_pair_kernel<1>, dim3(grid), dim3(576), V_SMEM_BYTES, st, mxh, mxl, mwh, mwl, bias, rows, K, N, summ, g_tc_dbg, g_tc_done));
    } else {
        if (!attr_set[2]) { VAG_CUDA(cudaFuncSetAttribute(vocab_top2_pair_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, V_SMEM_BYTES)); attr_set[2] = true; }
        VAG_CUDA(launch_pdl(PDL_TC, vocab_top2_pair_kernel<2>, dim3(grid), dim3(576), V_SMEM_BYTES, st, mxh, mxl, mwh, mwl, bias, rows, K, N, summ, g_tc_dbg, g_tc_done));
    }
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

// GRU cell fused into its contractions (gru_pair.cuh).  16-bit modes, rows > 32 (up to 32 rows the latency-shaped linear_rows32
// kernels are faster; between 33 and 128 rows the peer CTA of a pair works on zero-filled rows, still one launch instead of
// three), H % 32 == 0, K % 8 == 0.
bool tc_gru_supported(int rows, int H, int Kx, int Kh) {
    const int mode = gemm_mode();
    return (mode == 1 || mode == 2) && rows > 32 && H % 32 == 0 && H <= G_BIAS_MAX_H && Kh >= 32 && Kh % 8 == 0 && (Kx == 0 || (Kx >= 32 && Kx % 8 == 0));
}
int tc_gru(const GruCall& c, cudaStream_t st) {
    const int mode = gemm_mode();
    if (!tc_gru_supported(c.rows, c.H, c.Kx, c.Kh) || !c.out.hi || (c.out.ld & 7) || (reinterpret_cast<uintptr_t>(c.out.hi) & 15) ||
        (mode == 1 && (!c.out.lo || (reinterpret_cast<uintptr_t>(c.out.lo) & 15))) || (c.Kx == 0) != (c.g1 != nullptr) ||
        (c.y2 && ((reinterpret_cast<uintptr_t>(c.y2) & 15) || (c.ld_y2 & 3)))) {
        set_error("tc_gru: unsupported shape / mode / output planes");
        return VAG_ERR_UNSUPPORTED;
    }
    CUtensorMap mxh, mxl, mhh, mhl, mwih, mwil, mwhh, mwhl;
    VAG_TRY(make_map(&mhh, c.hh, c.rows, c.Kh, c.ldh, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mhl, mode == 2 ? c.hh : c.hl, c.rows, c.Kh, c.ldh, 128, true, Q_ROWB));
    VAG_TRY(make_map(&mwhh, c.whh_h, 3 * c.H, c.Kh, c.Kh, 48, true, Q_ROWB));
    VAG_TRY(make_map(&mwhl, mode == 2 ? c.whh_h : c.whh_l, 3 * c.H, c.Kh, c.Kh, 48, true, Q_ROWB));
    if (c.Kx) {
        VAG_TRY(make_map(&mxh, c.xh, c.rows, c.Kx, c.ldx, 128, true, Q_ROWB));
        VAG_TRY(make_map(&mxl, mode == 2 ? c.xh : c.xl, c.rows, c.Kx, c.ldx, 128, true, Q_ROWB));
        VAG_TRY(make_map(&mwih, c.wih_h, 3 * c.H, c.Kx, c.Kx, 48, true, Q_ROWB));
        VAG_TRY(make_map(&mwil, mode == 2 ? c.wih_h : c.wih_l, 3 * c.H, c.Kx, c.Kx, 48, true, Q_ROWB));
    } else {
        mxh = mhh; mxl = mhl; mwih = mwhh; mwil = mwhl;   // never dereferenced (no phase X)
    }
    GruArgs a;
    a.bias = c.bias4; a.g1 = c.g1; a.tokens = c.tokens; a.V = c.V; a.h_prev = c.h_prev; a.h_out = c.h_out;
    a.out_hi = c.out.hi; a.out_lo = c.out.lo; a.out_ld = c.out.ld; a.rows = c.rows; a.Kx = c.Kx; a.Kh = c.Kh; a.H = c.H; a.done = g_tc_done; a.y2 = c.y2; a.ld_y2 = c.ld_y2;
    static const bool wide_off = getenv("VAG_GRU_WIDE") && getenv("VAG_GRU_WIDE")[0] == '0';   // A/B runs: 128-bit epilogue accesses
    const uintptr_t al = reinterpret_cast<uintptr_t>(c.h_prev) | reinterpret_cast<uintptr_t>(c.h_out) | reinterpret_cast<uintptr_t>(c.g1) |
                         reinterpret_cast<uintptr_t>(c.out.hi) | (mode == 1 ? reinterpret_cast<uintptr_t>(c.out.lo) : 0) |
                         reinterpret_cast<uintptr_t>(c.y2);
    a.wide = (!wide_off && (al & 31) == 0 && (c.out.ld & 15) == 0 && (c.ld_y2 & 7) == 0) ? 1 : 0;
    const int n_tiles = (c.H / 32) * ceil_div(c.rows, 256);
    const int max_pairs = num_sms() / 2;
    const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
    // 8 epilogue warps (two passes of 8 units each) is the default; VAG_GRU_EW=16 selects sixteen single-pass warps — measured
    // SLOWER (1000 sentences: 52.9 vs 48.4 ms): 96 registers per thread at 576 threads spill the accumulator fragments
    static const bool ew8 = !(getenv("VAG_GRU_EW") && atoi(getenv("VAG_GRU_EW")) == 16);
#define VAG_GRU_LAUNCH(MODE_, NEW_)                                                                                            \
    do {                                                                                                                       \
        static bool attr_set = false;                                                                                          \
        if (!attr_set) { VAG_CUDA(cudaFuncSetAttribute(gru_pair_kernel<MODE_, NEW_>, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES)); attr_set = true; } \
        VAG_CUDA(launch_pdl(PDL_TC, gru_pair_kernel<MODE_, NEW_>, dim3(grid), dim3(64 + 32 * NEW_), G_SMEM_BYTES, st, mxh, mxl, mhh, mhl, mwih, mwil, mwhh, mwhl, a)); \
    } while (0)
    if (mode == 1) { if (ew8) VAG_GRU_LAUNCH(1, 8); else VAG_GRU_LAUNCH(1, 16); }
    else { if (ew8) VAG_GRU_LAUNCH(2, 8); else VAG_GRU_LAUNCH(2, 16); }
#undef VAG_GRU_LAUNCH
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
// Permuted operand planes of one GRU matrix w [3H, K] (hidden: the hidden-state matrix' tile order) and the unit-major biases.
int tc_gru_prepare_weight(void* hi, void* lo, const float* w, int H, int K, bool hidden, float* scratch, cudaStream_t st) {
    if (H % 32) { set_error("tc_gru_prepare_weight: H must be a multiple of 32"); return VAG_ERR_UNSUPPORTED; }
    gru_permute_rows_kernel<<<3 * H, 256, 0, st>>>(scratch, w, H, K, hidden ? 1 : 0);
    VAG_LAUNCH_CHECK();
    return tc_split(scratch, K, 3 * H, K, hi, lo, K, 0, st);
}
int tc_gru_prepare_bias(float* out4h, const float* b_ih, const float* b_hh, int H, bool with_ih, cudaStream_t st) {
    gru_bias_kernel<<<ceil_div(H, 128), 128, 0, st>>>(out4h, b_ih, b_hh, H, with_ih ? 1 : 0);
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}

int tc_gemm(float* y, int64_t ldy, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
            const float* bias, int rows, int K, int N, int flags, cudaStream_t st, float4* summ, int* summ_tile_w) {
    const bool f16 = use_f16_split();
    const int sms = num_sms();
    const char* e = getenv("VAG_TC_CFG");
    // Persistent kernel (default): needs a TMA-storable output (16-byte aligned rows) and no read-modify-write epilogue.
    const bool y_tma = ((ldy & 3) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && !(flags & VAG_LIN_ACCUMULATE);
    if (y_tma && rows > 128 && (!e || strcmp(e, "pair") == 0)) {
        // CTA-pair kernel: 256 x 128 tiles, half of the weight tile per CTA
        CUtensorMap mxh, mxl, mwh, mwl, my;
        VAG_TRY(make_map(&mxh, xh, rows, K, ldxs, 128, f16, Q_ROWB));
        VAG_TRY(make_map(&mxl, xl, rows, K, ldxs, 128, f16, Q_ROWB));
        VAG_TRY(make_map(&mwh, wh, N, K, ldws, 64, f16, Q_ROWB));
        VAG_TRY(make_map(&mwl, wl, N, K, ldws, 64, f16, Q_ROWB));
        VAG_TRY(make_out_map(&my, y, rows, N, ldy));
        if (summ_tile_w) *summ_tile_w = 128;
        static bool attr_set[3] = {false, false, false};
        const int n_tiles = ceil_div(N, 128) * ceil_div(rows, 256);
        const int max_pairs = sms / 2;
        const int grid = 2 * (n_tiles < max_pairs ? n_tiles : max_pairs);
        const int mode = gemm_mode();
#define VAG_PAIR(M)                                                                                                             \
    do {                                                                                                                        \
        if (!attr_set[M]) {                                                                                                     \
            VAG_CUDA(cudaFuncSetAttribute(linear_pair_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM_BYTES));    \
            attr_set[M] = true;                                                                                                 \
        }                                                                                                                       \
        VAG_CUDA(launch_pdl(PDL_TC, linear_pair_kernel<M>, dim3(grid), dim3(320), Q_SMEM_BYTES, st, mxh, mxl, mwh, mwl, my, my, my, bias, rows, K, N, flags & ~VAG_LIN_SPLIT_OUT, summ, g_tc_dbg, g_tc_done));          \
    } while (0)
        if (mode == 0) VAG_PAIR(0);
        else if (mode == 1) VAG_PAIR(1);
        else VAG_PAIR(2);
#undef VAG_PAIR
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    if (y_tma && (!e || strcmp(e, "persistent") == 0 || strcmp(e, "pair") == 0)) {
        CUtensorMap mxh, mxl, mwh, mwl, my;
        VAG_TRY(make_map(&mxh, xh, rows, K, ldxs, 128, f16, 64));
        VAG_TRY(make_map(&mxl, xl, rows, K, ldxs, 128, f16, 64));
        VAG_TRY(make_map(&mwh, wh, N, K, ldws, 128, f16, 64));
        VAG_TRY(make_map(&mwl, wl, N, K, ldws, 128, f16, 64));
        VAG_TRY(make_out_map(&my, y, rows, N, ldy));
        if (summ_tile_w) *summ_tile_w = 128;
        static bool attr_set[3] = {false, false, false};
        const int n_tiles = ceil_div(N, 128) * ceil_div(rows, 128);
        const int grid = n_tiles < sms ? n_tiles : sms;
        const int mode = gemm_mode();
#define VAG_PERSIST(M)                                                                                                          \
    do {                                                                                                                        \
        if (!attr_set[M]) {                                                                                                     \
            VAG_CUDA(cudaFuncSetAttribute(linear_split3_persistent_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES)); \
            attr_set[M] = true;                                                                                                 \
        }                                                                                                                       \
        linear_split3_persistent_kernel<M><<<grid, 320, P_SMEM_BYTES, st>>>(mxh, mxl, mwh, mwl, my, bias, rows, K, N, flags, summ); \
    } while (0)
        if (mode == 0) VAG_PERSIST(0);
        else if (mode == 1) VAG_PERSIST(1);
        else VAG_PERSIST(2);
#undef VAG_PERSIST
        VAG_LAUNCH_CHECK();
        return VAG_OK;
    }
    if (gemm_mode() == 2) {
        set_error("tc_gemm: the bf16 mode needs a TMA-storable output and no accumulate flag");
        return VAG_ERR_UNSUPPORTED;
    }
    // tile configuration of the one-tile-per-CTA kernel: VAG_TC_CFG=wide|narrow|dual (tuning / A-B runs)
    int cfg;  // 0 = 128x256 wide, 1 = 128x128 (128 B rows), 2 = 128x128 dual-CTA (64 B rows)
    if (e && strcmp(e, "wide") == 0) cfg = 0;
    else if (e && strcmp(e, "narrow") == 0) cfg = 1;
    else if (e && strcmp(e, "dual") == 0) cfg = 2;
    else {
        cfg = 2;  // measured on B200 (tools/tcbench.py): the dual-CTA tile wins or ties on every decode-step shape
        (void)sms;
    }
    const int bn = cfg == 0 ? 256 : 128;
    const int rowb = cfg == 2 ? 64 : 128;
    if (summ_tile_w) *summ_tile_w = bn;
    CUtensorMap mxh, mxl, mwh, mwl;
    VAG_TRY(make_map(&mxh, xh, rows, K, ldxs, 128, f16, rowb));
    VAG_TRY(make_map(&mxl, xl, rows, K, ldxs, 128, f16, rowb));
    VAG_TRY(make_map(&mwh, wh, N, K, ldws, bn, f16, rowb));
    VAG_TRY(make_map(&mwl, wl, N, K, ldws, bn, f16, rowb));
#define VAG_TC_GO(BN_, F16_, ROWB_) return launch_tc<BN_, F16_, ROWB_>(mxh, mxl, mwh, mwl, y, ldy, bias, rows, K, N, flags, summ, st)
    if (f16) {
        if (cfg == 0) VAG_TC_GO(256, true, 128);
        if (cfg == 1) VAG_TC_GO(128, true, 128);
        VAG_TC_GO(128, true, 64);
    }
    if (cfg == 0) VAG_TC_GO(256, false, 128);
    if (cfg == 1) VAG_TC_GO(128, false, 128);
    VAG_TC_GO(128, false, 64);
#undef VAG_TC_GO
}

// scratch: ≥ linear_tc_scratch_bytes(rows, K, N).  Splits both operands, then runs the tcgen05 kernel.
int linear_tc(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
              int K, int N, int flags, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (scratch_bytes < linear_tc_scratch_bytes(rows, K, N) || !scratch) {
        set_error("linear_tc: scratch %zu B too small", scratch_bytes);
        return VAG_ERR_WORKSPACE;
    }
    Arena ar(scratch, scratch_bytes);
    const size_t esz = (size_t)tc_elem_bytes();
    void* xh = ar.take<char>((size_t)rows * K * esz);
    void* xl = ar.take<char>((size_t)rows * K * esz);
    void* wh = ar.take<char>((size_t)N * K * esz);
    void* wl = ar.take<char>((size_t)N * K * esz);
    if (ar.overflow) {
        set_error("linear_tc: scratch overflow");
        return VAG_ERR_WORKSPACE;
    }
    VAG_TRY(tc_split(x, ldx, rows, K, xh, xl, K, 0, st));
    VAG_TRY(tc_split(w, ldw, N, K, wh, wl, K, 0, st));
    return tc_gemm(y, ldy, xh, xl, K, wh, wl, K, bias, rows, K, N, flags, st, nullptr, nullptr);
}

}  // namespace vag
