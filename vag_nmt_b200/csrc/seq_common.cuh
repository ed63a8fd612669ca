// Building blocks of the persistent time-loop kernels (enc_seq.cu, dec_seq.cu): flag barrier between co-resident CTAs, tensor-core
// fragments for the <= 32-row products (mma.sync m16n8k8 TF32).
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace vag {
namespace {

constexpr int ES_MAX_CTAS = 160;      // flags per barrier group

__device__ __forceinline__ float es_rbf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
__device__ __forceinline__ uint32_t es_tf32(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ void es_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ int es_ld_relaxed(const int* p) {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Warp 0 waits until every CTA of the direction has published step counter >= target (bounded); everybody leaves through the CTA
// barrier.  The polls are relaxed loads (one L2 round trip each); the acquire fence after the last one orders the peers' stores
// (released with their flags) before this CTA's loads of the exchange buffer.
__device__ __forceinline__ void es_wait(const int* flags, int n_cta, int target, int* gave_up) {
    if (threadIdx.x < 32 && !*gave_up) {
        int it = 0;
        for (;;) {
            bool ok = true;
            for (int c = threadIdx.x; c < n_cta; c += 32) ok &= es_ld_relaxed(flags + c) >= target;
            if (__all_sync(0xffffffffu, ok)) break;
            if (++it > (1 << 21)) {
                *gave_up = 1;
                break;
            }
        }
        asm volatile("fence.acquire.gpu;" ::: "memory");
    }
    __syncthreads();
}
__device__ __forceinline__ void es_arrive(int* flag, int value) {
    __syncthreads();                                // every thread's stores of this step precede thread 0's release
    if (threadIdx.x == 0) asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

// One k-step (8 contraction indices) of the 32-row product: A fragments of both 16-row tiles from a [32][pitch] tile.
template <bool RB>
__device__ __forceinline__ void es_load_a(uint32_t (&ah)[2][4], uint32_t (&al)[2][4], const float* xs, int pitch, int kk, int g) {
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const float* xa = xs + (mt * 16 + g) * pitch + kk;
        const float af[4] = {xa[0], xa[8 * pitch], xa[4], xa[8 * pitch + 4]};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            ah[mt][u] = RB ? __float_as_uint(af[u]) : es_tf32(af[u]);
            al[mt][u] = RB ? 0u : es_tf32(af[u] - __uint_as_float(ah[mt][u]));
        }
    }
}

}  // namespace
}  // namespace vag
