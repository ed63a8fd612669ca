// Fused operand split: producers of an activation (GRU gates, attention context, beam reorder, embedding gather,
// contraction epilogues) write the tensor-core operand planes themselves instead of leaving that to a separate
// split_f16_kernel launch.  Bit-identical to split_f16_kernel / round_bf16_kernel of linear_tc.cu:
//   mode 1 (FP32-exact): hi = rn_f16(v), lo = rn_f16((v - hi) * 2^11)        mode 2 (bf16): hi = rn_bf16(v), no lo plane
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace vag {

// internal epilogue flag (never part of the C ABI): write the output as tensor-core operand planes instead of fp32
constexpr int VAG_LIN_SPLIT_OUT = 0x100;
// internal: no output matrix, only the per-(row, tile) top-2 / Σexp summaries (vocabulary projection of the beam loop)
constexpr int VAG_LIN_TOP2 = 0x200;

struct SplitDst {
    uint16_t* hi = nullptr;   // nullptr: disabled
    uint16_t* lo = nullptr;
    int64_t ld = 0;           // row pitch in ELEMENTS (multiple of 8)
    int mode = 1;
};

__device__ __forceinline__ void split_one(int mode, float v, uint16_t& h, uint16_t& l) {
    if (mode == 2) {
        h = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        l = 0;
    } else {
        const __half hh = __float2half_rn(v);
        h = __half_as_ushort(hh);
        l = __half_as_ushort(__float2half_rn((v - __half2float(hh)) * 2048.0f));
    }
}

// four consecutive elements of one row (col a multiple of 4)
__device__ __forceinline__ void split_store4(const SplitDst& d, int64_t row, int col, const float4& v) {
    uint16_t h[4], l[4];
    split_one(d.mode, v.x, h[0], l[0]);
    split_one(d.mode, v.y, h[1], l[1]);
    split_one(d.mode, v.z, h[2], l[2]);
    split_one(d.mode, v.w, h[3], l[3]);
    const int64_t off = row * d.ld + col;
    *reinterpret_cast<uint2*>(d.hi + off) = make_uint2((uint32_t)h[0] | ((uint32_t)h[1] << 16), (uint32_t)h[2] | ((uint32_t)h[3] << 16));
    if (d.mode != 2)
        *reinterpret_cast<uint2*>(d.lo + off) = make_uint2((uint32_t)l[0] | ((uint32_t)l[1] << 16), (uint32_t)l[2] | ((uint32_t)l[3] << 16));
}

}  // namespace vag
