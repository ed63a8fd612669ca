// Shared helpers for libvagnmt.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/vag_nmt.h"

namespace vag {

void set_error(const char* fmt, ...);

#define VAG_REQUIRE(cond, ...)                 \
    do {                                       \
        if (!(cond)) {                         \
            ::vag::set_error(__VA_ARGS__);     \
            return VAG_ERR_INVALID;            \
        }                                      \
    } while (0)

#define VAG_CUDA(call)                                                                         \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            ::vag::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return VAG_ERR_CUDA;                                                               \
        }                                                                                      \
    } while (0)

// every kernel launch of the library goes through this macro; the counter backs vag_launch_count()
void count_launch();
#define VAG_LAUNCH_CHECK()               \
    do {                                 \
        ::vag::count_launch();           \
        VAG_CUDA(cudaGetLastError());    \
    } while (0)

#define VAG_TRY(call)              \
    do {                           \
        int s__ = (call);          \
        if (s__ != VAG_OK) return s__; \
    } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
    char* base;
    size_t size;
    size_t off;
    bool overflow;
    Arena(void* p, size_t n) : base((char*)p), size(n), off(0), overflow(false) {}
    template <typename T>
    T* take(size_t count) {
        size_t start = align_up(off, 256);
        size_t end = start + count * sizeof(T);
        if (end > size || base == nullptr) {
            overflow = true;
            off = end;
            return nullptr;
        }
        off = end;
        return (T*)(base + start);
    }
};
// Same arithmetic without memory: used by the *_workspace_bytes() functions.
struct ArenaSizer {
    size_t off = 0;
    template <typename T>
    void take(size_t count) {
        off = align_up(off, 256) + count * sizeof(T);
    }
    size_t total() const { return align_up(off, 256) + 256; }
};

int num_sms();

// Arithmetic mode of the contractions issued on behalf of the C-ABI call that is executing on this thread:
//   0 = TF32 hi/lo split (VAG_GEMM=tf32x3), 1 = FP16 hi/lo split (default of VAG_PREC_FP32), 2 = bf16 (VAG_PREC_BF16).
// Every extern "C" entry point that contracts opens a ModeScope from ITS OWN precision argument / struct field / flag and
// closes it on return: the mode is per call, the thread-local below is only how it reaches the kernels' launchers.
int gemm_mode();
struct ModeScope {
    int prev;
    explicit ModeScope(int precision);
    ~ModeScope();
    ModeScope(const ModeScope&) = delete;
    ModeScope& operator=(const ModeScope&) = delete;
};

// Programmatic dependent launch: a kernel launched with launch_pdl() may become resident while its predecessor on the stream
// is still draining; it must execute pdl_wait() before it touches global memory.  pdl_trigger() in the predecessor lets the
// dependent launch as soon as every CTA of the predecessor has started.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Kernel families for A/B runs: VAG_PDL_MASK is a bit mask of the families that launch programmatically.  Default TC | SMALL,
// from the graph-replayed training step (batch 32, bf16): none 2.144 ms, TC 2.127, TC+SMALL 2.052, TC+SMALL+ATTN 2.066,
// TC+SMALL+ROWS32 2.151 (early or late trigger alike: the recurrent 32-row kernels lose ~1 us per launch when launched
// programmatically, so they stay ordinary launches).
enum { PDL_TC = 1, PDL_ROWS32 = 2, PDL_ATTN = 4, PDL_SMALL = 8, PDL_BEAM = 16 /* selection + reorder kernels of the decode loop */ };
bool pdl_enabled(int family);
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(int family, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled(family) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float sigmoidf_precise(float x) { return 1.0f / (1.0f + expf(-x)); }
// exp(2x) = 2^(x·2·log2e) with the rounding of the product compensated: u = rn(x·c_hi), the exact residual of that product plus
// x·c_lo goes into a first-order correction, so the result carries only ex2.approx's own error (≤ 2 ulp) instead of an argument
// error that grows with |x| — six instructions against ~35 for the range-checked expf.  Overflows to +inf / underflows to 0 cleanly.
__device__ __forceinline__ float exp2x_comp(float x) {
    constexpr float c_hi = 2.88539004f;                 // rn(2·log2(e))
    constexpr float c_lo = 4.05197e-08f;                // 2·log2(e) − c_hi
    const float u = x * c_hi;
    const float err = fmaf(x, c_lo, fmaf(x, c_hi, -u));
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(u));
    return fmaf(e, err * 0.693147182f, e);
}
__device__ __forceinline__ float rcp_approx(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

}  // namespace vag
