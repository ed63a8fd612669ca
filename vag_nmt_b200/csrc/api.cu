// Library-level entry points: error string, ABI version, device probe, linear dispatch.
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include <atomic>

namespace vag {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int num_sms() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int linear_simt(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                int rows, int K, int N, int flags, cudaStream_t st);
int linear_tc(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias, int rows,
              int K, int N, int flags, void* scratch, size_t scratch_bytes, cudaStream_t st);
size_t linear_tc_scratch_bytes(int64_t rows, int64_t K, int64_t N);
bool linear_tc_eligible(const float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, int rows, int K, int N);
bool tc_call_supported(const float* y, int64_t ldy, int flags);

// VAG_GEMM=simt forces the FP32 FFMA path everywhere (A/B runs); default = the tcgen05 split-precision kernels where eligible.
bool tc_enabled();
bool tc_enabled() {
    const char* e = getenv("VAG_GEMM");
    return !(e && strcmp(e, "simt") == 0);
}

size_t gemm_scratch_bytes(int64_t rows, int64_t K, int64_t N) { return linear_tc_scratch_bytes(rows, K, N); }

// The one contraction entry point of the composites.  With a scratch region (for the hi/lo operand splits) and an
// eligible shape the tcgen05 kernel runs; otherwise the SIMT FP32 kernel.
int linear_dispatch(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw, const float* bias,
                    int rows, int K, int N, int flags, cudaStream_t st, void* scratch, size_t scratch_bytes) {
    if (rows == 0 || N == 0) return VAG_OK;
    if (!(flags & VAG_LIN_FORCE_SIMT) && scratch && tc_enabled() && linear_tc_eligible(y, ldy, x, ldx, w, ldw, rows, K, N) &&
        scratch_bytes >= linear_tc_scratch_bytes(rows, K, N))
        return linear_tc(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, scratch, scratch_bytes, st);
    if (flags & VAG_LIN_FORCE_TC) {
        set_error("linear: shape rows=%d K=%d N=%d (or its scratch) is not eligible for the tensor-core path", rows, K, N);
        return VAG_ERR_UNSUPPORTED;
    }
    return linear_simt(y, ldy, x, ldx, w, ldw, bias, rows, K, N, flags, st);
}

}  // namespace vag

using namespace vag;

extern "C" const char* vag_last_error(void) { return g_err; }
extern "C" int vag_abi_version(void) { return 1; }
extern "C" long long vag_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" int vag_device_supported(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return (major == 10 && minor == 0) ? 1 : 0;
}

extern "C" int vag_linear_f32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                              const float* bias, int rows, int in_dim, int out_dim, int flags, vag_stream_t stream) {
    VAG_REQUIRE(y && x && w, "vag_linear_f32: null pointer");
    VAG_REQUIRE(rows >= 0 && in_dim > 0 && out_dim > 0, "vag_linear_f32: bad shape rows=%d in=%d out=%d", rows, in_dim, out_dim);
    VAG_REQUIRE(ldx >= in_dim && ldw >= in_dim && ldy >= out_dim, "vag_linear_f32: leading dimension smaller than the row");
    ModeScope ms((flags & VAG_LIN_BF16) ? VAG_PREC_BF16 : VAG_PREC_FP32);
    return linear_dispatch(y, ldy, x, ldx, w, ldw, bias, rows, in_dim, out_dim, (flags & ~VAG_LIN_BF16) | VAG_LIN_FORCE_SIMT,
                           (cudaStream_t)stream, nullptr, 0);
}

namespace vag {
int tc_elem_bytes();
int tc_split(const float* x, int64_t ldx, int rows, int K, void* hi, void* lo, int64_t ld_out, int64_t col_off, cudaStream_t st);
int tc_gemm(float* y, int64_t ldy, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
            const float* bias, int rows, int K, int N, int flags, cudaStream_t st, float4* summ, int* summ_tile_w);
}

namespace vag {
bool pdl_enabled(int family) {
    static int mask = -1;
    if (mask < 0) {
        const char* e = getenv("VAG_PDL_MASK");
        mask = e ? atoi(e) : (PDL_TC | PDL_SMALL);
        const char* off = getenv("VAG_PDL");
        if (off && off[0] == '0') mask = 0;
    }
    return (mask & family) != 0;
}
}
extern "C" int vag_tc_elem_bytes(int precision) {
    ModeScope ms(precision);
    return tc_elem_bytes();
}
namespace vag { void set_tc_debug(long long* p); }
extern "C" int vag_tc_set_debug(void* device_i64x32) {
    set_tc_debug(reinterpret_cast<long long*>(device_i64x32));
    return VAG_OK;
}

extern "C" int vag_tc_split_f32(const float* x, int64_t ldx, int rows, int K, void* hi, void* lo, int64_t ld_out, int precision,
                                vag_stream_t stream) {
    ModeScope ms(precision);
    VAG_REQUIRE(x && hi && lo, "vag_tc_split_f32: null pointer");
    VAG_REQUIRE(rows > 0 && K > 0 && K % 8 == 0 && ldx % 4 == 0 && ld_out >= K && ld_out % 8 == 0,
                "vag_tc_split_f32: K and ld_out must be multiples of 8, ldx a multiple of 4");
    VAG_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)hi & 127) == 0 && ((uintptr_t)lo & 127) == 0, "vag_tc_split_f32: alignment");
    return tc_split(x, ldx, rows, K, hi, lo, ld_out, 0, (cudaStream_t)stream);
}

extern "C" int vag_tc_gemm_f32(float* y, int64_t ldy, const void* x_hi, const void* x_lo, int64_t ldx, const void* w_hi,
                               const void* w_lo, int64_t ldw, const float* bias, int rows, int in_dim, int out_dim, int flags,
                               vag_stream_t stream) {
    VAG_REQUIRE(y && x_hi && x_lo && w_hi && w_lo, "vag_tc_gemm_f32: null pointer");
    VAG_REQUIRE(rows > 0 && in_dim >= 32 && in_dim % 8 == 0 && out_dim > 0 && ldx % 8 == 0 && ldw % 8 == 0 && ldy >= out_dim,
                "vag_tc_gemm_f32: bad shape");
    ModeScope ms((flags & VAG_LIN_BF16) ? VAG_PREC_BF16 : VAG_PREC_FP32);
    return tc_gemm(y, ldy, x_hi, x_lo, ldx, w_hi, w_lo, ldw, bias, rows, in_dim, out_dim, flags & ~VAG_LIN_BF16, (cudaStream_t)stream,
                   nullptr, nullptr);
}

namespace vag {
int tc_gemm_top2(float4* summ, const void* xh, const void* xl, int64_t ldxs, const void* wh, const void* wl, int64_t ldws,
                 const float* bias, int rows, int K, int N, cudaStream_t st);
}
extern "C" int vag_tc_gemm_top2_f32(float* summ, const void* x_hi, const void* x_lo, int64_t ldx, const void* w_hi,
                                    const void* w_lo, int64_t ldw, const float* bias, int rows, int in_dim, int out_dim,
                                    int precision, vag_stream_t stream) {
    ModeScope ms(precision);
    VAG_REQUIRE(summ && x_hi && x_lo && w_hi && w_lo, "vag_tc_gemm_top2_f32: null pointer");
    VAG_REQUIRE(rows > 128 && in_dim >= 32 && in_dim % 8 == 0 && out_dim > 0 && out_dim < 65535 && ldx % 8 == 0 && ldw % 8 == 0,
                "vag_tc_gemm_top2_f32: bad shape (rows > 128, in %% 8 == 0, out < 65535)");
    VAG_REQUIRE(((uintptr_t)summ & 15) == 0, "vag_tc_gemm_top2_f32: summ must be 16-byte aligned");
    return tc_gemm_top2(reinterpret_cast<float4*>(summ), x_hi, x_lo, ldx, w_hi, w_lo, ldw, bias, rows, in_dim, out_dim, (cudaStream_t)stream);
}

extern "C" size_t vag_linear_tc_workspace_bytes(int rows, int in_dim, int out_dim) {
    return linear_tc_scratch_bytes(rows, in_dim, out_dim);
}

extern "C" int vag_linear_tc_f32(float* y, int64_t ldy, const float* x, int64_t ldx, const float* w, int64_t ldw,
                                 const float* bias, int rows, int in_dim, int out_dim, int flags, void* workspace,
                                 size_t workspace_bytes, vag_stream_t stream) {
    ModeScope ms((flags & VAG_LIN_BF16) ? VAG_PREC_BF16 : VAG_PREC_FP32);
    flags &= ~VAG_LIN_BF16;
    VAG_REQUIRE(y && x && w && workspace, "vag_linear_tc_f32: null pointer");
    VAG_REQUIRE(rows > 0 && in_dim > 0 && out_dim > 0, "vag_linear_tc_f32: bad shape rows=%d in=%d out=%d", rows, in_dim, out_dim);
    VAG_REQUIRE(ldx >= in_dim && ldw >= in_dim && ldy >= out_dim, "vag_linear_tc_f32: leading dimension smaller than the row");
    if (!linear_tc_eligible(y, ldy, x, ldx, w, ldw, rows, in_dim, out_dim) || !tc_call_supported(y, ldy, flags)) {
        set_error("vag_linear_tc_f32: needs rows >= 64, out >= 64, in >= 32 and a multiple of 4, 16-byte aligned operands");
        return VAG_ERR_UNSUPPORTED;
    }
    return linear_tc(y, ldy, x, ldx, w, ldw, bias, rows, in_dim, out_dim, flags, workspace, workspace_bytes, (cudaStream_t)stream);
}
