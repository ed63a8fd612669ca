// Data-parallel gradient exchange over NVLink peer memory (one process per GPU, one node).
//
// The training step ends with  all-reduce(64 MB of gradients) → Σ‖g‖² → clip → Adam.  With NCCL that is a library collective
// followed by two passes over the gradients.  Here every rank keeps its flat gradient buffer in a peer-visible allocation
// (cudaMalloc + CUDA IPC; NVSwitch gives every GPU full bandwidth to every peer) and the exchange is two kernels of plain
// peer loads around system-scope barriers:
//   1. reduce-scatter + norm: rank r owns slice r of the buffer; it reads that slice from ALL ranks (fixed order 0 … G-1, so the
//      sum is bit-identical whoever computes it), writes the average into its own copy and leaves the block partials of Σ avg²;
//   2. all-gather: every rank copies the other owners' averaged slices into its own buffer and adds up all owners' partials in
//      a fixed order — every replica ends with the same bits for the gradient AND the norm, which keeps the replicas in lockstep
//      through the (replicated) clip + Adam kernel that follows.
// Barriers: each rank owns an array flags[G] in its arena; a barrier with sequence number e = every rank stores e into ITS slot
// of every peer's array (system-scope release) and waits until all slots of its own array have reached e.  Waits are bounded:
// a rank that never arrives makes the others trap (a CUDA error) instead of hanging the GPU.
#include "common.cuh"
#include <string.h>

namespace vag {

constexpr int kDpMaxWorld = 16;
constexpr int kDpPartials = 1184;       // block partials of Σ avg² per owner (= blocks of the reduce-scatter kernel: 8 per SM)

struct DpPeers {
    char* base[kDpMaxWorld];            // arena base pointer of every rank (own pointer at [rank])
};

__global__ void dp_barrier_kernel(DpPeers peers, int world, int rank, int64_t flags_off, int seq) {
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();             // everything this GPU wrote before (earlier kernels of the stream) is visible system-wide
    volatile int* remote = reinterpret_cast<volatile int*>(peers.base[p] + flags_off) + rank;
    *remote = seq;
    volatile int* mine = reinterpret_cast<volatile int*>(peers.base[rank] + flags_off) + p;
    unsigned long long spins = 0;
    while (*mine < seq) {
        if (++spins > (1ull << 31)) __trap();   // ≈ seconds: a peer died or skipped the step
    }
    __threadfence_system();
}

// Slice r of n float4 elements: [r·per, min(n, (r+1)·per)), per = ceil(n / world)
__device__ __forceinline__ void dp_slice(int64_t n4, int world, int r, int64_t& lo, int64_t& hi) {
    const int64_t per = (n4 + world - 1) / world;
    lo = per * r;
    hi = lo + per < n4 ? lo + per : n4;
    if (lo > n4) lo = n4;
}

__global__ void __launch_bounds__(256)
dp_reduce_scatter_kernel(DpPeers peers, int world, int rank, int64_t grad_off, int64_t part_off, int64_t n4) {
    int64_t lo, hi;
    dp_slice(n4, world, rank, lo, hi);
    const float inv = 1.0f / (float)world;
    float4* mine = reinterpret_cast<float4*>(peers.base[rank] + grad_off);
    float ss = 0.f;
    // four elements per thread and trip: 4·world independent 16-byte peer loads in flight (NVLink latency ≈ 2 µs)
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += 4 * stride) {
        float4 acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < world; ++p) {     // fixed order: the sum does not depend on who computes it
            const float4* src = reinterpret_cast<const float4*>(peers.base[p] + grad_off);
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t i = i0 + u * stride;
                v[u] = i < hi ? __ldcg(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { acc[u].x += v[u].x; acc[u].y += v[u].y; acc[u].z += v[u].z; acc[u].w += v[u].w; }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            if (i < hi) {
                float4 a = acc[u];
                a.x *= inv; a.y *= inv; a.z *= inv; a.w *= inv;
                mine[i] = a;
                ss += (a.x * a.x + a.y * a.y) + (a.z * a.z + a.w * a.w);
            }
        }
    }
    // deterministic block reduction (fixed tree), one partial per block
    __shared__ float red[8];
    ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        reinterpret_cast<float*>(peers.base[rank] + part_off)[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(256)
dp_all_gather_kernel(DpPeers peers, int world, int rank, int64_t grad_off, int64_t part_off, int64_t n4, float* __restrict__ sumsq_out) {
    float4* mine = reinterpret_cast<float4*>(peers.base[rank] + grad_off);
    const int64_t per = (n4 + world - 1) / world;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < n4; i0 += 4 * stride) {
        float4 v[4];
        bool take[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int64_t i = i0 + u * stride;
            const int o = i < n4 ? (int)(i / per) : rank;
            take[u] = o != rank;
            if (take[u]) v[u] = __ldcg(reinterpret_cast<const float4*>(peers.base[o] + grad_off) + i);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (take[u]) mine[i0 + u * stride] = v[u];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && sumsq_out) {   // Σ over owners and blocks in a fixed order: the same bits on every rank
        float t = 0.f;
        for (int o = 0; o < world; ++o) {
            const float* part = reinterpret_cast<const float*>(peers.base[o] + part_off);
            for (int b = 0; b < kDpPartials; ++b) t += __ldcg(part + b);
        }
        *sumsq_out = t;
    }
}

}  // namespace vag

using namespace vag;

extern "C" int vag_p2p_alloc(size_t bytes, void** dev_ptr, unsigned char* handle_out_64) {
    VAG_REQUIRE(dev_ptr && handle_out_64 && bytes > 0, "vag_p2p_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    VAG_CUDA(cudaMalloc(&p, bytes));
    VAG_CUDA(cudaMemset(p, 0, bytes));
    cudaIpcMemHandle_t h;
    VAG_CUDA(cudaIpcGetMemHandle(&h, p));
    memcpy(handle_out_64, &h, 64);
    *dev_ptr = p;
    return VAG_OK;
}
extern "C" int vag_p2p_free(void* dev_ptr) {
    if (dev_ptr) VAG_CUDA(cudaFree(dev_ptr));
    return VAG_OK;
}
extern "C" int vag_p2p_open(const unsigned char* handle_64, void** dev_ptr) {
    VAG_REQUIRE(handle_64 && dev_ptr, "vag_p2p_open: null pointer");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle_64, 64);
    VAG_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VAG_OK;
}
extern "C" int vag_p2p_close(void* dev_ptr) {
    if (dev_ptr) VAG_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return VAG_OK;
}

extern "C" size_t vag_dp_arena_bytes(int64_t n_floats) {
    return align_up((size_t)n_floats * 4, 4096) + 4096 /*flags*/ + (size_t)kDpPartials * 4 + 4096;
}

extern "C" int vag_dp_allreduce_f32(const vag_dp_comm* c, int64_t n_floats, int step, float* sumsq_out, vag_stream_t stream) {
    VAG_REQUIRE(c && c->world >= 1 && c->world <= kDpMaxWorld && c->rank >= 0 && c->rank < c->world, "vag_dp_allreduce_f32: bad communicator");
    VAG_REQUIRE(n_floats > 0 && n_floats % 4 == 0 && step >= 0, "vag_dp_allreduce_f32: n_floats must be a positive multiple of 4");
    cudaStream_t st = (cudaStream_t)stream;
    DpPeers peers;
    for (int p = 0; p < c->world; ++p) {
        VAG_REQUIRE(c->peers[p] != nullptr, "vag_dp_allreduce_f32: peer %d not mapped", p);
        peers.base[p] = (char*)c->peers[p];
    }
    const int64_t grad_off = 0;
    const int64_t flags_off = (int64_t)align_up((size_t)n_floats * 4, 4096);
    const int64_t part_off = flags_off + 4096;
    const int64_t n4 = n_floats / 4;
    const int seq = step * 3;
    dp_barrier_kernel<<<1, 32, 0, st>>>(peers, c->world, c->rank, flags_off, seq + 1);      // every rank's gradients are complete
    VAG_LAUNCH_CHECK();
    dp_reduce_scatter_kernel<<<kDpPartials, 256, 0, st>>>(peers, c->world, c->rank, grad_off, part_off, n4);
    VAG_LAUNCH_CHECK();
    dp_barrier_kernel<<<1, 32, 0, st>>>(peers, c->world, c->rank, flags_off, seq + 2);      // every owner's slice is averaged
    VAG_LAUNCH_CHECK();
    dp_all_gather_kernel<<<8 * num_sms(), 256, 0, st>>>(peers, c->world, c->rank, grad_off, part_off, n4, sumsq_out);
    VAG_LAUNCH_CHECK();
    dp_barrier_kernel<<<1, 32, 0, st>>>(peers, c->world, c->rank, flags_off, seq + 3);      // nobody still reads this step's slices
    VAG_LAUNCH_CHECK();
    return VAG_OK;
}
