// GRU cell fused into its contractions (included by linear_tc.cu, which holds the PTX wrappers and tensor-map builders).
//
// One decoder step used to run a GRU cell as   gi = x·W_ihᵀ (contraction, [rows, 3H] fp32 written)  gh = h·W_hhᵀ (contraction,
// [rows, 3H] written)  gates (element-wise kernel: both matrices read back).  At 12 000 rows that is 4 x 73 MB of HBM traffic
// per cell for numbers that are consumed once.  Here a CTA pair owns 256 rows x 32 HIDDEN UNITS and keeps, per unit, four FP32
// accumulator columns in TMEM, laid out   [ gi_n (0-31) | S_r (32-63) | S_z (64-95) | gh_n (96-127) ]   with
//        S_r = gi_r + gh_r        S_z = gi_z + gh_z                                   (PyTorch gate order r | z | n)
// Phase X streams x against the cell's input matrix  with ONE MMA of N = 96 per K step into columns  0-95  (gi_n | r | z),
// phase H streams h against its hidden matrix        with ONE MMA of N = 96 per K step into columns 32-127 (r | z | gh_n),
// accumulating on top of phase X in the r / z columns — exactly the FLOPs of the two separate contractions, and no narrow MMA:
// a tcgen05.mma re-reads its 128 x 16 A slab from shared memory whatever N is, so N = 32 / 64 instructions are bound by
// shared-memory bandwidth ((4096 + 16·N) bytes per N/2 cycles: 288 B/clk at N = 32 against 128 B/clk available; N = 96 needs 117).
// One instruction has one accumulate flag, and phase H must ADD to r / z while it STARTS gh_n: the epilogue warps therefore leave
// the gh_n columns of an accumulator buffer zeroed (tcgen05.st) when they hand it back, and phase H always accumulates.
// The epilogue finishes the cell on the TMEM drain,
//        r = σ(S_r + b_r)   z = σ(S_z + b_z)   n = tanh(gi_n + b_in + r·(gh_n + b_hn))   h' = (1 − z)·n + z·h_prev
// and writes h' once, as fp32 and as tensor-core operand planes for the contractions that consume it.
// gru_1 of the decoder has no phase X: its input pre-activations come from the per-token table (Emb·W_ihᵀ + b_ih) and are added
// in the epilogue (row gather by token id).
//
// Weight layout (made once per weight set, vag_decoder_prepare_f32): rows permuted into tile order, 96 rows per tile of 32 units,
//   input  matrix, tile t (units 32t…32t+31):  [ n(32) | r(32) | z(32) ]        hidden matrix:  [ r(32) | z(32) | n(32) ]
// so that ONE TMA box of 48 rows per CTA and K block delivers what the pair's MMA needs (cta_group::2 takes the first half of an
// instruction's N rows from the leader's shared memory, the second half from the peer's).
constexpr int G_A_BYTES = 128 * Q_ROWB;                 // 128 activation rows x 128 B
constexpr int G_B_BYTES = 48 * Q_ROWB;                  // 48 weight rows x 128 B
constexpr int G_STAGES_SPLIT = 4;                       // 4 x 44 KB (hi + lo planes); bf16 mode: 8 x 22 KB
constexpr int G_BIAS_MAX_H = 1024;                      // biases of the whole layer staged in shared memory: 4 x H floats
constexpr int G_SMEM_BYTES = G_STAGES_SPLIT * (2 * G_A_BYTES + 2 * G_B_BYTES) + 256 /*barriers*/ + 4 * G_BIAS_MAX_H * 4 /*biases*/ + 1024 /*align*/;

#define VAG_TMEM_LD8(v, addr)                                                                                                   \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"                                \
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])               \
                 : "r"(addr)                                                                                                    \
                 : "memory")

// 256-bit global accesses (sm_100: LDG / STG .256).  The epilogue's accesses are one row per lane — 32 cache lines per warp
// instruction whatever its width, and the LSU pays per line — so moving a lane's 64 contiguous bytes with two instructions
// instead of four halves the epilogue's LSU time.  Addresses must be 32-byte aligned.
__device__ __forceinline__ void ldg256(float* d, const float* p) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p));
}
__device__ __forceinline__ void ldg256_nc(float* d, const float* p) {
    asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3]), "=f"(d[4]), "=f"(d[5]), "=f"(d[6]), "=f"(d[7]) : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void stg256_b32(void* p, const uint32_t* v) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}

struct GruArgs {
    const float* bias;        // [4][H]: b_r, b_z, b_in, b_hn (unit-major; made by gru_bias_kernel)
    const float* g1;          // optional per-token table [V, 3H] (gate-major r | z | n, b_ih included): replaces phase X
    const int64_t* tokens;    // [rows] row → table row (out of range → 0, like the embedding); nullptr: table row = row
    int64_t V;
    const float* h_prev;      // [rows, H] fp32
    float* h_out;             // [rows, H] fp32
    uint16_t* out_hi;         // operand planes of h' (pitch out_ld elements; lo unused in bf16 mode)
    uint16_t* out_lo;
    int64_t out_ld;
    int rows, Kx, Kh, H;
    const int* done;
    float* y2;                // optional second fp32 copy of h' (row pitch ld_y2): the encoder's context slab
    int64_t ld_y2;
    int wide;                 // every row base above is 32-byte aligned: 256-bit epilogue accesses (set by the launcher)
};

// NEW = epilogue warps per CTA: 8 (each warp finishes 16 units of its 32 rows in two passes) or 16 (8 units, one pass).  The
// epilogue of a tile is a chain of latencies (token → table row → accumulator → gate math); twice the warps halve it.
template <int MODE, int NEW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * NEW, 1)
gru_pair_kernel(const __grid_constant__ CUtensorMap map_xh, const __grid_constant__ CUtensorMap map_xl,
                const __grid_constant__ CUtensorMap map_hh, const __grid_constant__ CUtensorMap map_hl,
                const __grid_constant__ CUtensorMap map_wih, const __grid_constant__ CUtensorMap map_wil,
                const __grid_constant__ CUtensorMap map_whh, const __grid_constant__ CUtensorMap map_whl, const GruArgs args) {
    constexpr bool SPLIT = MODE != 2;
    constexpr int BMP = 256, BM = 128, UNITS = 32, BK = Q_ROWB / 2, UK = 16;
    constexpr uint32_t FMT = MODE == 1 ? 0u : 1u;   // F16 = 0, BF16 = 1
    constexpr uint32_t IDESC_BASE = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BMP >> 4) << 24);
    constexpr uint32_t IDESC96 = IDESC_BASE | ((uint32_t)(96 >> 3) << 17), IDESC64 = IDESC_BASE | ((uint32_t)(64 >> 3) << 17),
                       IDESC32 = IDESC_BASE | ((uint32_t)(32 >> 3) << 17);
    constexpr int NST = SPLIT ? G_STAGES_SPLIT : 2 * G_STAGES_SPLIT;
    constexpr int A_USED = SPLIT ? 2 * G_A_BYTES : G_A_BYTES, B_USED = SPLIT ? 2 * G_B_BYTES : G_B_BYTES;
    constexpr int STG = A_USED + B_USED;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + G_STAGES_SPLIT * (2 * G_A_BYTES + 2 * G_B_BYTES));
    uint64_t* empty_bar = full_bar + NST;
    uint64_t* tfull_bar = empty_bar + NST;    // [2]
    uint64_t* tempty_bar = tfull_bar + 2;     // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    float* bias_s = reinterpret_cast<float*>(smem + G_STAGES_SPLIT * (2 * G_A_BYTES + 2 * G_B_BYTES) + 256);   // [4][H]: the whole layer's biases

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
    const int H = args.H, rows = args.rows;
    const int tiles_n = H / UNITS, tiles_m = (rows + BMP - 1) / BMP;
    const int n_tiles = tiles_n * tiles_m;
    const int nkb_x = (args.Kx + BK - 1) / BK, nkb_h = (args.Kh + BK - 1) / BK;
    const int nkb = nkb_x + nkb_h;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hh) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whh) : "memory");
        if (SPLIT) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_hl) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_whl) : "memory");
        }
        if (nkb_x) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_xh) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wih) : "memory");
        }
        for (int s = 0; s < NST; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], rank == 0 ? NEW + 1 : NEW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    // the layer's biases (weights: independent of the predecessor kernel) go to shared memory once; the epilogue used to fetch
    // them per tile and pass with same-address global loads whose L2 round trip sat between the TMEM loads and the gate math
    // (only when a CTA works through several tiles: for a small batch the 8 KB copy and its barrier cost more than they save —
    // measured +2.5 us per decoder step at 1500 rows)
    const bool stage_bias = n_tiles > 2 * n_pairs;
    if (stage_bias) {
        for (int i = threadIdx.x; i < 4 * args.H; i += blockDim.x) bias_s[i] = args.bias[i];
        __syncthreads();
    }
    const float* bias_src = stage_bias ? bias_s : args.bias;   // generic pointer: shared or global
    tcgen05_fence_before();
    cluster_sync_all();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_trigger();
    pdl_wait();
    const bool dead = args.done && *reinterpret_cast<const volatile int*>(args.done) != 0;   // see linear_pair_kernel

    if (dead) {
    } else if (warp == 0) {
        if (lane == 0) {
            uint32_t g = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs) {
                const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
                const int m0 = tm * BMP + (int)rank * BM, w0 = tn * 96 + (int)rank * 48;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    mbar_wait(&empty_bar[s], ((g / NST) & 1) ^ 1);
                    uint8_t* st = smem + s * STG;
                    if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * STG);
                    const uint32_t fb = mapa_u32(smem_u32(&full_bar[s]), 0);
                    const bool px = kb < nkb_x;
                    const int k0 = (px ? kb : kb - nkb_x) * BK;
                    tma_load_2d_pair(st, px ? &map_xh : &map_hh, fb, k0, m0);
                    if (SPLIT) tma_load_2d_pair(st + G_A_BYTES, px ? &map_xl : &map_hl, fb, k0, m0);
                    tma_load_2d_pair(st + A_USED, px ? &map_wih : &map_whh, fb, k0, w0);
                    if (SPLIT) tma_load_2d_pair(st + A_USED + G_B_BYTES, px ? &map_wil : &map_whl, fb, k0, w0);
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {   // the whole warp walks the loops (uniform values); one elected lane issues the tcgen05 instructions
            uint32_t g = 0, it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
                const uint32_t a = it & 1;
                // completion k of tempty[a] = both CTAs' epilogues have drained (and re-zeroed) buffer a after its (k-1)-th use;
                // completion 0 is the initial zeroing, so every use waits for a real completion: use k waits for parity k & 1
                mbar_wait(&tempty_bar[a], (it >> 1) & 1);
                tcgen05_fence_after();
                const uint32_t d_main = tmem_base + a * 256, d_cross = d_main + 128;
                for (int kb = 0; kb < nkb; ++kb, ++g) {
                    const int s = g % NST;
                    mbar_wait(&full_bar[s], (g / NST) & 1);
                    tcgen05_fence_after();
                    const uint32_t st = smem_u32(smem + s * STG);
                    const uint64_t d_ah = make_smem_desc<Q_ROWB>(st), d_al = make_smem_desc<Q_ROWB>(st + G_A_BYTES);
                    const uint64_t d_bh = make_smem_desc<Q_ROWB>(st + A_USED), d_bl = make_smem_desc<Q_ROWB>(st + A_USED + G_B_BYTES);
                    const bool px = kb < nkb_x;
                    const int kbh = kb - nkb_x;
                    if (elect_one_lane()) {
                        // phase X: columns 0-95 (gi_n | r | z), started by its first MMA; phase H: columns 32-127 (r | z | gh_n),
                        // always accumulating when a phase X came before (gh_n was left zeroed by the epilogue)
                        const uint32_t col = px ? 0u : 32u;
#pragma unroll
                        for (int j = 0; j < BK / UK; ++j) {
                            const uint64_t adv = (uint64_t)((j * 32) >> 4);
                            const uint32_t acc = px ? (uint32_t)((kb | j) != 0) : (uint32_t)((nkb_x | kbh | j) != 0);
                            if (SPLIT) {
                                umma_pair<true>(d_cross + col, d_al + adv, d_bh + adv, IDESC96, acc);
                                umma_pair<true>(d_cross + col, d_ah + adv, d_bl + adv, IDESC96, 1);
                            }
                            umma_pair<true>(d_main + col, d_ah + adv, d_bh + adv, IDESC96, acc);
                        }
                        tcgen05_commit_pair(&empty_bar[s]);
                        if (kb == nkb - 1) tcgen05_commit_pair(&tfull_bar[a]);
                    }
                    __syncwarp();
                }
            }
        } else if (rank == 1 && lane == 0) {
            // forward "my eight epilogue warps have handed back buffer a" to the leader: the two initial completions (zeroing),
            // then one per tile — same order as the leader's MMA warp consumes them
            uint32_t it = 0;
            for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
                const uint32_t a = it & 1;
                mbar_wait(&tempty_bar[a], (it >> 1) & 1);
                mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[a]), 0));
            }
        }
    } else {
        // ---- epilogue warps 2..: TMEM lane group lg = warp % 4 (rows), unit group uh (UPW of the tile's 32 hidden units)
        constexpr int UPW = 128 / NEW;                // units per warp: 16 (8 warps) or 8 (16 warps)
        constexpr int NP = UPW / 8;                   // passes of 8 units
        const int ew = warp - 2, lg = warp & 3, uh = ew >> 2;
        const bool table = args.g1 != nullptr;
        constexpr uint32_t C_GIN = 0, C_R = 32, C_Z = 64, C_GHN = 96;   // accumulator columns of a buffer (cross accumulator: + 128)
        const uint32_t lane_base = tmem_base + ((uint32_t)(lg * 32) << 16);
        // zero this warp's UPW gh_n columns (main and cross) of accumulator buffer a
        auto zero_ghn = [&](uint32_t a) {
            const uint32_t z = 0u;
#pragma unroll
            for (int q = 0; q < NP; ++q) {
                const uint32_t t0 = lane_base + a * 256 + C_GHN + (uint32_t)(uh * UPW + q * 8);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t0), "r"(z) : "memory");
                if (SPLIT)
                    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(t0 + 128u), "r"(z) : "memory");
            }
        };
        // completion 0 of both tempty barriers: the buffers start with gh_n = 0
        if (nkb_x) { zero_ghn(0); zero_ghn(1); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
        tcgen05_fence_before();
        if (lane == 0) { mbar_arrive(&tempty_bar[0]); mbar_arrive(&tempty_bar[1]); }
        __syncwarp();
        uint32_t it = 0;
        // the token of a row is the head of the epilogue's dependency chain (token → table row → loads): fetched one tile ahead
        auto token_of = [&](int tile) -> int64_t {
            const int r = (tile / tiles_n) * BMP + (int)rank * BM + lg * 32 + lane;
            return (table && tile < n_tiles && r < rows) ? (args.tokens ? args.tokens[r] : (int64_t)r) : 0;
        };
        int64_t id_next = token_of(pair);
        for (int tile = pair; tile < n_tiles; tile += n_pairs, ++it) {
            const int tm = tile / tiles_n, tn = tile - tm * tiles_n;
            const int row = tm * BMP + (int)rank * BM + lg * 32 + lane;
            const bool live = row < rows;
            const uint32_t a = it & 1;
            const float* hp_row = args.h_prev + (int64_t)(live ? row : 0) * H;
            const float* g1_row = nullptr;
            if (table) {
                int64_t id = id_next;
                if (id < 0 || id >= args.V) id = 0;
                g1_row = args.g1 + id * 3 * (int64_t)H;
            }
            // Operands that do not depend on the accumulators — this row's previous state and, for gru_1, its token's table row —
            // are requested for all unit passes before the wait for the tile's MMAs, so their L2 latency (every lane reads its own
            // row: 32 lines per load instruction) overlaps the tensor work.
            const int ub = tn * UNITS + uh * UPW;         // first of this warp's units
            float hp[UPW], tr[UPW], tz[UPW], tq[UPW];
            const bool wide = args.wide != 0;             // kernel-uniform
            if (wide) {
#pragma unroll
                for (int q8 = 0; q8 < NP; ++q8) {
                    ldg256(hp + 8 * q8, hp_row + ub + 8 * q8);
                    if (table) {
                        ldg256_nc(tr + 8 * q8, g1_row + ub + 8 * q8);
                        ldg256_nc(tz + 8 * q8, g1_row + H + ub + 8 * q8);
                        ldg256_nc(tq + 8 * q8, g1_row + 2 * H + ub + 8 * q8);
                    }
                }
            } else {
#pragma unroll
                for (int q4 = 0; q4 < 2 * NP; ++q4) {
                    *reinterpret_cast<float4*>(hp + 4 * q4) = *reinterpret_cast<const float4*>(hp_row + ub + 4 * q4);
                    if (table) {
                        *reinterpret_cast<float4*>(tr + 4 * q4) = *reinterpret_cast<const float4*>(g1_row + ub + 4 * q4);
                        *reinterpret_cast<float4*>(tz + 4 * q4) = *reinterpret_cast<const float4*>(g1_row + H + ub + 4 * q4);
                        *reinterpret_cast<float4*>(tq + 4 * q4) = *reinterpret_cast<const float4*>(g1_row + 2 * H + ub + 4 * q4);
                    }
                }
            }
            id_next = token_of(tile + n_pairs);
            uint32_t keep_h[4], keep_l[4];                // operand planes of pass 0, stored together with pass 1's (wide mode, two passes)
            mbar_wait(&tfull_bar[a], (it >> 1) & 1);
            tcgen05_fence_after();
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const int uc = uh * UPW + p * 8;           // first unit of this pass inside the tile
                const int u0 = tn * UNITS + uc;            // … and in the layer
                uint32_t mr[8], mz[8], mi[8], mh[8], cr[8], cz[8], ci[8], chn[8];
                const uint32_t taddr = lane_base + a * 256 + (uint32_t)uc;
                VAG_TMEM_LD8(mr, taddr + C_R);
                VAG_TMEM_LD8(mz, taddr + C_Z);
                if (!table) VAG_TMEM_LD8(mi, taddr + C_GIN);
                VAG_TMEM_LD8(mh, taddr + C_GHN);
                if (SPLIT) {
                    VAG_TMEM_LD8(cr, taddr + 128u + C_R);
                    VAG_TMEM_LD8(cz, taddr + 128u + C_Z);
                    if (!table) VAG_TMEM_LD8(ci, taddr + 128u + C_GIN);
                    VAG_TMEM_LD8(chn, taddr + 128u + C_GHN);
                }
                float br[8], bz[8], bi[8], bh[8];          // biases: the same address for every lane (shared-memory broadcast / one L1 line)
#pragma unroll
                for (int q4 = 0; q4 < 2; ++q4) {
                    *reinterpret_cast<float4*>(br + 4 * q4) = *reinterpret_cast<const float4*>(bias_src + u0 + 4 * q4);
                    *reinterpret_cast<float4*>(bz + 4 * q4) = *reinterpret_cast<const float4*>(bias_src + H + u0 + 4 * q4);
                    *reinterpret_cast<float4*>(bi + 4 * q4) = *reinterpret_cast<const float4*>(bias_src + 2 * H + u0 + 4 * q4);
                    *reinterpret_cast<float4*>(bh + 4 * q4) = *reinterpret_cast<const float4*>(bias_src + 3 * H + u0 + 4 * q4);
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (p == NP - 1) {   // all TMEM reads of this warp for this tile are complete: re-zero gh_n, hand the buffer back
                    if (nkb_x) { zero_ghn(a); asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
                    tcgen05_fence_before();
                    if (lane == 0) mbar_arrive(&tempty_bar[a]);
                }
                float out[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    constexpr float kc = 1.0f / 2048.0f;
                    float sr = __uint_as_float(mr[u]), sz = __uint_as_float(mz[u]), gh = __uint_as_float(mh[u]);
                    float gi = table ? 0.f : __uint_as_float(mi[u]);
                    if (SPLIT) {
                        sr += __uint_as_float(cr[u]) * kc;
                        sz += __uint_as_float(cz[u]) * kc;
                        gh += __uint_as_float(chn[u]) * kc;
                        if (!table) gi += __uint_as_float(ci[u]) * kc;
                    }
                    sr += br[u];
                    sz += bz[u];
                    gh += bh[u];
                    gi += bi[u];
                    if (table) { sr += tr[8 * p + u]; sz += tz[8 * p + u]; gi = tq[8 * p + u]; }
                    // σ(x) = 1 / (1 + e^-x) and tanh(y) = 1 − 2 / (1 + e^2y) from the compensated exponential (≤ 2 ulp) and the
                    // hardware reciprocal: absolute error ≤ 2e-7, FP32 rounding level of the gate values they feed — at a tenth of
                    // the instructions of expf / tanhf, which would make this epilogue the kernel's critical path
                    const float r = rcp_approx(1.0f + exp2x_comp(-0.5f * sr));
                    const float z = rcp_approx(1.0f + exp2x_comp(-0.5f * sz));
                    const float n = fmaf(-2.0f, rcp_approx(1.0f + exp2x_comp(fmaf(r, gh, gi))), 1.0f);
                    out[u] = fmaf(z, hp[8 * p + u] - n, n);       // (1 − z)·n + z·h_prev
                }
                uint32_t hw[4], lw[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint16_t h0, l0, h1, l1;
                    split_one(MODE, out[2 * u], h0, l0);
                    split_one(MODE, out[2 * u + 1], h1, l1);
                    hw[u] = (uint32_t)h0 | ((uint32_t)h1 << 16);
                    lw[u] = (uint32_t)l0 | ((uint32_t)l1 << 16);
                }
                if (live && args.y2) {
                    float* y2 = args.y2 + (int64_t)row * args.ld_y2 + u0;
                    if (wide) stg256(y2, out);
                    else {
                        *reinterpret_cast<float4*>(y2) = *reinterpret_cast<float4*>(out);
                        *reinterpret_cast<float4*>(y2 + 4) = *reinterpret_cast<float4*>(out + 4);
                    }
                }
                if (live && wide && NP == 2) {
                    stg256(args.h_out + (int64_t)row * H + u0, out);
                    if (p == 0) {
#pragma unroll
                        for (int u = 0; u < 4; ++u) { keep_h[u] = hw[u]; keep_l[u] = lw[u]; }
                    } else {                               // 16 units = 32 bytes per plane
                        const int64_t po = (int64_t)row * args.out_ld + u0 - 8;
                        const uint32_t h8[8] = {keep_h[0], keep_h[1], keep_h[2], keep_h[3], hw[0], hw[1], hw[2], hw[3]};
                        stg256_b32(args.out_hi + po, h8);
                        if (SPLIT) {
                            const uint32_t l8[8] = {keep_l[0], keep_l[1], keep_l[2], keep_l[3], lw[0], lw[1], lw[2], lw[3]};
                            stg256_b32(args.out_lo + po, l8);
                        }
                    }
                } else if (live) {
                    float* ho = args.h_out + (int64_t)row * H + u0;
                    if (wide) stg256(ho, out);
                    else {
                        *reinterpret_cast<float4*>(ho) = *reinterpret_cast<float4*>(out);
                        *reinterpret_cast<float4*>(ho + 4) = *reinterpret_cast<float4*>(out + 4);
                    }
                    const int64_t po = (int64_t)row * args.out_ld + u0;
                    *reinterpret_cast<uint4*>(args.out_hi + po) = make_uint4(hw[0], hw[1], hw[2], hw[3]);
                    if (SPLIT) *reinterpret_cast<uint4*>(args.out_lo + po) = make_uint4(lw[0], lw[1], lw[2], lw[3]);
                }
            }
        }
    }
    tcgen05_fence_before();
    cluster_sync_all();   // the peer's shared memory and the leader's barriers stay alive until both CTAs are done
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

// ---- weight preparation: permuted rows (see the layout comment above) and unit-major bias vectors
// dst [3H, K] fp32 = rows of src [3H, K] (gate-major r | z | n) in tile order; hidden != 0 selects the hidden-matrix order.
__global__ void __launch_bounds__(256)
gru_permute_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int H, int K, int hidden) {
    const int drow = blockIdx.x;                    // destination row
    const int t = drow / 96, p = drow % 96;
    const int slot = p / 32, u = p % 32;
    const int gate = hidden ? slot : (slot == 0 ? 2 : slot - 1);      // hidden: r | z | n      input: n | r | z
    const float* s = src + (int64_t)(gate * H + t * 32 + u) * K;
    float* d = dst + (int64_t)drow * K;
    for (int c = threadIdx.x; c < K; c += blockDim.x) d[c] = s[c];
}
// out [4][H] = b_r, b_z, b_in, b_hn.  with_ih: b_r = b_ih_r + b_hh_r, b_z likewise, b_in = b_ih_n (gru_2); otherwise the input
// side (bias included) arrives through the per-token table: b_r = b_hh_r, b_z = b_hh_z, b_in = 0 (gru_1).
__global__ void gru_bias_kernel(float* __restrict__ out, const float* __restrict__ b_ih, const float* __restrict__ b_hh, int H, int with_ih) {
    const int u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= H) return;
    const float hr = b_hh ? b_hh[u] : 0.f, hz = b_hh ? b_hh[H + u] : 0.f, hn = b_hh ? b_hh[2 * H + u] : 0.f;
    const float ir = (with_ih && b_ih) ? b_ih[u] : 0.f, iz = (with_ih && b_ih) ? b_ih[H + u] : 0.f, in_ = (with_ih && b_ih) ? b_ih[2 * H + u] : 0.f;
    out[u] = ir + hr;
    out[H + u] = iz + hz;
    out[2 * H + u] = in_;
    out[3 * H + u] = hn;
}
