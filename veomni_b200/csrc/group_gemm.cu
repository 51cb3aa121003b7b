// Ragged MoE GroupGEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA) for sm_100a.
//
// Reference: group_gemm_same_nk / group_gemm_same_mn
//   (veomni/ops/kernels/moe/_kernels/kernel/group_gemm.py:157-234, 357-397; Triton kernels :54-154, :241-354)
//   used by the fused MoE autograd functions (veomni/ops/kernels/moe/group_gemm.py:269-444) and the EP
//   variants (veomni/distributed/moe/moe_layer.py:140-441):
//     mode NT  (fwd)   C[rows g] = A[rows g] (M_g x K) * B[g]^T,  B: [G, N, K]   (transpose_b=True)
//     mode NN  (dgrad) C[rows g] = A[rows g] (M_g x K) * B[g],    B: [G, K, N]   (transpose_b=False)
//     mode TN  (wgrad) C[g] (M x N) = A[rows g]^T (M x K_g) * B[rows g] (K_g x N), zero when K_g == 0
//   bf16 operands, fp32 accumulation, one rounding of the result to bf16.  Rows past cumsum[G-1] are
//   never written (the reference leaves them unspecified as well).
//
// Design: persistent kernel, one CTA per SM, warp-specialised:
//   warp 0   TMA producer: SWIZZLE_128B boxes into a 4-stage smem ring of (128x64 A + 256x64 B) tiles (mbarrier tx-count)
//   warp 1   MMA issuer: one elected lane issues tcgen05.mma (M=128, N=256, K=16, cta_group::1); operands
//            are read from smem through UMMA descriptors (K-major or MN-major), accumulators live in TMEM
//            (2 x 256 columns = all of TMEM, double-buffered against the epilogue); tcgen05.commit frees smem stages and
//            hands the accumulator over
//   warps 2-5 epilogue: tcgen05.ld (32 lanes x 32 columns per instruction) -> bf16 -> guarded global stores
// The tile list (group, m-tile, n-tile) is derived on the device from the cumsum tensor, so no host
// synchronisation is needed (the reference launches a max_M-sized grid and early-exits, :97-98).
// Ragged edges: M tails are handled by row guards at the store; the ragged-K tail of mode TN is zeroed in
// shared memory (A operand rows past the group's end) before the MMA consumes the stage.
#include <cstdlib>

#include "umma.cuh"

namespace vb {

constexpr int GG_BM = 128, GG_BN = 256, GG_BK = 64, GG_STAGES = 4;
constexpr int GG_STAGE_BYTES = (GG_BM + GG_BN) * GG_BK * 2;  // 48 KB (a 128x256 tile halves the L2->smem bytes per FLOP of 128x128)
constexpr int GG_MAX_G = 1024;
constexpr int GG_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2-9 epilogue (two per TMEM lane quadrant)
constexpr int GG_EPI_WARPS = 8;

enum { GG_NT = 0, GG_NN = 1, GG_TN = 2 };

struct GGParams {
    const int* cumsum;  // [G] inclusive
    int G;
    int M, N, K;  // NT/NN: N, K = per-group GEMM dims (M unused); TN: M, N = output dims (K unused)
    __nv_bfloat16* C;
    int staged_epi;  // group_gemm_kernel: epilogue rows leave through a per-warp smem transpose (coalesced 64-byte row segments)
};
constexpr int GG_EPI_ROW = 64;                       // bytes per staged row (32 bf16); its four 16-byte pieces are XOR-swizzled
constexpr int GG_EPI_STAGE = 32 * GG_EPI_ROW;        // one warp's 32 rows x 32 bf16 columns

// Tile bookkeeping shared by the three roles.
struct TileInfo {
    int g, mt, nt;       // group, m-tile, n-tile
    int row0;            // NT/NN: first A/C row of the tile; TN: first reduction row of the group
    int rows_valid;      // NT/NN: valid rows in this m-tile (1..128); TN: K_g
};

template <int MODE>
__device__ __forceinline__ bool get_tile(int t, const int* tile_start /*smem [G+1]*/, const GGParams& p, int n_tiles_n,
                                         int m_tiles_tn, TileInfo& ti) {
    const int total = tile_start[p.G];
    if (t >= total) return false;
    if (MODE == GG_TN) {
        const int per_g = m_tiles_tn * n_tiles_n;
        ti.g = t / per_g;
        const int r = t - ti.g * per_g;
        ti.mt = r / n_tiles_n;
        ti.nt = r - ti.mt * n_tiles_n;
        const int s = ti.g ? p.cumsum[ti.g - 1] : 0;
        ti.row0 = s;
        ti.rows_valid = p.cumsum[ti.g] - s;
        return true;
    }
    const int mtg = t / n_tiles_n;  // global m-tile index
    ti.nt = t - mtg * n_tiles_n;
    int lo = 0, hi = p.G - 1;       // find g with tile_start[g] <= mtg*n_tiles_n < tile_start[g+1]
    const int key = mtg * n_tiles_n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid + 1] <= key) lo = mid + 1;
        else hi = mid;
    }
    ti.g = lo;
    ti.mt = mtg - tile_start[lo] / n_tiles_n;
    const int s = lo ? p.cumsum[lo - 1] : 0;
    ti.row0 = s + ti.mt * GG_BM;
    ti.rows_valid = min(GG_BM, p.cumsum[lo] - ti.row0);
    return true;
}

template <int MODE>
__global__ void __launch_bounds__(GG_THREADS, 1)
group_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GGParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stages = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GG_STAGES * GG_STAGE_BYTES);
    uint64_t* empty = full + GG_STAGES;
    uint64_t* tfull = empty + GG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int* tile_start = reinterpret_cast<int*>(tmem_slot + 2);  // [G+1]
    uint8_t* epi_stage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tile_start + p.G + 1) + 15) & ~uintptr_t(15));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int n_tiles_n = (p.N + GG_BN - 1) / GG_BN;
    const int m_tiles_tn = (p.M + GG_BM - 1) / GG_BM;
    // every thread helps building the tile prefix (G is small)
    if (threadIdx.x == 0) {
        int run = 0;
        for (int g = 0; g < p.G; ++g) {
            tile_start[g] = run;
            if (MODE == GG_TN) run += m_tiles_tn * n_tiles_n;
            else {
                const int rows = p.cumsum[g] - (g ? p.cumsum[g - 1] : 0);
                run += ((rows + GG_BM - 1) / GG_BM) * n_tiles_n;
            }
        }
        tile_start[p.G] = run;
        for (int s = 0; s < GG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], GG_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * GG_BN);  // two fp32 accumulators of GG_BN columns
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            int s = 0;
            uint32_t ph = 0;
            TileInfo ti;
            for (int t = blockIdx.x; get_tile<MODE>(t, tile_start, p, n_tiles_n, m_tiles_tn, ti); t += gridDim.x) {
                const int kblocks = (MODE == GG_TN) ? (ti.rows_valid + GG_BK - 1) / GG_BK : (p.K + GG_BK - 1) / GG_BK;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = stages + s * GG_STAGE_BYTES;
                    uint8_t* sb = sa + GG_BM * GG_BK * 2;
                    mbar_expect_tx(&full[s], GG_STAGE_BYTES);
                    if (MODE == GG_TN) {
                        const int kr = ti.row0 + kb * GG_BK;
                        tma_load_2d(sa, &tmA, ti.mt * GG_BM, kr, &full[s]);
                        tma_load_2d(sa + GG_BK * 128, &tmA, ti.mt * GG_BM + 64, kr, &full[s]);
#pragma unroll
                        for (int i = 0; i < GG_BN / 64; ++i)
                            tma_load_2d(sb + i * GG_BK * 128, &tmB, ti.nt * GG_BN + i * 64, kr, &full[s]);
                    } else {
                        tma_load_2d(sa, &tmA, kb * GG_BK, ti.row0, &full[s]);
                        if (MODE == GG_NT) {
                            tma_load_3d(sb, &tmB, kb * GG_BK, ti.nt * GG_BN, ti.g, &full[s]);
                        } else {
#pragma unroll
                            for (int i = 0; i < GG_BN / 64; ++i)
                                tma_load_3d(sb + i * GG_BK * 128, &tmB, ti.nt * GG_BN + i * 64, kb * GG_BK, ti.g, &full[s]);
                        }
                    }
                    if (++s == GG_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc = umma_idesc(MODE == GG_TN, MODE != GG_NT, GG_BM, GG_BN);
        int s = 0, acc = 0;
        uint32_t ph = 0, aph = 0;
        TileInfo ti;
        for (int t = blockIdx.x; get_tile<MODE>(t, tile_start, p, n_tiles_n, m_tiles_tn, ti); t += gridDim.x) {
            const int kblocks = (MODE == GG_TN) ? (ti.rows_valid + GG_BK - 1) / GG_BK : (p.K + GG_BK - 1) / GG_BK;
            mbar_wait(&tempty[acc], aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * GG_BN;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                uint8_t* sa = stages + s * GG_STAGE_BYTES;
                uint8_t* sb = sa + GG_BM * GG_BK * 2;
                if (MODE == GG_TN) {
                    const int valid = ti.rows_valid - kb * GG_BK;  // reduction rows of this block inside the group
                    if (valid < GG_BK) {
                        // zero A's rows [valid, 64): rows of the next group must not contribute
                        for (int r = valid + (lane >> 4); r < GG_BK; r += 2) {
                            const int c = lane & 15;  // 16 x 16 B = two 128-byte box rows
                            *reinterpret_cast<uint4*>(sa + (c >> 3) * GG_BK * 128 + r * 128 + (c & 7) * 16) = make_uint4(0, 0, 0, 0);
                        }
                        fence_async_smem();
                    }
                    __syncwarp();
                }
                if (elect_one_sync()) {
                    const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
                    for (int k = 0; k < GG_BK / 16; ++k) {
                        // A: MN-major for TN (wgrad), else K-major; B: K-major for NT (fwd), else MN-major
                        const uint32_t a_off = MODE == GG_TN ? k * 16 * 128 : k * 32, a_lbo = MODE == GG_TN ? GG_BK * 128 : 16;
                        const uint32_t b_off = MODE == GG_NT ? k * 32 : k * 16 * 128, b_lbo = MODE == GG_NT ? 16 : GG_BK * 128;
                        umma_f16_bo(tmem_d, a_addr >> 4, a_off, a_lbo, 1024, b_addr >> 4, b_off, b_lbo, 1024, idesc,
                                    (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == kblocks - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++s == GG_STAGES) { s = 0; ph ^= 1; }
            }
            if (kblocks == 0 && lane == 0) mbar_arrive(&tfull[acc]);  // empty group (TN): epilogue writes zeros
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        // ===== epilogue (warps 2..9; TMEM lane quadrant = warp % 4; warps w and w+4 split the tile's columns) =====
        // With short reductions (wgrad: K_g ~ 256 tokens = 4 k-blocks, ~2000 clk of MMA per tile) the epilogue of a
        // 128x256 tile is longer than its mainloop; eight warps halve it.
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0;
        uint32_t aph = 0;
        TileInfo ti;
        for (int t = blockIdx.x; get_tile<MODE>(t, tile_start, p, n_tiles_n, m_tiles_tn, ti); t += gridDim.x) {
            const int kblocks = (MODE == GG_TN) ? (ti.rows_valid + GG_BK - 1) / GG_BK : (p.K + GG_BK - 1) / GG_BK;
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const int r = q * 32 + lane;  // row inside the tile
            bool row_ok;
            __nv_bfloat16* crow;
            if (MODE == GG_TN) {
                const int m = ti.mt * GG_BM + r;
                row_ok = m < p.M;
                crow = p.C + ((int64_t)ti.g * p.M + m) * p.N + ti.nt * GG_BN;
            } else {
                row_ok = r < ti.rows_valid;
                crow = p.C + (int64_t)(ti.row0 + r) * p.N + ti.nt * GG_BN;
            }
            const int ncols = min(GG_BN, p.N - ti.nt * GG_BN);
#pragma unroll 1
            for (int c = half * (GG_BN / 64); c < (half + 1) * (GG_BN / 64); ++c) {
                uint32_t v[32];
                if (kblocks > 0) {
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GG_BN + c * 32, v);
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = 0u;
                }
                if (p.staged_epi && c * 32 + 32 <= ncols) {
                    // A thread owns a ROW of the accumulator (TMEM lane): storing its 64 bytes directly makes every store
                    // instruction of the warp touch 32 different 128-byte lines, 16 bytes each — 4096 such line requests
                    // per 128x256 tile, which is what bounded the wgrad (K_g ~ 256: 2048 clk of MMA per tile, ~7900 clk
                    // measured per tile, tensor pipe 30 %; profiles/r02_topkernels_ncu.txt). Transposed through a
                    // swizzled per-warp staging block, an instruction writes 8 rows x 64 contiguous bytes instead.
                    uint8_t* stg = epi_stage + (warp - 2) * GG_EPI_STAGE;
                    __syncwarp();  // the previous chunk's read-back is complete
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        uint4 o;
                        o.x = f2_to_bf2(__uint_as_float(v[i * 8 + 0]), __uint_as_float(v[i * 8 + 1]));
                        o.y = f2_to_bf2(__uint_as_float(v[i * 8 + 2]), __uint_as_float(v[i * 8 + 3]));
                        o.z = f2_to_bf2(__uint_as_float(v[i * 8 + 4]), __uint_as_float(v[i * 8 + 5]));
                        o.w = f2_to_bf2(__uint_as_float(v[i * 8 + 6]), __uint_as_float(v[i * 8 + 7]));
                        // piece i of row `lane` sits at slot i ^ ((lane >> 1) & 3): the 8 lanes of a quarter-warp then cover
                        // all 32 banks both here (8 rows, same piece) and in the read-back (2 rows, 4 pieces)
                        *reinterpret_cast<uint4*>(stg + lane * GG_EPI_ROW + ((i ^ ((lane >> 1) & 3)) << 4)) = o;
                    }
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int rr = j * 8 + (lane >> 2), ch = lane & 3;  // row inside the warp's 32, 16-byte piece of its 64
                        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * GG_EPI_ROW + ((ch ^ ((rr >> 1) & 3)) << 4));
                        const int r2 = q * 32 + rr;
                        if (MODE == GG_TN) {
                            const int m2 = ti.mt * GG_BM + r2;
                            if (m2 < p.M)
                                __stcs(reinterpret_cast<uint4*>(p.C + ((int64_t)ti.g * p.M + m2) * p.N + ti.nt * GG_BN + c * 32 + ch * 8), val);
                        } else if (r2 < ti.rows_valid) {
                            *reinterpret_cast<uint4*>(p.C + (int64_t)(ti.row0 + r2) * p.N + ti.nt * GG_BN + c * 32 + ch * 8) = val;
                        }
                    }
                } else if (row_ok) {
                    if (c * 32 + 32 <= ncols) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            uint4 o;
                            o.x = f2_to_bf2(__uint_as_float(v[i * 8 + 0]), __uint_as_float(v[i * 8 + 1]));
                            o.y = f2_to_bf2(__uint_as_float(v[i * 8 + 2]), __uint_as_float(v[i * 8 + 3]));
                            o.z = f2_to_bf2(__uint_as_float(v[i * 8 + 4]), __uint_as_float(v[i * 8 + 5]));
                            o.w = f2_to_bf2(__uint_as_float(v[i * 8 + 6]), __uint_as_float(v[i * 8 + 7]));
                            *reinterpret_cast<uint4*>(crow + c * 32 + i * 8) = o;
                        }
                    } else {
                        for (int i = 0; i < 32; ++i)
                            if (c * 32 + i < ncols) crow[c * 32 + i] = __float2bfloat16_rn(__uint_as_float(v[i]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * GG_BN);
    }
}


// ================================================================================================================
// Swapped-operand variant for the ragged-M modes (NT forward, NN dgrad):  C^T tile = W tile (M side) x tokens^T (N side).
//   D[128 weight rows (output features), ntok tokens] = W_g[128, K] * X[rows of group g, K]^T
// Why: the tensor core's M is 128 rows or nothing, its N is any multiple of 16 up to 256. With tokens on M (the layout of
// group_gemm_kernel above, and of the reference's Triton kernel) an expert with 260 tokens costs three 128-row tiles
// (+48 % padding work; Qwen3-30B-A3B at T=4096 averages 256 tokens per expert: ~25 % of all MMA work is padding);
// with tokens on N the same expert is one 256-token tile plus one 16-token tile (N = 16), and every weight tile is
// streamed from HBM/L2 once per <= 256 tokens instead of once per 128.
// Output features sit on TMEM lanes, tokens on columns: the epilogue stores C[token][feature] with the 32 lanes of a
// warp on 32 consecutive features (64-byte segments).
// ================================================================================================================
constexpr int GS_BM = 128;   // weight rows (output features) per tile
constexpr int GS_TOK = 256;  // tokens per tile (MMA N <= 256)
constexpr int GS_STAGE_BYTES = (GS_BM + GS_TOK) * GG_BK * 2;  // 48 KB

struct SwapTile {
    int g, tt, wt;
    int row0;        // first token row of the tile
    int rows_valid;  // 1..256 tokens of the group in this tile
    int ntok;        // MMA N: rows_valid rounded up to 16
};

__device__ __forceinline__ bool get_swap_tile(int t, const int* tile_start, const GGParams& p, int n_wt, SwapTile& ti) {
    if (t >= tile_start[p.G]) return false;
    int lo = 0, hi = p.G - 1;  // tile_start[lo] <= t < tile_start[lo + 1]
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid + 1] <= t) lo = mid + 1;
        else hi = mid;
    }
    ti.g = lo;
    const int r = t - tile_start[lo];
    ti.tt = r / n_wt;
    ti.wt = r - ti.tt * n_wt;  // weight tile fastest: CTAs working on consecutive t share the token tile in L2
    const int s = lo ? p.cumsum[lo - 1] : 0;
    ti.row0 = s + ti.tt * GS_TOK;
    ti.rows_valid = min(GS_TOK, p.cumsum[lo] - ti.row0);
    ti.ntok = (ti.rows_valid + 15) & ~15;
    return true;
}

// NN = false: W is [G, N, K] (transpose_b=True, forward);  NN = true: W is [G, K, N] (dgrad)
template <bool NN>
__global__ void __launch_bounds__(GG_THREADS, 1)
group_gemm_swap_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const GGParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stages = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GG_STAGES * GS_STAGE_BYTES);
    uint64_t* empty = full + GG_STAGES;
    uint64_t* tfull = empty + GG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int* tile_start = reinterpret_cast<int*>(tmem_slot + 2);  // [G+1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_wt = (p.N + GS_BM - 1) / GS_BM;
    const int kblocks = (p.K + GG_BK - 1) / GG_BK;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int g = 0; g < p.G; ++g) {
            tile_start[g] = run;
            const int rows = p.cumsum[g] - (g ? p.cumsum[g - 1] : 0);
            run += ((rows + GS_TOK - 1) / GS_TOK) * n_wt;
        }
        tile_start[p.G] = run;
        for (int s = 0; s < GG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], GG_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 2 * GS_TOK);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            int s = 0;
            uint32_t ph = 0;
            SwapTile ti;
            for (int t = blockIdx.x; get_swap_tile(t, tile_start, p, n_wt, ti); t += gridDim.x) {
                const int nbox = (ti.ntok + 63) >> 6;  // 64-token boxes actually needed
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = stages + s * GS_STAGE_BYTES;  // weight tile [128 features x 64 k]
                    uint8_t* sb = sa + GS_BM * GG_BK * 2;       // token tile  [<=256 tokens x 64 k]
                    mbar_expect_tx(&full[s], GS_BM * GG_BK * 2 + nbox * 64 * GG_BK * 2);
                    if (!NN) {
                        tma_load_3d(sa, &tmW, kb * GG_BK, ti.wt * GS_BM, ti.g, &full[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < GS_BM / 64; ++i)
                            tma_load_3d(sa + i * GG_BK * 128, &tmW, ti.wt * GS_BM + i * 64, kb * GG_BK, ti.g, &full[s]);
                    }
                    for (int i = 0; i < nbox; ++i)
                        tma_load_2d(sb + i * 64 * 128, &tmX, kb * GG_BK, ti.row0 + i * 64, &full[s]);
                    if (++s == GG_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        int s = 0, acc = 0;
        uint32_t ph = 0, aph = 0;
        SwapTile ti;
        for (int t = blockIdx.x; get_swap_tile(t, tile_start, p, n_wt, ti); t += gridDim.x) {
            const uint32_t idesc = umma_idesc(NN ? 1 : 0, 0, GS_BM, ti.ntok);
            mbar_wait(&tempty[acc], aph ^ 1);
            tc_fence_after();
            const uint32_t tmem_d = tmem_base + acc * GS_TOK;
            for (int kb = 0; kb < kblocks; ++kb) {
                mbar_wait(&full[s], ph);
                tc_fence_after();
                uint8_t* sa = stages + s * GS_STAGE_BYTES;
                uint8_t* sb = sa + GS_BM * GG_BK * 2;
                if (elect_one_sync()) {
                    const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
                    for (int k = 0; k < GG_BK / 16; ++k) {
                        // A = weights: K-major (forward) or MN-major (dgrad, two 64-feature boxes); B = tokens, K-major
                        const uint32_t a_off = NN ? k * 16 * 128 : k * 32, a_lbo = NN ? GG_BK * 128 : 16;
                        umma_f16_bo(tmem_d, a_addr >> 4, a_off, a_lbo, 1024, b_addr >> 4, k * 32, 16, 1024, idesc, (kb | k) ? 1u : 0u);
                    }
                    umma_commit(&empty[s]);
                    if (kb == kblocks - 1) umma_commit(&tfull[acc]);
                }
                __syncwarp();
                if (++s == GG_STAGES) { s = 0; ph ^= 1; }
            }
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    } else {
        // ===== epilogue (warps 2..9; TMEM lane quadrant = warp % 4): lane = output feature, column = token; warps w and
        // w+4 take alternate 32-token chunks =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0;
        uint32_t aph = 0;
        SwapTile ti;
        for (int t = blockIdx.x; get_swap_tile(t, tile_start, p, n_wt, ti); t += gridDim.x) {
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const int f = ti.wt * GS_BM + q * 32 + lane;
            const bool f_ok = f < p.N;
            __nv_bfloat16* cbase = p.C + (int64_t)ti.row0 * p.N + f;
            const int nch = (ti.rows_valid + 31) >> 5;
#pragma unroll 1
            for (int c = half; c < nch; c += 2) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GS_TOK + c * 32, v);
                if (f_ok) {
                    const int left = ti.rows_valid - c * 32;
                    __nv_bfloat16* cp = cbase + (int64_t)c * 32 * p.N;
                    if (left >= 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) cp[(int64_t)j * p.N] = __float2bfloat16_rn(__uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < left) cp[(int64_t)j * p.N] = __float2bfloat16_rn(__uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 2 * GS_TOK);
    }
}


// ================================================================================================================
// 2-CTA ("cta_group::2") form of the swapped-operand kernel: a cluster of two CTAs (one TPC) computes a
// 256-feature x <=256-token tile. Each CTA stages HALF of both operands — its 128 weight rows and half of the token rows —
// and the leader CTA's single thread issues tcgen05.mma.cta_group::2 (M = 256), which reads both CTAs' shared memory and
// writes each CTA's 128 x ntok half of the accumulator into that CTA's TMEM. Per FLOP a CTA moves (16 + <=16) KB per
// k-block instead of (16 + <=32): the 1-CTA kernels sit on the L2 -> shared-memory path (ncu), this is the
// lever cuBLAS / CUTLASS pull with their 256x256 2-SM tiles.
// Barriers: full[s] lives in the leader only (both CTAs' TMA loads complete on it: cp.async.bulk.tensor ...cta_group::2 with
// the barrier's shared::cluster address of CTA 0); empty[s] and tfull[a] exist in both CTAs and are signalled by the leader's
// tcgen05.commit ...multicast::cluster; tempty[a] lives in the leader and counts the epilogue warps of BOTH CTAs (the peer's
// arrive remotely). The CTAs of a cluster stay together until the last MMA has retired (cluster barrier at both ends).
// ================================================================================================================
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t* slot_in_smem, uint32_t ncols) {  // whole warp, in both CTAs
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_in_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t base, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols) : "memory");
}
// TMA loads whose completion bytes go to the barrier at the same shared-memory offset in CTA 0 of the pair
__device__ __forceinline__ void tma_load_2d_2sm(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_f16_bo_2sm(uint32_t tmem_d, uint32_t a16, uint32_t a_off, uint32_t a_lbo, uint32_t a_sbo,
                                                uint32_t b16, uint32_t b_off, uint32_t b_lbo, uint32_t b_sbo, uint32_t idesc,
                                                uint32_t accumulate) {
    const uint32_t a_lo = a16 + ((a_off >> 4) + (((a_lbo >> 4) & 0x3FFFu) << 16));
    const uint32_t b_lo = b16 + ((b_off >> 4) + (((b_lbo >> 4) & 0x3FFFu) << 16));
    const uint32_t a_hi = ((a_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t b_hi = ((b_sbo >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}"
        ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {  // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta0(uint64_t* bar) {  // arrive on the barrier at this offset in CTA 0
    asm volatile(
        "{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, 0;\n\tmbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
        ::"r"(smem_u32(bar)) : "memory");
}

constexpr int G2_BM = 256;  // weight rows per cluster tile (128 per CTA)

__device__ __forceinline__ bool get_swap2_tile(int t, const int* tile_start, const GGParams& p, int n_wt, SwapTile& ti) {
    if (t >= tile_start[p.G]) return false;
    int lo = 0, hi = p.G - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tile_start[mid + 1] <= t) lo = mid + 1;
        else hi = mid;
    }
    ti.g = lo;
    const int r = t - tile_start[lo];
    ti.tt = r / n_wt;
    ti.wt = r - ti.tt * n_wt;
    const int s = lo ? p.cumsum[lo - 1] : 0;
    ti.row0 = s + ti.tt * GS_TOK;
    ti.rows_valid = min(GS_TOK, p.cumsum[lo] - ti.row0);
    ti.ntok = (ti.rows_valid + 31) & ~31;  // MMA N: a multiple of 32 so that each CTA's half is a multiple of 16
    return true;
}

template <bool NN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GG_THREADS, 1)
group_gemm_swap2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const GGParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* stages = smem;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + GG_STAGES * GS_STAGE_BYTES);
    uint64_t* empty = full + GG_STAGES;
    uint64_t* tfull = empty + GG_STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    int* tile_start = reinterpret_cast<int*>(tmem_slot + 2);  // [G+1]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
    const int n_wt = (p.N + G2_BM - 1) / G2_BM;
    const int kblocks = (p.K + GG_BK - 1) / GG_BK;
    if (threadIdx.x == 0) {
        int run = 0;
        for (int g = 0; g < p.G; ++g) {
            tile_start[g] = run;
            const int rows = p.cumsum[g] - (g ? p.cumsum[g - 1] : 0);
            run += ((rows + GS_TOK - 1) / GS_TOK) * n_wt;
        }
        tile_start[p.G] = run;
        for (int s = 0; s < GG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], 2 * GG_EPI_WARPS); }
        mbar_fence_init();
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
    }
    if (warp == 1) tmem_alloc2(tmem_slot, 2 * GS_TOK);
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit / 2-SM TMA
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own 128 weight rows, own half of the token rows) =====
        if (elect_one_sync()) {
            int s = 0;
            uint32_t ph = 0;
            SwapTile ti;
            for (int t = cluster_id; get_swap2_tile(t, tile_start, p, n_wt, ti); t += n_clusters) {
                const int half_tok = ti.ntok >> 1;                  // token rows staged by each CTA (multiple of 16)
                const int nbox = (half_tok + 63) >> 6;               // 64-row boxes per CTA
                const uint32_t bytes_cta = GS_BM * GG_BK * 2 + nbox * 64 * GG_BK * 2;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* sa = stages + s * GS_STAGE_BYTES;
                    uint8_t* sb = sa + GS_BM * GG_BK * 2;
                    if (crank == 0) mbar_expect_tx(&full[s], 2 * bytes_cta);  // the pair's bytes land on the leader's barrier
                    const int w0 = ti.wt * G2_BM + (int)crank * GS_BM;
                    if (!NN) {
                        tma_load_3d_2sm(sa, &tmW, kb * GG_BK, w0, ti.g, &full[s]);
                    } else {
#pragma unroll
                        for (int i = 0; i < GS_BM / 64; ++i)
                            tma_load_3d_2sm(sa + i * GG_BK * 128, &tmW, w0 + i * 64, kb * GG_BK, ti.g, &full[s]);
                    }
                    const int r0 = ti.row0 + (int)crank * half_tok;
                    for (int i = 0; i < nbox; ++i)
                        tma_load_2d_2sm(sb + i * 64 * 128, &tmX, kb * GG_BK, r0 + i * 64, &full[s]);
                    if (++s == GG_STAGES) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (crank == 0) {
            int s = 0, acc = 0;
            uint32_t ph = 0, aph = 0;
            SwapTile ti;
            for (int t = cluster_id; get_swap2_tile(t, tile_start, p, n_wt, ti); t += n_clusters) {
                const uint32_t idesc = umma_idesc(NN ? 1 : 0, 0, G2_BM, ti.ntok);
                mbar_wait(&tempty[acc], aph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + acc * GS_TOK;
                for (int kb = 0; kb < kblocks; ++kb) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    uint8_t* sa = stages + s * GS_STAGE_BYTES;
                    uint8_t* sb = sa + GS_BM * GG_BK * 2;
                    if (elect_one_sync()) {
                        const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
#pragma unroll
                        for (int k = 0; k < GG_BK / 16; ++k) {
                            const uint32_t a_off = NN ? k * 16 * 128 : k * 32, a_lbo = NN ? GG_BK * 128 : 16;
                            umma_f16_bo_2sm(tmem_d, a_addr >> 4, a_off, a_lbo, 1024, b_addr >> 4, k * 32, 16, 1024, idesc, (kb | k) ? 1u : 0u);
                        }
                        umma_commit_2sm(&empty[s]);
                        if (kb == kblocks - 1) umma_commit_2sm(&tfull[acc]);
                    }
                    __syncwarp();
                    if (++s == GG_STAGES) { s = 0; ph ^= 1; }
                }
                if (++acc == 2) { acc = 0; aph ^= 1; }
            }
        }
    } else {
        // ===== epilogue (both CTAs): lane = output feature of this CTA's 128, column = token =====
        const int q = warp & 3;
        const int half = (warp - 2) >> 2;
        int acc = 0;
        uint32_t aph = 0;
        SwapTile ti;
        for (int t = cluster_id; get_swap2_tile(t, tile_start, p, n_wt, ti); t += n_clusters) {
            mbar_wait(&tfull[acc], aph);
            tc_fence_after();
            const int f = ti.wt * G2_BM + (int)crank * GS_BM + q * 32 + lane;
            const bool f_ok = f < p.N;
            __nv_bfloat16* cbase = p.C + (int64_t)ti.row0 * p.N + f;
            const int nch = (ti.rows_valid + 31) >> 5;
#pragma unroll 1
            for (int c = half; c < nch; c += 2) {
                uint32_t v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * GS_TOK + c * 32, v);
                if (f_ok) {
                    const int left = ti.rows_valid - c * 32;
                    __nv_bfloat16* cp = cbase + (int64_t)c * 32 * p.N;
                    if (left >= 32) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) cp[(int64_t)j * p.N] = __float2bfloat16_rn(__uint_as_float(v[j]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < left) cp[(int64_t)j * p.N] = __float2bfloat16_rn(__uint_as_float(v[j]));
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cta0(&tempty[acc]);
            if (++acc == 2) { acc = 0; aph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();  // the peer's shared memory and TMEM stay valid until the leader's last MMA has been consumed
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc2(tmem_base, 2 * GS_TOK);
    }
}

static size_t gg_smem_bytes(int G) {  // stages | barriers | tmem slot | tile_start[G+1] | epilogue staging (16-byte aligned)
    return (size_t)GG_STAGES * GG_STAGE_BYTES + (2 * GG_STAGES + 4) * 8 + 16 + (size_t)(G + 1) * 4 + 64 +
           (size_t)GG_EPI_WARPS * GG_EPI_STAGE;
}

}  // namespace vb

using namespace vb;

// a: NT/NN [sumM, K]; TN [sumK, M].   b: NT [G, N, K]; NN [G, K, N]; TN [sumK, N].   c: NT/NN [sumM, N]; TN [G, M, N].
extern "C" int vb200_group_gemm(int32_t mode, const void* a, const void* b, void* c, const int32_t* cumsum,
                                int32_t num_groups, int64_t total_rows, int32_t m, int32_t n, int32_t k, void* stream) {
    const int variant = (mode >> 8) & 3;  // bits 8-9: 0 = default, 1 = tokens on the MMA M side (classic), 2 = tokens on N (swapped)
    mode &= 0xff;
    if (mode < 0 || mode > 2) return vb200_set_error(VB200_EINVAL, "group_gemm: mode must be 0 (NT), 1 (NN) or 2 (TN)");
    if (num_groups < 1 || num_groups > GG_MAX_G) return vb200_set_error(VB200_EINVAL, "group_gemm: 1..1024 groups");
    if (n <= 0 || (n & 7)) return vb200_set_error(VB200_EINVAL, "group_gemm: N must be a positive multiple of 8");
    if (mode != GG_TN && (k <= 0 || (k & 7))) return vb200_set_error(VB200_EINVAL, "group_gemm: K must be a positive multiple of 8");
    if (mode == GG_TN && (m <= 0 || (m & 7))) return vb200_set_error(VB200_EINVAL, "group_gemm: M must be a positive multiple of 8");
    if (total_rows <= 0 && mode != GG_TN) return VB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    CUtensorMap tmA, tmB;
    int rc;
    GGParams p{};
    p.cumsum = cumsum; p.G = num_groups; p.M = m; p.N = n; p.K = k; p.C = (__nv_bfloat16*)c;
    const uint64_t rows = (uint64_t)(total_rows > 0 ? total_rows : 1);
    const size_t smem = gg_smem_bytes(num_groups);
    static int staged = -1;
    if (staged < 0) {
        const char* e = getenv("VB200_GG_EPI_STAGED");
        staged = (e && e[0] == '0') ? 0 : 1;  // 0: every thread stores its own accumulator row (A/B runs)
    }
    p.staged_epi = staged;
    static int swap_mode = -1;
    if (swap_mode < 0) {
        const char* e = getenv("VB200_GG_SWAP");
        swap_mode = (e && e[0] == '0') ? 0 : 1;  // default on (validated on B200: bit-compatible results, fc1 +4 %, fc2 +11 %, dgrad +6 %)
    }
    static int two_cta = -1;
    if (two_cta < 0) {
        const char* e = getenv("VB200_GG_2CTA");
        two_cta = (e && e[0] == '1') ? 1 : 0;  // default off until validated on hardware
    }
    if (mode != GG_TN && (variant == 3 || (variant == 0 && two_cta))) {
        const size_t smem_sw = (size_t)GG_STAGES * GS_STAGE_BYTES + (2 * GG_STAGES + 4) * 8 + 16 + (size_t)(num_groups + 1) * 4 + 64;
        if ((rc = make_tmap_2d(&tmA, a, k, rows, k, 64))) return rc;  // tokens [rows, K], box 64 x 64
        if (mode == GG_NT) {
            if ((rc = make_tmap_3d_box(&tmB, b, k, n, num_groups, k, (uint64_t)n * k, 64, GS_BM))) return rc;
            static bool attr = false;
            if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_swap2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
            group_gemm_swap2_kernel<false><<<kNumSMs, GG_THREADS, smem_sw, st>>>(tmA, tmB, p);
        } else {
            if ((rc = make_tmap_3d_box(&tmB, b, n, k, num_groups, n, (uint64_t)n * k, 64, GG_BK))) return rc;
            static bool attr = false;
            if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_swap2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
            group_gemm_swap2_kernel<true><<<kNumSMs, GG_THREADS, smem_sw, st>>>(tmA, tmB, p);
        }
        vb200_count_launch(1);
        VB_HOST_CHECK_LAUNCH();
        return VB200_OK;
    }
    if (mode != GG_TN && (variant == 2 || (variant == 0 && swap_mode))) {
        // tokens on the MMA N side (see group_gemm_swap_kernel)
        const size_t smem_sw = (size_t)GG_STAGES * GS_STAGE_BYTES + (2 * GG_STAGES + 4) * 8 + 16 + (size_t)(num_groups + 1) * 4 + 64;
        if ((rc = make_tmap_2d(&tmA, a, k, rows, k, 64))) return rc;  // tokens [rows, K], box 64 x 64
        if (mode == GG_NT) {
            if ((rc = make_tmap_3d_box(&tmB, b, k, n, num_groups, k, (uint64_t)n * k, 64, GS_BM))) return rc;
            static bool attr = false;
            if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_swap_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
            group_gemm_swap_kernel<false><<<kNumSMs, GG_THREADS, smem_sw, st>>>(tmA, tmB, p);
        } else {
            if ((rc = make_tmap_3d_box(&tmB, b, n, k, num_groups, n, (uint64_t)n * k, 64, GG_BK))) return rc;
            static bool attr = false;
            if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_swap_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
            group_gemm_swap_kernel<true><<<kNumSMs, GG_THREADS, smem_sw, st>>>(tmA, tmB, p);
        }
        vb200_count_launch(1);
        VB_HOST_CHECK_LAUNCH();
        return VB200_OK;
    }
    if (mode == GG_NT) {
        if ((rc = make_tmap_2d(&tmA, a, k, rows, k, GG_BM))) return rc;
        if ((rc = make_tmap_3d_box(&tmB, b, k, n, num_groups, k, (uint64_t)n * k, 64, GG_BN))) return rc;
        static bool attr = false;
        if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_kernel<GG_NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
        group_gemm_kernel<GG_NT><<<kNumSMs, GG_THREADS, smem, st>>>(tmA, tmB, p);
    } else if (mode == GG_NN) {
        if ((rc = make_tmap_2d(&tmA, a, k, rows, k, GG_BM))) return rc;
        if ((rc = make_tmap_3d_box(&tmB, b, n, k, num_groups, n, (uint64_t)n * k, 64, GG_BK))) return rc;
        static bool attr = false;
        if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_kernel<GG_NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
        group_gemm_kernel<GG_NN><<<kNumSMs, GG_THREADS, smem, st>>>(tmA, tmB, p);
    } else {
        if ((rc = make_tmap_2d(&tmA, a, m, rows, m, GG_BK))) return rc;
        if ((rc = make_tmap_2d(&tmB, b, n, rows, n, GG_BK))) return rc;
        static bool attr = false;
        if (!attr) { VB_CUDA_TRY(cudaFuncSetAttribute(group_gemm_kernel<GG_TN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); attr = true; }
        group_gemm_kernel<GG_TN><<<kNumSMs, GG_THREADS, smem, st>>>(tmA, tmB, p);
    }
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
