// Packed varlen causal attention FORWARD on the 5th-generation tensor cores (tcgen05 + TMEM), head_dim 128.
//
// Same contract as attn_fwd_kernel in attention.cu (reference call site: flash_attn_varlen_func behind
// veomni/ops/kernels/attention/__init__.py:304-320); this is the Blackwell-native pipeline:
//
//   warp 0      TMA producer   Q once; K_j / V_j tiles (128 kv rows) into 2-stage smem rings (SWIZZLE_128B boxes)
//   warp 1      MMA issuer     S_j  = Q K_j^T   tcgen05.mma M128 N128 K16 x8, K-major A/B      -> TMEM S[j&1]
//                              O   += P_j V_j   tcgen05.mma, A = P (smem, K-major), B = V (MN-major) -> TMEM O
//                              (S_{j+1} is issued before PV_j so the tensor pipe runs ahead of the softmax)
//   warps 2..5  softmax        thread == query row (TMEM lane): S read once by tcgen05.ld, row max, exp2 -> bf16 P written
//                              to swizzled smem; O stays in TMEM (PV accumulates into it) and is only rescaled when a
//                              row max grows by more than 2^8 (lazy rescale); LSE saved for the backward
// TMEM: S[2] (2x128 columns) + O (128 columns).  smem: Q 32 KB + K 2x32 KB + V 2x32 KB + P 32 KB.
// All hand-offs are mbarriers (TMA tx-count, tcgen05.commit, thread arrives) — no __syncthreads in the loop.
#include "umma.cuh"

namespace vb {

struct AttnTcParams {
    const int* cu_seqlens;
    int Hq, Hk, total;
    int tiles, nseq;  // 1-D grid = tiles x heads x nseq, heaviest tiles first (see tile_of_block)
    float scale;
    int causal;
    __nv_bfloat16* o;
    int64_t o_stride_tok, o_stride_head;
    float* lse;
};

constexpr int TC_BM = 128, TC_BN = 128, TC_D = 128;

// Block order. With a causal mask the work of a tile grows (Q tiles) or shrinks (KV tiles) linearly with its
// index, and the hardware hands CTAs to SMs in linear block order. A (tile, head, seq) 3-D grid runs one head's
// tiles heavy-to-light and then starts the next head's heaviest tile late: list-scheduling that order on 148 SMs
// costs 264 (fwd, dQ) and 376 (dK/dV) tile-units against an ideal 228 at T=4096, 32/8 heads. Ordering the 1-D grid
// by tile first — every head's heaviest tile, then every head's second heaviest, ... (longest-processing-time
// first) — gives 232 and 256. `order` = position in that list; returns the slot's (order, head, seq).
__device__ __forceinline__ void tile_of_block(int heads, int nseq, int& order, int& head, int& seq) {
    const int per = heads * nseq;
    order = (int)blockIdx.x / per;
    const int rem = (int)blockIdx.x - order * per;
    seq = rem / heads;
    head = rem - seq * heads;
}
constexpr int TC_TILE = TC_BM * TC_D * 2;  // 32 KB
constexpr int TC_VSTAGES = 3;              // V is held until PV_j retires (one tile later than K): deeper ring
constexpr float kLog2eTc = 1.4426950408889634f;

// Row max of NC 32-column chunks of an S tile held in registers (first column n0); MASK applies the causal /
// sequence-end mask in place.
template <bool MASK, int NC>
__device__ __forceinline__ float tile_row_max(uint32_t (&sv)[NC][32], int n0, int m, int L, int causal) {
    float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int c = 0; c < NC; ++c) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            if (MASK) {
                const int n = n0 + c * 32 + i;
                if (n >= L || (causal && n > m)) sv[c][i] = 0xff800000u;  // -inf
            }
            mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[c][i]));
        }
    }
    return fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
}

enum {  // barrier indices
    B_Q = 0, B_KFULL = 1, B_VFULL = 3, B_KEMPTY = 6, B_VEMPTY = 8, B_SFULL = 11, B_SEMPTY = 13, B_PFULL = 15,
    B_PVDONE = 16, B_COUNT = 17
};

// Forward v2 (FA4-style data flow): O stays resident in TMEM for the whole KV sweep (the PV MMAs accumulate
// into it), S is read from TMEM exactly once per tile, and the running max used in the exponentials is only
// refreshed — with a TMEM read-modify-write of O — when a row's new max exceeds it by more than 2^8
// ("lazy rescale"); TMEM read bandwidth, not the tensor pipe, is the scarce resource in this kernel.
//
// W8 = true: eight softmax warps instead of four — warps w and w+4 share a TMEM lane quadrant (a row) and take the
// left / right 64 columns of every S tile; the two half-row maxima are exchanged through shared memory (one named
// barrier per tile), each thread keeps the partial row sum of its columns and rescales / writes its half of O.
// Motivation (tools/ubench_sm100.cu, profiles/r01_ubench_sm100.jsonl): with four warps `tcgen05.ld` delivers 88 B/clk
// and MUFU.EX2 11.9 lanes/clk per SM — 740 + 1377 clk of the 2800 clk a 128x128 tile takes — against 155 B/clk and
// 16 lanes/clk with eight.  Default (attention.FWD_W8): 173 vs 178 us at T=4096 on B200, tests/test_attention_gpu.py.
template <bool W8>
__global__ void __launch_bounds__(W8 ? 320 : 192, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
    constexpr int NC = W8 ? 2 : 4;  // 32-column chunks of S per thread
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sK = sQ + TC_TILE;          // 2 stages
    uint8_t* sV = sK + 2 * TC_TILE;      // TC_VSTAGES stages
    uint8_t* sP = sV + TC_VSTAGES * TC_TILE;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sP + TC_TILE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + B_COUNT);
    float* sHalf = reinterpret_cast<float*>(tmem_slot + 4);  // W8: [2 tile parities][2 halves][128 rows] half-row maxima / sums
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int order, h, seq;
    tile_of_block(p.Hq, p.nseq, order, h, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int mblk = p.tiles - 1 - order;  // causal: the last Q tile is the heaviest
    const int m0 = mblk * TC_BM;
    if (m0 >= L) return;
    const int hk = h / (p.Hq / p.Hk);
    const int kv_end = p.causal ? min(L, m0 + TC_BM) : L;
    const int n_tiles = (kv_end + TC_BN - 1) / TC_BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_COUNT; ++i) {
            const bool by_warps = (i >= B_SEMPTY && i < B_SEMPTY + 2) || i == B_PFULL;
            mbar_init(&bar[i], by_warps ? (W8 ? 8 : 4) : 1);
        }
        mbar_fence_init();
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S[2] 0/128, O 256..383

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one_sync()) {
            mbar_expect_tx(&bar[B_Q], TC_TILE);
            tma_load_3d(sQ, &tmQ, 0, h, s0 + m0, &bar[B_Q]);
            tma_load_3d(sQ + TC_BM * 128, &tmQ, 64, h, s0 + m0, &bar[B_Q]);
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&bar[B_KEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[B_KFULL + st], TC_TILE);
                tma_load_3d(sK + st * TC_TILE, &tmK, 0, hk, s0 + j * TC_BN, &bar[B_KFULL + st]);
                tma_load_3d(sK + st * TC_TILE + TC_BN * 128, &tmK, 64, hk, s0 + j * TC_BN, &bar[B_KFULL + st]);
                const int vs = j % TC_VSTAGES;
                const uint32_t vph = (uint32_t)(j / TC_VSTAGES) & 1u;
                mbar_wait(&bar[B_VEMPTY + vs], vph ^ 1);
                mbar_expect_tx(&bar[B_VFULL + vs], TC_TILE);
                tma_load_3d(sV + vs * TC_TILE, &tmV, 0, hk, s0 + j * TC_BN, &bar[B_VFULL + vs]);
                tma_load_3d(sV + vs * TC_TILE + TC_BN * 128, &tmV, 64, hk, s0 + j * TC_BN, &bar[B_VFULL + vs]);
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        constexpr uint32_t idesc_qk = umma_idesc(0, 0, TC_BM, TC_BN);  // A, B K-major
        constexpr uint32_t idesc_pv = umma_idesc(0, 1, TC_BM, TC_D);   // A K-major (P), B MN-major (V)
        const uint32_t q_addr = smem_u32(sQ), p_addr = smem_u32(sP);
        mbar_wait(&bar[B_Q], 0);
        for (int j = 0; j <= n_tiles; ++j) {
            if (j < n_tiles) {  // S_j = Q K_j^T
                const int st = j & 1;
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                mbar_wait(&bar[B_SEMPTY + st], ph ^ 1);
                mbar_wait(&bar[B_KFULL + st], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t k_addr = smem_u32(sK + st * TC_TILE);
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + st * TC_BN, q_addr >> 4, off, 16, 1024, k_addr >> 4, off, 16, 1024,
                                    idesc_qk, k ? 1u : 0u);
                    }
                    umma_commit(&bar[B_KEMPTY + st]);
                    umma_commit(&bar[B_SFULL + st]);
                }
                __syncwarp();
            }
            if (j >= 1) {  // O += P_{j-1} V_{j-1}
                const int i = j - 1, vs = i % TC_VSTAGES;
                const uint32_t vph = (uint32_t)(i / TC_VSTAGES) & 1u;
                mbar_wait(&bar[B_VFULL + vs], vph);
                mbar_wait(&bar[B_PFULL], (uint32_t)i & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t v_addr = smem_u32(sV + vs * TC_TILE);
#pragma unroll
                    for (int k = 0; k < TC_BN / 16; ++k) {
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + 256, p_addr >> 4, a_off, 16, 1024, v_addr >> 4, k * 16 * 128, TC_BN * 128, 1024,
                                    idesc_pv, (i | k) ? 1u : 0u);
                    }
                    umma_commit(&bar[B_VEMPTY + vs]);
                    umma_commit(&bar[B_PVDONE]);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== softmax: thread == query row (W8: one half of its columns) =====
        const int q = warp & 3;
        const int cw = W8 ? (warp - 2) >> 2 : 0;  // column half of this warp
        const int r = q * 32 + lane;
        const int m = m0 + r;  // sequence-relative query index
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float sl2 = p.scale * kLog2eTc;
        float m_ref = 0.f, l_i = 0.f;  // m_ref: the (possibly stale) max the exponentials are taken against
        const uint32_t sP_a = smem_u32(sP);
        for (int j = 0; j < n_tiles; ++j) {
            const int st = j & 1;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            mbar_wait(&bar[B_SFULL + st], ph);
            tc_fence_after();
            uint32_t sv[NC][32];
#pragma unroll
            for (int c = 0; c < NC; ++c) tmem_ld32_nowait(lane_base + st * TC_BN + (cw * NC + c) * 32, sv[c]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_SEMPTY + st]);  // S[st] is in registers: the next QK^T may overwrite it

            const bool need_mask = (j * TC_BN + TC_BN > L) || (p.causal && j * TC_BN + TC_BN > m0);
            const int nfirst = j * TC_BN + cw * NC * 32;
            float mx = need_mask ? tile_row_max<true, NC>(sv, nfirst, m, L, p.causal)
                                 : tile_row_max<false, NC>(sv, nfirst, m, L, p.causal);
            if (W8) {  // the row's maximum over both halves
                float* slot = sHalf + (j & 1) * 256;
                slot[cw * 128 + r] = mx;
                asm volatile("bar.sync 1, 256;" ::: "memory");
                mx = fmaxf(mx, slot[(cw ^ 1) * 128 + r]);
            }
            // lazy rescale: refresh m_ref only when this tile exceeds it by more than 2^8 (or on the first tile)
            float alpha = 1.f;
            bool refresh = false;
            if (j == 0) {
                m_ref = (mx == -INFINITY) ? 0.f : mx;
            } else if (mx != -INFINITY && (mx - m_ref) * sl2 > 8.0f) {
                alpha = exp2f((m_ref - mx) * sl2);
                m_ref = mx;
                refresh = true;
            }
            // P buffer free <=> PV_{j-1} committed; the same barrier also means O holds tiles 0..j-1
            if (j >= 1) mbar_wait(&bar[B_PVDONE], (uint32_t)(j - 1) & 1u);
            if (__any_sync(0xffffffffu, refresh)) {  // warp-collective TMEM read-modify-write of (this warp's half of) O
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < NC; ++c) {
                    uint32_t ov[32];
                    tmem_ld32(lane_base + 256 + (cw * NC + c) * 32, ov);
#pragma unroll
                    for (int i = 0; i < 32; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
                    tmem_st32(lane_base + 256 + (cw * NC + c) * 32, ov);
                }
                tmem_wait_st();
                l_i *= alpha;
            }
            const float mb = m_ref * sl2;
            float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float e0 = exp2f(__uint_as_float(sv[c][i]) * sl2 - mb);      // -inf -> 0
                    const float e1 = exp2f(__uint_as_float(sv[c][i + 1]) * sl2 - mb);
                    sum4[(i >> 1) & 3] += e0 + e1;
                    pk[i >> 1] = f2_to_bf2(e0, e1);
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t addr = swz_addr(sP_a, TC_BM, r, (cw * NC + c) * 4 + g);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                                 "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
                }
            }
            l_i += (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[B_PFULL]);
        }
        mbar_wait(&bar[B_PVDONE], (uint32_t)(n_tiles - 1) & 1u);
        tc_fence_after();
        if (W8) {  // row sum = the two halves' partial sums
            float* slot = sHalf + (n_tiles & 1) * 256;
            slot[cw * 128 + r] = l_i;
            asm volatile("bar.sync 1, 256;" ::: "memory");
            l_i += slot[(cw ^ 1) * 128 + r];
        }
        const float inv = l_i > 0.f ? 1.f / l_i : 0.f;
        if (m < L && cw == 0) p.lse[(int64_t)h * p.total + s0 + m] = (l_i > 0.f) ? m_ref * p.scale + logf(l_i) : -INFINITY;
        __nv_bfloat16* orow = p.o + (int64_t)(s0 + m) * p.o_stride_tok + (int64_t)h * p.o_stride_head;
#pragma unroll 1
        for (int c = cw * NC; c < cw * NC + NC; ++c) {
            uint32_t ov[32];
            tmem_ld32(lane_base + 256 + c * 32, ov);
            if (m < L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w;
                    w.x = f2_to_bf2(__uint_as_float(ov[g * 8 + 0]) * inv, __uint_as_float(ov[g * 8 + 1]) * inv);
                    w.y = f2_to_bf2(__uint_as_float(ov[g * 8 + 2]) * inv, __uint_as_float(ov[g * 8 + 3]) * inv);
                    w.z = f2_to_bf2(__uint_as_float(ov[g * 8 + 4]) * inv, __uint_as_float(ov[g * 8 + 5]) * inv);
                    w.w = f2_to_bf2(__uint_as_float(ov[g * 8 + 6]) * inv, __uint_as_float(ov[g * 8 + 7]) * inv);
                    *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ================================================================================================
// backward on tcgen05: dQ kernel (Q-row tiles, streams K/V) and dK/dV kernel (KV-row tiles, streams
// Q/dO of every q head of the GQA group).  No atomics: same deterministic split as attention.cu.
//   dQ   : S = Q K^T, dP = dO V^T (TMEM, N=64) -> dS = P o (dP - delta) (softmax warps, bf16 -> smem)
//          dQ += dS K   (TMEM accumulator for the whole KV sweep, read once at the end)
//   dKdV : S^T = K Q^T, dP^T = V dO^T (TMEM, N=64) -> P^T, dS^T (bf16 -> smem)
//          dV += P^T dO, dK += dS^T Q   (TMEM accumulators for the whole sweep)
// ================================================================================================
struct AttnTcBwdParams {
    const int* cu_seqlens;
    int Hq, Hk, total;
    int tiles, nseq;
    float scale;
    int causal;
    const float* lse;    // [Hq, total]
    const float* delta;  // [Hq, total]
    __nv_bfloat16 *dq, *dk, *dv;
    int64_t dq_st, dq_sh, dk_st, dk_sh, dv_st, dv_sh;
    const __nv_bfloat16 *q, *dout;  // raw pointers: the dQ kernel keeps its Q / dO tile in TMEM (A operands)
    int64_t q_st, q_sh, do_st, do_sh;
    long long* trace;  // debugging: clock64 stamps of block 0's hand-offs ([event][tile], 64 tiles); nullptr in production
};

// Debug timeline of the dQ kernel (tools/attn_trace.py): events 0..5 of tile j < 64 in block 0.
constexpr int kTraceTiles = 64;
__device__ long long g_attn_trace[8 * kTraceTiles];
#define VB_TRACE(ev, j) do { if (p.trace && blockIdx.x == 0 && (j) < kTraceTiles) p.trace[(ev) * kTraceTiles + (j)] = clock64(); } while (0)

constexpr int TB_N = 64;                      // streamed tile rows
constexpr int TB_SMALL = TB_N * TC_D * 2;     // 16 KB: a [64][128] bf16 tile (two 8 KB boxes)
constexpr int TB_DS = 128 * TB_N * 2;         // 16 KB: a [128][64] bf16 tile (one box)

// K/V smem ring depth of the dQ kernel: K_j is held from S_j until dQ_j retires, and the refill is a ~1 us TMA round
// trip — 2 stages left the tensor pipe idle 75 % of the time. With the Q / dO tile in TMEM (TS) their 64 KB of smem
// go to the ring.
constexpr int TB_KV_SS = 4, TB_KV_TS = 6, TB_KV_MAX = 6;
enum { Q_LOAD = 0, Q_KFULL = 1, Q_VFULL = Q_KFULL + TB_KV_MAX, Q_KEMPTY = Q_VFULL + TB_KV_MAX, Q_VEMPTY = Q_KEMPTY + TB_KV_MAX,
       Q_SPFULL = Q_VEMPTY + TB_KV_MAX, Q_SEMPTY = Q_SPFULL + 2, Q_DSFULL = Q_SEMPTY + 2, Q_DSEMPTY = Q_DSFULL + 2,
       Q_DONE = Q_DSEMPTY + 2, Q_COUNT = Q_DONE + 1 };

// TS = true: the A operands of S = Q K^T and dP = dO V^T (the CTA's resident Q and dO tiles) live in TMEM instead of
// shared memory. The SS version moves ~176 KB through shared memory per 128x64 tile (A re-read for every tile) for
// 768 clk of MMA and is bound by the 128 B/clk shared-memory pipe (ncu: 57 % of it with the tensor pipe at 37 %);
// with A in TMEM it is ~112 KB and the N=64 MMAs run at their 32-clk floor instead of 48.
// PP = true ("ping-pong"): the two softmax warpgroups (warps 2-5 / 6-9) take alternate K/V tiles — each thread does
// all 64 columns of its row for its tiles and owns one S/dP buffer — instead of all eight warps splitting the columns
// of the same tile and then waiting together for the next S: while one warpgroup waits for its MMAs the other computes.
// P16 = true (implies TS, !PP): SIXTEEN softmax warps in two groups of eight. A group works like the eight warps of the
// default kernel (warps w and w+4 of the group share a TMEM lane quadrant and take one 32-column chunk each) but the two
// groups take alternate K/V tiles, each owning one S/dP/dS buffer: the per-tile chain wait -> tcgen05.ld -> exponentials ->
// store -> fence -> arrive (~2000 clk in the default kernel against 770-1140 clk of MMA, the softmax warps never idle) runs
// twice concurrently, with four warps per scheduler instead of two to hide its latencies.
template <bool TS, bool PP, bool P16 = false>
__global__ void __launch_bounds__(P16 ? 576 : 320, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const AttnTcBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int TB_KV = TS ? TB_KV_TS : TB_KV_SS;
    uint8_t* sQ = smem;                           // 32 KB (SS only)
    uint8_t* sdO = sQ + (TS ? 0 : TC_TILE);       // 32 KB (SS only)
    uint8_t* sK = sdO + (TS ? 0 : TC_TILE);       // TB_KV x 16 KB
    uint8_t* sV = sK + TB_KV * TB_SMALL;          // TB_KV x 16 KB
    uint8_t* sdS = sV + TB_KV * TB_SMALL;         // 2 x 16 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sdS + 2 * TB_DS);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + Q_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int order, h, seq;
    tile_of_block(p.Hq, p.nseq, order, h, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int mblk = p.tiles - 1 - order;  // causal: the last Q tile is the heaviest
    const int m0 = mblk * TC_BM;
    if (m0 >= L) return;
    const int hk = h / (p.Hq / p.Hk);
    const int kv_end = p.causal ? min(L, m0 + TC_BM) : L;
    const int n_tiles = (kv_end + TB_N - 1) / TB_N;

    if (threadIdx.x == 0) {
        for (int i = 0; i < Q_COUNT; ++i) {
            const bool by_warps = (i >= Q_SEMPTY && i < Q_SEMPTY + 2) || (i >= Q_DSFULL && i < Q_DSFULL + 2) ||
                                  (TS && i == Q_LOAD);
            mbar_init(&bar[i], by_warps ? ((PP && i != Q_LOAD) ? 4 : 8) : 1);  // P16: 8 warps per group / per barrier
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S[2] 0/64, dP[2] 128/192, dQ 256..383; TS: Q (bf16 A operand) 384..447, dO 448..511

    if (warp == 0) {
        if (elect_one_sync()) {
            if (!TS) {
                mbar_expect_tx(&bar[Q_LOAD], 2 * TC_TILE);
                for (int hf = 0; hf < 2; ++hf) {
                    tma_load_3d(sQ + hf * TC_BM * 128, &tmQ, hf * 64, h, s0 + m0, &bar[Q_LOAD]);
                    tma_load_3d(sdO + hf * TC_BM * 128, &tmdO, hf * 64, h, s0 + m0, &bar[Q_LOAD]);
                }
            }
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % TB_KV;
                const uint32_t ph = (uint32_t)(j / TB_KV) & 1u;
                mbar_wait(&bar[Q_KEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[Q_KFULL + st], TB_SMALL);
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_3d(sK + st * TB_SMALL + hf * TB_N * 128, &tmK, hf * 64, hk, s0 + j * TB_N, &bar[Q_KFULL + st]);
                mbar_wait(&bar[Q_VEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[Q_VFULL + st], TB_SMALL);
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_3d(sV + st * TB_SMALL + hf * TB_N * 128, &tmV, hf * 64, hk, s0 + j * TB_N, &bar[Q_VFULL + st]);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_nt = umma_idesc(0, 0, TC_BM, TB_N);   // S, dP: N = 64
        constexpr uint32_t idesc_dq = umma_idesc(0, 1, TC_BM, TC_D);   // dQ: A = dS (K-major), B = K (MN-major)
        const uint32_t q_addr = smem_u32(sQ), do_addr = smem_u32(sdO);
        mbar_wait(&bar[Q_LOAD], 0);
        for (int j = 0; j <= n_tiles; ++j) {
            if (j < n_tiles) {
                const int st = j & 1;                       // TMEM S/dP buffer
                const uint32_t ph = (uint32_t)(j >> 1) & 1u;
                const int ks = j % TB_KV;                   // smem K/V stage
                const uint32_t kph = (uint32_t)(j / TB_KV) & 1u;
                mbar_wait(&bar[Q_SEMPTY + st], ph ^ 1);
                mbar_wait(&bar[Q_KFULL + ks], kph);
                mbar_wait(&bar[Q_VFULL + ks], kph);
                tc_fence_after();
                if (elect_one_sync()) {
                    VB_TRACE(0, j);  // S/dP of tile j issued
                    const uint32_t k_addr = smem_u32(sK + ks * TB_SMALL), v_addr = smem_u32(sV + ks * TB_SMALL);
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        const uint32_t b_off = (uint32_t)(k >> 2) * (TB_N * 128) + (uint32_t)(k & 3) * 32;
                        if (TS)  // A = Q rows in TMEM: 16 bf16 of K per step = 8 columns
                            umma_f16_ts(tmem + st * TB_N, tmem + 384 + k * 8, k_addr >> 4, b_off, 16, 1024, idesc_nt, k ? 1u : 0u);
                        else
                            umma_f16_bo(tmem + st * TB_N, q_addr >> 4, a_off, 16, 1024, k_addr >> 4, b_off, 16, 1024,
                                        idesc_nt, k ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        const uint32_t b_off = (uint32_t)(k >> 2) * (TB_N * 128) + (uint32_t)(k & 3) * 32;
                        if (TS)
                            umma_f16_ts(tmem + 128 + st * TB_N, tmem + 448 + k * 8, v_addr >> 4, b_off, 16, 1024, idesc_nt, k ? 1u : 0u);
                        else
                            umma_f16_bo(tmem + 128 + st * TB_N, do_addr >> 4, a_off, 16, 1024, v_addr >> 4, b_off, 16, 1024,
                                        idesc_nt, k ? 1u : 0u);
                    }
                    umma_commit(&bar[Q_VEMPTY + ks]);
                    umma_commit(&bar[Q_SPFULL + st]);
                }
                __syncwarp();
            }
            if (j >= 1) {
                const int i = j - 1, st = i & 1, ks = i % TB_KV;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(&bar[Q_DSFULL + st], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    VB_TRACE(1, i);  // dS of tile i seen by the MMA warp, dQ MMAs issued
                    const uint32_t ds_addr = smem_u32(sdS + st * TB_DS), k_addr = smem_u32(sK + ks * TB_SMALL);
#pragma unroll
                    for (int k = 0; k < TB_N / 16; ++k)
                        umma_f16_bo(tmem + 256, ds_addr >> 4, k * 32, 16, 1024, k_addr >> 4, k * 16 * 128, TB_N * 128, 1024,
                                    idesc_dq, (i | k) ? 1u : 0u);
                    umma_commit(&bar[Q_KEMPTY + ks]);
                    umma_commit(&bar[Q_DSEMPTY + st]);
                    if (i == n_tiles - 1) umma_commit(&bar[Q_DONE]);
                }
                __syncwarp();
            }
        }
    } else {
        // 8 softmax warps: warps w and w+4 share TMEM lane quadrant (w & 3) and split the tile's two 32-column chunks
        static_assert(!P16 || (TS && !PP), "P16 builds on the TMEM-operand kernel and replaces PP");
        const int q = warp & 3;
        const int grp = P16 ? (warp - 2) >> 3 : 0;   // P16: softmax group (alternate tiles)
        const int cw = ((warp - 2) >> 2) & 1;        // column chunk of this warp
        const int r = q * 32 + lane;
        const int m = m0 + r;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float sl2 = p.scale * kLog2eTc;
        float lse2 = 0.f, dl = 0.f;
        if (m < L) {
            const float l = p.lse[(int64_t)h * p.total + s0 + m];
            lse2 = (l == -INFINITY) ? 0.f : l * kLog2eTc;
            dl = p.delta[(int64_t)h * p.total + s0 + m];
        }
        if (TS && grp == 0) {
            // this thread's Q row (warps 2-5) or dO row (warps 6-9): 256 contiguous bytes, global -> registers -> TMEM;
            // two bf16 per 32-bit column is exactly the K-major A-operand layout
            const __nv_bfloat16* src = cw == 0 ? p.q + (int64_t)(s0 + m) * p.q_st + (int64_t)h * p.q_sh
                                               : p.dout + (int64_t)(s0 + m) * p.do_st + (int64_t)h * p.do_sh;
            const uint32_t dst = lane_base + 384 + cw * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t w[32];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (m < L) v = *reinterpret_cast<const uint4*>(src + half * 64 + g * 8);
                    w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
                }
                tmem_st32(dst + half * 32, w);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[Q_LOAD]);
        }
        for (int j = P16 ? grp : (PP ? cw : 0); j < n_tiles; j += (PP || P16) ? 2 : 1) {
            const int st = j & 1;
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            mbar_wait(&bar[Q_SPFULL + st], ph);
            tc_fence_after();
            if (warp == 2 && lane == 0) VB_TRACE(2, j);  // S/dP of tile j complete (seen by softmax warp 2)
            mbar_wait(&bar[Q_DSEMPTY + st], ph ^ 1);  // dS[st] free (dQ_{j-2} committed)
            if (warp == 2 && lane == 0) VB_TRACE(3, j);
            const uint32_t ds_a = smem_u32(sdS + st * TB_DS);
            // tiles fully below the diagonal and inside the sequence need no per-element predicate
            const bool need_mask = (j * TB_N + TB_N > L) || (m0 + TC_BM > L) || (p.causal && j * TB_N + TB_N > m0);
#pragma unroll 1
            for (int c = PP ? 0 : cw; c < (PP ? 2 : cw + 1); ++c) {
                uint32_t sv[32], dv[32];
                tmem_ld32_nowait(lane_base + st * TB_N + c * 32, sv);
                tmem_ld32_nowait(lane_base + 128 + st * TB_N + c * 32, dv);
                tmem_wait_ld();
                if (warp == 2 && lane == 0) VB_TRACE(4, j);  // S/dP chunk in registers
                // S/dP[st] are in registers: release the TMEM buffer now so the MMA warp can run S/dP of tile j+2
                // while this tile's exponentials are still being computed
                if (!PP || c == 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[Q_SEMPTY + st]);
                }
                uint32_t pk[16];
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const int n = j * TB_N + c * 32 + i;
                        const bool ok0 = n < L && m < L && (!p.causal || n <= m);
                        const bool ok1 = n + 1 < L && m < L && (!p.causal || n + 1 <= m);
                        const float p0 = ok0 ? exp2f(__uint_as_float(sv[i]) * sl2 - lse2) : 0.f;
                        const float p1 = ok1 ? exp2f(__uint_as_float(sv[i + 1]) * sl2 - lse2) : 0.f;
                        pk[i >> 1] = f2_to_bf2(p0 * (__uint_as_float(dv[i]) - dl), p1 * (__uint_as_float(dv[i + 1]) - dl));
                    }
                } else {
                    const float2 sl2v = make_float2(sl2, sl2), nlse = make_float2(-lse2, -lse2), ndl = make_float2(-dl, -dl);
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {  // packed pairs: FFMA2, 2x MUFU, FADD2, FMUL2, F2FP
                        const float2 t = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, nlse);
                        const float2 u = fadd2(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), ndl);
                        const float2 w = fmul2(make_float2(exp2f(t.x), exp2f(t.y)), u);
                        pk[i >> 1] = f2_to_bf2(w.x, w.y);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const uint32_t addr = swz_addr(ds_a, TC_BM, r, c * 4 + g);
                    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                                 "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
                }
            }
            if (warp == 2 && lane == 0) VB_TRACE(5, j);  // dS chunk stored
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[Q_DSFULL + st]);
            if (warp == 2 && lane == 0) VB_TRACE(6, j);  // arrived
        }
        mbar_wait(&bar[Q_DONE], 0);
        tc_fence_after();
        __nv_bfloat16* row = p.dq + (int64_t)(s0 + m) * p.dq_st + (int64_t)h * p.dq_sh;
#pragma unroll 1
        for (int c = P16 ? grp * 2 + cw : cw * 2; c < (P16 ? grp * 2 + cw + 1 : cw * 2 + 2); ++c) {
            uint32_t v[32];
            tmem_ld32(lane_base + 256 + c * 32, v);
            if (m < L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w;
                    w.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * p.scale, __uint_as_float(v[g * 8 + 1]) * p.scale);
                    w.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * p.scale, __uint_as_float(v[g * 8 + 3]) * p.scale);
                    w.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * p.scale, __uint_as_float(v[g * 8 + 5]) * p.scale);
                    w.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * p.scale, __uint_as_float(v[g * 8 + 7]) * p.scale);
                    *reinterpret_cast<uint4*>(row + c * 32 + g * 8) = w;
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// ================================================================================================================
// dQ kernel with 128-row K/V tiles ("N128").
// Timeline of the 64-row kernel above (tools/attn_trace.py, profiles/r02_attn_trace.txt): 1528 clk per 128x64 tile in
// steady state against ~770 clk of tensor-core math — the 20 small MMAs per tile (16 with N = 64) plus five commits keep
// the MMA stream itself at ~1140 clk (tools/ubench_sm100.cu), and doubling the softmax warps (P16) changed nothing: the
// stage is bound by the MMA stream, not by the exponentials. This variant halves the instruction count per FLOP:
//   S = Q K^T and dP = dO V^T as 8 + 8 MMAs with N = 128 into SINGLE TMEM buffers (S cols 0..127, dP 128..255): the
//       sixteen softmax warps (four per TMEM lane quadrant, 32 columns each) pull a tile into registers right after its
//       commit and release the buffers at once, so the next tile's S/dP run while this tile's exponentials are computed;
//   dQ += dS K as 8 MMAs (K = 128) from ONE 32 KB dS buffer (the wait for its previous reader sits right before the
//       stores, i.e. after the exponentials);
//   K/V ring: 3 stages of 2 x 32 KB. One diagonal (masked) tile per CTA instead of two.
// Same arithmetic in the same order as the 64-row kernel (the TMEM accumulator sees the same k-steps): bit-identical dQ.
// ================================================================================================================
constexpr int N8_BN = 128, N8_KV = 3;
constexpr int N8_TILE = N8_BN * TC_D * 2;  // 32 KB: a [128][128] bf16 tile (two [128][64] boxes)
enum { N8_LOAD = 0, N8_KFULL = 1, N8_VFULL = N8_KFULL + N8_KV, N8_KEMPTY = N8_VFULL + N8_KV, N8_VEMPTY = N8_KEMPTY + N8_KV,
       N8_SPFULL = N8_VEMPTY + N8_KV, N8_SEMPTY, N8_DSFULL, N8_DSEMPTY, N8_DONE, N8_COUNT };

__global__ void __launch_bounds__(576, 1)
attn_bwd_dq_n128_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnTcBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sK = smem;                      // N8_KV x 32 KB
    uint8_t* sV = sK + N8_KV * N8_TILE;      // N8_KV x 32 KB
    uint8_t* sdS = sV + N8_KV * N8_TILE;     // 32 KB
    uint64_t* bar = reinterpret_cast<uint64_t*>(sdS + N8_TILE);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + N8_COUNT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int order, h, seq;
    tile_of_block(p.Hq, p.nseq, order, h, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int mblk = p.tiles - 1 - order;  // causal: the last Q tile is the heaviest
    const int m0 = mblk * TC_BM;
    if (m0 >= L) return;
    const int hk = h / (p.Hq / p.Hk);
    const int kv_end = p.causal ? min(L, m0 + TC_BM) : L;
    const int n_tiles = (kv_end + N8_BN - 1) / N8_BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < N8_COUNT; ++i) {
            const int cnt = (i == N8_SEMPTY || i == N8_DSFULL) ? 16 : (i == N8_LOAD ? 8 : 1);
            mbar_init(&bar[i], cnt);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S 0..127, dP 128..255, dQ 256..383, Q (bf16 A operand) 384..447, dO 448..511

    if (warp == 0) {
        if (elect_one_sync()) {
            for (int j = 0; j < n_tiles; ++j) {
                const int st = j % N8_KV;
                const uint32_t ph = (uint32_t)(j / N8_KV) & 1u;
                mbar_wait(&bar[N8_KEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[N8_KFULL + st], N8_TILE);
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_3d(sK + st * N8_TILE + hf * N8_BN * 128, &tmK, hf * 64, hk, s0 + j * N8_BN, &bar[N8_KFULL + st]);
                mbar_wait(&bar[N8_VEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[N8_VFULL + st], N8_TILE);
                for (int hf = 0; hf < 2; ++hf)
                    tma_load_3d(sV + st * N8_TILE + hf * N8_BN * 128, &tmV, hf * 64, hk, s0 + j * N8_BN, &bar[N8_VFULL + st]);
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_nt = umma_idesc(0, 0, TC_BM, N8_BN);  // S, dP: N = 128, both operands K-major
        constexpr uint32_t idesc_dq = umma_idesc(0, 1, TC_BM, TC_D);   // dQ: A = dS (K-major), B = K (MN-major)
        mbar_wait(&bar[N8_LOAD], 0);
        for (int j = 0; j <= n_tiles; ++j) {
            if (j < n_tiles) {
                const uint32_t ph = (uint32_t)j & 1u;
                const int ks = j % N8_KV;
                const uint32_t kph = (uint32_t)(j / N8_KV) & 1u;
                mbar_wait(&bar[N8_SEMPTY], ph ^ 1);  // tile j-1's S/dP are in the softmax warps' registers
                mbar_wait(&bar[N8_KFULL + ks], kph);
                mbar_wait(&bar[N8_VFULL + ks], kph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t k_addr = smem_u32(sK + ks * N8_TILE), v_addr = smem_u32(sV + ks * N8_TILE);
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {  // A = Q rows in TMEM: 16 bf16 of K per step = 8 columns
                        const uint32_t b_off = (uint32_t)(k >> 2) * (N8_BN * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_ts(tmem, tmem + 384 + k * 8, k_addr >> 4, b_off, 16, 1024, idesc_nt, k ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t b_off = (uint32_t)(k >> 2) * (N8_BN * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_ts(tmem + 128, tmem + 448 + k * 8, v_addr >> 4, b_off, 16, 1024, idesc_nt, k ? 1u : 0u);
                    }
                    umma_commit(&bar[N8_VEMPTY + ks]);
                    umma_commit(&bar[N8_SPFULL]);
                }
                __syncwarp();
            }
            if (j >= 1) {
                const int i = j - 1, ks = i % N8_KV;
                mbar_wait(&bar[N8_DSFULL], (uint32_t)i & 1u);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t ds_addr = smem_u32(sdS), k_addr = smem_u32(sK + ks * N8_TILE);
#pragma unroll
                    for (int k = 0; k < N8_BN / 16; ++k) {
                        // A = dS [128 q][128 kv], K-major, two [128][64] halves; B = K tile [128 kv][128 d] read MN-major:
                        // k-step k = kv rows 16k.., the two 64-wide d halves N8_BN*128 bytes apart
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + 256, ds_addr >> 4, a_off, 16, 1024, k_addr >> 4, k * 16 * 128, N8_BN * 128, 1024, idesc_dq,
                                    (i | k) ? 1u : 0u);
                    }
                    umma_commit(&bar[N8_KEMPTY + ks]);
                    umma_commit(&bar[N8_DSEMPTY]);
                    if (i == n_tiles - 1) umma_commit(&bar[N8_DONE]);
                }
                __syncwarp();
            }
        }
    } else {
        // 16 softmax warps: four per TMEM lane quadrant (warp & 3), one 32-column chunk of the 128-column tile each
        const int q = warp & 3;
        const int cw = (warp - 2) >> 2;  // 0..3
        const int r = q * 32 + lane;
        const int m = m0 + r;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float sl2 = p.scale * kLog2eTc;
        float lse2 = 0.f, dl = 0.f;
        if (m < L) {
            const float l = p.lse[(int64_t)h * p.total + s0 + m];
            lse2 = (l == -INFINITY) ? 0.f : l * kLog2eTc;
            dl = p.delta[(int64_t)h * p.total + s0 + m];
        }
        if (cw < 2) {
            // this thread's Q row (cw 0) or dO row (cw 1): 256 contiguous bytes, global -> registers -> TMEM
            const __nv_bfloat16* src = cw == 0 ? p.q + (int64_t)(s0 + m) * p.q_st + (int64_t)h * p.q_sh
                                               : p.dout + (int64_t)(s0 + m) * p.do_st + (int64_t)h * p.do_sh;
            const uint32_t dst = lane_base + 384 + cw * 64;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t w[32];
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    uint4 v = make_uint4(0u, 0u, 0u, 0u);
                    if (m < L) v = *reinterpret_cast<const uint4*>(src + half * 64 + g * 8);
                    w[g * 4] = v.x; w[g * 4 + 1] = v.y; w[g * 4 + 2] = v.z; w[g * 4 + 3] = v.w;
                }
                tmem_st32(dst + half * 32, w);
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[N8_LOAD]);
        }
        const uint32_t ds_a = smem_u32(sdS);
        for (int j = 0; j < n_tiles; ++j) {
            const uint32_t ph = (uint32_t)j & 1u;
            mbar_wait(&bar[N8_SPFULL], ph);
            tc_fence_after();
            uint32_t sv[32], dv[32];
            tmem_ld32_nowait(lane_base + cw * 32, sv);
            tmem_ld32_nowait(lane_base + 128 + cw * 32, dv);
            tmem_wait_ld();
            tc_fence_before();  // S/dP are in registers: release both buffers for tile j+1
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[N8_SEMPTY]);
            const bool need_mask = (j * N8_BN + N8_BN > L) || (m0 + TC_BM > L) || (p.causal && j * N8_BN + N8_BN > m0);
            uint32_t pk[16];
            if (need_mask) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const int n = j * N8_BN + cw * 32 + i;
                    const bool ok0 = n < L && m < L && (!p.causal || n <= m);
                    const bool ok1 = n + 1 < L && m < L && (!p.causal || n + 1 <= m);
                    const float p0 = ok0 ? exp2f(__uint_as_float(sv[i]) * sl2 - lse2) : 0.f;
                    const float p1 = ok1 ? exp2f(__uint_as_float(sv[i + 1]) * sl2 - lse2) : 0.f;
                    pk[i >> 1] = f2_to_bf2(p0 * (__uint_as_float(dv[i]) - dl), p1 * (__uint_as_float(dv[i + 1]) - dl));
                }
            } else {
                const float2 sl2v = make_float2(sl2, sl2), nlse = make_float2(-lse2, -lse2), ndl = make_float2(-dl, -dl);
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                    const float2 t = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, nlse);
                    const float2 u = fadd2(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), ndl);
                    const float2 w = fmul2(make_float2(exp2f(t.x), exp2f(t.y)), u);
                    pk[i >> 1] = f2_to_bf2(w.x, w.y);
                }
            }
            mbar_wait(&bar[N8_DSEMPTY], ph ^ 1);  // the dS buffer's previous reader (dQ_{j-1}) has retired
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                const uint32_t addr = swz_addr(ds_a, TC_BM, r, cw * 4 + g);
                asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                             "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[N8_DSFULL]);
        }
        mbar_wait(&bar[N8_DONE], 0);
        tc_fence_after();
        __nv_bfloat16* row = p.dq + (int64_t)(s0 + m) * p.dq_st + (int64_t)h * p.dq_sh;
        {
            const int c = cw;
            uint32_t v[32];
            tmem_ld32(lane_base + 256 + c * 32, v);
            if (m < L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w;
                    w.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]) * p.scale, __uint_as_float(v[g * 8 + 1]) * p.scale);
                    w.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]) * p.scale, __uint_as_float(v[g * 8 + 3]) * p.scale);
                    w.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]) * p.scale, __uint_as_float(v[g * 8 + 5]) * p.scale);
                    w.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]) * p.scale, __uint_as_float(v[g * 8 + 7]) * p.scale);
                    *reinterpret_cast<uint4*>(row + c * 32 + g * 8) = w;
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

constexpr int TB_QS_SS = 3, TB_QS_TS = 5, TB_QS_MAX = 5;  // Q/dO smem ring depth of the dK/dV kernel
enum { K_LOAD = 0, K_QFULL = 1, K_QEMPTY = K_QFULL + TB_QS_MAX, K_STFULL = K_QEMPTY + TB_QS_MAX, K_STEMPTY = K_STFULL + 2,
       K_PFULL = K_STEMPTY + 2, K_PEMPTY = K_PFULL + 2, K_DONE = K_PEMPTY + 2, K_COUNT = K_DONE + 1 };

// TS = true: P^T and dS^T — the A operands of dV += P^T dO and dK += dS^T Q — are written by the softmax warps into
// TMEM (tcgen05.st, two bf16 per column, over the S^T / dP^T buffer they were computed from) instead of shared memory.
// ncu on the SS version: the shared-memory pipe is 83 % busy (tensor-core operand reads 56 %, P/dS stores + statistics
// 27 %) with the tensor pipe at 50 %: ~270 KB cross it per 128x64 tile for 1024 clk of MMA. TS removes the 32 KB of
// stores and the 32 KB of A reads per tile, the proxy fence, and frees 64 KB for a deeper Q/dO ring. No extra
// synchronisation is needed for the aliasing: S^T_{j+2} is issued after dV/dK_j and tcgen05.mma executes in issue order.

// P16 (implies TS, !PP): sixteen softmax warps in two groups of eight on alternate jobs, as in the dQ kernel; each thread then
// works on 16-column halves of its chunk (tcgen05.ld/st .x16/.x8) so that 576 threads fit the register file.
template <bool TS, bool PP, bool P16 = false>  // PP: the two softmax warpgroups take alternate (head, q tile) jobs, see the dQ kernel
__global__ void __launch_bounds__(P16 ? 576 : 320, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                        const AttnTcBwdParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sK = smem;                    // 32 KB
    uint8_t* sV = sK + TC_TILE;            // 32 KB
    constexpr int TB_QS = TS ? TB_QS_TS : TB_QS_SS;
    uint8_t* sQ = sV + TC_TILE;            // TB_QS x 16 KB
    uint8_t* sdO = sQ + TB_QS * TB_SMALL;  // TB_QS x 16 KB
    uint8_t* sPt = sdO + TB_QS * TB_SMALL; // 2 x 16 KB (SS only)
    uint8_t* sdSt = sPt + (TS ? 0 : 2 * TB_DS);       // 2 x 16 KB (SS only)
    uint64_t* bar = reinterpret_cast<uint64_t*>(sdSt + (TS ? 0 : 2 * TB_DS));
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + K_COUNT);
    float* sStat = reinterpret_cast<float*>(tmem_slot + 4);  // [2 stages (PP: 2 warpgroups x 2)][lse2 x64 | delta x64]
    static_assert(!PP || TS, "the ping-pong variant is built on the TMEM-operand kernel");
    static_assert(!P16 || (TS && !PP), "P16 builds on the TMEM-operand kernel and replaces PP");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int order, hk, seq;
    tile_of_block(p.Hk, p.nseq, order, hk, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int n0 = order * TC_BM;  // causal: the first KV tile is the heaviest
    if (n0 >= L) return;
    const int G = p.Hq / p.Hk;
    const int i_start = p.causal ? n0 / TB_N : 0;
    const int nq = (L + TB_N - 1) / TB_N - i_start;
    const int jobs = nq * G;

    if (threadIdx.x == 0) {
        for (int i = 0; i < K_COUNT; ++i) {
            const bool by_warps = (i >= K_STEMPTY && i < K_STEMPTY + 2) || (i >= K_PFULL && i < K_PFULL + 2);
            mbar_init(&bar[i], by_warps ? (PP ? 4 : 8) : 1);
        }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    // TMEM columns: S^T[2] 0/64, dP^T[2] 128/192, dV 256..383, dK 384..511

    if (warp == 0) {
        if (elect_one_sync()) {
            mbar_expect_tx(&bar[K_LOAD], 2 * TC_TILE);
            for (int hf = 0; hf < 2; ++hf) {
                tma_load_3d(sK + hf * TC_BM * 128, &tmK, hf * 64, hk, s0 + n0, &bar[K_LOAD]);
                tma_load_3d(sV + hf * TC_BM * 128, &tmV, hf * 64, hk, s0 + n0, &bar[K_LOAD]);
            }
            for (int jb = 0; jb < jobs; ++jb) {
                const int st = jb % TB_QS;
                const uint32_t ph = (uint32_t)(jb / TB_QS) & 1u;
                const int hh = hk * G + jb / nq, qi = i_start + jb % nq;
                mbar_wait(&bar[K_QEMPTY + st], ph ^ 1);
                mbar_expect_tx(&bar[K_QFULL + st], 2 * TB_SMALL);
                for (int hf = 0; hf < 2; ++hf) {
                    tma_load_3d(sQ + st * TB_SMALL + hf * TB_N * 128, &tmQ, hf * 64, hh, s0 + qi * TB_N, &bar[K_QFULL + st]);
                    tma_load_3d(sdO + st * TB_SMALL + hf * TB_N * 128, &tmdO, hf * 64, hh, s0 + qi * TB_N, &bar[K_QFULL + st]);
                }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc_nt = umma_idesc(0, 0, TC_BM, TB_N);
        constexpr uint32_t idesc_acc = umma_idesc(0, 1, TC_BM, TC_D);
        const uint32_t k_addr = smem_u32(sK), v_addr = smem_u32(sV);
        mbar_wait(&bar[K_LOAD], 0);
        for (int jb = 0; jb <= jobs; ++jb) {
            if (jb < jobs) {
                const int st = jb & 1;                       // TMEM S^T/dP^T buffer
                const uint32_t ph = (uint32_t)(jb >> 1) & 1u;
                const int qs = jb % TB_QS;                   // smem Q/dO stage
                const uint32_t qph = (uint32_t)(jb / TB_QS) & 1u;
                if (!TS) mbar_wait(&bar[K_STEMPTY + st], ph ^ 1);  // TS: implied by PFULL_{jb-2} + in-order MMA execution
                mbar_wait(&bar[K_QFULL + qs], qph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t q_addr = smem_u32(sQ + qs * TB_SMALL), do_addr = smem_u32(sdO + qs * TB_SMALL);
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        const uint32_t b_off = (uint32_t)(k >> 2) * (TB_N * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + st * TB_N, k_addr >> 4, a_off, 16, 1024, q_addr >> 4, b_off, 16, 1024,
                                    idesc_nt, k ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < TC_D / 16; ++k) {
                        const uint32_t a_off = (uint32_t)(k >> 2) * (TC_BM * 128) + (uint32_t)(k & 3) * 32;
                        const uint32_t b_off = (uint32_t)(k >> 2) * (TB_N * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + 128 + st * TB_N, v_addr >> 4, a_off, 16, 1024, do_addr >> 4, b_off, 16, 1024,
                                    idesc_nt, k ? 1u : 0u);
                    }
                    umma_commit(&bar[K_STFULL + st]);
                }
                __syncwarp();
            }
            if (jb >= 1) {
                const int i = jb - 1, st = i & 1, qs = i % TB_QS;
                const uint32_t ph = (uint32_t)(i >> 1) & 1u;
                mbar_wait(&bar[K_PFULL + st], ph);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint32_t pt_addr = smem_u32(sPt + st * TB_DS), dst_addr = smem_u32(sdSt + st * TB_DS);
                    const uint32_t q_addr = smem_u32(sQ + qs * TB_SMALL), do_addr = smem_u32(sdO + qs * TB_SMALL);
#pragma unroll
                    for (int k = 0; k < TB_N / 16; ++k) {
                        if (TS)  // A = P^T in TMEM over S^T[st]: 16 q (K) per step = 8 columns; q 32..63 start at column 32
                            umma_f16_ts(tmem + 256, tmem + st * TB_N + (k >> 1) * 32 + (k & 1) * 8, do_addr >> 4, k * 16 * 128, TB_N * 128, 1024, idesc_acc,
                                        (i | k) ? 1u : 0u);
                        else
                            umma_f16_bo(tmem + 256, pt_addr >> 4, k * 32, 16, 1024, do_addr >> 4, k * 16 * 128, TB_N * 128, 1024,
                                        idesc_acc, (i | k) ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < TB_N / 16; ++k) {
                        if (TS)
                            umma_f16_ts(tmem + 384, tmem + 128 + st * TB_N + (k >> 1) * 32 + (k & 1) * 8, q_addr >> 4, k * 16 * 128, TB_N * 128, 1024, idesc_acc,
                                        (i | k) ? 1u : 0u);
                        else
                            umma_f16_bo(tmem + 384, dst_addr >> 4, k * 32, 16, 1024, q_addr >> 4, k * 16 * 128, TB_N * 128, 1024,
                                        idesc_acc, (i | k) ? 1u : 0u);
                    }
                    umma_commit(&bar[K_QEMPTY + qs]);
                    umma_commit(&bar[K_PEMPTY + st]);
                    if (i == jobs - 1) umma_commit(&bar[K_DONE]);
                }
                __syncwarp();
            }
        }
    } else {
        // 8 softmax warps: warps w and w+4 share TMEM lane quadrant (w & 3) and split the tile's two 32-column chunks
        const int q = warp & 3;
        const int grp = P16 ? (warp - 2) >> 3 : 0;  // P16: softmax group (alternate jobs)
        const int cw = ((warp - 2) >> 2) & 1;
        const int r = q * 32 + lane;
        const int tid = PP ? ((warp - 2) & 3) * 32 + lane   // 0..127 inside this warpgroup
                           : ((warp - 2) & 7) * 32 + lane;  // 0..255 over the (group's) eight softmax warps
        const int n = n0 + r;  // kv index of this thread's row
        // per-column statistics of the first job are fetched up front; each later job's are prefetched one job ahead
        auto load_stat = [&](int jb) -> float {
            if (jb >= jobs || tid >= 128) return 0.f;
            const int hh = hk * G + jb / nq, qi = i_start + jb % nq;
            const int mm = qi * TB_N + (tid & 63);
            if (mm >= L) return 0.f;
            // stored negated: the strip holds -lse*log2(e) and -delta so the packed FMA / ADD take them as addends
            if (tid < 64) {
                const float l = p.lse[(int64_t)hh * p.total + s0 + mm];
                return (l == -INFINITY) ? 0.f : -l * kLog2eTc;
            }
            return -p.delta[(int64_t)hh * p.total + s0 + mm];
        };
        float stat_next = load_stat(P16 ? grp : (PP ? cw : 0));
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        const float sl2 = p.scale * kLog2eTc;
        for (int jb = P16 ? grp : (PP ? cw : 0); jb < jobs; jb += (PP || P16) ? 2 : 1) {
            const int st = jb & 1;
            const uint32_t ph = (uint32_t)(jb >> 1) & 1u;
            const int qi = i_start + jb % nq;
            // per-column (q index) softmax statistics of this q tile: 64 lse2 + 64 delta values through smem
            // (PP: a strip pair per warpgroup, alternating with the warpgroup's job count)
            float* strip = sStat + (P16 ? (grp * 2 + (int)ph) : PP ? (cw * 2 + (int)ph) : st) * 128;
            if (tid < 128) strip[tid] = stat_next;
            if (PP) asm volatile("bar.sync %0, 128;" ::"r"(1 + cw) : "memory");  // this warpgroup only
            else if (P16) asm volatile("bar.sync %0, 256;" ::"r"(1 + grp) : "memory");  // this group's eight warps
            else asm volatile("bar.sync 1, 256;" ::: "memory");                  // the 8 softmax warps only
            stat_next = load_stat(jb + ((PP || P16) ? 2 : 1));
            const float* lse_s = strip;
            const float* dl_s = lse_s + 64;
            mbar_wait(&bar[K_STFULL + st], ph);
            tc_fence_after();
            if (!TS) mbar_wait(&bar[K_PEMPTY + st], ph ^ 1);
            const uint32_t pt_a = smem_u32(sPt + st * TB_DS), dst_a = smem_u32(sdSt + st * TB_DS);
            const bool need_mask = (qi * TB_N + TB_N > L) || (n0 + TC_BM > L) || (p.causal && n0 + TC_BM > qi * TB_N);
            if (P16) {
                // two 16-column halves of this warp's 32-column chunk; P^T / dS^T go back over columns the chunk has
                // already given up (packed: half h lands on columns [c*32 + 8h, +8), all of them read in half 0)
                const int c = cw;
#pragma unroll 1
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t sv[16], dv[16], pk[8], dk[8];
                    tmem_ld16_nowait(lane_base + st * TB_N + c * 32 + hf * 16, sv);
                    tmem_ld16_nowait(lane_base + 128 + st * TB_N + c * 32 + hf * 16, dv);
                    tmem_wait_ld();
                    const int c0 = c * 32 + hf * 16;
                    if (need_mask) {
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            float pe[2], de[2];
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int m = qi * TB_N + c0 + i + e;  // q index (column)
                                const bool ok = n < L && m < L && (!p.causal || n <= m);
                                pe[e] = ok ? exp2f(__uint_as_float(sv[i + e]) * sl2 + lse_s[c0 + i + e]) : 0.f;
                                de[e] = pe[e] * (__uint_as_float(dv[i + e]) + dl_s[c0 + i + e]);
                            }
                            pk[i >> 1] = f2_to_bf2(pe[0], pe[1]);
                            dk[i >> 1] = f2_to_bf2(de[0], de[1]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 l4 = *reinterpret_cast<const float4*>(lse_s + c0 + i);
                            const float4 d4 = *reinterpret_cast<const float4*>(dl_s + c0 + i);
                            const float2 sl2v = make_float2(sl2, sl2);
                            const float2 t0 = ffma2(make_float2(__uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1])), sl2v, make_float2(l4.x, l4.y));
                            const float2 t1 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, make_float2(l4.z, l4.w));
                            const float2 e0 = make_float2(exp2f(t0.x), exp2f(t0.y)), e1 = make_float2(exp2f(t1.x), exp2f(t1.y));
                            const float2 u0 = fadd2(make_float2(__uint_as_float(dv[i + 0]), __uint_as_float(dv[i + 1])), make_float2(d4.x, d4.y));
                            const float2 u1 = fadd2(make_float2(__uint_as_float(dv[i + 2]), __uint_as_float(dv[i + 3])), make_float2(d4.z, d4.w));
                            const float2 w0 = fmul2(e0, u0), w1 = fmul2(e1, u1);
                            pk[i >> 1] = f2_to_bf2(e0.x, e0.y);
                            pk[(i >> 1) + 1] = f2_to_bf2(e1.x, e1.y);
                            dk[i >> 1] = f2_to_bf2(w0.x, w0.y);
                            dk[(i >> 1) + 1] = f2_to_bf2(w1.x, w1.y);
                        }
                    }
                    tmem_st8(lane_base + st * TB_N + c * 32 + hf * 8, pk);
                    tmem_st8(lane_base + 128 + st * TB_N + c * 32 + hf * 8, dk);
                }
                tmem_wait_st();
            } else
#pragma unroll 1
            for (int c = PP ? 0 : cw; c < (PP ? 2 : cw + 1); ++c) {
                uint32_t sv[32], dv[32];
                tmem_ld32_nowait(lane_base + st * TB_N + c * 32, sv);
                tmem_ld32_nowait(lane_base + 128 + st * TB_N + c * 32, dv);
                tmem_wait_ld();
                if (!TS) {
                    tc_fence_before();  // early release of S^T/dP^T[st] (see the dQ kernel)
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar[K_STEMPTY + st]);
                }
                uint32_t pk[16], dk[16];
                if (need_mask) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        float pe[2], de[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int m = qi * TB_N + c * 32 + i + e;  // q index (column)
                            const bool ok = n < L && m < L && (!p.causal || n <= m);
                            pe[e] = ok ? exp2f(__uint_as_float(sv[i + e]) * sl2 + lse_s[c * 32 + i + e]) : 0.f;
                            de[e] = pe[e] * (__uint_as_float(dv[i + e]) + dl_s[c * 32 + i + e]);
                        }
                        pk[i >> 1] = f2_to_bf2(pe[0], pe[1]);
                        dk[i >> 1] = f2_to_bf2(de[0], de[1]);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const float4 l4 = *reinterpret_cast<const float4*>(lse_s + c * 32 + i);  // -lse2 of 4 q columns
                        const float4 d4 = *reinterpret_cast<const float4*>(dl_s + c * 32 + i);   // -delta
                        const float2 sl2v = make_float2(sl2, sl2);
                        const float2 t0 = ffma2(make_float2(__uint_as_float(sv[i + 0]), __uint_as_float(sv[i + 1])), sl2v, make_float2(l4.x, l4.y));
                        const float2 t1 = ffma2(make_float2(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])), sl2v, make_float2(l4.z, l4.w));
                        const float2 e0 = make_float2(exp2f(t0.x), exp2f(t0.y)), e1 = make_float2(exp2f(t1.x), exp2f(t1.y));
                        const float2 u0 = fadd2(make_float2(__uint_as_float(dv[i + 0]), __uint_as_float(dv[i + 1])), make_float2(d4.x, d4.y));
                        const float2 u1 = fadd2(make_float2(__uint_as_float(dv[i + 2]), __uint_as_float(dv[i + 3])), make_float2(d4.z, d4.w));
                        const float2 w0 = fmul2(e0, u0), w1 = fmul2(e1, u1);
                        pk[i >> 1] = f2_to_bf2(e0.x, e0.y);
                        pk[(i >> 1) + 1] = f2_to_bf2(e1.x, e1.y);
                        dk[i >> 1] = f2_to_bf2(w0.x, w0.y);
                        dk[(i >> 1) + 1] = f2_to_bf2(w1.x, w1.y);
                    }
                }
                if (TS) {
                    // P^T / dS^T rows of this thread for q columns [c*32, c*32+32): 16 packed columns, written over the first
                    // half of the 32 S^T / dP^T columns THIS warp just read (the other half-tile belongs to warp w+-4)
                    tmem_st16(lane_base + st * TB_N + c * 32, pk);
                    tmem_st16(lane_base + 128 + st * TB_N + c * 32, dk);
                    tmem_wait_st();
                } else {
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const uint32_t a1 = swz_addr(pt_a, TC_BM, r, c * 4 + g), a2 = swz_addr(dst_a, TC_BM, r, c * 4 + g);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a1), "r"(pk[g * 4]), "r"(pk[g * 4 + 1]),
                                     "r"(pk[g * 4 + 2]), "r"(pk[g * 4 + 3]) : "memory");
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a2), "r"(dk[g * 4]), "r"(dk[g * 4 + 1]),
                                     "r"(dk[g * 4 + 2]), "r"(dk[g * 4 + 3]) : "memory");
                    }
                }
            }
            tc_fence_before();
            if (!TS) fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar[K_PFULL + st]);
        }
        mbar_wait(&bar[K_DONE], 0);
        tc_fence_after();
        __nv_bfloat16* vrow = p.dv + (int64_t)(s0 + n) * p.dv_st + (int64_t)hk * p.dv_sh;
        __nv_bfloat16* krow = p.dk + (int64_t)(s0 + n) * p.dk_st + (int64_t)hk * p.dk_sh;
#pragma unroll 1
        for (int c = P16 ? grp * 2 + cw : cw * 2; c < (P16 ? grp * 2 + cw + 1 : cw * 2 + 2); ++c) {
            uint32_t v[32], kk[32];
            tmem_ld32(lane_base + 256 + c * 32, v);
            tmem_ld32(lane_base + 384 + c * 32, kk);
            if (n < L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint4 w, x;
                    w.x = f2_to_bf2(__uint_as_float(v[g * 8 + 0]), __uint_as_float(v[g * 8 + 1]));
                    w.y = f2_to_bf2(__uint_as_float(v[g * 8 + 2]), __uint_as_float(v[g * 8 + 3]));
                    w.z = f2_to_bf2(__uint_as_float(v[g * 8 + 4]), __uint_as_float(v[g * 8 + 5]));
                    w.w = f2_to_bf2(__uint_as_float(v[g * 8 + 6]), __uint_as_float(v[g * 8 + 7]));
                    x.x = f2_to_bf2(__uint_as_float(kk[g * 8 + 0]) * p.scale, __uint_as_float(kk[g * 8 + 1]) * p.scale);
                    x.y = f2_to_bf2(__uint_as_float(kk[g * 8 + 2]) * p.scale, __uint_as_float(kk[g * 8 + 3]) * p.scale);
                    x.z = f2_to_bf2(__uint_as_float(kk[g * 8 + 4]) * p.scale, __uint_as_float(kk[g * 8 + 5]) * p.scale);
                    x.w = f2_to_bf2(__uint_as_float(kk[g * 8 + 6]) * p.scale, __uint_as_float(kk[g * 8 + 7]) * p.scale);
                    *reinterpret_cast<uint4*>(vrow + c * 32 + g * 8) = w;
                    *reinterpret_cast<uint4*>(krow + c * 32 + g * 8) = x;
                }
            }
        }
        tc_fence_before();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_attn_varlen_fwd_tc(const void* q, const void* k, const void* v, void* o, float* lse,
                                        const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int32_t total,
                                        int32_t q_heads, int32_t k_heads, int32_t head_dim, const int64_t* st, float scale,
                                        int32_t causal, void* stream) {
    const bool w8 = (causal >> 8) & 1;  // bit 8: eight softmax warps (experimental, see attn_fwd_tc_kernel)
    causal &= 1;
    if (head_dim != 128) return vb200_set_error(VB200_EINVAL, "attn_fwd_tc: head_dim must be 128");
    if (q_heads <= 0 || k_heads <= 0 || q_heads % k_heads) return vb200_set_error(VB200_EINVAL, "attn_fwd_tc: Hq % Hk != 0");
    for (int i = 0; i < 8; ++i)
        if (st[i] & 7) return vb200_set_error(VB200_EINVAL, "attn_fwd_tc: strides must be multiples of 8 elements");
    if (total <= 0 || num_seqs <= 0 || max_seqlen <= 0) return VB200_OK;
    CUtensorMap tmQ, tmK, tmV;
    int rc;
    if ((rc = make_tmap_3d(&tmQ, q, 128, q_heads, total, st[1], st[0], 128))) return rc;
    if ((rc = make_tmap_3d(&tmK, k, 128, k_heads, total, st[3], st[2], 128))) return rc;
    if ((rc = make_tmap_3d(&tmV, v, 128, k_heads, total, st[5], st[4], 128))) return rc;
    AttnTcParams p{};
    p.cu_seqlens = cu_seqlens; p.Hq = q_heads; p.Hk = k_heads; p.total = total; p.scale = scale; p.causal = causal;
    p.o = (__nv_bfloat16*)o; p.o_stride_tok = st[6]; p.o_stride_head = st[7]; p.lse = lse;
    const size_t smem = (4 + TC_VSTAGES) * TC_TILE + B_COUNT * 8 + 16 + 2 * 256 * 4 + 64;  // + the W8 half-row exchange strip
    static bool attr = false;
    if (!attr) {
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    p.tiles = (max_seqlen + TC_BM - 1) / TC_BM;
    p.nseq = num_seqs;
    dim3 grid(p.tiles * q_heads * num_seqs);
    if (w8) attn_fwd_tc_kernel<true><<<grid, 320, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
    else attn_fwd_tc_kernel<false><<<grid, 192, smem, (cudaStream_t)stream>>>(tmQ, tmK, tmV, p);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

// dq/dk/dv through the tcgen05 kernels; `delta` must already hold rowsum(dO*O) (vb200_attn_varlen_bwd computes it
// the same way: call vb200_attn_bwd_delta first).
extern "C" int vb200_attn_varlen_bwd_tc(const void* q, const void* k, const void* v, const void* dout, const float* lse,
                                        const float* delta, void* dq, void* dk, void* dv, const int32_t* cu_seqlens,
                                        int32_t num_seqs, int32_t max_seqlen, int32_t total, int32_t q_heads,
                                        int32_t k_heads, int32_t head_dim, const int64_t* st, float scale, int32_t causal,
                                        void* stream) {
    // st: (tok, head) strides of q, k, v, dout, dq, dk, dv.  causal bit 0 = causal; bits 8/9 = run only the
    // dQ / only the dK-dV kernel (used to time the two kernels separately; 0 = both); bit 10 = SS-operand dQ kernel.
    const int only = (causal >> 8) & 3;
    const bool ss_operands = (causal >> 10) & 1;  // bit 10: all MMA operands from shared memory (cross-check variant)
    const bool pingpong = (causal >> 11) & 1;     // bit 11: softmax warpgroups on alternate tiles
    const bool p16 = (causal >> 12) & 1;          // bit 12: sixteen softmax warps, two groups of eight on alternate tiles
    const bool trace = (causal >> 13) & 1;        // bit 13: debug timeline of block 0 of the dQ kernel (vb200_attn_debug_trace)
    const bool dq_n128 = (causal >> 14) & 1;      // bit 14: dQ kernel with 128-row K/V tiles (attn_bwd_dq_n128_kernel)
    causal &= 1;
    if (head_dim != 128) return vb200_set_error(VB200_EINVAL, "attn_bwd_tc: head_dim must be 128");
    if (q_heads <= 0 || k_heads <= 0 || q_heads % k_heads) return vb200_set_error(VB200_EINVAL, "attn_bwd_tc: Hq % Hk != 0");
    for (int i = 0; i < 14; ++i)
        if (st[i] & 7) return vb200_set_error(VB200_EINVAL, "attn_bwd_tc: strides must be multiples of 8 elements");
    if (total <= 0 || num_seqs <= 0 || max_seqlen <= 0) return VB200_OK;
    CUtensorMap tmQ128, tmdO128, tmK64, tmV64, tmK128, tmV128, tmQ64, tmdO64;
    int rc;
    if ((rc = make_tmap_3d(&tmQ128, q, 128, q_heads, total, st[1], st[0], 128))) return rc;
    if ((rc = make_tmap_3d(&tmdO128, dout, 128, q_heads, total, st[7], st[6], 128))) return rc;
    if ((rc = make_tmap_3d(&tmK64, k, 128, k_heads, total, st[3], st[2], 64))) return rc;
    if ((rc = make_tmap_3d(&tmV64, v, 128, k_heads, total, st[5], st[4], 64))) return rc;
    if ((rc = make_tmap_3d(&tmK128, k, 128, k_heads, total, st[3], st[2], 128))) return rc;
    if ((rc = make_tmap_3d(&tmV128, v, 128, k_heads, total, st[5], st[4], 128))) return rc;
    if ((rc = make_tmap_3d(&tmQ64, q, 128, q_heads, total, st[1], st[0], 64))) return rc;
    if ((rc = make_tmap_3d(&tmdO64, dout, 128, q_heads, total, st[7], st[6], 64))) return rc;
    AttnTcBwdParams p{};
    p.cu_seqlens = cu_seqlens; p.Hq = q_heads; p.Hk = k_heads; p.total = total; p.scale = scale; p.causal = causal;
    p.lse = lse; p.delta = delta;
    p.dq = (__nv_bfloat16*)dq; p.dq_st = st[8]; p.dq_sh = st[9];
    p.dk = (__nv_bfloat16*)dk; p.dk_st = st[10]; p.dk_sh = st[11];
    p.dv = (__nv_bfloat16*)dv; p.dv_st = st[12]; p.dv_sh = st[13];
    p.q = (const __nv_bfloat16*)q; p.q_st = st[0]; p.q_sh = st[1];
    p.dout = (const __nv_bfloat16*)dout; p.do_st = st[6]; p.do_sh = st[7];
    p.trace = nullptr;
    if (trace) {
        void* sym = nullptr;
        VB_CUDA_TRY(cudaGetSymbolAddress(&sym, g_attn_trace));
        VB_CUDA_TRY(cudaMemsetAsync(sym, 0, sizeof(long long) * 8 * kTraceTiles, (cudaStream_t)stream));
        p.trace = (long long*)sym;
    }
    const size_t smem_dq = 2 * TC_TILE + 2 * TB_KV_SS * TB_SMALL + 2 * TB_DS + Q_COUNT * 8 + 16 + 64;
    const size_t smem_dq_ts = 2 * TB_KV_TS * TB_SMALL + 2 * TB_DS + Q_COUNT * 8 + 16 + 64;
    const size_t smem_dq_n128 = (2 * N8_KV + 1) * (size_t)N8_TILE + N8_COUNT * 8 + 16 + 64;
    const size_t smem_kv = 2 * TC_TILE + 2 * TB_QS_SS * TB_SMALL + 4 * TB_DS + K_COUNT * 8 + 16 + 2 * 128 * 4 + 64;
    const size_t smem_kv_ts = 2 * TC_TILE + 2 * TB_QS_TS * TB_SMALL + K_COUNT * 8 + 16 + 4 * 128 * 4 + 64;
    static bool attr = false;
    if (!attr) {
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_tc_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv_ts));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_n128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq_n128));
        attr = true;
    }
    cudaStream_t s = (cudaStream_t)stream;
    p.tiles = (max_seqlen + TC_BM - 1) / TC_BM;
    p.nseq = num_seqs;
    dim3 gq(p.tiles * q_heads * num_seqs);
    if (only != 2) {
        if (dq_n128) attn_bwd_dq_n128_kernel<<<gq, 576, smem_dq_n128, s>>>(tmK128, tmV128, p);
        else if (ss_operands) attn_bwd_dq_tc_kernel<false, false><<<gq, 320, smem_dq, s>>>(tmQ128, tmdO128, tmK64, tmV64, p);
        else if (p16) attn_bwd_dq_tc_kernel<true, false, true><<<gq, 576, smem_dq_ts, s>>>(tmQ128, tmdO128, tmK64, tmV64, p);
        else if (pingpong) attn_bwd_dq_tc_kernel<true, true><<<gq, 320, smem_dq_ts, s>>>(tmQ128, tmdO128, tmK64, tmV64, p);
        else attn_bwd_dq_tc_kernel<true, false><<<gq, 320, smem_dq_ts, s>>>(tmQ128, tmdO128, tmK64, tmV64, p);
        vb200_count_launch(1);
        VB_HOST_CHECK_LAUNCH();
    }
    dim3 gk(p.tiles * k_heads * num_seqs);
    if (only != 1) {
        if (ss_operands) attn_bwd_dkdv_tc_kernel<false, false><<<gk, 320, smem_kv, s>>>(tmK128, tmV128, tmQ64, tmdO64, p);
        else if (p16) attn_bwd_dkdv_tc_kernel<true, false, true><<<gk, 576, smem_kv_ts, s>>>(tmK128, tmV128, tmQ64, tmdO64, p);
        else if (pingpong) attn_bwd_dkdv_tc_kernel<true, true><<<gk, 320, smem_kv_ts, s>>>(tmK128, tmV128, tmQ64, tmdO64, p);
        else attn_bwd_dkdv_tc_kernel<true, false><<<gk, 320, smem_kv_ts, s>>>(tmK128, tmV128, tmQ64, tmdO64, p);
        vb200_count_launch(1);
    }
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

// Debugging: copy the last dQ-kernel timeline (8 events x 64 tiles of clock64 stamps, block 0) to host memory.
extern "C" int vb200_attn_debug_trace(int64_t* out512) {
    if (!out512) return vb200_set_error(VB200_EINVAL, "attn_debug_trace: null output");
    VB_CUDA_TRY(cudaDeviceSynchronize());
    VB_CUDA_TRY(cudaMemcpyFromSymbol(out512, g_attn_trace, sizeof(long long) * 8 * kTraceTiles));
    return VB200_OK;
}
