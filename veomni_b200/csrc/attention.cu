// Packed (varlen) causal attention forward / backward for sm_100a.
//
// Reference call site: flash_attention_forward -> _flash_attention_forward -> flash_attn_varlen_func
//   (veomni/ops/kernels/attention/__init__.py:304-320; HF modeling_flash_attention_utils padding-free
//   branch): q [T,Hq,D], k/v [T,Hkv,D] packed over sequences delimited by cu_seqlens (int32), causal
//   within each sequence, GQA by head replication, softmax in fp32, output bf16.
//
// Design (one CTA = 8 warps, each warp owns 16 rows of the CTA's row tile):
//   * all operand tiles are staged by the TMA engine (cp.async.bulk.tensor, SWIZZLE_128B) into
//     shared memory and signalled through mbarriers; K/V (fwd, dQ) or Q/dO (dK/dV) tiles are
//     double-buffered so the next tile streams in while the current one is consumed;
//   * matmuls use bf16 mma.sync m16n8k16 fragments fed by ldmatrix from the swizzled tiles, fp32
//     accumulation, exp2-based online softmax with warp-quad shuffles for the row reductions;
//   * backward is split into a dQ kernel (row tiles of Q) and a dK/dV kernel (row tiles of K/V,
//     looping over the q heads of the GQA group) so that no atomics are needed: results are
//     deterministic and bit-reproducible, which the reference's e2e tests rely on
//     (tests/tools/training_utils.py:238-240 enable_full_determinism).
// Roofline: tensor-pipe bound; algorithmic FLOPs = 4*sum(L_i^2)/2*D*Hq forward (causal),
// 2.5x that backward; the split backward executes 3.5x (S and dP recomputed).
#include "tma.cuh"

namespace vb {

constexpr float kLog2e = 1.4426950408889634f;

struct AttnParams {
    const int* cu_seqlens;
    int num_seqs, Hq, Hk, total;
    int tiles;  // 1-D grid = tiles x heads x num_seqs, ordered heaviest tile first (attention_tc.cu: tile_of_block)
    float scale;
    int causal;
    // forward
    __nv_bfloat16* o;
    int64_t o_stride_tok, o_stride_head;
    float* lse;  // [Hq, total]
    // backward
    const float* delta;  // [Hq, total]
    __nv_bfloat16 *dq, *dk, *dv;
    int64_t dq_stride_tok, dq_stride_head, dk_stride_tok, dk_stride_head, dv_stride_tok, dv_stride_head;
};

// linear block id -> (position in the heavy-to-light tile order, head, sequence); see attention_tc.cu
__device__ __forceinline__ void tile_of_block(int heads, int nseq, int& order, int& head, int& seq) {
    const int per = heads * nseq;
    order = (int)blockIdx.x / per;
    const int rem = (int)blockIdx.x - order * per;
    seq = rem / heads;
    head = rem - seq * heads;
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Stage a warp's [16 x HD] fp32 accumulator as bf16 into its rows of a swizzled tile, then write the
// valid rows to global memory with 16-byte stores.
template <int HD>
__device__ __forceinline__ void store_rows(float (&acc)[HD / 8][4], uint32_t tile_base, uint8_t* tile_ptr,
                                           int tile_rows, int warp_r0, int lane, __nv_bfloat16* gbase,
                                           int64_t stride_tok, int valid_rows /* rows of this warp that exist */) {
#pragma unroll
    for (int nb = 0; nb < HD / 8; ++nb) {
        const uint32_t lo = swz_addr(tile_base, tile_rows, warp_r0 + (lane >> 2), nb) + (lane & 3) * 4;
        const uint32_t hi = swz_addr(tile_base, tile_rows, warp_r0 + (lane >> 2) + 8, nb) + (lane & 3) * 4;
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(lo), "r"(f2_to_bf2(acc[nb][0], acc[nb][1])) : "memory");
        asm volatile("st.shared.b32 [%0], %1;" ::"r"(hi), "r"(f2_to_bf2(acc[nb][2], acc[nb][3])) : "memory");
    }
    __syncwarp();
    constexpr int CH = HD / 8;
    for (int idx = lane; idx < 16 * CH; idx += 32) {
        const int r = idx / CH, c = idx % CH;
        if (r < valid_rows) {
            const uint32_t off = swz_addr(0, tile_rows, warp_r0 + r, c);
            const uint4 v = *reinterpret_cast<const uint4*>(tile_ptr + off);
            *reinterpret_cast<uint4*>(gbase + (int64_t)r * stride_tok + c * 8) = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    constexpr int BM = 128, BN = 64, NH = HD / 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;                         // BM*HD*2
    uint8_t* sK = sQ + BM * HD * 2;             // 2 stages * BN*HD*2
    uint8_t* sV = sK + 2 * BN * HD * 2;         // 2 stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * BN * HD * 2);  // barQ, fullK[2], fullV[2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    int order, h, seq;
    tile_of_block(p.Hq, p.num_seqs, order, h, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int mblk = p.tiles - 1 - order;  // heaviest (latest) row tiles first
    const int m0 = mblk * BM;
    if (m0 >= L) return;
    const int hk = h / (p.Hq / p.Hk);
    const int kv_end = p.causal ? min(L, m0 + BM) : L;
    const int n_tiles = (kv_end + BN - 1) / BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], BM * HD * 2);
        for (int hf = 0; hf < NH; ++hf) tma_load_3d(sQ + hf * BM * 128, &tmQ, hf * 64, h, s0 + m0, &bars[0]);
        for (int st = 0; st < 2 && st < n_tiles; ++st) {
            mbar_expect_tx(&bars[1 + st], BN * HD * 2);
            mbar_expect_tx(&bars[3 + st], BN * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sK + st * BN * HD * 2 + hf * BN * 128, &tmK, hf * 64, hk, s0 + st * BN, &bars[1 + st]);
                tma_load_3d(sV + st * BN * HD * 2 + hf * BN * 128, &tmV, hf * 64, hk, s0 + st * BN, &bars[3 + st]);
            }
        }
    }
    float o_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) o_acc[i][0] = o_acc[i][1] = o_acc[i][2] = o_acc[i][3] = 0.f;
    float m_i[2] = {-INFINITY, -INFINITY}, l_i[2] = {0.f, 0.f};
    const float sl2 = p.scale * kLog2e;
    const int row_lo = m0 + warp * 16 + (lane >> 2);  // sequence-relative q index of c0/c1 (c2/c3: +8)
    const uint32_t sQ_a = smem_u32(sQ);

    mbar_wait(&bars[0], 0);
    for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const uint32_t sK_a = smem_u32(sK + st * BN * HD * 2), sV_a = smem_u32(sV + st * BN * HD * 2);
        float s[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        mbar_wait(&bars[1 + st], ph);
        gemm_nt<HD, BN / 8>(s, sQ_a, BM, warp * 16, sK_a, BN, 0, lane);

        // mask + row max
        const int n_base = j * BN + (lane & 3) * 2;
        const bool need_mask = (j * BN + BN > L) || (p.causal && j * BN + BN > m0 + warp * 16);
        float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (need_mask) {
                    const int n = n_base + nb * 8 + (e & 1);
                    const int m = row_lo + ((e >> 1) << 3);
                    if (n >= L || (p.causal && n > m)) s[nb][e] = -INFINITY;
                }
                mx[e >> 1] = fmaxf(mx[e >> 1], s[nb][e]);
            }
        }
        float alpha[2], m_use[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float m_new = fmaxf(m_i[r], quad_max(mx[r]));
            m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
            alpha[r] = exp2f((m_i[r] - m_use[r]) * sl2);  // m_i = -inf -> 0
            m_i[r] = m_new;
            l_i[r] *= alpha[r];
        }
        uint32_t pf[BN / 16][4];
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
            float e0 = exp2f(s[nb][0] * sl2 - m_use[0] * sl2), e1 = exp2f(s[nb][1] * sl2 - m_use[0] * sl2);
            float e2 = exp2f(s[nb][2] * sl2 - m_use[1] * sl2), e3 = exp2f(s[nb][3] * sl2 - m_use[1] * sl2);
            l_i[0] += e0 + e1;
            l_i[1] += e2 + e3;
            pf[nb >> 1][(nb & 1) * 2 + 0] = f2_to_bf2(e0, e1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = f2_to_bf2(e2, e3);
        }
#pragma unroll
        for (int i = 0; i < HD / 8; ++i) {
            o_acc[i][0] *= alpha[0]; o_acc[i][1] *= alpha[0];
            o_acc[i][2] *= alpha[1]; o_acc[i][3] *= alpha[1];
        }
        mbar_wait(&bars[3 + st], ph);
        gemm_rt<HD, BN>(o_acc, pf, sV_a, BN, 0, lane);
        __syncthreads();  // every warp is done with stage st
        if (threadIdx.x == 0 && j + 2 < n_tiles) {
            mbar_expect_tx(&bars[1 + st], BN * HD * 2);
            mbar_expect_tx(&bars[3 + st], BN * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sK + st * BN * HD * 2 + hf * BN * 128, &tmK, hf * 64, hk, s0 + (j + 2) * BN, &bars[1 + st]);
                tma_load_3d(sV + st * BN * HD * 2 + hf * BN * 128, &tmV, hf * 64, hk, s0 + (j + 2) * BN, &bars[3 + st]);
            }
        }
    }
    // epilogue
    float inv[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        l_i[r] = quad_sum(l_i[r]);
        inv[r] = l_i[r] > 0.f ? 1.f / l_i[r] : 0.f;
        const int m = row_lo + r * 8;
        if ((lane & 3) == 0 && m < L)
            p.lse[(int64_t)h * p.total + s0 + m] = (l_i[r] > 0.f) ? m_i[r] * p.scale + logf(l_i[r]) : -INFINITY;
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        o_acc[i][0] *= inv[0]; o_acc[i][1] *= inv[0];
        o_acc[i][2] *= inv[1]; o_acc[i][3] *= inv[1];
    }
    // (all warps passed the final __syncthreads of the loop: sQ is free to be reused by its owner warp)
    const int valid = min(16, L - (m0 + warp * 16));
    if (valid > 0)
        store_rows<HD>(o_acc, sQ_a, sQ, BM, warp * 16, lane,
                       p.o + (int64_t)(s0 + m0 + warp * 16) * p.o_stride_tok + (int64_t)h * p.o_stride_head,
                       p.o_stride_tok, valid);
}

// ------------------------------------------------------------------------------------------------
// backward: delta = rowsum(dO * O)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                      float* __restrict__ delta, int total, int Hq, int64_t o_st, int64_t o_sh, int64_t do_st,
                      int64_t do_sh) {
    constexpr int LPR = HD / 8;  // lanes per (token, head) row
    const int64_t rows = (int64_t)total * Hq;
    const int sub = threadIdx.x % LPR;
    const int64_t per_iter = (int64_t)gridDim.x * blockDim.x / LPR;
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPR;
    const int64_t padded = (rows + per_iter - 1) / per_iter * per_iter;
    for (int64_t r = first; r < padded; r += per_iter) {
        float acc = 0.f;
        const bool ok = r < rows;
        const int64_t t = ok ? r / Hq : 0;
        const int hh = ok ? (int)(r % Hq) : 0;
        if (ok) {
            float a[8], b[8];
            unpack8(ldg_stream(o + t * o_st + hh * o_sh + sub * 8), a);
            unpack8(ldg_stream(dout + t * do_st + hh * do_sh + sub * 8), b);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc = fmaf(a[i], b[i], acc);
        }
        acc = group_sum<LPR>(acc);
        if (ok && sub == 0) delta[(int64_t)hh * total + t] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// backward: dQ  (row tiles of Q; streams K/V)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                   const AttnParams p) {
    constexpr int BM = 128, BN = 64, NH = HD / 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + BM * HD * 2;
    uint8_t* sK = sdO + BM * HD * 2;
    uint8_t* sV = sK + 2 * BN * HD * 2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sV + 2 * BN * HD * 2);  // barQ, fullK[2], fullV[2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int order, h, seq;
    tile_of_block(p.Hq, p.num_seqs, order, h, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int mblk = p.tiles - 1 - order;
    const int m0 = mblk * BM;
    if (m0 >= L) return;
    const int hk = h / (p.Hq / p.Hk);
    const int kv_end = p.causal ? min(L, m0 + BM) : L;
    const int n_tiles = (kv_end + BN - 1) / BN;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 5; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 2 * BM * HD * 2);
        for (int hf = 0; hf < NH; ++hf) {
            tma_load_3d(sQ + hf * BM * 128, &tmQ, hf * 64, h, s0 + m0, &bars[0]);
            tma_load_3d(sdO + hf * BM * 128, &tmdO, hf * 64, h, s0 + m0, &bars[0]);
        }
        for (int st = 0; st < 2 && st < n_tiles; ++st) {
            mbar_expect_tx(&bars[1 + st], BN * HD * 2);
            mbar_expect_tx(&bars[3 + st], BN * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sK + st * BN * HD * 2 + hf * BN * 128, &tmK, hf * 64, hk, s0 + st * BN, &bars[1 + st]);
                tma_load_3d(sV + st * BN * HD * 2 + hf * BN * 128, &tmV, hf * 64, hk, s0 + st * BN, &bars[3 + st]);
            }
        }
    }
    float dq_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) dq_acc[i][0] = dq_acc[i][1] = dq_acc[i][2] = dq_acc[i][3] = 0.f;
    const float sl2 = p.scale * kLog2e;
    const int row_lo = m0 + warp * 16 + (lane >> 2);
    float lse2[2], dl[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int m = row_lo + r * 8;
        const bool ok = m < L;
        const float l = ok ? p.lse[(int64_t)h * p.total + s0 + m] : 0.f;
        lse2[r] = (l == -INFINITY) ? 0.f : l * kLog2e;
        dl[r] = ok ? p.delta[(int64_t)h * p.total + s0 + m] : 0.f;
    }
    const uint32_t sQ_a = smem_u32(sQ), sdO_a = smem_u32(sdO);
    mbar_wait(&bars[0], 0);
    for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const uint32_t sK_a = smem_u32(sK + st * BN * HD * 2), sV_a = smem_u32(sV + st * BN * HD * 2);
        float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
        for (int i = 0; i < BN / 8; ++i) {
            s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
            dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
        }
        mbar_wait(&bars[1 + st], ph);
        gemm_nt<HD, BN / 8>(s, sQ_a, BM, warp * 16, sK_a, BN, 0, lane);
        mbar_wait(&bars[3 + st], ph);
        gemm_nt<HD, BN / 8>(dp, sdO_a, BM, warp * 16, sV_a, BN, 0, lane);
        const int n_base = j * BN + (lane & 3) * 2;
        uint32_t dsf[BN / 16][4];
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
            float ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int n = n_base + nb * 8 + (e & 1);
                const int m = row_lo + ((e >> 1) << 3);
                const bool valid = n < L && m < L && (!p.causal || n <= m);
                const float pe = valid ? exp2f(s[nb][e] * sl2 - lse2[e >> 1]) : 0.f;
                ds[e] = pe * (dp[nb][e] - dl[e >> 1]);
            }
            dsf[nb >> 1][(nb & 1) * 2 + 0] = f2_to_bf2(ds[0], ds[1]);
            dsf[nb >> 1][(nb & 1) * 2 + 1] = f2_to_bf2(ds[2], ds[3]);
        }
        gemm_rt<HD, BN>(dq_acc, dsf, sK_a, BN, 0, lane);
        __syncthreads();
        if (threadIdx.x == 0 && j + 2 < n_tiles) {
            mbar_expect_tx(&bars[1 + st], BN * HD * 2);
            mbar_expect_tx(&bars[3 + st], BN * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sK + st * BN * HD * 2 + hf * BN * 128, &tmK, hf * 64, hk, s0 + (j + 2) * BN, &bars[1 + st]);
                tma_load_3d(sV + st * BN * HD * 2 + hf * BN * 128, &tmV, hf * 64, hk, s0 + (j + 2) * BN, &bars[3 + st]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        dq_acc[i][0] *= p.scale; dq_acc[i][1] *= p.scale; dq_acc[i][2] *= p.scale; dq_acc[i][3] *= p.scale;
    }
    const int valid = min(16, L - (m0 + warp * 16));
    if (valid > 0)
        store_rows<HD>(dq_acc, sQ_a, sQ, BM, warp * 16, lane,
                       p.dq + (int64_t)(s0 + m0 + warp * 16) * p.dq_stride_tok + (int64_t)h * p.dq_stride_head,
                       p.dq_stride_tok, valid);
}

// ------------------------------------------------------------------------------------------------
// backward: dK, dV  (row tiles of K/V; streams Q/dO of every q head in the GQA group)
// ------------------------------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(256, 1)
attn_bwd_dkdv_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                     const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                     const AttnParams p) {
    constexpr int BNK = 128, BMQ = 64, NH = HD / 64;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sK = smem;
    uint8_t* sV = sK + BNK * HD * 2;
    uint8_t* sQ = sV + BNK * HD * 2;             // 2 stages * BMQ*HD*2
    uint8_t* sdO = sQ + 2 * BMQ * HD * 2;        // 2 stages
    uint64_t* bars = reinterpret_cast<uint64_t*>(sdO + 2 * BMQ * HD * 2);  // barKV, full[2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int order, hk, seq;
    tile_of_block(p.Hk, p.num_seqs, order, hk, seq);
    const int s0 = p.cu_seqlens[seq], L = p.cu_seqlens[seq + 1] - s0;
    const int n0 = order * BNK;
    if (n0 >= L) return;
    const int G = p.Hq / p.Hk;
    const int i_start = p.causal ? n0 / BMQ : 0;
    const int nq = (L + BMQ - 1) / BMQ - i_start;  // q tiles per head
    const int jobs = nq * G;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bars[0], 2 * BNK * HD * 2);
        for (int hf = 0; hf < NH; ++hf) {
            tma_load_3d(sK + hf * BNK * 128, &tmK, hf * 64, hk, s0 + n0, &bars[0]);
            tma_load_3d(sV + hf * BNK * 128, &tmV, hf * 64, hk, s0 + n0, &bars[0]);
        }
        for (int st = 0; st < 2 && st < jobs; ++st) {
            const int hh = hk * G + st / nq, qi = i_start + st % nq;
            mbar_expect_tx(&bars[1 + st], 2 * BMQ * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sQ + st * BMQ * HD * 2 + hf * BMQ * 128, &tmQ, hf * 64, hh, s0 + qi * BMQ, &bars[1 + st]);
                tma_load_3d(sdO + st * BMQ * HD * 2 + hf * BMQ * 128, &tmdO, hf * 64, hh, s0 + qi * BMQ, &bars[1 + st]);
            }
        }
    }
    float dk_acc[HD / 8][4], dv_acc[HD / 8][4];
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        dk_acc[i][0] = dk_acc[i][1] = dk_acc[i][2] = dk_acc[i][3] = 0.f;
        dv_acc[i][0] = dv_acc[i][1] = dv_acc[i][2] = dv_acc[i][3] = 0.f;
    }
    const float sl2 = p.scale * kLog2e;
    const int kv_lo = n0 + warp * 16 + (lane >> 2);  // kv index of c0/c1 (c2/c3: +8)
    const uint32_t sK_a = smem_u32(sK), sV_a = smem_u32(sV);
    mbar_wait(&bars[0], 0);
    for (int jb = 0; jb < jobs; ++jb) {
        const int st = jb & 1;
        const uint32_t ph = (uint32_t)(jb >> 1) & 1u;
        const int hh = hk * G + jb / nq, qi = i_start + jb % nq;
        const uint32_t sQ_a = smem_u32(sQ + st * BMQ * HD * 2), sdO_a = smem_u32(sdO + st * BMQ * HD * 2);
        // per-column (q index) softmax statistics; issued early, consumed after the first GEMM
        float lse2[BMQ / 8][2], dl[BMQ / 8][2];
        const int m_base = qi * BMQ + (lane & 3) * 2;
#pragma unroll
        for (int nb = 0; nb < BMQ / 8; ++nb) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int m = m_base + nb * 8 + e;
                const bool ok = m < L;
                const float l = ok ? p.lse[(int64_t)hh * p.total + s0 + m] : 0.f;
                lse2[nb][e] = (l == -INFINITY) ? 0.f : l * kLog2e;
                dl[nb][e] = ok ? p.delta[(int64_t)hh * p.total + s0 + m] : 0.f;
            }
        }
        float st_acc[BMQ / 8][4], dpt[BMQ / 8][4];
#pragma unroll
        for (int i = 0; i < BMQ / 8; ++i) {
            st_acc[i][0] = st_acc[i][1] = st_acc[i][2] = st_acc[i][3] = 0.f;
            dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
        mbar_wait(&bars[1 + st], ph);
        gemm_nt<HD, BMQ / 8>(st_acc, sK_a, BNK, warp * 16, sQ_a, BMQ, 0, lane);   // S^T = K Q^T
        gemm_nt<HD, BMQ / 8>(dpt, sV_a, BNK, warp * 16, sdO_a, BMQ, 0, lane);     // dP^T = V dO^T
        uint32_t pf[BMQ / 16][4], dsf[BMQ / 16][4];
#pragma unroll
        for (int nb = 0; nb < BMQ / 8; ++nb) {
            float pe[4], ds[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m_base + nb * 8 + (e & 1);
                const int n = kv_lo + ((e >> 1) << 3);
                const bool valid = n < L && m < L && (!p.causal || n <= m);
                pe[e] = valid ? exp2f(st_acc[nb][e] * sl2 - lse2[nb][e & 1]) : 0.f;
                ds[e] = pe[e] * (dpt[nb][e] - dl[nb][e & 1]);
            }
            pf[nb >> 1][(nb & 1) * 2 + 0] = f2_to_bf2(pe[0], pe[1]);
            pf[nb >> 1][(nb & 1) * 2 + 1] = f2_to_bf2(pe[2], pe[3]);
            dsf[nb >> 1][(nb & 1) * 2 + 0] = f2_to_bf2(ds[0], ds[1]);
            dsf[nb >> 1][(nb & 1) * 2 + 1] = f2_to_bf2(ds[2], ds[3]);
        }
        gemm_rt<HD, BMQ>(dv_acc, pf, sdO_a, BMQ, 0, lane);   // dV += P^T dO
        gemm_rt<HD, BMQ>(dk_acc, dsf, sQ_a, BMQ, 0, lane);   // dK += dS^T Q
        __syncthreads();
        if (threadIdx.x == 0 && jb + 2 < jobs) {
            const int h2 = hk * G + (jb + 2) / nq, q2 = i_start + (jb + 2) % nq;
            mbar_expect_tx(&bars[1 + st], 2 * BMQ * HD * 2);
            for (int hf = 0; hf < NH; ++hf) {
                tma_load_3d(sQ + st * BMQ * HD * 2 + hf * BMQ * 128, &tmQ, hf * 64, h2, s0 + q2 * BMQ, &bars[1 + st]);
                tma_load_3d(sdO + st * BMQ * HD * 2 + hf * BMQ * 128, &tmdO, hf * 64, h2, s0 + q2 * BMQ, &bars[1 + st]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < HD / 8; ++i) {
        dk_acc[i][0] *= p.scale; dk_acc[i][1] *= p.scale; dk_acc[i][2] *= p.scale; dk_acc[i][3] *= p.scale;
    }
    const int valid = min(16, L - (n0 + warp * 16));
    if (valid > 0) {
        store_rows<HD>(dk_acc, sK_a, sK, BNK, warp * 16, lane,
                       p.dk + (int64_t)(s0 + n0 + warp * 16) * p.dk_stride_tok + (int64_t)hk * p.dk_stride_head,
                       p.dk_stride_tok, valid);
        store_rows<HD>(dv_acc, sV_a, sV, BNK, warp * 16, lane,
                       p.dv + (int64_t)(s0 + n0 + warp * 16) * p.dv_stride_tok + (int64_t)hk * p.dv_stride_head,
                       p.dv_stride_tok, valid);
    }
}

template <int HD>
static int fwd_impl(const void* q, const void* k, const void* v, void* o, float* lse, const int* cu, int num_seqs,
                    int max_seqlen, int total, int Hq, int Hk, const int64_t* st, float scale, int causal,
                    cudaStream_t stream) {
    CUtensorMap tmQ, tmK, tmV;
    int rc;
    if ((rc = make_tmap_3d(&tmQ, q, HD, Hq, total, st[1], st[0], 128))) return rc;
    if ((rc = make_tmap_3d(&tmK, k, HD, Hk, total, st[3], st[2], 64))) return rc;
    if ((rc = make_tmap_3d(&tmV, v, HD, Hk, total, st[5], st[4], 64))) return rc;
    AttnParams p{};
    p.cu_seqlens = cu; p.num_seqs = num_seqs; p.Hq = Hq; p.Hk = Hk; p.total = total; p.scale = scale; p.causal = causal;
    p.o = (__nv_bfloat16*)o; p.o_stride_tok = st[6]; p.o_stride_head = st[7]; p.lse = lse;
    const size_t smem = 128 * HD * 2 + 4 * 64 * HD * 2 + 64;
    static bool attr = false;
    if (!attr) {
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr = true;
    }
    p.tiles = (max_seqlen + 127) / 128;
    dim3 grid(p.tiles * Hq * num_seqs);
    attn_fwd_kernel<HD><<<grid, 256, smem, stream>>>(tmQ, tmK, tmV, p);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

template <int HD>
static int bwd_impl(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse,
                    float* delta, void* dq, void* dk, void* dv, const int* cu, int num_seqs, int max_seqlen,
                    int total, int Hq, int Hk, const int64_t* st, float scale, int causal, cudaStream_t stream) {
    // st: q(tok,head) k v o do dq dk dv
    attn_bwd_delta_kernel<HD><<<vb::kNumSMs * 4, 256, 0, stream>>>(
        (const __nv_bfloat16*)o, (const __nv_bfloat16*)dout, delta, total, Hq, st[6], st[7], st[8], st[9]);
    VB_HOST_CHECK_LAUNCH();
    CUtensorMap tmQ128, tmdO128, tmK64, tmV64, tmK128, tmV128, tmQ64, tmdO64;
    int rc;
    if ((rc = make_tmap_3d(&tmQ128, q, HD, Hq, total, st[1], st[0], 128))) return rc;
    if ((rc = make_tmap_3d(&tmdO128, dout, HD, Hq, total, st[9], st[8], 128))) return rc;
    if ((rc = make_tmap_3d(&tmK64, k, HD, Hk, total, st[3], st[2], 64))) return rc;
    if ((rc = make_tmap_3d(&tmV64, v, HD, Hk, total, st[5], st[4], 64))) return rc;
    if ((rc = make_tmap_3d(&tmK128, k, HD, Hk, total, st[3], st[2], 128))) return rc;
    if ((rc = make_tmap_3d(&tmV128, v, HD, Hk, total, st[5], st[4], 128))) return rc;
    if ((rc = make_tmap_3d(&tmQ64, q, HD, Hq, total, st[1], st[0], 64))) return rc;
    if ((rc = make_tmap_3d(&tmdO64, dout, HD, Hq, total, st[9], st[8], 64))) return rc;
    AttnParams p{};
    p.cu_seqlens = cu; p.num_seqs = num_seqs; p.Hq = Hq; p.Hk = Hk; p.total = total; p.scale = scale; p.causal = causal;
    p.lse = const_cast<float*>(lse); p.delta = delta;
    p.dq = (__nv_bfloat16*)dq; p.dq_stride_tok = st[10]; p.dq_stride_head = st[11];
    p.dk = (__nv_bfloat16*)dk; p.dk_stride_tok = st[12]; p.dk_stride_head = st[13];
    p.dv = (__nv_bfloat16*)dv; p.dv_stride_tok = st[14]; p.dv_stride_head = st[15];
    const size_t smem_dq = 2 * 128 * HD * 2 + 4 * 64 * HD * 2 + 64;
    const size_t smem_kv = 2 * 128 * HD * 2 + 4 * 64 * HD * 2 + 64;
    static bool attr = false;
    if (!attr) {
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_dq));
        VB_CUDA_TRY(cudaFuncSetAttribute(attn_bwd_dkdv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_kv));
        attr = true;
    }
    p.tiles = (max_seqlen + 127) / 128;
    dim3 gq(p.tiles * Hq * num_seqs);
    attn_bwd_dq_kernel<HD><<<gq, 256, smem_dq, stream>>>(tmQ128, tmdO128, tmK64, tmV64, p);
    VB_HOST_CHECK_LAUNCH();
    dim3 gk(p.tiles * Hk * num_seqs);
    attn_bwd_dkdv_kernel<HD><<<gk, 256, smem_kv, stream>>>(tmK128, tmV128, tmQ64, tmdO64, p);
    vb200_count_launch(3);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

}  // namespace vb

using namespace vb;

static int check_common(int head_dim, int Hq, int Hk, const int64_t* st, int nst) {
    if (head_dim != 64 && head_dim != 128) return vb200_set_error(VB200_EINVAL, "attention: head_dim must be 64 or 128");
    if (Hq <= 0 || Hk <= 0 || Hq % Hk) return vb200_set_error(VB200_EINVAL, "attention: Hq must be a multiple of Hk");
    for (int i = 0; i < nst; ++i)
        if (st[i] & 7) return vb200_set_error(VB200_EINVAL, "attention: strides must be multiples of 8 elements");
    return 0;
}

extern "C" int vb200_attn_varlen_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                                     const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int32_t total,
                                     int32_t q_heads, int32_t k_heads, int32_t head_dim, const int64_t* strides,
                                     float scale, int32_t causal, void* stream) {
    int rc = check_common(head_dim, q_heads, k_heads, strides, 8);
    if (rc) return rc;
    if (total <= 0 || num_seqs <= 0 || max_seqlen <= 0) return VB200_OK;
    if (head_dim == 128)
        return fwd_impl<128>(q, k, v, o, lse, cu_seqlens, num_seqs, max_seqlen, total, q_heads, k_heads, strides, scale,
                             causal, (cudaStream_t)stream);
    return fwd_impl<64>(q, k, v, o, lse, cu_seqlens, num_seqs, max_seqlen, total, q_heads, k_heads, strides, scale,
                        causal, (cudaStream_t)stream);
}

extern "C" int vb200_attn_varlen_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                                     const float* lse, float* delta, void* dq, void* dk, void* dv,
                                     const int32_t* cu_seqlens, int32_t num_seqs, int32_t max_seqlen, int32_t total,
                                     int32_t q_heads, int32_t k_heads, int32_t head_dim, const int64_t* strides,
                                     float scale, int32_t causal, void* stream) {
    int rc = check_common(head_dim, q_heads, k_heads, strides, 16);
    if (rc) return rc;
    if (total <= 0 || num_seqs <= 0 || max_seqlen <= 0) return VB200_OK;
    if (head_dim == 128)
        return bwd_impl<128>(q, k, v, o, dout, lse, delta, dq, dk, dv, cu_seqlens, num_seqs, max_seqlen, total, q_heads,
                             k_heads, strides, scale, causal, (cudaStream_t)stream);
    return bwd_impl<64>(q, k, v, o, dout, lse, delta, dq, dk, dv, cu_seqlens, num_seqs, max_seqlen, total, q_heads,
                        k_heads, strides, scale, causal, (cudaStream_t)stream);
}

// delta[h, t] = sum_d dO[t,h,d] * O[t,h,d]  (the softmax-backward row term; shared by both backward paths)
extern "C" int vb200_attn_bwd_delta(const void* o, const void* dout, float* delta, int32_t total, int32_t q_heads,
                                    int32_t head_dim, int64_t o_st, int64_t o_sh, int64_t do_st, int64_t do_sh, void* stream) {
    if (head_dim != 64 && head_dim != 128) return vb200_set_error(VB200_EINVAL, "attn_bwd_delta: head_dim must be 64 or 128");
    if (total <= 0) return VB200_OK;
    cudaStream_t s = (cudaStream_t)stream;
    if (head_dim == 128)
        attn_bwd_delta_kernel<128><<<vb::kNumSMs * 4, 256, 0, s>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)dout, delta,
                                                                  total, q_heads, o_st, o_sh, do_st, do_sh);
    else
        attn_bwd_delta_kernel<64><<<vb::kNumSMs * 4, 256, 0, s>>>((const __nv_bfloat16*)o, (const __nv_bfloat16*)dout, delta,
                                                                 total, q_heads, o_st, o_sh, do_st, do_sh);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
