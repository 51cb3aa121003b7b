// RoPE and fused per-head q/k RMSNorm + RoPE for sm_100a.
//
// Reference semantics:
//   apply_rotary_pos_emb (veomni/models/transformers/qwen3/generated/
//   patched_modeling_qwen3_gpu.py:208-223):  out = x*cos + rotate_half(x)*sin with
//   rotate_half(x) = cat(-x[D/2:], x[:D/2]) (:196-200); cos/sin are [tokens, D] bf16 produced by
//   Qwen3RotaryEmbedding.forward (:181-192).  Bound on GPU to liger_rotary_pos_emb
//   (veomni/ops/liger/__init__.py:102-113).
//   q_norm / k_norm (patched_modeling_qwen3_gpu.py:305-306) are Qwen3RMSNorm over head_dim.
//
// Roofline: HBM stream, 2 * tokens * (Hq+Hk) * D * 2 B per pass (+ tokens*D*4 B of cos/sin).
// Layout: D/16 lanes cooperate on one (token, head) row; lane `sub` owns the 8 elements at
// [sub*8, sub*8+8) and their rotation partners at [D/2 + sub*8, ...), so the rotate_half pairing
// never leaves the thread. One work item = (token, chunk of 8 heads): cos/sin for the token are
// loaded once into registers and reused for the 8 heads.
#include "common.cuh"

namespace vb {

// Work item = (token, kHeadChunk heads). These are 15-40 us kernels on 296 resident CTAs: with 8-head items the T=4096,
// 40-head problem is 2.16 items per lane group, i.e. a third pass that is 16 % full (72 % of the achievable rate; the
// backward's capped grid did worse); 2-head items make it 8.65 -> 9 passes (96 %). cos/sin rows are re-read per item,
// from L2 (1 MB in total).
constexpr int kHeadChunk = 2;  // heads per work item
constexpr int kSub = 2;        // heads in flight per thread (all of an item: 8 x 16 B loads + the 4 table loads)

struct RopeTables {
    float c_lo[8], c_hi[8], s_lo[8], s_hi[8];
};

__device__ __forceinline__ void load_tables(RopeTables& t, const __nv_bfloat16* cos, const __nv_bfloat16* sin,
                                            int64_t tok, int D, int sub) {
    const int half = D >> 1;
    unpack8(*reinterpret_cast<const uint4*>(cos + tok * D + sub * 8), t.c_lo);
    unpack8(*reinterpret_cast<const uint4*>(cos + tok * D + half + sub * 8), t.c_hi);
    unpack8(*reinterpret_cast<const uint4*>(sin + tok * D + sub * 8), t.s_lo);
    unpack8(*reinterpret_cast<const uint4*>(sin + tok * D + half + sub * 8), t.s_hi);
}

// forward rotation:  o_lo = x_lo*c_lo - x_hi*s_lo ; o_hi = x_hi*c_hi + x_lo*s_hi
__device__ __forceinline__ void rotate_fwd(const RopeTables& t, const float (&lo)[8], const float (&hi)[8],
                                           float (&olo)[8], float (&ohi)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        olo[i] = lo[i] * t.c_lo[i] - hi[i] * t.s_lo[i];
        ohi[i] = hi[i] * t.c_hi[i] + lo[i] * t.s_hi[i];
    }
}
// transposed rotation (vector-Jacobian product of rotate_fwd):
//   dx_lo = g_lo*c_lo + g_hi*s_hi ; dx_hi = g_hi*c_hi - g_lo*s_lo
__device__ __forceinline__ void rotate_bwd(const RopeTables& t, const float (&glo)[8], const float (&ghi)[8],
                                           float (&dlo)[8], float (&dhi)[8]) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        dlo[i] = glo[i] * t.c_lo[i] + ghi[i] * t.s_hi[i];
        dhi[i] = ghi[i] * t.c_hi[i] - glo[i] * t.s_lo[i];
    }
}

template <int LPH>
__global__ void __launch_bounds__(256, 2)
rope_kernel(const __nv_bfloat16* __restrict__ q_in, __nv_bfloat16* __restrict__ q_out,
            const __nv_bfloat16* __restrict__ k_in, __nv_bfloat16* __restrict__ k_out,
            const __nv_bfloat16* __restrict__ cos, const __nv_bfloat16* __restrict__ sin, int64_t tokens,
            int Hq, int Hk, int64_t qs_t, int64_t qs_h, int64_t ks_t, int64_t ks_h, int64_t qos_t,
            int64_t qos_h, int64_t kos_t, int64_t kos_h, int inverse) {
    constexpr int D = LPH * 16, HALF = D / 2;
    const int H = Hq + Hk;
    const int nchunks = (H + kHeadChunk - 1) / kHeadChunk;
    const int64_t items = tokens * nchunks;
    const int sub = threadIdx.x % LPH;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPH;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x / LPH;
    for (int64_t item = gid; item < items; item += gstride) {
        const int64_t tok = item / nchunks;
        const int h0 = (int)(item % nchunks) * kHeadChunk;
        RopeTables t;
        load_tables(t, cos, sin, tok, D, sub);
#pragma unroll 1
        for (int jb = 0; jb < kHeadChunk; jb += kSub) {
            uint4 lo_v[kSub], hi_v[kSub];
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                const int h = h0 + jb + j;
                if (h < H) {
                    const __nv_bfloat16* src = h < Hq ? q_in + tok * qs_t + (int64_t)h * qs_h
                                                      : k_in + tok * ks_t + (int64_t)(h - Hq) * ks_h;
                    lo_v[j] = ldg_stream(src + sub * 8);
                    hi_v[j] = ldg_stream(src + HALF + sub * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                const int h = h0 + jb + j;
                if (h < H) {
                    float lo[8], hi[8], olo[8], ohi[8];
                    unpack8(lo_v[j], lo);
                    unpack8(hi_v[j], hi);
                    if (inverse) rotate_bwd(t, lo, hi, olo, ohi);
                    else rotate_fwd(t, lo, hi, olo, ohi);
                    __nv_bfloat16* dst = h < Hq ? q_out + tok * qos_t + (int64_t)h * qos_h
                                                : k_out + tok * kos_t + (int64_t)(h - Hq) * kos_h;
                    stg_stream(dst + sub * 8, pack8(olo));
                    stg_stream(dst + HALF + sub * 8, pack8(ohi));
                }
            }
        }
    }
}

// ---- fused q/k-norm + RoPE ---------------------------------------------------------------
template <int LPH>
__global__ void __launch_bounds__(256, 2)
qknorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ q_in, const __nv_bfloat16* __restrict__ k_in,
                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ cos, const __nv_bfloat16* __restrict__ sin,
                       __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out,
                       float* __restrict__ rstd_q, float* __restrict__ rstd_k, int64_t tokens, int Hq, int Hk,
                       float eps) {
    constexpr int D = LPH * 16, HALF = D / 2;
    const int H = Hq + Hk;
    const int nchunks = (H + kHeadChunk - 1) / kHeadChunk;
    const int64_t items = tokens * nchunks;
    const int sub = threadIdx.x % LPH;
    // weights stay packed (bf16x8) to keep the register budget at 2 CTAs/SM
    const uint4 wq_lo_p = *reinterpret_cast<const uint4*>(wq + sub * 8);
    const uint4 wq_hi_p = *reinterpret_cast<const uint4*>(wq + HALF + sub * 8);
    const uint4 wk_lo_p = *reinterpret_cast<const uint4*>(wk + sub * 8);
    const uint4 wk_hi_p = *reinterpret_cast<const uint4*>(wk + HALF + sub * 8);
    const int64_t per_iter = (int64_t)gridDim.x * blockDim.x / LPH;
    const int64_t first = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPH;
    const int64_t padded = (items + per_iter - 1) / per_iter * per_iter;
    const float inv_d = 1.0f / (float)D;
    for (int64_t item = first; item < padded; item += per_iter) {
        const bool item_ok = item < items;
        const int64_t tok = item_ok ? item / nchunks : 0;
        const int h0 = item_ok ? (int)(item % nchunks) * kHeadChunk : 0;
        RopeTables t;
        load_tables(t, cos, sin, tok, D, sub);
#pragma unroll 1
        for (int jb = 0; jb < kHeadChunk; jb += kSub) {
            uint4 lo_v[kSub], hi_v[kSub];
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                const int h = h0 + jb + j;
                if (item_ok && h < H) {
                    const __nv_bfloat16* src =
                        h < Hq ? q_in + (tok * Hq + h) * D : k_in + (tok * Hk + (h - Hq)) * D;
                    lo_v[j] = ldg_stream(src + sub * 8);
                    hi_v[j] = ldg_stream(src + HALF + sub * 8);
                } else {
                    lo_v[j] = make_uint4(0, 0, 0, 0);
                    hi_v[j] = make_uint4(0, 0, 0, 0);
                }
            }
#pragma unroll
            for (int j = 0; j < kSub; ++j) {
                const int h = h0 + jb + j;
                const bool ok = item_ok && h < H;
                const bool is_q = h < Hq;
                float lo[8], hi[8];
                unpack8(lo_v[j], lo);
                unpack8(hi_v[j], hi);
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 8; ++i) ss = fmaf(lo[i], lo[i], fmaf(hi[i], hi[i], ss));
                ss = group_sum<LPH>(ss);
                const float rs = rsqrtf(ss * inv_d + eps);
                if (ok) {
                    float olo[8], ohi[8], w_lo[8], w_hi[8];
                    unpack8(is_q ? wq_lo_p : wk_lo_p, w_lo);
                    unpack8(is_q ? wq_hi_p : wk_hi_p, w_hi);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        // y = bf16(w * bf16(x*rstd)): the reference materialises q_norm's output in bf16
                        lo[i] = round_bf16(w_lo[i] * round_bf16(lo[i] * rs));
                        hi[i] = round_bf16(w_hi[i] * round_bf16(hi[i] * rs));
                    }
                    rotate_fwd(t, lo, hi, olo, ohi);
                    __nv_bfloat16* dst = is_q ? q_out + (tok * Hq + h) * D : k_out + (tok * Hk + (h - Hq)) * D;
                    stg_stream(dst + sub * 8, pack8(olo));
                    stg_stream(dst + HALF + sub * 8, pack8(ohi));
                    if (sub == 0) {
                        if (is_q) rstd_q[tok * Hq + h] = rs;
                        else rstd_k[tok * Hk + (h - Hq)] = rs;
                    }
                }
            }
        }
    }
}

template <int LPH>
__global__ void __launch_bounds__(256, 2)
qknorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ dq_out, const __nv_bfloat16* __restrict__ dk_out,
                       const __nv_bfloat16* __restrict__ q_in, const __nv_bfloat16* __restrict__ k_in,
                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ cos, const __nv_bfloat16* __restrict__ sin,
                       const float* __restrict__ rstd_q, const float* __restrict__ rstd_k,
                       __nv_bfloat16* __restrict__ dq_in, __nv_bfloat16* __restrict__ dk_in,
                       float* __restrict__ dw_partial, int64_t tokens, int Hq, int Hk) {
    constexpr int D = LPH * 16, HALF = D / 2, GROUPS = 256 / LPH;
    __shared__ float acc_s[GROUPS][2 * D + 1];
    const int H = Hq + Hk;
    const int nchunks = (H + kHeadChunk - 1) / kHeadChunk;
    const int64_t items = tokens * nchunks;
    const int sub = threadIdx.x % LPH, grp = threadIdx.x / LPH;
    // weights stay packed (bf16x8) to keep the register budget at 2 CTAs/SM
    const uint4 wq_lo_p = *reinterpret_cast<const uint4*>(wq + sub * 8);
    const uint4 wq_hi_p = *reinterpret_cast<const uint4*>(wq + HALF + sub * 8);
    const uint4 wk_lo_p = *reinterpret_cast<const uint4*>(wk + sub * 8);
    const uint4 wk_hi_p = *reinterpret_cast<const uint4*>(wk + HALF + sub * 8);
    float aq_lo[8] = {0, 0, 0, 0, 0, 0, 0, 0}, aq_hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float ak_lo[8] = {0, 0, 0, 0, 0, 0, 0, 0}, ak_hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const int64_t per_iter = (int64_t)gridDim.x * GROUPS;
    const int64_t first = (int64_t)blockIdx.x * GROUPS + grp;
    const int64_t padded = (items + per_iter - 1) / per_iter * per_iter;
    const float inv_d = 1.0f / (float)D;
    for (int64_t item = first; item < padded; item += per_iter) {
        const bool item_ok = item < items;
        const int64_t tok = item_ok ? item / nchunks : 0;
        const int h0 = item_ok ? (int)(item % nchunks) * kHeadChunk : 0;
        // cos/sin stay packed (bf16x8) until used: the unpacked tables would cost 32 registers next to the 32 dw accumulators
        const uint4 tcl = *reinterpret_cast<const uint4*>(cos + tok * D + sub * 8), tch = *reinterpret_cast<const uint4*>(cos + tok * D + HALF + sub * 8);
        const uint4 tsl = *reinterpret_cast<const uint4*>(sin + tok * D + sub * 8), tsh = *reinterpret_cast<const uint4*>(sin + tok * D + HALF + sub * 8);
        // all loads of the item first (8 x 16 B in flight per thread), then the arithmetic
        uint4 gv[kHeadChunk][2], xv[kHeadChunk][2];
        float rsv[kHeadChunk];
#pragma unroll
        for (int j = 0; j < kHeadChunk; ++j) {
            const int h = h0 + j;
            const bool is_q = h < Hq;
            const int64_t row = is_q ? (tok * Hq + h) : (tok * Hk + (h - Hq));
            gv[j][0] = gv[j][1] = xv[j][0] = xv[j][1] = make_uint4(0, 0, 0, 0);
            rsv[j] = 0.f;
            if (item_ok && h < H) {
                const __nv_bfloat16* gsrc = (is_q ? dq_out : dk_out) + row * D;
                const __nv_bfloat16* xsrc = (is_q ? q_in : k_in) + row * D;
                gv[j][0] = ldg_stream(gsrc + sub * 8);
                gv[j][1] = ldg_stream(gsrc + HALF + sub * 8);
                xv[j][0] = ldg_stream(xsrc + sub * 8);
                xv[j][1] = ldg_stream(xsrc + HALF + sub * 8);
                rsv[j] = is_q ? rstd_q[row] : rstd_k[row];
            }
        }
#pragma unroll
        for (int j = 0; j < kHeadChunk; ++j) {
            const int h = h0 + j;
            const bool ok = item_ok && h < H;
            const bool is_q = h < Hq;
            const int64_t row = is_q ? (tok * Hq + h) : (tok * Hk + (h - Hq));
            float glo[8], ghi[8], xlo[8], xhi[8];
            unpack8(gv[j][0], glo);
            unpack8(gv[j][1], ghi);
            unpack8(xv[j][0], xlo);
            unpack8(xv[j][1], xhi);
            const float rs = rsv[j];
            float dlo[8], dhi[8];
            {
                RopeTables t;
                unpack8(tcl, t.c_lo); unpack8(tch, t.c_hi); unpack8(tsl, t.s_lo); unpack8(tsh, t.s_hi);
                rotate_bwd(t, glo, ghi, dlo, dhi);  // dy of the norm
            }
            float dot = 0.f, w_lo[8], w_hi[8];
            unpack8(is_q ? wq_lo_p : wk_lo_p, w_lo);
            unpack8(is_q ? wq_hi_p : wk_hi_p, w_hi);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float wl = w_lo[i], wh = w_hi[i];
                xlo[i] *= rs;  // xhat
                xhi[i] *= rs;
                const float rl = round_bf16(xlo[i]), rh = round_bf16(xhi[i]);
                if (is_q) {
                    aq_lo[i] = fmaf(dlo[i], rl, aq_lo[i]);
                    aq_hi[i] = fmaf(dhi[i], rh, aq_hi[i]);
                } else {
                    ak_lo[i] = fmaf(dlo[i], rl, ak_lo[i]);
                    ak_hi[i] = fmaf(dhi[i], rh, ak_hi[i]);
                }
                dlo[i] *= wl;  // g = dy * w
                dhi[i] *= wh;
                dot = fmaf(dlo[i], xlo[i], fmaf(dhi[i], xhi[i], dot));
            }
            dot = group_sum<LPH>(dot);
            const float c = dot * inv_d;
            if (ok) {
                float olo[8], ohi[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    olo[i] = rs * (dlo[i] - xlo[i] * c);
                    ohi[i] = rs * (dhi[i] - xhi[i] * c);
                }
                __nv_bfloat16* dst = (is_q ? dq_in : dk_in) + row * D;
                stg_stream(dst + sub * 8, pack8(olo));
                stg_stream(dst + HALF + sub * 8, pack8(ohi));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        acc_s[grp][sub * 8 + i] = aq_lo[i];
        acc_s[grp][HALF + sub * 8 + i] = aq_hi[i];
        acc_s[grp][D + sub * 8 + i] = ak_lo[i];
        acc_s[grp][D + HALF + sub * 8 + i] = ak_hi[i];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * D; c += 256) {
        float s = 0.f;
#pragma unroll 4
        for (int g = 0; g < GROUPS; ++g) s += acc_s[g][c];
        dw_partial[(int64_t)blockIdx.x * 2 * D + c] = s;
    }
}

__global__ void __launch_bounds__(256)
colsum2_kernel(const float* __restrict__ partial, float* __restrict__ out_a, float* __restrict__ out_b, int64_t nparts,
               int D) {
    __shared__ float red[8][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + x;
    float t = 0.f;
    if (c < 2 * D)
        for (int64_t p = y; p < nparts; p += 8) t += partial[p * 2 * D + c];
    red[y][x] = t;
    __syncthreads();
    if (y == 0 && c < 2 * D) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) a += red[i][x];
        if (c < D) out_a[c] = a;
        else out_b[c - D] = a;
    }
}

static int items_grid(int64_t tokens, int H, int lph, int max_ctas) {
    const int nchunks = (H + kHeadChunk - 1) / kHeadChunk;
    const int64_t items = tokens * nchunks;
    const int groups = 256 / lph;
    int64_t g = (items + groups - 1) / groups;
    if (g > max_ctas) g = max_ctas;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_rope(const void* q_in, void* q_out, const void* k_in, void* k_out, const void* cos,
                          const void* sin, int64_t tokens, int32_t q_heads, int32_t k_heads, int32_t head_dim,
                          int64_t qs_t, int64_t qs_h, int64_t ks_t, int64_t ks_h, int64_t qos_t, int64_t qos_h,
                          int64_t kos_t, int64_t kos_h, int32_t inverse, void* stream) {
    if (head_dim != 64 && head_dim != 128 && head_dim != 256)
        return vb200_set_error(VB200_EINVAL, "rope: head_dim must be 64, 128 or 256");
    if ((qs_t | qs_h | ks_t | ks_h | qos_t | qos_h | kos_t | kos_h) & 7)
        return vb200_set_error(VB200_EINVAL, "rope: strides must be multiples of 8 elements");
    if (tokens <= 0 || q_heads + k_heads <= 0) return VB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int lph = head_dim / 16;
    const int g = items_grid(tokens, q_heads + k_heads, lph, 2 * kNumSMs);
#define GO(L)                                                                                            \
    rope_kernel<L><<<g, 256, 0, st>>>((const __nv_bfloat16*)q_in, (__nv_bfloat16*)q_out,                 \
                                      (const __nv_bfloat16*)k_in, (__nv_bfloat16*)k_out,                 \
                                      (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, tokens,      \
                                      q_heads, k_heads, qs_t, qs_h, ks_t, ks_h, qos_t, qos_h, kos_t, kos_h, \
                                      inverse)
    if (lph == 4) GO(4);
    else if (lph == 8) GO(8);
    else GO(16);
#undef GO
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_qknorm_rope_fwd(const void* q_in, const void* k_in, const void* wq, const void* wk,
                                     const void* cos, const void* sin, void* q_out, void* k_out, float* rstd_q,
                                     float* rstd_k, int64_t tokens, int32_t q_heads, int32_t k_heads,
                                     int32_t head_dim, float eps, void* stream) {
    if (head_dim != 64 && head_dim != 128 && head_dim != 256)
        return vb200_set_error(VB200_EINVAL, "qknorm_rope: head_dim must be 64, 128 or 256");
    if (tokens <= 0) return VB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int lph = head_dim / 16;
    const int g = items_grid(tokens, q_heads + k_heads, lph, 2 * kNumSMs);  // = the resident CTAs (launch bounds 256 x 2)
#define GO(L)                                                                                           \
    qknorm_rope_fwd_kernel<L><<<g, 256, 0, st>>>(                                                       \
        (const __nv_bfloat16*)q_in, (const __nv_bfloat16*)k_in, (const __nv_bfloat16*)wq,               \
        (const __nv_bfloat16*)wk, (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin,                 \
        (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_out, rstd_q, rstd_k, tokens, q_heads, k_heads, eps)
    if (lph == 4) GO(4);
    else if (lph == 8) GO(8);
    else GO(16);
#undef GO
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int64_t vb200_qknorm_rope_bwd_partials(int64_t tokens) {
    // upper bound over head counts / head dims: the bwd grid is capped at 4 CTAs per SM
    (void)tokens;
    return 4 * kNumSMs;
}

extern "C" int vb200_qknorm_rope_bwd(const void* dq_out, const void* dk_out, const void* q_in, const void* k_in,
                                     const void* wq, const void* wk, const void* cos, const void* sin,
                                     const float* rstd_q, const float* rstd_k, void* dq_in, void* dk_in,
                                     float* dw_partial, float* dwq, float* dwk, int64_t tokens, int32_t q_heads,
                                     int32_t k_heads, int32_t head_dim, void* stream) {
    if (head_dim != 64 && head_dim != 128 && head_dim != 256)
        return vb200_set_error(VB200_EINVAL, "qknorm_rope: head_dim must be 64, 128 or 256");
    cudaStream_t st = (cudaStream_t)stream;
    if (tokens <= 0) {
        VB_CUDA_TRY(cudaMemsetAsync(dwq, 0, sizeof(float) * head_dim, st));
        VB_CUDA_TRY(cudaMemsetAsync(dwk, 0, sizeof(float) * head_dim, st));
        return VB200_OK;
    }
    const int lph = head_dim / 16;
    const int g = items_grid(tokens, q_heads + k_heads, lph, 2 * kNumSMs);  // = the resident CTAs
#define GO(L)                                                                                           \
    qknorm_rope_bwd_kernel<L><<<g, 256, 0, st>>>(                                                       \
        (const __nv_bfloat16*)dq_out, (const __nv_bfloat16*)dk_out, (const __nv_bfloat16*)q_in,         \
        (const __nv_bfloat16*)k_in, (const __nv_bfloat16*)wq, (const __nv_bfloat16*)wk,                 \
        (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, rstd_q, rstd_k, (__nv_bfloat16*)dq_in,    \
        (__nv_bfloat16*)dk_in, dw_partial, tokens, q_heads, k_heads)
    if (lph == 4) GO(4);
    else if (lph == 8) GO(8);
    else GO(16);
#undef GO
    VB_HOST_CHECK_LAUNCH();
    colsum2_kernel<<<(2 * head_dim + 31) / 32, 256, 0, st>>>(dw_partial, dwq, dwk, g, head_dim);
    vb200_count_launch(2);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
