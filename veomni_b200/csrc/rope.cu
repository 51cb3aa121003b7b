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
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace vb {

// Work item = (token, kHeadChunk heads). These are 15-40 us kernels on 296 resident CTAs: with 8-head items the T=4096,
// 40-head problem is 2.16 items per lane group, i.e. a third pass that is 16 % full (72 % of the achievable rate; the
// backward's capped grid did worse); 2-head items make it 8.65 -> 9 passes (96 %). cos/sin rows are re-read per item,
// from L2 (1 MB in total).
constexpr int kHeadChunk = 2;  // heads per work item (backward; the forward kernels take it as a template parameter)

// cos/sin of one token for this lane's 8 + 8 columns, as packed fp32 pairs (the arithmetic below is FMUL2 / FFMA2: one issue
// slot per two elements). ns_lo = -sin_lo, so both rotations are a multiply feeding a fused multiply-add.
struct RopeTables {
    float2 c_lo[4], c_hi[4], s_hi[4], ns_lo[4];
};

__device__ __forceinline__ void unpack4(const uint4& u, float2 (&f)[4]) {
    f[0] = bf2_to_f2(u.x); f[1] = bf2_to_f2(u.y); f[2] = bf2_to_f2(u.z); f[3] = bf2_to_f2(u.w);
}
__device__ __forceinline__ uint4 pack4(const float2 (&f)[4]) {
    return make_uint4(f2_to_bf2(f[0].x, f[0].y), f2_to_bf2(f[1].x, f[1].y), f2_to_bf2(f[2].x, f[2].y), f2_to_bf2(f[3].x, f[3].y));
}

__device__ __forceinline__ void load_tables(RopeTables& t, const __nv_bfloat16* cos, const __nv_bfloat16* sin,
                                            int64_t tok, int D, int sub) {
    const int half = D >> 1;
    float2 s_lo[4];
    unpack4(*reinterpret_cast<const uint4*>(cos + tok * D + sub * 8), t.c_lo);
    unpack4(*reinterpret_cast<const uint4*>(cos + tok * D + half + sub * 8), t.c_hi);
    unpack4(*reinterpret_cast<const uint4*>(sin + tok * D + sub * 8), s_lo);
    unpack4(*reinterpret_cast<const uint4*>(sin + tok * D + half + sub * 8), t.s_hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) t.ns_lo[i] = make_float2(-s_lo[i].x, -s_lo[i].y);
}

// forward rotation:  o_lo = x_lo*c_lo - x_hi*s_lo ; o_hi = x_hi*c_hi + x_lo*s_hi
__device__ __forceinline__ void rotate_fwd(const RopeTables& t, const float2 (&lo)[4], const float2 (&hi)[4],
                                           float2 (&olo)[4], float2 (&ohi)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        olo[i] = ffma2(lo[i], t.c_lo[i], fmul2(hi[i], t.ns_lo[i]));
        ohi[i] = ffma2(hi[i], t.c_hi[i], fmul2(lo[i], t.s_hi[i]));
    }
}
// transposed rotation (vector-Jacobian product of rotate_fwd):
//   dx_lo = g_lo*c_lo + g_hi*s_hi ; dx_hi = g_hi*c_hi - g_lo*s_lo
__device__ __forceinline__ void rotate_bwd(const RopeTables& t, const float2 (&glo)[4], const float2 (&ghi)[4],
                                           float2 (&dlo)[4], float2 (&dhi)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        dlo[i] = ffma2(glo[i], t.c_lo[i], fmul2(ghi[i], t.s_hi[i]));
        dhi[i] = ffma2(ghi[i], t.c_hi[i], fmul2(glo[i], t.ns_lo[i]));
    }
}

template <int LPH, int HC>
__global__ void __launch_bounds__(256, 2)
rope_kernel(const __nv_bfloat16* __restrict__ q_in, __nv_bfloat16* __restrict__ q_out,
            const __nv_bfloat16* __restrict__ k_in, __nv_bfloat16* __restrict__ k_out,
            const __nv_bfloat16* __restrict__ cos, const __nv_bfloat16* __restrict__ sin, int64_t tokens,
            int Hq, int Hk, int64_t qs_t, int64_t qs_h, int64_t ks_t, int64_t ks_h, int64_t qos_t,
            int64_t qos_h, int64_t kos_t, int64_t kos_h, int inverse) {
    constexpr int D = LPH * 16, HALF = D / 2;
    const int H = Hq + Hk;
    const int nchunks = (H + HC - 1) / HC;
    const int64_t items = tokens * nchunks;
    const int sub = threadIdx.x % LPH;
    const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LPH;
    const int64_t gstride = (int64_t)gridDim.x * blockDim.x / LPH;
    for (int64_t item = gid; item < items; item += gstride) {
        const int64_t tok = item / nchunks;
        const int h0 = (int)(item % nchunks) * HC;
        RopeTables t;
        load_tables(t, cos, sin, tok, D, sub);
#pragma unroll 1
        for (int jb = 0; jb < HC; jb += HC) {
            uint4 lo_v[HC], hi_v[HC];
#pragma unroll
            for (int j = 0; j < HC; ++j) {
                const int h = h0 + jb + j;
                if (h < H) {
                    const __nv_bfloat16* src = h < Hq ? q_in + tok * qs_t + (int64_t)h * qs_h
                                                      : k_in + tok * ks_t + (int64_t)(h - Hq) * ks_h;
                    lo_v[j] = ldg_stream(src + sub * 8);
                    hi_v[j] = ldg_stream(src + HALF + sub * 8);
                }
            }
#pragma unroll
            for (int j = 0; j < HC; ++j) {
                const int h = h0 + jb + j;
                if (h < H) {
                    float2 lo[4], hi[4], olo[4], ohi[4];
                    unpack4(lo_v[j], lo);
                    unpack4(hi_v[j], hi);
                    if (inverse) rotate_bwd(t, lo, hi, olo, ohi);
                    else rotate_fwd(t, lo, hi, olo, ohi);
                    __nv_bfloat16* dst = h < Hq ? q_out + tok * qos_t + (int64_t)h * qos_h
                                                : k_out + tok * kos_t + (int64_t)(h - Hq) * kos_h;
                    stg_stream(dst + sub * 8, pack4(olo));
                    stg_stream(dst + HALF + sub * 8, pack4(ohi));
                }
            }
        }
    }
}

// ---- fused q/k-norm + RoPE ---------------------------------------------------------------
template <int LPH, int HC>
__global__ void __launch_bounds__(256, 2)
qknorm_rope_fwd_kernel(const __nv_bfloat16* __restrict__ q_in, const __nv_bfloat16* __restrict__ k_in,
                       const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                       const __nv_bfloat16* __restrict__ cos, const __nv_bfloat16* __restrict__ sin,
                       __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out,
                       float* __restrict__ rstd_q, float* __restrict__ rstd_k, int64_t tokens, int Hq, int Hk,
                       float eps) {
    constexpr int D = LPH * 16, HALF = D / 2;
    const int H = Hq + Hk;
    const int nchunks = (H + HC - 1) / HC;
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
        const int h0 = item_ok ? (int)(item % nchunks) * HC : 0;
        RopeTables t;
        load_tables(t, cos, sin, tok, D, sub);
#pragma unroll 1
        for (int jb = 0; jb < HC; jb += HC) {
            uint4 lo_v[HC], hi_v[HC];
#pragma unroll
            for (int j = 0; j < HC; ++j) {
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
            for (int j = 0; j < HC; ++j) {
                const int h = h0 + jb + j;
                const bool ok = item_ok && h < H;
                const bool is_q = h < Hq;
                float2 lo[4], hi[4];
                unpack4(lo_v[j], lo);
                unpack4(hi_v[j], hi);
                float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    a0 = ffma2(lo[i], lo[i], a0);
                    a1 = ffma2(hi[i], hi[i], a1);
                }
                const float ss = group_sum<LPH>((a0.x + a0.y) + (a1.x + a1.y));
                const float rs = rsqrtf(ss * inv_d + eps);
                if (ok) {
                    // y = bf16(w * bf16(x*rstd)): the reference materialises q_norm's output in bf16. The second product is a
                    // packed bf16 multiply on the still-packed weights (exact product, one rounding).
                    const float2 rs2 = make_float2(rs, rs);
                    const uint4 wl = is_q ? wq_lo_p : wk_lo_p, wh = is_q ? wq_hi_p : wk_hi_p;
                    const uint32_t wlu[4] = {wl.x, wl.y, wl.z, wl.w}, whu[4] = {wh.x, wh.y, wh.z, wh.w};
                    float2 olo[4], ohi[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 tl = fmul2(lo[i], rs2), th = fmul2(hi[i], rs2);
                        lo[i] = bf2_to_f2(bf2_mul(wlu[i], f2_to_bf2(tl.x, tl.y)));
                        hi[i] = bf2_to_f2(bf2_mul(whu[i], f2_to_bf2(th.x, th.y)));
                    }
                    rotate_fwd(t, lo, hi, olo, ohi);
                    __nv_bfloat16* dst = is_q ? q_out + (tok * Hq + h) * D : k_out + (tok * Hk + (h - Hq)) * D;
                    stg_stream(dst + sub * 8, pack4(olo));
                    stg_stream(dst + HALF + sub * 8, pack4(ohi));
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
    float2 aq_lo[4], aq_hi[4], ak_lo[4], ak_hi[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) aq_lo[i] = aq_hi[i] = ak_lo[i] = ak_hi[i] = make_float2(0.f, 0.f);
    const int64_t per_iter = (int64_t)gridDim.x * GROUPS;
    const int64_t first = (int64_t)blockIdx.x * GROUPS + grp;
    const int64_t padded = (items + per_iter - 1) / per_iter * per_iter;
    const float inv_d = 1.0f / (float)D;
    for (int64_t item = first; item < padded; item += per_iter) {
        const bool item_ok = item < items;
        const int64_t tok = item_ok ? item / nchunks : 0;
        const int h0 = item_ok ? (int)(item % nchunks) * kHeadChunk : 0;
        RopeTables t;
        load_tables(t, cos, sin, tok, D, sub);
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
            // Operands are unpacked pair by pair so that only xhat and g (= dy * w) stay live across the row reduction.
            float2 xlo[4], xhi[4], dlo[4], dhi[4];
            const float rs = rsv[j];
            const float2 rs2 = make_float2(rs, rs);
            const uint4 wl = is_q ? wq_lo_p : wk_lo_p, wh = is_q ? wq_hi_p : wk_hi_p;
            const uint32_t glu[4] = {gv[j][0].x, gv[j][0].y, gv[j][0].z, gv[j][0].w}, ghu[4] = {gv[j][1].x, gv[j][1].y, gv[j][1].z, gv[j][1].w};
            const uint32_t xlu[4] = {xv[j][0].x, xv[j][0].y, xv[j][0].z, xv[j][0].w}, xhu[4] = {xv[j][1].x, xv[j][1].y, xv[j][1].z, xv[j][1].w};
            const uint32_t wlu[4] = {wl.x, wl.y, wl.z, wl.w}, whu[4] = {wh.x, wh.y, wh.z, wh.w};
            float2 dot2 = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 gl = bf2_to_f2(glu[i]), gh = bf2_to_f2(ghu[i]);
                // transposed rotation (dy of the norm): dlo = g_lo*c_lo + g_hi*s_hi ; dhi = g_hi*c_hi - g_lo*s_lo
                const float2 dyl = ffma2(gl, t.c_lo[i], fmul2(gh, t.s_hi[i])), dyh = ffma2(gh, t.c_hi[i], fmul2(gl, t.ns_lo[i]));
                xlo[i] = fmul2(bf2_to_f2(xlu[i]), rs2);  // xhat
                xhi[i] = fmul2(bf2_to_f2(xhu[i]), rs2);
                const float2 rl = bf2_to_f2(f2_to_bf2(xlo[i].x, xlo[i].y)), rh = bf2_to_f2(f2_to_bf2(xhi[i].x, xhi[i].y));
                if (is_q) {
                    aq_lo[i] = ffma2(dyl, rl, aq_lo[i]);
                    aq_hi[i] = ffma2(dyh, rh, aq_hi[i]);
                } else {
                    ak_lo[i] = ffma2(dyl, rl, ak_lo[i]);
                    ak_hi[i] = ffma2(dyh, rh, ak_hi[i]);
                }
                dlo[i] = fmul2(dyl, bf2_to_f2(wlu[i]));  // g = dy * w
                dhi[i] = fmul2(dyh, bf2_to_f2(whu[i]));
                dot2 = ffma2(dlo[i], xlo[i], ffma2(dhi[i], xhi[i], dot2));
            }
            const float dot = group_sum<LPH>(dot2.x + dot2.y);
            if (ok) {
                const float nc = -dot * inv_d;
                const float2 nc2 = make_float2(nc, nc);
                float2 olo[4], ohi[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    olo[i] = fmul2(rs2, ffma2(xlo[i], nc2, dlo[i]));  // rs * (g - xhat * c)
                    ohi[i] = fmul2(rs2, ffma2(xhi[i], nc2, dhi[i]));
                }
                __nv_bfloat16* dst = (is_q ? dq_in : dk_in) + row * D;
                stg_stream(dst + sub * 8, pack4(olo));
                stg_stream(dst + HALF + sub * 8, pack4(ohi));
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        acc_s[grp][sub * 8 + 2 * i] = aq_lo[i].x;
        acc_s[grp][sub * 8 + 2 * i + 1] = aq_lo[i].y;
        acc_s[grp][HALF + sub * 8 + 2 * i] = aq_hi[i].x;
        acc_s[grp][HALF + sub * 8 + 2 * i + 1] = aq_hi[i].y;
        acc_s[grp][D + sub * 8 + 2 * i] = ak_lo[i].x;
        acc_s[grp][D + sub * 8 + 2 * i + 1] = ak_lo[i].y;
        acc_s[grp][D + HALF + sub * 8 + 2 * i] = ak_hi[i].x;
        acc_s[grp][D + HALF + sub * 8 + 2 * i + 1] = ak_hi[i].y;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * D; c += 256) {
        float s = 0.f;
#pragma unroll 4
        for (int g = 0; g < GROUPS; ++g) s += acc_s[g][c];
        dw_partial[(int64_t)blockIdx.x * 2 * D + c] = s;
    }
}

__global__ void __launch_bounds__(1024)
colsum2_kernel(const float* __restrict__ partial, float* __restrict__ out_a, float* __restrict__ out_b, int64_t nparts,
               int D) {
    __shared__ float red[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + x;
    float t = 0.f;
    griddep_wait();  // launched with programmatic stream serialization right behind qknorm_rope_bwd_kernel
    if (c < 2 * D)
        for (int64_t p = y; p < nparts; p += 256) {  // eight independent loads per round trip (see colsum_kernel)
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = p + 32 * j;
                v[j] = q < nparts ? partial[q * 2 * D + c] : 0.f;
            }
            t += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
    red[y][x] = t;
    __syncthreads();
    if (y == 0 && c < 2 * D) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += red[i][x];
        if (c < D) out_a[c] = a;
        else out_b[c - D] = a;
    }
}

static int items_grid(int64_t tokens, int H, int lph, int64_t max_ctas, int hc = kHeadChunk) {
    const int nchunks = (H + hc - 1) / hc;
    const int64_t items = tokens * nchunks;
    const int groups = 256 / lph;
    int64_t g = (items + groups - 1) / groups;
    if (g > max_ctas) g = max_ctas;
    return (int)(g < 1 ? 1 : g);
}

// Forward-side launch shape: heads per work item (2 or 4: 4 or 8 x 16-byte loads in flight per thread) and the grid cap in
// CTAs per SM (0 = one CTA per 32 lane groups of work, no grid-stride loop). VB200_ROPE_CFG="heads,ctas_per_sm" overrides
// the defaults for tuning runs (tools/hbm_sweep.sh).
struct RopeCfg {
    int hc, cps;
};
static RopeCfg rope_cfg() {
    static RopeCfg c = {0, 0};
    if (!c.hc) {
        int hc = 4, cps = 2;  // 4 heads / item, 2 CTAs per SM: profiles/r02_hbm_sweep.txt
        if (const char* e = getenv("VB200_ROPE_CFG")) sscanf(e, "%d,%d", &hc, &cps);
        c.hc = hc == 2 ? 2 : 4;
        c.cps = cps < 0 ? 2 : cps;
    }
    return c;
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
    const RopeCfg rc = rope_cfg();
    const int g = items_grid(tokens, q_heads + k_heads, lph, rc.cps ? (int64_t)rc.cps * kNumSMs : (int64_t)1 << 30, rc.hc);
#define GO(L)                                                                                            \
    do {                                                                                                 \
        if (rc.hc == 4) GO2(L, 4);                                                                       \
        else GO2(L, 2);                                                                                  \
    } while (0)
#define GO2(L, C)                                                                                        \
    rope_kernel<L, C><<<g, 256, 0, st>>>((const __nv_bfloat16*)q_in, (__nv_bfloat16*)q_out,              \
                                         (const __nv_bfloat16*)k_in, (__nv_bfloat16*)k_out,              \
                                         (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin, tokens,   \
                                         q_heads, k_heads, qs_t, qs_h, ks_t, ks_h, qos_t, qos_h, kos_t, kos_h, \
                                         inverse)
    if (lph == 4) GO(4);
    else if (lph == 8) GO(8);
    else GO(16);
#undef GO2
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
    const RopeCfg rc = rope_cfg();
    const int g = items_grid(tokens, q_heads + k_heads, lph, rc.cps ? (int64_t)rc.cps * kNumSMs : (int64_t)1 << 30, rc.hc);
#define GO(L)                                                                                           \
    do {                                                                                                \
        if (rc.hc == 4) GO2(L, 4);                                                                      \
        else GO2(L, 2);                                                                                 \
    } while (0)
#define GO2(L, C)                                                                                       \
    qknorm_rope_fwd_kernel<L, C><<<g, 256, 0, st>>>(                                                    \
        (const __nv_bfloat16*)q_in, (const __nv_bfloat16*)k_in, (const __nv_bfloat16*)wq,               \
        (const __nv_bfloat16*)wk, (const __nv_bfloat16*)cos, (const __nv_bfloat16*)sin,                 \
        (__nv_bfloat16*)q_out, (__nv_bfloat16*)k_out, rstd_q, rstd_k, tokens, q_heads, k_heads, eps)
    if (lph == 4) GO(4);
    else if (lph == 8) GO(8);
    else GO(16);
#undef GO2
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
    {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)((2 * head_dim + 31) / 32));
        cfg.blockDim = dim3(1024);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        VB_CUDA_TRY(cudaLaunchKernelEx(&cfg, colsum2_kernel, (const float*)dw_partial, dwq, dwk, (int64_t)g, (int)head_dim));
    }
    vb200_count_launch(2);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
