// RMSNorm forward / backward for sm_100a.
//
// Reference semantics: Qwen3RMSNorm.forward
//   (veomni/models/transformers/qwen3/generated/patched_modeling_qwen3_gpu.py:88-97)
//     x32 = x.float(); var = mean(x32^2); xhat = x32 * rsqrt(var + eps)
//     y   = w * xhat.to(bf16)                      (product rounded to bf16)
//   bound at run time to liger's LigerRMSNormFunction "llama" casting mode
//   (veomni/ops/liger/__init__.py:28-59), which has the same rounding points.
//
// Roofline: pure HBM stream. Forward moves 2*rows*cols*2 B (+cols*2 weight, +rows*4 rstd);
// backward reads dy,x and writes dx: 3*rows*cols*2 B.
//
// Forward, wide rows (cols > 256): persistent warp-per-row kernel. Each warp owns a ring of
// STAGES row buffers in shared memory filled by 1-D bulk-async copies (TMA engine, SASS UBLKCP)
// that complete on an mbarrier; the warp reduces sum(x^2) with shuffles, overwrites the row in
// place with the normalised bf16 values and hands the buffer back to the TMA engine with a
// bulk shared->global store. No register staging of the row, 2 rows in flight per warp, 12 warps per SM.
// Forward, narrow rows (cols <= 256, e.g. per-head q/k norm): a group of lanes per row.
// Backward: threads own columns and walk over rows, so dw accumulates in registers; one
// __syncthreads per row (double-buffered partials), deterministic two-pass dw reduction.
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

namespace vb {

// ------------------------------------------------------------------------------------------
// forward, bulk-async staged
// ------------------------------------------------------------------------------------------
constexpr int kFwdMaxWarps = 12;  // 12 warps x 2 rows in flight x 8 KB (H = 4096) fills the SM's shared memory

__global__ void __launch_bounds__(kFwdMaxWarps * 32, 1)
rmsnorm_fwd_bulk_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                        __nv_bfloat16* __restrict__ y, float* __restrict__ rstd, int64_t rows, int cols,
                        float eps, int warps_per_cta, int kFwdStages) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)cols * 2u;
    const int nvec = cols >> 3;
    // layout: [w row][warps * STAGES rows][mbarriers]
    uint4* w_s = reinterpret_cast<uint4*>(smem);
    uint8_t* bufs = smem + row_bytes;
    uint64_t* bars =
        reinterpret_cast<uint64_t*>(smem + (size_t)row_bytes * (1 + (size_t)warps_per_cta * kFwdStages));

    for (int v = threadIdx.x; v < nvec; v += blockDim.x) w_s[v] = reinterpret_cast<const uint4*>(w)[v];
    if (lane == 0) {
        for (int s = 0; s < kFwdStages; ++s) mbar_init(&bars[warp * kFwdStages + s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const int64_t gw = (int64_t)blockIdx.x * warps_per_cta + warp;
    const int64_t GW = (int64_t)gridDim.x * warps_per_cta;
    uint8_t* my_bufs = bufs + (size_t)warp * kFwdStages * row_bytes;
    uint64_t* my_bars = bars + warp * kFwdStages;

    if (lane == 0) {
        for (int s = 0; s < kFwdStages; ++s) {
            int64_t r = gw + (int64_t)s * GW;
            if (r < rows) {
                mbar_expect_tx(&my_bars[s], row_bytes);
                bulk_g2s(my_bufs + (size_t)s * row_bytes, x + r * cols, row_bytes, &my_bars[s]);
            }
        }
    }
    const float inv_cols = 1.0f / (float)cols;
    int it = 0, s = 0;
    uint32_t parity = 0;
    for (int64_t r = gw; r < rows; r += GW, ++it) {
        uint4* buf = reinterpret_cast<uint4*>(my_bufs + (size_t)s * row_bytes);
        // Refill the buffer stored one iteration ago BEFORE working on this row: the load of the next row then overlaps
        // this row's arithmetic (the store only has to have been read out of shared memory, which takes well under a
        // microsecond, not to have landed).
        if (lane == 0 && it >= 1) {
            bulk_wait_read<0>();
            const int ps = s == 0 ? kFwdStages - 1 : s - 1;
            const int64_t nr = r + (int64_t)(kFwdStages - 1) * GW;
            if (nr < rows) {
                mbar_expect_tx(&my_bars[ps], row_bytes);
                bulk_g2s(my_bufs + (size_t)ps * row_bytes, x + nr * cols, row_bytes, &my_bars[ps]);
            }
        }
        mbar_wait(&my_bars[s], parity);

        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        for (int v = lane; v < nvec; v += 32) sumsq8(buf[v], acc0, acc1);
        const float ss = warp_sum((acc0.x + acc0.y) + (acc1.x + acc1.y));
        const float rs = rsqrtf(ss * inv_cols + eps);
        const float2 rs2 = make_float2(rs, rs);
        for (int v = lane; v < nvec; v += 32) buf[v] = norm_scale8(buf[v], w_s[v], rs2);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) {
            rstd[r] = rs;
            bulk_s2g(y + r * cols, buf, row_bytes);
            bulk_commit();
        }
        __syncwarp();
        if (++s == kFwdStages) { s = 0; parity ^= 1u; }
    }
    if (lane == 0) bulk_wait_all<0>();
}

// ------------------------------------------------------------------------------------------
// forward, narrow rows: TPR lanes per row (TPR power of two <= 32), one 16-byte vector per lane
// ------------------------------------------------------------------------------------------
template <int TPR>
__global__ void __launch_bounds__(256)
rmsnorm_fwd_small_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w,
                         __nv_bfloat16* __restrict__ y, float* __restrict__ rstd, int64_t rows, int cols,
                         float eps) {
    const int nvec = cols >> 3;
    const int sub = threadIdx.x % TPR;
    const int64_t rows_per_iter = (int64_t)gridDim.x * (blockDim.x / TPR);
    const bool active = sub < nvec;
    float wf[8];
    if (active) unpack8(reinterpret_cast<const uint4*>(w)[sub], wf);
    const float inv_cols = 1.0f / (float)cols;
    // Loop bound is uniform per warp when rows_per_iter divides the padded row range; inactive
    // rows still take part in the shuffles with zeros.
    const int64_t first = (int64_t)blockIdx.x * (blockDim.x / TPR) + threadIdx.x / TPR;
    const int64_t padded = (rows + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
    for (int64_t r = first; r < padded; r += rows_per_iter) {
        const bool ok = active && r < rows;
        float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (ok) unpack8(ldg_stream(x + r * cols + sub * 8), f);
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
        ss = group_sum<TPR>(ss);
        const float rs = rsqrtf(ss * inv_cols + eps);
        if (ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) f[i] = wf[i] * round_bf16(f[i] * rs);
            stg_stream(y + r * cols + sub * 8, pack8(f));
            if (sub == 0) rstd[r] = rs;
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
// Wide rows: the whole CTA works on one row at a time; thread t owns vectors t, t+THREADS, ...
template <int THREADS, int V>
__global__ void __launch_bounds__(THREADS)
rmsnorm_bwd_wide_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                        const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                        __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_partial, int64_t rows,
                        int cols) {
    constexpr int NW = THREADS / 32;
    __shared__ float red[2][NW];
    const int nvec = cols >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wf[V][8], dwacc[V][8];
    bool act[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int v = threadIdx.x + k * THREADS;
        act[k] = v < nvec;
#pragma unroll
        for (int i = 0; i < 8; ++i) { wf[k][i] = 0.f; dwacc[k][i] = 0.f; }
        if (act[k]) unpack8(reinterpret_cast<const uint4*>(w)[v], wf[k]);
    }
    const float inv_cols = 1.0f / (float)cols;
    uint4 dyv[V], xv[V];
    int64_t r = blockIdx.x;
    if (r < rows) {
#pragma unroll
        for (int k = 0; k < V; ++k)
            if (act[k]) {
                const int64_t off = r * cols + (int64_t)(threadIdx.x + k * THREADS) * 8;
                dyv[k] = ldg_stream(dy + off);
                xv[k] = ldg_stream(x + off);
            }
    }
    int it = 0;
    for (; r < rows; r += gridDim.x, ++it) {
        const float rs = rstd[r];
        float g[V][8], xh[V][8];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (act[k]) {
                float d[8], xx[8];
                unpack8(dyv[k], d);
                unpack8(xv[k], xx);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xh[k][i] = xx[i] * rs;
                    g[k][i] = d[i] * wf[k][i];
                    dot = fmaf(g[k][i], xh[k][i], dot);
                    dwacc[k][i] = fmaf(d[i], round_bf16(xh[k][i]), dwacc[k][i]);
                }
            }
        }
        // prefetch the next row before the reduction barrier
        const int64_t nr = r + gridDim.x;
        if (nr < rows) {
#pragma unroll
            for (int k = 0; k < V; ++k)
                if (act[k]) {
                    const int64_t off = nr * cols + (int64_t)(threadIdx.x + k * THREADS) * 8;
                    dyv[k] = ldg_stream(dy + off);
                    xv[k] = ldg_stream(x + off);
                }
        }
        dot = warp_sum(dot);
        if (lane == 0) red[it & 1][warp] = dot;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) tot += red[it & 1][i];
        const float c = tot * inv_cols;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (act[k]) {
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = rs * (g[k][i] - xh[k][i] * c);
                stg_stream(dx + r * cols + (int64_t)(threadIdx.x + k * THREADS) * 8, pack8(o));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        if (act[k]) {
            float* dst = dw_partial + (int64_t)blockIdx.x * cols + (int64_t)(threadIdx.x + k * THREADS) * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(dwacc[k][0], dwacc[k][1], dwacc[k][2], dwacc[k][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(dwacc[k][4], dwacc[k][5], dwacc[k][6], dwacc[k][7]);
        }
    }
}

// ---- fused residual-add + RMSNorm (default in the decoder layer, tests/test_ops_gpu.py) -------------------
// Forward: h = bf16(x + residual) (the new residual stream, as the reference's `hidden_states = residual +
// hidden_states` rounds it), y = bf16(w * bf16(h * rstd)) — one pass over x and residual instead of an elementwise add
// kernel followed by the norm. One warp per row, the row's h kept packed in registers between the two sweeps;
// H = NCH * 256 elements.
template <int NCH, bool ADD>
__global__ void __launch_bounds__(256)
add_rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                       const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ h_out, __nv_bfloat16* __restrict__ y,
                       float* __restrict__ rstd, int64_t rows, float eps) {
    constexpr int H = NCH * 256;
    const int lane = threadIdx.x & 31;
    const int64_t wid = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nw = (int64_t)gridDim.x * 8;
    for (int64_t row = wid; row < rows; row += nw) {
        const __nv_bfloat16 *xr = x + row * H, *rr = res + row * H;
        uint4 hv[NCH];
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            uint4 h = ldg_stream(xr + (k * 32 + lane) * 8);
            if (ADD) {  // h = bf16(x + residual): fp32 add, one rounding (as torch adds two bf16 tensors)
                const uint4 a = h, b = ldg_stream(rr + (k * 32 + lane) * 8);
                float2 t;
                t = fadd2(bf2_to_f2(a.x), bf2_to_f2(b.x)); h.x = f2_to_bf2(t.x, t.y);
                t = fadd2(bf2_to_f2(a.y), bf2_to_f2(b.y)); h.y = f2_to_bf2(t.x, t.y);
                t = fadd2(bf2_to_f2(a.z), bf2_to_f2(b.z)); h.z = f2_to_bf2(t.x, t.y);
                t = fadd2(bf2_to_f2(a.w), bf2_to_f2(b.w)); h.w = f2_to_bf2(t.x, t.y);
                stg_stream(h_out + row * H + (k * 32 + lane) * 8, h);
            }
            sumsq8(h, acc0, acc1);
            hv[k] = h;
        }
        const float ss = warp_sum((acc0.x + acc0.y) + (acc1.x + acc1.y));
        const float rs = rsqrtf(ss * (1.0f / (float)H) + eps);
        if (lane == 0) rstd[row] = rs;
        const float2 rs2 = make_float2(rs, rs);
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const uint4 wv = *reinterpret_cast<const uint4*>(w + (k * 32 + lane) * 8);
            stg_stream(y + row * H + (k * 32 + lane) * 8, norm_scale8(hv[k], wv, rs2));
        }
    }
}

// Backward of the fused op: dx = dresidual = rmsnorm_bwd(dy; h, w, rstd) + dh, where dh is the gradient that reaches
// h through the residual stream (summed and rounded to bf16 as autograd's accumulation of two bf16 gradients does).
// Same structure as rmsnorm_bwd_wide_kernel, plus the dh stream.
template <int THREADS, int V>
__global__ void __launch_bounds__(THREADS)
rmsnorm_bwd_wide_add_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                            const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                            const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx,
                            float* __restrict__ dw_partial, int64_t rows, int cols) {
    constexpr int NW = THREADS / 32;
    __shared__ float red[2][NW];
    const int nvec = cols >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float wf[V][8], dwacc[V][8];
    bool act[V];
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int v = threadIdx.x + k * THREADS;
        act[k] = v < nvec;
#pragma unroll
        for (int i = 0; i < 8; ++i) { wf[k][i] = 0.f; dwacc[k][i] = 0.f; }
        if (act[k]) unpack8(reinterpret_cast<const uint4*>(w)[v], wf[k]);
    }
    const float inv_cols = 1.0f / (float)cols;
    int it = 0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x, ++it) {
        const float rs = rstd[r];
        float g[V][8], xh[V][8];
        uint4 dr[V];
        float dot = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (act[k]) {
                const int64_t off = r * cols + (int64_t)(threadIdx.x + k * THREADS) * 8;
                float d[8], xx[8];
                unpack8(ldg_stream(dy + off), d);
                unpack8(ldg_stream(x + off), xx);
                dr[k] = ldg_stream(dres + off);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    xh[k][i] = xx[i] * rs;
                    g[k][i] = d[i] * wf[k][i];
                    dot = fmaf(g[k][i], xh[k][i], dot);
                    dwacc[k][i] = fmaf(d[i], round_bf16(xh[k][i]), dwacc[k][i]);
                }
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) red[it & 1][warp] = dot;
        __syncthreads();
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) tot += red[it & 1][i];
        const float c = tot * inv_cols;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (act[k]) {
                float o[8], e[8];
                unpack8(dr[k], e);
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = round_bf16(rs * (g[k][i] - xh[k][i] * c)) + e[i];  // bf16 + bf16 -> bf16
                stg_stream(dx + r * cols + (int64_t)(threadIdx.x + k * THREADS) * 8, pack8(o));
            }
        }
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        if (act[k]) {
            float* dst = dw_partial + (int64_t)blockIdx.x * cols + (int64_t)(threadIdx.x + k * THREADS) * 8;
            *reinterpret_cast<float4*>(dst) = make_float4(dwacc[k][0], dwacc[k][1], dwacc[k][2], dwacc[k][3]);
            *reinterpret_cast<float4*>(dst + 4) = make_float4(dwacc[k][4], dwacc[k][5], dwacc[k][6], dwacc[k][7]);
        }
    }
}

// Wide rows, bulk-async staged (the default for cols <= 8192): as above the CTA owns one row at a time and thread t owns
// vector t, but the dy / x (/ dh) rows arrive through a ring of `stages` shared-memory slots filled by 1-D bulk copies
// (TMA engine, SASS UBLKCP) that complete on one mbarrier per slot, so `stages` rows of loads are in flight per CTA
// instead of the one row a register prefetch could hold (the register version reached 0.40-0.53 of the HBM stream).
// The row's reduction barrier doubles as the slot's "consumed" signal: thread 0 refills it right after. Arithmetic is
// on packed fp32 pairs (FMUL2/FFMA2); dx leaves through 16-byte streaming stores.
constexpr int kBwdMaxStages = 8;

template <int THREADS, bool ADD>
__global__ void __launch_bounds__(THREADS)
rmsnorm_bwd_ring_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                        const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                        const __nv_bfloat16* __restrict__ dres, __nv_bfloat16* __restrict__ dx,
                        float* __restrict__ dw_partial, int64_t rows, int cols, int stages) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NW = THREADS / 32;
    constexpr int NSRC = ADD ? 3 : 2;
    __shared__ float red[2][NW];
    __shared__ __align__(8) uint64_t full[kBwdMaxStages];
    griddep_launch_dependents();  // the column-sum grid may be scheduled early; it waits for this grid's results
    const int nvec = cols >> 3;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t row_bytes = (uint32_t)cols * 2u, stage_bytes = NSRC * row_bytes;
    const bool act = (int)threadIdx.x < nvec;
    const int64_t G = gridDim.x;
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();
    auto issue = [&](int s, int64_t r) {
        uint8_t* dst = smem + (size_t)s * stage_bytes;
        mbar_expect_tx(&full[s], stage_bytes);
        bulk_g2s(dst, dy + r * cols, row_bytes, &full[s]);
        bulk_g2s(dst + row_bytes, x + r * cols, row_bytes, &full[s]);
        if (ADD) bulk_g2s(dst + 2 * row_bytes, dres + r * cols, row_bytes, &full[s]);
    };
    if (threadIdx.x == 0)
        for (int s = 0; s < stages; ++s) {
            const int64_t r = blockIdx.x + (int64_t)s * G;
            if (r < rows) issue(s, r);
        }
    float2 wf[4], dwacc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) wf[i] = dwacc[i] = make_float2(0.f, 0.f);
    if (act) {
        const uint4 wv = reinterpret_cast<const uint4*>(w)[threadIdx.x];
        wf[0] = bf2_to_f2(wv.x); wf[1] = bf2_to_f2(wv.y); wf[2] = bf2_to_f2(wv.z); wf[3] = bf2_to_f2(wv.w);
    }
    const float inv_cols = 1.0f / (float)cols;
    int it = 0, s = 0;
    uint32_t parity = 0;
    for (int64_t r = blockIdx.x; r < rows; r += G, ++it) {
        const float rs = rstd[r];
        mbar_wait(&full[s], parity);
        const uint4* slot = reinterpret_cast<const uint4*>(smem + (size_t)s * stage_bytes);
        float2 g[4], xh[4];
        uint4 dr = make_uint4(0u, 0u, 0u, 0u);
        float2 dot2 = make_float2(0.f, 0.f);
        if (act) {
            const uint4 dv = slot[threadIdx.x], xv = slot[nvec + threadIdx.x];
            if (ADD) dr = slot[2 * nvec + threadIdx.x];
            const uint32_t du[4] = {dv.x, dv.y, dv.z, dv.w}, xu[4] = {xv.x, xv.y, xv.z, xv.w};
            const float2 rs2 = make_float2(rs, rs);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 d = bf2_to_f2(du[i]);
                xh[i] = fmul2(bf2_to_f2(xu[i]), rs2);
                g[i] = fmul2(d, wf[i]);
                dot2 = ffma2(g[i], xh[i], dot2);
                dwacc[i] = ffma2(d, bf2_to_f2(f2_to_bf2(xh[i].x, xh[i].y)), dwacc[i]);  // dy * bf16(xhat)
            }
        }
        float dot = warp_sum(dot2.x + dot2.y);
        if (lane == 0) red[it & 1][warp] = dot;
        __syncthreads();  // every thread has read slot s
        if (threadIdx.x == 0) {
            const int64_t nr = r + (int64_t)stages * G;
            if (nr < rows) issue(s, nr);
        }
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < NW; ++i) tot += red[it & 1][i];
        if (act) {
            const float2 nc2 = make_float2(-tot * inv_cols, -tot * inv_cols), rs2 = make_float2(rs, rs);
            uint32_t o[4];
            const uint32_t e[4] = {dr.x, dr.y, dr.z, dr.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 t = fmul2(rs2, ffma2(xh[i], nc2, g[i]));  // rs * (g - xhat * c)
                o[i] = f2_to_bf2(t.x, t.y);
                if (ADD) {  // bf16(bf16(dx) + dh): autograd's accumulation of two bf16 gradients
                    const float2 u = fadd2(bf2_to_f2(o[i]), bf2_to_f2(e[i]));
                    o[i] = f2_to_bf2(u.x, u.y);
                }
            }
            stg_stream(dx + r * cols + (int64_t)threadIdx.x * 8, make_uint4(o[0], o[1], o[2], o[3]));
        }
        if (++s == stages) { s = 0; parity ^= 1u; }
    }
    if (act) {
        float* dst = dw_partial + (int64_t)blockIdx.x * cols + (int64_t)threadIdx.x * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(dwacc[0].x, dwacc[0].y, dwacc[1].x, dwacc[1].y);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(dwacc[2].x, dwacc[2].y, dwacc[3].x, dwacc[3].y);
    }
}

// Narrow rows: TPR lanes per row, CTA of 256 threads handles 256/TPR rows per iteration.
template <int TPR>
__global__ void __launch_bounds__(256)
rmsnorm_bwd_small_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                         const __nv_bfloat16* __restrict__ w, const float* __restrict__ rstd,
                         __nv_bfloat16* __restrict__ dx, float* __restrict__ dw_partial, int64_t rows,
                         int cols) {
    constexpr int GROUPS = 256 / TPR;
    __shared__ float acc_s[GROUPS][TPR * 8 + 1];
    const int nvec = cols >> 3;
    const int sub = threadIdx.x % TPR, grp = threadIdx.x / TPR;
    const bool active = sub < nvec;
    float wf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, dwacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active) unpack8(reinterpret_cast<const uint4*>(w)[sub], wf);
    const float inv_cols = 1.0f / (float)cols;
    const int64_t rows_per_iter = (int64_t)gridDim.x * GROUPS;
    const int64_t first = (int64_t)blockIdx.x * GROUPS + grp;
    const int64_t padded = (rows + rows_per_iter - 1) / rows_per_iter * rows_per_iter;
    for (int64_t r = first; r < padded; r += rows_per_iter) {
        const bool ok = active && r < rows;
        float d[8] = {0, 0, 0, 0, 0, 0, 0, 0}, xx[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        float rs = 0.f;
        if (ok) {
            unpack8(ldg_stream(dy + r * cols + sub * 8), d);
            unpack8(ldg_stream(x + r * cols + sub * 8), xx);
            rs = rstd[r];
        }
        float g[8], xh[8], dot = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            xh[i] = xx[i] * rs;
            g[i] = d[i] * wf[i];
            dot = fmaf(g[i], xh[i], dot);
            dwacc[i] = fmaf(d[i], round_bf16(xh[i]), dwacc[i]);
        }
        dot = group_sum<TPR>(dot);
        const float c = dot * inv_cols;
        if (ok) {
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = rs * (g[i] - xh[i] * c);
            stg_stream(dx + r * cols + sub * 8, pack8(o));
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc_s[grp][sub * 8 + i] = dwacc[i];
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += 256) {
        float t = 0.f;
#pragma unroll 4
        for (int gI = 0; gI < GROUPS; ++gI) t += acc_s[gI][c];
        dw_partial[(int64_t)blockIdx.x * cols + c] = t;
    }
}

// out[c] = sum_p partial[p][c].  Block = 32 columns x 32 partial-lanes (1024 threads), grid = cols / 32: every thread sums
// nparts / 32 rows with its loads issued eight at a time (the 256-thread version walked 37 dependent round trips and cost
// 5.5 us per call — a fifth of the backward it follows), coalesced 128-byte rows, combined in a fixed order
// (deterministic).
__global__ void __launch_bounds__(1024)
colsum_kernel(const float* __restrict__ partial, float* __restrict__ out, int64_t nparts, int cols) {
    __shared__ float red[32][33];
    const int x = threadIdx.x & 31, y = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + x;
    float t = 0.f;
    griddep_wait();  // launched with programmatic stream serialization right behind the kernel that writes `partial`
    if (c < cols)
        for (int64_t p = y; p < nparts; p += 256) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int64_t q = p + 32 * j;
                v[j] = q < nparts ? partial[q * cols + c] : 0.f;
            }
            t += ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        }
    red[y][x] = t;
    __syncthreads();
    if (y == 0 && c < cols) {
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) a += red[i][x];
        out[c] = a;
    }
}

// Wide-row backward configuration: threads per CTA (one 16-byte vector per thread), CTAs per SM and ring depth.
struct BwdCfg {
    int threads, ctas_per_sm, stages;
    size_t smem;
};
static BwdCfg bwd_cfg(int cols, bool add) {
    const int nvec = cols >> 3;
    BwdCfg c;
    c.threads = nvec <= 128 ? 128 : nvec <= 256 ? 256 : nvec <= 512 ? 512 : 1024;
    c.ctas_per_sm = 1024 / c.threads;  // 64 registers per thread: 1024 threads per SM
    if (c.ctas_per_sm > 8) c.ctas_per_sm = 8;
    if (c.ctas_per_sm < 1) c.ctas_per_sm = 1;
    static int max_stages = 0;
    if (!max_stages) {
        const char* e = getenv("VB200_RMS_BWD_STAGES");  // tuning override
        max_stages = e ? atoi(e) : 2;  // deeper rings measured slower on 4096 x 4096 (profiles/r02_hbm_sweep.txt)
        if (max_stages < 2 || max_stages > kBwdMaxStages) max_stages = 2;
    }
    const size_t stage = (size_t)(add ? 3 : 2) * cols * 2;
    const int st = (int)((200 * 1024 / c.ctas_per_sm) / stage);
    c.stages = st > max_stages ? max_stages : st < 2 ? 2 : st;
    c.smem = stage * c.stages;
    return c;
}

static int bwd_grid(int64_t rows, int cols) {
    const int nvec = cols >> 3;
    int64_t g;
    if (nvec <= 32) {
        int tpr = 1;
        while (tpr < nvec) tpr <<= 1;
        const int groups = 256 / tpr;
        g = (rows + groups - 1) / groups;
        if (g > 4 * kNumSMs) g = 4 * kNumSMs;
    } else if (nvec <= 1024) {
        const int64_t cap = (int64_t)bwd_cfg(cols, false).ctas_per_sm * kNumSMs;  // the fused-add variant uses the same grid
        g = rows < cap ? rows : cap;
    } else {
        g = rows < 2 * kNumSMs ? rows : 2 * kNumSMs;
    }
    return (int)(g < 1 ? 1 : g);
}

// out = column sums of the [nparts, cols] partials, launched as a programmatic dependent of the kernel before it.
static cudaError_t launch_colsum(const float* partial, float* out, int64_t nparts, int cols, cudaStream_t st) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((cols + 31) / 32));
    cfg.blockDim = dim3(1024);
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, colsum_kernel, partial, out, nparts, cols);
}

template <int THREADS, bool ADD>
static cudaError_t launch_bwd_ring(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                                   float* dw_partial, int64_t rows, int cols, int g, cudaStream_t st) {
    const BwdCfg c = bwd_cfg(cols, ADD);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(rmsnorm_bwd_ring_kernel<THREADS, ADD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    rmsnorm_bwd_ring_kernel<THREADS, ADD><<<g, THREADS, c.smem, st>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, rstd, (const __nv_bfloat16*)dres,
        (__nv_bfloat16*)dx, dw_partial, rows, cols, c.stages);
    return cudaGetLastError();
}

template <bool ADD>
static cudaError_t launch_bwd_ring_any(const void* dy, const void* x, const void* w, const float* rstd, const void* dres,
                                       void* dx, float* dw_partial, int64_t rows, int cols, int g, cudaStream_t st) {
    switch (bwd_cfg(cols, ADD).threads) {
        case 128: return launch_bwd_ring<128, ADD>(dy, x, w, rstd, dres, dx, dw_partial, rows, cols, g, st);
        case 256: return launch_bwd_ring<256, ADD>(dy, x, w, rstd, dres, dx, dw_partial, rows, cols, g, st);
        case 512: return launch_bwd_ring<512, ADD>(dy, x, w, rstd, dres, dx, dw_partial, rows, cols, g, st);
        default: return launch_bwd_ring<1024, ADD>(dy, x, w, rstd, dres, dx, dw_partial, rows, cols, g, st);
    }
}

}  // namespace vb

using namespace vb;

template <int TPR>
static void launch_fwd_small(const void* x, const void* w, void* y, float* rstd, int64_t rows, int cols,
                             float eps, cudaStream_t st) {
    const int groups = 256 / TPR;
    int64_t g = (rows + groups - 1) / groups;
    if (g > 8 * kNumSMs) g = 8 * kNumSMs;
    rmsnorm_fwd_small_kernel<TPR><<<(int)g, 256, 0, st>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, rstd, rows, cols, eps);
}

extern "C" int vb200_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, int64_t rows,
                                 int64_t cols, float eps, void* stream) {
    if (rows < 0 || cols <= 0 || (cols & 7) || cols > 16384)
        return vb200_set_error(VB200_EINVAL, "rmsnorm_fwd: cols must be a multiple of 8 in (0,16384]");
    if (rows == 0) return VB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int nvec = (int)(cols >> 3);
    if (nvec <= 32) {
        if (nvec <= 1) launch_fwd_small<1>(x, w, y, rstd, rows, (int)cols, eps, st);
        else if (nvec <= 2) launch_fwd_small<2>(x, w, y, rstd, rows, (int)cols, eps, st);
        else if (nvec <= 4) launch_fwd_small<4>(x, w, y, rstd, rows, (int)cols, eps, st);
        else if (nvec <= 8) launch_fwd_small<8>(x, w, y, rstd, rows, (int)cols, eps, st);
        else if (nvec <= 16) launch_fwd_small<16>(x, w, y, rstd, rows, (int)cols, eps, st);
        else launch_fwd_small<32>(x, w, y, rstd, rows, (int)cols, eps, st);
    } else if ((cols == 1024 || cols == 2048 || cols == 4096 || cols == 5120) && !getenv("VB200_RMS_FWD_CFG")) {
        // The decoder hidden sizes: a warp keeps its row packed in registers between the two sweeps (no shared-memory
        // staging, no per-launch barrier set-up) — 16.6 -> see profiles/r02_microbench_hbm_final.jsonl at 4096 x 4096; the
        // bulk-staged kernel below keeps the general widths.
        int64_t blocks = (rows + 7) / 8;
        if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
#define GO(N)                                                                                                          \
    add_rmsnorm_fwd_kernel<N, false><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, nullptr,               \
                                                                       (const __nv_bfloat16*)w, nullptr,               \
                                                                       (__nv_bfloat16*)y, rstd, rows, eps)
        if (cols == 1024) GO(4);
        else if (cols == 2048) GO(8);
        else if (cols == 4096) GO(16);
        else GO(20);
#undef GO
    } else {
        const size_t row_bytes = (size_t)cols * 2;
        // (warps per CTA, rows in flight per warp): tuned on 4096 x 4096 (tools/hbm_sweep.sh); VB200_RMS_FWD_CFG="warps,stages"
        // overrides for tuning runs.
        static int cfg_warps = 0, cfg_stages = 0;
        if (!cfg_warps) {
            int w_ = kFwdMaxWarps, s_ = 2;
            if (const char* e = getenv("VB200_RMS_FWD_CFG")) sscanf(e, "%d,%d", &w_, &s_);
            cfg_warps = w_ < 1 ? 1 : w_ > kFwdMaxWarps ? kFwdMaxWarps : w_;
            cfg_stages = s_ < 2 ? 2 : s_ > 8 ? 8 : s_;
        }
        const int stages = cfg_stages;
        int warps = (int)((200 * 1024 - row_bytes) / (row_bytes * stages));
        if (warps > cfg_warps) warps = cfg_warps;
        if (warps < 1) return vb200_set_error(VB200_EINVAL, "rmsnorm_fwd: row too wide for smem staging");
        const size_t smem = row_bytes * (1 + (size_t)warps * stages) + sizeof(uint64_t) * warps * stages;
        static bool attr_set = false;
        if (!attr_set) {
            VB_CUDA_TRY(cudaFuncSetAttribute(rmsnorm_fwd_bulk_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            attr_set = true;
        }
        int64_t g = (rows + warps - 1) / warps;
        if (g > kNumSMs) g = kNumSMs;
        rmsnorm_fwd_bulk_kernel<<<(int)g, warps * 32, smem, st>>>(
            (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (__nv_bfloat16*)y, rstd, rows, (int)cols, eps,
            warps, stages);
    }
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int64_t vb200_rmsnorm_bwd_partials(int64_t rows, int64_t cols) {
    if (cols <= 0 || (cols & 7)) return 0;
    return bwd_grid(rows, (int)cols);
}

extern "C" int vb200_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx,
                                 float* dw_partial, float* dw, int64_t rows, int64_t cols, void* stream) {
    if (rows < 0 || cols <= 0 || (cols & 7) || cols > 16384)
        return vb200_set_error(VB200_EINVAL, "rmsnorm_bwd: cols must be a multiple of 8 in (0,16384]");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        VB_CUDA_TRY(cudaMemsetAsync(dw, 0, sizeof(float) * cols, st));
        return VB200_OK;
    }
    const int nvec = (int)(cols >> 3);
    const int g = bwd_grid(rows, (int)cols);
    const __nv_bfloat16 *dy_ = (const __nv_bfloat16*)dy, *x_ = (const __nv_bfloat16*)x,
                        *w_ = (const __nv_bfloat16*)w;
    __nv_bfloat16* dx_ = (__nv_bfloat16*)dx;
#define SMALL(T) rmsnorm_bwd_small_kernel<T><<<g, 256, 0, st>>>(dy_, x_, w_, rstd, dx_, dw_partial, rows, (int)cols)
    if (nvec <= 1) SMALL(1);
    else if (nvec <= 2) SMALL(2);
    else if (nvec <= 4) SMALL(4);
    else if (nvec <= 8) SMALL(8);
    else if (nvec <= 16) SMALL(16);
    else if (nvec <= 32) SMALL(32);
    else if (nvec <= 1024) VB_CUDA_TRY(launch_bwd_ring_any<false>(dy, x, w, rstd, nullptr, dx, dw_partial, rows, (int)cols, g, st));
    else rmsnorm_bwd_wide_kernel<512, 4><<<g, 512, 0, st>>>(dy_, x_, w_, rstd, dx_, dw_partial, rows, (int)cols);
#undef SMALL
    VB_HOST_CHECK_LAUNCH();
    VB_CUDA_TRY(launch_colsum(dw_partial, dw, g, (int)cols, st));
    vb200_count_launch(2);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_add_rmsnorm_fwd(const void* x, const void* residual, const void* w, void* h_out, void* y, float* rstd,
                                     int64_t rows, int64_t cols, float eps, void* stream) {
    if (rows < 0 || !x || !residual || !w || !h_out || !y || !rstd) return vb200_set_error(VB200_EINVAL, "add_rmsnorm_fwd: bad arguments");
    if (rows == 0) return VB200_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (rows + 7) / 8;
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
#define GO(N)                                                                                                               \
    add_rmsnorm_fwd_kernel<N, true><<<(unsigned)blocks, 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)residual, \
                                                                      (const __nv_bfloat16*)w, (__nv_bfloat16*)h_out,         \
                                                                      (__nv_bfloat16*)y, rstd, rows, eps)
    switch (cols) {
        case 1024: GO(4); break;
        case 2048: GO(8); break;
        case 4096: GO(16); break;
        case 5120: GO(20); break;
        case 8192: GO(32); break;
        default: return vb200_set_error(VB200_EINVAL, "add_rmsnorm_fwd: hidden size must be 1024, 2048, 4096, 5120 or 8192");
    }
#undef GO
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_rmsnorm_bwd_add(const void* dy, const void* x, const void* w, const float* rstd, const void* dres, void* dx,
                                     float* dw_partial, float* dw, int64_t rows, int64_t cols, void* stream) {
    if (!dres) return vb200_rmsnorm_bwd(dy, x, w, rstd, dx, dw_partial, dw, rows, cols, stream);
    if (rows < 0 || cols < 264 || (cols & 7) || cols > 16384)
        return vb200_set_error(VB200_EINVAL, "rmsnorm_bwd_add: cols must be a multiple of 8 in [264,16384]");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) {
        VB_CUDA_TRY(cudaMemsetAsync(dw, 0, sizeof(float) * cols, st));
        return VB200_OK;
    }
    const int nvec = (int)(cols >> 3);
    const int g = bwd_grid(rows, (int)cols);
    if (nvec <= 1024)
        VB_CUDA_TRY(launch_bwd_ring_any<true>(dy, x, w, rstd, dres, dx, dw_partial, rows, (int)cols, g, st));
    else
        rmsnorm_bwd_wide_add_kernel<512, 4><<<g, 512, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                             (const __nv_bfloat16*)w, rstd, (const __nv_bfloat16*)dres,
                                                             (__nv_bfloat16*)dx, dw_partial, rows, (int)cols);
    VB_HOST_CHECK_LAUNCH();
    VB_CUDA_TRY(launch_colsum(dw_partial, dw, g, (int)cols, st));
    vb200_count_launch(2);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
