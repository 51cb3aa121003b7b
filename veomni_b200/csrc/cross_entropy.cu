// Row-wise softmax cross-entropy for large vocabularies (V = 151936 for Qwen3), forward and gradient.
//
// Replaces the arithmetic of eager_cross_entropy / fixed_cross_entropy
// (veomni/ops/kernels/cross_entropy/eager.py:23-38) and of the liger fused-linear-cross-entropy element kernel
// (veomni/ops/kernels/cross_entropy/liger.py) on the logits of one row chunk:
//     loss_row = logsumexp(x) - x[label]                 (0 when label == ignore_index)
//     grad     = (softmax(x) - onehot(label)) * scale    (0 for ignored rows)
//
// HBM-bound. One CTA of 1024 threads owns a row: pass 1 streams the row once and keeps a per-thread online
// (max, sum) pair — one exp2 per element plus one per 8-element vector for the running rescale; pass 2 re-reads
// the row and writes the gradient over it (in place when grad == logits). With 2 CTAs per SM, 296 rows of 304 KB
// (bf16) are in flight = 90 MB, which the 126 MB L2 holds, so the second read is an L2 hit and DRAM sees the
// algorithmic minimum: one read and one write of the chunk. fp32 rows (608 KB) spill L2 and pay a second DRAM read.
#include "common.cuh"

namespace vb {

constexpr int CE_THREADS = 1024;
constexpr float kLog2eCe = 1.4426950408889634f;
constexpr float kLn2Ce = 0.6931471805599453f;

template <typename T>
struct CeVec;
template <>
struct CeVec<__nv_bfloat16> {
    static constexpr int N = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 u = *reinterpret_cast<const uint4*>(p);
        float2 a = bf2_to_f2(u.x), b = bf2_to_f2(u.y), c = bf2_to_f2(u.z), d = bf2_to_f2(u.w);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint4 u;
        u.x = f2_to_bf2(v[0], v[1]); u.y = f2_to_bf2(v[2], v[3]); u.z = f2_to_bf2(v[4], v[5]); u.w = f2_to_bf2(v[6], v[7]);
        *reinterpret_cast<uint4*>(p) = u;
    }
};
template <>
struct CeVec<float> {
    static constexpr int N = 4;
    __device__ static void load(const float* p, float (&v)[4]) {
        const float4 u = *reinterpret_cast<const float4*>(p);
        v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};

__device__ __forceinline__ float ce_to_float(__nv_bfloat16 x) { return __bfloat162float(x); }
__device__ __forceinline__ float ce_to_float(float x) { return x; }
__device__ __forceinline__ void ce_from_float(__nv_bfloat16* p, float x) { *p = __float2bfloat16_rn(x); }
__device__ __forceinline__ void ce_from_float(float* p, float x) { *p = x; }

// x: [rows, vocab] with row stride `stride` (elements). `lse` is an output when lse_given == 0 and an input
// otherwise (backward-only call). `grad` may be null (forward only) or alias `x`.
template <typename T>
__global__ void __launch_bounds__(CE_THREADS, 2)
cross_entropy_kernel(const T* __restrict__ x, int64_t stride, int64_t vocab, const int64_t* __restrict__ labels,
                     int64_t ignore_index, float* __restrict__ loss_rows, float* __restrict__ lse, int lse_given,
                     T* grad, int64_t grad_stride, float scale_host, const float* __restrict__ scale_dev,
                     const float* __restrict__ upstream) {
    constexpr int N = CeVec<T>::N;
    __shared__ float red_m[32], red_s[32];
    __shared__ float row_lse2;
    const int64_t row = blockIdx.x;
    const T* xr = x + row * stride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t label = labels[row];
    const bool ignored = label == ignore_index;
    const bool in_range = label >= 0 && label < vocab;  // anything else (and not ignore_index) yields a NaN loss
    // rows need not be 16-byte aligned relative to the vector width if stride % N != 0: peel a scalar head
    const int64_t mis = (reinterpret_cast<uintptr_t>(xr) & 15) / sizeof(T);
    const int64_t head = mis ? ((vocab < (int64_t)(N - mis)) ? vocab : (int64_t)(N - mis)) : 0;  // N == 16 / sizeof(T)
    const int64_t nvec = (vocab - head) / N;
    const int64_t tail0 = head + nvec * N;

    float lse2;  // log2-domain logsumexp: lse2 = log2(sum exp(x)) = max*log2e + log2(sum exp2((x - max)*log2e))
    if (!lse_given) {
        float m = -INFINITY, s = 0.f;  // running max (already multiplied by log2e) and sum of exp2(x*log2e - m)
        auto fold = [&](float v) {
            const float y = v * kLog2eCe;
            if (y == -INFINITY) return;
            if (y > m) {
                s = s * exp2f(m - y) + 1.f;
                m = y;
            } else {
                s += exp2f(y - m);
            }
        };
        for (int64_t i = tid; i < head; i += CE_THREADS) fold(ce_to_float(xr[i]));
        for (int64_t i = tail0 + tid; i < vocab; i += CE_THREADS) fold(ce_to_float(xr[i]));
        auto fold_vec = [&](const float (&v)[N]) {
            float vm = v[0];
#pragma unroll
            for (int e = 1; e < N; ++e) vm = fmaxf(vm, v[e]);
            vm *= kLog2eCe;
            if (vm > m) {  // one rescale per vector instead of per element
                s *= exp2f(m - vm);
                m = vm;
            }
            if (m != -INFINITY) {
#pragma unroll
                for (int e = 0; e < N; ++e) s += exp2f(fmaf(v[e], kLog2eCe, -m));
            }
        };
        int64_t j = tid;
        for (; j + CE_THREADS < nvec; j += 2 * CE_THREADS) {  // two 16-byte loads in flight per thread
            float v0[N], v1[N];
            CeVec<T>::load(xr + head + j * N, v0);
            CeVec<T>::load(xr + head + (j + CE_THREADS) * N, v1);
            fold_vec(v0);
            fold_vec(v1);
        }
        if (j < nvec) {
            float v0[N];
            CeVec<T>::load(xr + head + j * N, v0);
            fold_vec(v0);
        }
        // block reduction of (m, s) pairs, fixed order -> deterministic
        float wm = warp_max(m);
        float ws = (m == -INFINITY) ? 0.f : s * exp2f(m - wm);
        ws = warp_sum(ws);
        if (lane == 0) {
            red_m[warp] = wm;
            red_s[warp] = ws;
        }
        __syncthreads();
        if (warp == 0) {
            const float pm = red_m[lane], ps = red_s[lane];  // CE_THREADS / 32 == 32 warps
            const float bm = warp_max(pm);
            float bs = (pm == -INFINITY) ? 0.f : ps * exp2f(pm - bm);
            bs = warp_sum(bs);
            if (lane == 0) {
                const float l2 = bm + log2f(bs);
                row_lse2 = l2;
                const float l = l2 * kLn2Ce;
                if (lse) lse[row] = l;
                if (loss_rows) loss_rows[row] = ignored ? 0.f : (in_range ? l - ce_to_float(xr[label]) : NAN);
            }
        }
        __syncthreads();
        lse2 = row_lse2;
    } else {
        lse2 = lse[row] * kLog2eCe;
    }
    if (grad == nullptr) return;

    float scale = scale_host;
    if (scale_dev) scale *= *scale_dev;
    if (upstream) scale *= *upstream;
    T* gr = grad + row * grad_stride;
    if (ignored) {
        float z[N];
#pragma unroll
        for (int e = 0; e < N; ++e) z[e] = 0.f;
        for (int64_t i = tid; i < head; i += CE_THREADS) ce_from_float(gr + i, 0.f);
        for (int64_t i = tail0 + tid; i < vocab; i += CE_THREADS) ce_from_float(gr + i, 0.f);
        for (int64_t j = tid; j < nvec; j += CE_THREADS) CeVec<T>::store(gr + head + j * N, z);
        return;
    }
    for (int64_t i = tid; i < head; i += CE_THREADS) {
        const float p = exp2f(fmaf(ce_to_float(xr[i]), kLog2eCe, -lse2));
        ce_from_float(gr + i, (p - (i == label ? 1.f : 0.f)) * scale);
    }
    for (int64_t i = tail0 + tid; i < vocab; i += CE_THREADS) {
        const float p = exp2f(fmaf(ce_to_float(xr[i]), kLog2eCe, -lse2));
        ce_from_float(gr + i, (p - (i == label ? 1.f : 0.f)) * scale);
    }
    const int64_t lj = (in_range && label >= head && label < tail0) ? (label - head) / N : -1;
    const int le = (int)((label - head) % N);
    auto grad_vec = [&](int64_t j, float (&v)[N]) {
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] = exp2f(fmaf(v[e], kLog2eCe, -lse2));
        if (j == lj) {
#pragma unroll
            for (int e = 0; e < N; ++e)
                if (e == le) v[e] -= 1.f;
        }
#pragma unroll
        for (int e = 0; e < N; ++e) v[e] *= scale;
        CeVec<T>::store(gr + head + j * N, v);
    };
    int64_t j = tid;
    for (; j + CE_THREADS < nvec; j += 2 * CE_THREADS) {
        float v0[N], v1[N];
        CeVec<T>::load(xr + head + j * N, v0);
        CeVec<T>::load(xr + head + (j + CE_THREADS) * N, v1);
        grad_vec(j, v0);
        grad_vec(j + CE_THREADS, v1);
    }
    if (j < nvec) {
        float v0[N];
        CeVec<T>::load(xr + head + j * N, v0);
        grad_vec(j, v0);
    }
}

// 1 / count(labels != ignore_index) as a device scalar (0 valid labels -> 0, so every gradient is 0 and the
// loss sum times it is 0, matching "mean over nothing" being masked out by the callers' token weighting).
__global__ void ce_valid_recip_kernel(const int64_t* __restrict__ labels, int64_t n, int64_t ignore_index,
                                      float* __restrict__ out) {
    __shared__ int red[32];
    int c = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) c += labels[i] != ignore_index;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x < 32) {
        c = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if (threadIdx.x == 0) {
            out[0] = c > 0 ? 1.f / (float)c : 0.f;
            out[1] = (float)c;
        }
    }
}

}  // namespace vb

using namespace vb;

extern "C" int vb200_cross_entropy(const void* logits, int32_t dtype, int64_t rows, int64_t vocab, int64_t row_stride,
                                   const int64_t* labels, int64_t ignore_index, float* loss_rows, float* lse,
                                   int32_t lse_given, void* grad, int64_t grad_stride, float scale,
                                   const float* scale_dev, const float* upstream, void* stream) {
    if (rows == 0) return VB200_OK;
    if (rows < 0 || vocab <= 0 || !logits || !labels) return vb200_set_error(VB200_EINVAL, "cross_entropy: bad arguments");
    if (dtype != 0 && dtype != 1) return vb200_set_error(VB200_EINVAL, "cross_entropy: dtype must be 0 (bf16) or 1 (f32)");
    if (lse_given && !lse) return vb200_set_error(VB200_EINVAL, "cross_entropy: lse_given without lse");
    if (!lse_given && !grad && !loss_rows && !lse) return vb200_set_error(VB200_EINVAL, "cross_entropy: nothing to compute");
    const size_t esz = dtype == 0 ? 2 : 4;
    // the gradient row must share the logits row's 16-byte phase so both use the same head/vector split
    if (grad && (((uintptr_t)grad ^ (uintptr_t)logits) & 15) != 0)
        return vb200_set_error(VB200_EINVAL, "cross_entropy: grad and logits must have the same 16-byte alignment phase");
    if (grad && ((grad_stride - row_stride) * (int64_t)esz) % 16 != 0)
        return vb200_set_error(VB200_EINVAL, "cross_entropy: grad/logits row strides must differ by a multiple of 16 bytes");
    if (rows > 0x7fffffffLL) return vb200_set_error(VB200_EINVAL, "cross_entropy: too many rows");
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == 0)
        cross_entropy_kernel<__nv_bfloat16><<<(unsigned)rows, CE_THREADS, 0, s>>>(
            (const __nv_bfloat16*)logits, row_stride, vocab, labels, ignore_index, loss_rows, lse, lse_given,
            (__nv_bfloat16*)grad, grad_stride, scale, scale_dev, upstream);
    else
        cross_entropy_kernel<float><<<(unsigned)rows, CE_THREADS, 0, s>>>(
            (const float*)logits, row_stride, vocab, labels, ignore_index, loss_rows, lse, lse_given, (float*)grad,
            grad_stride, scale, scale_dev, upstream);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_count_valid_labels(const int64_t* labels, int64_t n, int64_t ignore_index, float* out2,
                                        void* stream) {
    if (!labels || !out2 || n < 0) return vb200_set_error(VB200_EINVAL, "count_valid_labels: bad arguments");
    ce_valid_recip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(labels, n, ignore_index, out2);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
