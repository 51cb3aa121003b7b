// Multi-tensor L2 norm and in-place scaling for gradient clipping.
//
// Replaces the two passes of torch.nn.utils.clip_grad_norm_ as called by veomni_clip_grad_norm
// (veomni/distributed/clip_grad_norm.py:7-20 -> fsdp2/clip_grad_norm.py:21-51,73-153): `_foreach_norm` over every
// local gradient shard (measured 1 TB/s on the ~400 Qwen3-8B shards) and `_foreach_mul_` by the clip coefficient.
// One launch each over a device table of (pointer, numel) entries — the host splits every tensor into pieces of at most
// 2^20 elements so work per entry is bounded: grid.x = entry, grid.y = kMtBlocks blocks striding over it with 16-byte
// loads, so a 600 M-element lm_head shard and a 4096-element norm weight stream side by side without a tail. The sum of squares
// is reduced in a fixed order (thread -> warp -> block partial -> one finishing block), hence bit-reproducible.
// The coefficient is read from device memory (max_norm / (total + 1e-6), clamped to 1, computed by the caller after
// the cross-rank all-reduce of the scalar), so the pass needs no host synchronisation; a coefficient of exactly 1
// returns without touching memory (x * 1.0f == x).
#include "common.cuh"

namespace vb {

constexpr int kMtBlocks = 4;     // blocks per table entry (the host splits tensors into entries of <= 2^20 elements)
constexpr int kMtThreads = 512;

__device__ __forceinline__ float block_sum_512(float v, float* red) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        v = threadIdx.x < (kMtThreads >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_sum(v);
    }
    return v;  // valid in thread 0
}

template <typename T>
__global__ void __launch_bounds__(kMtThreads)
multi_sumsq_kernel(const void* const* __restrict__ ptrs, const int64_t* __restrict__ numels, float* __restrict__ partials) {
    __shared__ float red[32];
    const T* x = reinterpret_cast<const T*>(ptrs[blockIdx.x]);
    const int64_t n = numels[blockIdx.x];
    constexpr int V = 16 / sizeof(T);
    const int64_t tid = (int64_t)blockIdx.y * kMtThreads + threadIdx.x;
    const int64_t nthr = (int64_t)kMtBlocks * kMtThreads;
    // scalar head until 16-byte alignment, vector body, scalar tail
    int64_t head = ((16 - ((uintptr_t)x & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    const int64_t nvec = (n - head) / V;
    const int64_t tail0 = head + nvec * V;
    float acc = 0.f;
    for (int64_t j = tid; j < head + (n - tail0); j += nthr) {
        const int64_t i = j < head ? j : tail0 + (j - head);
        float v;
        if constexpr (sizeof(T) == 4) v = reinterpret_cast<const float*>(x)[i];
        else v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(x)[i]);
        acc = fmaf(v, v, acc);
    }
    const uint4* xv = reinterpret_cast<const uint4*>(x + head);
    int64_t v0 = tid;
    for (; v0 + 3 * nthr < nvec; v0 += 4 * nthr) {  // four 16-byte loads in flight per thread
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream(xv + v0 + u * nthr);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if constexpr (sizeof(T) == 4) {
                const float a = __uint_as_float(r[u].x), b = __uint_as_float(r[u].y), c = __uint_as_float(r[u].z), d = __uint_as_float(r[u].w);
                acc = fmaf(a, a, acc); acc = fmaf(b, b, acc); acc = fmaf(c, c, acc); acc = fmaf(d, d, acc);
            } else {
                const float2 a = bf2_to_f2(r[u].x), b = bf2_to_f2(r[u].y), c = bf2_to_f2(r[u].z), d = bf2_to_f2(r[u].w);
                acc = fmaf(a.x, a.x, acc); acc = fmaf(a.y, a.y, acc); acc = fmaf(b.x, b.x, acc); acc = fmaf(b.y, b.y, acc);
                acc = fmaf(c.x, c.x, acc); acc = fmaf(c.y, c.y, acc); acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
            }
        }
    }
    for (; v0 < nvec; v0 += nthr) {
        const uint4 r = ldg_stream(xv + v0);
        if constexpr (sizeof(T) == 4) {
            const float a = __uint_as_float(r.x), b = __uint_as_float(r.y), c = __uint_as_float(r.z), d = __uint_as_float(r.w);
            acc = fmaf(a, a, acc); acc = fmaf(b, b, acc); acc = fmaf(c, c, acc); acc = fmaf(d, d, acc);
        } else {
            const float2 a = bf2_to_f2(r.x), b = bf2_to_f2(r.y), c = bf2_to_f2(r.z), d = bf2_to_f2(r.w);
            acc = fmaf(a.x, a.x, acc); acc = fmaf(a.y, a.y, acc); acc = fmaf(b.x, b.x, acc); acc = fmaf(b.y, b.y, acc);
            acc = fmaf(c.x, c.x, acc); acc = fmaf(c.y, c.y, acc); acc = fmaf(d.x, d.x, acc); acc = fmaf(d.y, d.y, acc);
        }
    }
    const float s = block_sum_512(acc, red);
    if (threadIdx.x == 0) partials[(int64_t)blockIdx.x * kMtBlocks + blockIdx.y] = s;
}

// out[0] = sum of all partials, out[1 + t] = sum of tensor t's partials (optional), fixed order.
__global__ void __launch_bounds__(kMtThreads)
multi_sumsq_finish_kernel(const float* __restrict__ partials, int n_tensors, float* __restrict__ total, float* __restrict__ per_tensor) {
    __shared__ float red[32];
    float acc = 0.f;
    for (int t = threadIdx.x; t < n_tensors; t += kMtThreads) {
        float s = 0.f;
#pragma unroll
        for (int b = 0; b < kMtBlocks; ++b) s += partials[(int64_t)t * kMtBlocks + b];
        if (per_tensor) per_tensor[t] = s;
        acc += s;
    }
    const float s = block_sum_512(acc, red);
    if (threadIdx.x == 0) total[0] = s;
}

template <typename T>
__global__ void __launch_bounds__(kMtThreads)
multi_scale_kernel(void* const* __restrict__ ptrs, const int64_t* __restrict__ numels, const float* __restrict__ coef_dev) {
    const float coef = *coef_dev;
    if (coef == 1.0f) return;
    T* x = reinterpret_cast<T*>(ptrs[blockIdx.x]);
    const int64_t n = numels[blockIdx.x];
    constexpr int V = 16 / sizeof(T);
    const int64_t tid = (int64_t)blockIdx.y * kMtThreads + threadIdx.x;
    const int64_t nthr = (int64_t)kMtBlocks * kMtThreads;
    int64_t head = ((16 - ((uintptr_t)x & 15)) & 15) / sizeof(T);
    if (head > n) head = n;
    const int64_t nvec = (n - head) / V;
    const int64_t tail0 = head + nvec * V;
    for (int64_t j = tid; j < head + (n - tail0); j += nthr) {
        const int64_t i = j < head ? j : tail0 + (j - head);
        if constexpr (sizeof(T) == 4) reinterpret_cast<float*>(x)[i] *= coef;
        else reinterpret_cast<__nv_bfloat16*>(x)[i] = __float2bfloat16_rn(__bfloat162float(reinterpret_cast<__nv_bfloat16*>(x)[i]) * coef);
    }
    uint4* xv = reinterpret_cast<uint4*>(x + head);
    auto scale = [&](uint4 r) {
        if constexpr (sizeof(T) == 4) {
            r.x = __float_as_uint(__uint_as_float(r.x) * coef); r.y = __float_as_uint(__uint_as_float(r.y) * coef);
            r.z = __float_as_uint(__uint_as_float(r.z) * coef); r.w = __float_as_uint(__uint_as_float(r.w) * coef);
        } else {
            float2 a = bf2_to_f2(r.x), b = bf2_to_f2(r.y), c = bf2_to_f2(r.z), d = bf2_to_f2(r.w);
            r.x = f2_to_bf2(a.x * coef, a.y * coef); r.y = f2_to_bf2(b.x * coef, b.y * coef);
            r.z = f2_to_bf2(c.x * coef, c.y * coef); r.w = f2_to_bf2(d.x * coef, d.y * coef);
        }
        return r;
    };
    int64_t v0 = tid;
    for (; v0 + 3 * nthr < nvec; v0 += 4 * nthr) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) r[u] = ldg_stream(xv + v0 + u * nthr);
#pragma unroll
        for (int u = 0; u < 4; ++u) stg_stream(xv + v0 + u * nthr, scale(r[u]));
    }
    for (; v0 < nvec; v0 += nthr) stg_stream(xv + v0, scale(ldg_stream(xv + v0)));
}


// AdamW step over a device table of (param, grad, exp_avg, exp_avg_sq, numel) entries, fp32 state and parameters.
// Arithmetic of torch.optim.AdamW(fused=True) (aten/src/ATen/native/cuda/fused_adam_utils.cuh, ADAMW mode, no
// amsgrad / maximize):  p -= lr*wd*p;  m = m + (1-b1)*(g - m);  v = b2*v + (1-b2)*g*g;
//                       p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps).
// One pass: 16 B read of each of p, g, m, v and 16 B write of p, m, v per 4 elements (28 B/element), which PyTorch's
// multi_tensor_apply version moves at ~4.5 TB/s on the ~400 Qwen3-8B shards. grad_scale (optional device scalar)
// multiplies the gradient first, so a clip coefficient can ride along instead of a separate scaling pass.
struct AdamWArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2_sqrt;
};

// G = gradient type (float or bf16). LP: also write the updated parameter, rounded to bf16, to a second table
// ("master weights": the fp32 parameter lives in the optimizer, the model computes on the bf16 copy).
template <typename G, bool LP>
__global__ void __launch_bounds__(kMtThreads)
multi_adamw_kernel(void* const* __restrict__ params, const void* const* __restrict__ grads, void* const* __restrict__ exp_avgs,
                   void* const* __restrict__ exp_avg_sqs, void* const* __restrict__ lp_params,
                   const int64_t* __restrict__ numels, AdamWArgs a, const float* __restrict__ grad_scale) {
    float* p = reinterpret_cast<float*>(params[blockIdx.x]);
    const G* g = reinterpret_cast<const G*>(grads[blockIdx.x]);
    float* m = reinterpret_cast<float*>(exp_avgs[blockIdx.x]);
    float* v = reinterpret_cast<float*>(exp_avg_sqs[blockIdx.x]);
    __nv_bfloat16* lp = LP ? reinterpret_cast<__nv_bfloat16*>(lp_params[blockIdx.x]) : nullptr;
    const int64_t n = numels[blockIdx.x];
    const float gs = grad_scale ? *grad_scale : 1.0f;
    const float step_size = a.lr / a.bias_correction1, decay = 1.0f - a.lr * a.weight_decay;
    const float omb1 = 1.0f - a.beta1, omb2 = 1.0f - a.beta2, inv_bc2s = 1.0f / a.bias_correction2_sqrt;
    auto upd = [&](float& pp, float gg, float& mm, float& vv) {
        gg *= gs;
        pp *= decay;
        mm = fmaf(omb1, gg - mm, mm);
        vv = fmaf(a.beta2, vv, omb2 * gg * gg);
        pp -= step_size * mm / (sqrtf(vv) * inv_bc2s + a.eps);
    };
    auto gload = [&](int64_t i) -> float {
        if constexpr (sizeof(G) == 4) return reinterpret_cast<const float*>(g)[i];
        else return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(g)[i]);
    };
    const int64_t tid = (int64_t)blockIdx.y * kMtThreads + threadIdx.x;
    const int64_t nthr = (int64_t)kMtBlocks * kMtThreads;
    // all arrays of an entry share their 16-byte phase only if the caller guarantees it (entries are cut at
    // multiples of 2^20 elements from 256-byte aligned allocations); otherwise the scalar path runs
    bool aligned = ((((uintptr_t)p | (uintptr_t)m | (uintptr_t)v) & 15) == 0) && (((uintptr_t)g & (sizeof(G) == 4 ? 15 : 7)) == 0);
    if (LP) aligned = aligned && (((uintptr_t)lp & 7) == 0);
    if (aligned) {
        const int64_t nvec = n >> 2;
        for (int64_t i = tid; i < nvec; i += nthr) {
            float4 pv = reinterpret_cast<float4*>(p)[i], mv = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
            float g0, g1, g2, g3;
            if constexpr (sizeof(G) == 4) {
                const uint4 gr = ldg_stream(reinterpret_cast<const uint4*>(g) + i);
                g0 = __uint_as_float(gr.x); g1 = __uint_as_float(gr.y); g2 = __uint_as_float(gr.z); g3 = __uint_as_float(gr.w);
            } else {
                uint2 gr;
                asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(gr.x), "=r"(gr.y) : "l"(reinterpret_cast<const uint2*>(g) + i));
                const float2 a0 = bf2_to_f2(gr.x), a1 = bf2_to_f2(gr.y);
                g0 = a0.x; g1 = a0.y; g2 = a1.x; g3 = a1.y;
            }
            upd(pv.x, g0, mv.x, vv.x);
            upd(pv.y, g1, mv.y, vv.y);
            upd(pv.z, g2, mv.z, vv.z);
            upd(pv.w, g3, mv.w, vv.w);
            reinterpret_cast<float4*>(p)[i] = pv;
            reinterpret_cast<float4*>(m)[i] = mv;
            reinterpret_cast<float4*>(v)[i] = vv;
            if (LP) {
                uint2 o;
                o.x = f2_to_bf2(pv.x, pv.y);
                o.y = f2_to_bf2(pv.z, pv.w);
                reinterpret_cast<uint2*>(lp)[i] = o;
            }
        }
        for (int64_t i = (nvec << 2) + tid; i < n; i += nthr) {
            upd(p[i], gload(i), m[i], v[i]);
            if (LP) lp[i] = __float2bfloat16_rn(p[i]);
        }
    } else {
        for (int64_t i = tid; i < n; i += nthr) {
            upd(p[i], gload(i), m[i], v[i]);
            if (LP) lp[i] = __float2bfloat16_rn(p[i]);
        }
    }
}

}  // namespace vb

using namespace vb;

extern "C" int64_t vb200_multi_sumsq_partials(int32_t n_tensors) { return (int64_t)n_tensors * kMtBlocks; }

extern "C" int vb200_multi_sumsq(const void* const* ptrs_dev, const int64_t* numels_dev, int32_t n_tensors, int32_t dtype,
                                 float* partials, float* total, float* per_tensor, void* stream) {
    if (n_tensors < 0 || (n_tensors > 0 && (!ptrs_dev || !numels_dev || !partials)) || !total)
        return vb200_set_error(VB200_EINVAL, "multi_sumsq: bad arguments");
    if (dtype != 0 && dtype != 1) return vb200_set_error(VB200_EINVAL, "multi_sumsq: dtype must be 0 (bf16) or 1 (f32)");
        cudaStream_t s = (cudaStream_t)stream;
    if (n_tensors > 0) {
        dim3 grid(n_tensors, kMtBlocks);
        if (dtype == 1) multi_sumsq_kernel<float><<<grid, kMtThreads, 0, s>>>(ptrs_dev, numels_dev, partials);
        else multi_sumsq_kernel<__nv_bfloat16><<<grid, kMtThreads, 0, s>>>(ptrs_dev, numels_dev, partials);
        VB_HOST_CHECK_LAUNCH();
    }
    multi_sumsq_finish_kernel<<<1, kMtThreads, 0, s>>>(partials, n_tensors, total, per_tensor);
    vb200_count_launch(n_tensors > 0 ? 2 : 1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_multi_scale(void* const* ptrs_dev, const int64_t* numels_dev, int32_t n_tensors, int32_t dtype,
                                 const float* coef_dev, void* stream) {
    if (n_tensors == 0) return VB200_OK;
    if (n_tensors < 0 || !ptrs_dev || !numels_dev || !coef_dev) return vb200_set_error(VB200_EINVAL, "multi_scale: bad arguments");
    if (dtype != 0 && dtype != 1) return vb200_set_error(VB200_EINVAL, "multi_scale: dtype must be 0 (bf16) or 1 (f32)");
    dim3 grid(n_tensors, kMtBlocks);
    cudaStream_t s = (cudaStream_t)stream;
    if (dtype == 1) multi_scale_kernel<float><<<grid, kMtThreads, 0, s>>>(ptrs_dev, numels_dev, coef_dev);
    else multi_scale_kernel<__nv_bfloat16><<<grid, kMtThreads, 0, s>>>(ptrs_dev, numels_dev, coef_dev);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_multi_adamw(void* const* params_dev, const void* const* grads_dev, void* const* exp_avgs_dev,
                                 void* const* exp_avg_sqs_dev, void* const* lp_params_dev, const int64_t* numels_dev,
                                 int32_t n_entries, int32_t grad_dtype, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, float bias_correction1, float bias_correction2_sqrt,
                                 const float* grad_scale_dev, void* stream) {
    if (n_entries == 0) return VB200_OK;
    if (n_entries < 0 || !params_dev || !grads_dev || !exp_avgs_dev || !exp_avg_sqs_dev || !numels_dev)
        return vb200_set_error(VB200_EINVAL, "multi_adamw: bad arguments");
    if (grad_dtype != 0 && grad_dtype != 1) return vb200_set_error(VB200_EINVAL, "multi_adamw: grad_dtype must be 0 (bf16) or 1 (f32)");
    if (!(bias_correction1 > 0.f) || !(bias_correction2_sqrt > 0.f))
        return vb200_set_error(VB200_EINVAL, "multi_adamw: bias corrections must be positive (step >= 1)");
    AdamWArgs a{lr, beta1, beta2, eps, weight_decay, bias_correction1, bias_correction2_sqrt};
    dim3 grid(n_entries, kMtBlocks);
    cudaStream_t st = (cudaStream_t)stream;
#define VB_ADAMW(G, LP) multi_adamw_kernel<G, LP><<<grid, kMtThreads, 0, st>>>(params_dev, grads_dev, exp_avgs_dev, exp_avg_sqs_dev, lp_params_dev, numels_dev, a, grad_scale_dev)
    if (grad_dtype == 1) {
        if (lp_params_dev) VB_ADAMW(float, true); else VB_ADAMW(float, false);
    } else {
        if (lp_params_dev) VB_ADAMW(__nv_bfloat16, true); else VB_ADAMW(__nv_bfloat16, false);
    }
#undef VB_ADAMW
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
