// MoE routing: expert histogram, stable scatter index, row scatter / weighted gather for sm_100a.
//
// Reference (non-EP fused MoE, veomni/ops/kernels/moe/group_gemm.py:277-345):
//   splits        = expert_histogram(expert_index, E)                      (_kernels/kernel/moe.py:53-82)
//   scatter_index = expert_index.flatten().argsort(stable=True).argsort()  (group_gemm.py:44,287)
//   scatter_out   = moe_scatter(hidden, scatter_index)                     (moe.py:253-333)
//   out           = moe_gather(fc2_out, scatter_index)                     (moe.py:87-159, fp32 acc over top-k)
// The same indices drive the EP permutation (veomni/distributed/moe/moe_utils.py:19-41: expert-major,
// token order inside an expert == the stable sort, because a token's top-k experts are distinct).
//
// Integer results are bit-exact by construction: scatter_index[i] = (#slots with a smaller expert id)
// + (#earlier slots with the same expert id). Three tiny kernels: per-chunk histograms, a scan over
// (expert, chunk), and a per-chunk stable rank using warp match/ballot — no sort, no atomics to
// global memory, deterministic.
// Row scatter/gather are HBM streams: 2*T*K*H*2 B (scatter writes K copies, gather reads K rows).
#include "common.cuh"

namespace vb {

constexpr int kChunk = 1024;      // routing slots per CTA
constexpr int kMaxExperts = 1024;

template <typename IdxT>
__global__ void __launch_bounds__(kChunk)
route_hist_kernel(const IdxT* __restrict__ idx, int64_t n, int E, int* __restrict__ chunk_hist /*[chunks][E]*/) {
    extern __shared__ int hist[];
    for (int e = threadIdx.x; e < E; e += blockDim.x) hist[e] = 0;
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kChunk + threadIdx.x;
    if (i < n) {
        const int e = (int)idx[i];
        if (e >= 0 && e < E) atomicAdd(&hist[e], 1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < E; e += blockDim.x) chunk_hist[(int64_t)blockIdx.x * E + e] = hist[e];
}

// One CTA: for every expert, exclusive scan over chunks; then inclusive scan over experts.
__global__ void __launch_bounds__(1024)
route_scan_kernel(int* __restrict__ chunk_hist, int nchunks, int E, int* __restrict__ splits,
                  int* __restrict__ cumsum /* inclusive, [E] */) {
    __shared__ int tot[kMaxExperts];
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        int run = 0;
        for (int c = 0; c < nchunks; ++c) {
            const int v = chunk_hist[(int64_t)c * E + e];
            chunk_hist[(int64_t)c * E + e] = run;  // becomes the chunk's offset inside the expert
            run += v;
        }
        tot[e] = run;
        splits[e] = run;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int e = 0; e < E; ++e) {
            run += tot[e];
            cumsum[e] = run;
        }
    }
}

template <typename IdxT>
__global__ void __launch_bounds__(kChunk)
route_rank_kernel(const IdxT* __restrict__ idx, int64_t n, int E, const int* __restrict__ chunk_off,
                  const int* __restrict__ cumsum, int* __restrict__ scatter_index) {
    extern __shared__ int run[];  // next free row of every expert for this chunk
    for (int e = threadIdx.x; e < E; e += blockDim.x)
        run[e] = (e ? cumsum[e - 1] : 0) + chunk_off[(int64_t)blockIdx.x * E + e];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kChunk + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool ok = i < n;
    int e = ok ? (int)idx[i] : -1;
    if (e < 0 || e >= E) e = -1;
    const unsigned active = __ballot_sync(0xffffffffu, e >= 0);
    // warps claim their rows in order => stable inside the chunk
    for (int w = 0; w < kChunk / 32; ++w) {
        if (warp == w && e >= 0) {
            const unsigned peers = __match_any_sync(active, e);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (lane == leader) {
                base = run[e];
                run[e] = base + __popc(peers);
            }
            base = __shfl_sync(peers, base, leader);
            scatter_index[i] = base + __popc(peers & ((1u << lane) - 1));
        }
        __syncthreads();
    }
}

// out[index[t,k]] = x[t]; with w_in: either also scatter the per-slot weight (w_out != null) or, when
// w_out == null, scale the copied row by w_in[t,k] (rounded to bf16) — the backward of the weighted combine.
__global__ void __launch_bounds__(256)
moe_scatter_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ index, __nv_bfloat16* __restrict__ out,
                   const __nv_bfloat16* __restrict__ w_in, __nv_bfloat16* __restrict__ w_out, int64_t T, int K,
                   int vec_per_row) {
    const bool scale = (w_in != nullptr) && (w_out == nullptr);
    for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
        for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
            const uint4 val = ldg_stream(x + t * vec_per_row * 8 + v * 8);
            float f[8];
            if (scale) unpack8(val, f);
            for (int k = 0; k < K; ++k) {
                const int64_t r = index[t * K + k];
                if (scale) {
                    const float wk = __bfloat162float(w_in[t * K + k]);
                    float g[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) g[i] = f[i] * wk;
                    stg_stream(out + r * vec_per_row * 8 + v * 8, pack8(g));
                } else {
                    stg_stream(out + r * vec_per_row * 8 + v * 8, val);
                }
            }
        }
        if (w_out != nullptr && w_in != nullptr && threadIdx.x < K) w_out[index[t * K + threadIdx.x]] = w_in[t * K + threadIdx.x];
    }
}

// out[t] = sum_k (w ? bf16(x[index[t,k]] * w[t,k]) : x[index[t,k]])   fp32 accumulation in k order
template <int KMAX>
__global__ void __launch_bounds__(256)
moe_gather_kernel(const __nv_bfloat16* __restrict__ x, const int* __restrict__ index, const __nv_bfloat16* __restrict__ w,
                  __nv_bfloat16* __restrict__ out, int64_t T, int K, int vec_per_row) {
    for (int64_t t = blockIdx.x; t < T; t += gridDim.x) {
        for (int v = threadIdx.x; v < vec_per_row; v += blockDim.x) {
            float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            uint4 vals[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) vals[k] = ldg_stream(x + (int64_t)index[t * K + k] * vec_per_row * 8 + v * 8);
#pragma unroll
            for (int k = 0; k < KMAX; ++k)
                if (k < K) {
                    float f[8];
                    unpack8(vals[k], f);
                    if (w != nullptr) {
                        const float wk = __bfloat162float(w[t * K + k]);
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] += round_bf16(f[i] * wk);
                    } else {
#pragma unroll
                        for (int i = 0; i < 8; ++i) acc[i] += f[i];
                    }
                }
            stg_stream(out + t * vec_per_row * 8 + v * 8, pack8(acc));
        }
    }
}

// grad_w[t,k] = <g[t,:], x[index[t,k],:]> in fp32 (backward of the weighted combine with respect to the routing weights,
// veomni/distributed/moe/moe_utils.py:44-72). One warp per token: g[t] stays in registers for the K rows it is dotted with.
template <int VPL>  // 16-byte vectors per lane: hidden = VPL * 256
__global__ void __launch_bounds__(256)
moe_weight_grad_kernel(const __nv_bfloat16* __restrict__ g, const __nv_bfloat16* __restrict__ x, const int* __restrict__ index,
                       float* __restrict__ out, int64_t T, int K) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5), nwarps = (int64_t)gridDim.x * 8;
    constexpr int H = VPL * 256;
    for (int64_t t = warp; t < T; t += nwarps) {
        uint4 gv[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) gv[i] = ldg_stream(g + t * H + (i * 32 + lane) * 8);
        for (int k = 0; k < K; ++k) {
            const __nv_bfloat16* row = x + (int64_t)index[t * K + k] * H;
            uint4 xv[VPL];
#pragma unroll
            for (int i = 0; i < VPL; ++i) xv[i] = ldg_stream(row + (i * 32 + lane) * 8);
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < VPL; ++i) {
                acc = ffma2(bf2_to_f2(gv[i].x), bf2_to_f2(xv[i].x), acc);
                acc = ffma2(bf2_to_f2(gv[i].y), bf2_to_f2(xv[i].y), acc);
                acc = ffma2(bf2_to_f2(gv[i].z), bf2_to_f2(xv[i].z), acc);
                acc = ffma2(bf2_to_f2(gv[i].w), bf2_to_f2(xv[i].w), acc);
            }
            const float d = warp_sum(acc.x + acc.y);
            if (lane == 0) out[t * K + k] = d;
        }
    }
}

}  // namespace vb

using namespace vb;

extern "C" int64_t vb200_moe_route_workspace(int64_t num_slots, int32_t num_experts) {
    const int64_t chunks = (num_slots + kChunk - 1) / kChunk;
    return (chunks > 0 ? chunks : 1) * (int64_t)num_experts * 4;
}

extern "C" int vb200_moe_route(const void* expert_index, int32_t index_is_int64, int64_t num_slots, int32_t num_experts,
                               int32_t* splits, int32_t* cumsum, int32_t* scatter_index, void* workspace, void* stream) {
    if (num_experts <= 0 || num_experts > kMaxExperts)
        return vb200_set_error(VB200_EINVAL, "moe_route: num_experts must be in [1,1024]");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (int)((num_slots + kChunk - 1) / kChunk);
    int* ws = (int*)workspace;
    if (chunks == 0) {
        VB_CUDA_TRY(cudaMemsetAsync(splits, 0, 4 * num_experts, st));
        VB_CUDA_TRY(cudaMemsetAsync(cumsum, 0, 4 * num_experts, st));
        return VB200_OK;
    }
    const size_t sm = (size_t)num_experts * 4;
    if (index_is_int64) route_hist_kernel<int64_t><<<chunks, kChunk, sm, st>>>((const int64_t*)expert_index, num_slots, num_experts, ws);
    else route_hist_kernel<int32_t><<<chunks, kChunk, sm, st>>>((const int32_t*)expert_index, num_slots, num_experts, ws);
    VB_HOST_CHECK_LAUNCH();
    route_scan_kernel<<<1, 1024, 0, st>>>(ws, chunks, num_experts, splits, cumsum);
    VB_HOST_CHECK_LAUNCH();
    if (index_is_int64) route_rank_kernel<int64_t><<<chunks, kChunk, sm, st>>>((const int64_t*)expert_index, num_slots, num_experts, ws, cumsum, scatter_index);
    else route_rank_kernel<int32_t><<<chunks, kChunk, sm, st>>>((const int32_t*)expert_index, num_slots, num_experts, ws, cumsum, scatter_index);
    vb200_count_launch(3);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_moe_scatter(const void* x, const int32_t* scatter_index, void* out, const void* w_in, void* w_out,
                                 int64_t tokens, int32_t topk, int64_t hidden, void* stream) {
    if (hidden <= 0 || (hidden & 7)) return vb200_set_error(VB200_EINVAL, "moe_scatter: hidden must be a multiple of 8");
    if (tokens <= 0) return VB200_OK;
    const int grid = (int)(tokens < 16 * kNumSMs ? tokens : 16 * kNumSMs);
    moe_scatter_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, scatter_index, (__nv_bfloat16*)out,
                                                              (const __nv_bfloat16*)w_in, (__nv_bfloat16*)w_out, tokens,
                                                              topk, (int)(hidden >> 3));
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_moe_gather(const void* x, const int32_t* scatter_index, const void* weights, void* out,
                                int64_t tokens, int32_t topk, int64_t hidden, void* stream) {
    if (hidden <= 0 || (hidden & 7)) return vb200_set_error(VB200_EINVAL, "moe_gather: hidden must be a multiple of 8");
    if (topk < 1 || topk > 16) return vb200_set_error(VB200_EINVAL, "moe_gather: topk must be in [1,16]");
    if (tokens <= 0) return VB200_OK;
    const int grid = (int)(tokens < 16 * kNumSMs ? tokens : 16 * kNumSMs);
    cudaStream_t st = (cudaStream_t)stream;
#define GO(KM)                                                                                                     \
    moe_gather_kernel<KM><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, scatter_index, (const __nv_bfloat16*)weights, \
                                                (__nv_bfloat16*)out, tokens, topk, (int)(hidden >> 3))
    if (topk <= 2) GO(2);
    else if (topk <= 4) GO(4);
    else if (topk <= 8) GO(8);
    else GO(16);
#undef GO
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_moe_weight_grad(const void* g, const void* x, const int32_t* scatter_index, float* out, int64_t tokens,
                                     int32_t topk, int64_t hidden, void* stream) {
    if (hidden <= 0 || (hidden & 255) || hidden > 8192)
        return vb200_set_error(VB200_EINVAL, "moe_weight_grad: hidden must be a multiple of 256 up to 8192");
    if (topk < 1) return vb200_set_error(VB200_EINVAL, "moe_weight_grad: topk must be positive");
    if (tokens <= 0) return VB200_OK;
    const int64_t want = (tokens + 7) / 8;
    const int grid = (int)(want < 8 * kNumSMs ? want : 8 * kNumSMs);
    cudaStream_t st = (cudaStream_t)stream;
#define GO(V)                                                                                                            \
    moe_weight_grad_kernel<V><<<grid, 256, 0, st>>>((const __nv_bfloat16*)g, (const __nv_bfloat16*)x, scatter_index, out, \
                                                    tokens, topk)
    switch ((int)(hidden >> 8)) {
        case 1: GO(1); break;
        case 2: GO(2); break;
        case 3: GO(3); break;
        case 4: GO(4); break;
        case 5: GO(5); break;
        case 6: GO(6); break;
        case 7: GO(7); break;
        case 8: GO(8); break;
        case 16: GO(16); break;
        case 20: GO(20); break;
        case 32: GO(32); break;
        default: return vb200_set_error(VB200_EINVAL, "moe_weight_grad: unsupported hidden size");
    }
#undef GO
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
