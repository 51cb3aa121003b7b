// SwiGLU elementwise (silu(gate) * up) forward / backward for sm_100a.
//
// Reference semantics: Qwen3MLP.forward (patched_modeling_qwen3_gpu.py:121-127)
//   down_proj(act_fn(gate_proj(x)) * up_proj(x)); bound on GPU to LigerSiLUMulFunction
//   (veomni/ops/liger/__init__.py:119-142). silu is evaluated in fp32 and rounded to bf16 before
//   the product, as both eager bf16 and liger do. The MoE expert path uses the same op on the two
//   halves of a merged fc1 output (veomni/distributed/moe/moe_layer.py:339-346,
//   veomni/ops/kernels/moe/group_gemm.py:300-304), hence the row strides.
//
// Roofline: HBM stream; fwd 3 * rows*cols*2 B, bwd 5 * rows*cols*2 B.
#include "common.cuh"

namespace vb {

__device__ __forceinline__ float sigmoidf_fast(float x) { return 1.0f / (1.0f + __expf(-x)); }

__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up,
                  __nv_bfloat16* __restrict__ out, int64_t rows, int vec_per_row, int64_t in_stride,
                  int64_t out_stride) {
    const int64_t total = rows * vec_per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / vec_per_row;
        const int c = (int)(i % vec_per_row) * 8;
        float g[8], u[8], o[8];
        unpack8(ldg_stream(gate + r * in_stride + c), g);
        unpack8(ldg_stream(up + r * in_stride + c), u);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = round_bf16(g[k] * sigmoidf_fast(g[k])) * u[k];
        stg_stream(out + r * out_stride + c, pack8(o));
    }
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ gate,
                  const __nv_bfloat16* __restrict__ up, __nv_bfloat16* __restrict__ dgate,
                  __nv_bfloat16* __restrict__ dup, int64_t rows, int vec_per_row, int64_t in_stride,
                  int64_t dout_stride, int64_t dgrad_stride) {
    const int64_t total = rows * vec_per_row;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / vec_per_row;
        const int c = (int)(i % vec_per_row) * 8;
        float d[8], g[8], u[8], dg[8], du[8];
        unpack8(ldg_stream(dout + r * dout_stride + c), d);
        unpack8(ldg_stream(gate + r * in_stride + c), g);
        unpack8(ldg_stream(up + r * in_stride + c), u);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float s = sigmoidf_fast(g[k]);
            const float silu = g[k] * s;
            du[k] = d[k] * silu;
            dg[k] = d[k] * u[k] * (s + silu * (1.0f - s));
        }
        stg_stream(dgate + r * dgrad_stride + c, pack8(dg));
        stg_stream(dup + r * dgrad_stride + c, pack8(du));
    }
}

}  // namespace vb

using namespace vb;

static int ew_grid(int64_t total_vec) {
    int64_t g = (total_vec + 255) / 256;
    const int64_t cap = (int64_t)kNumSMs * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

extern "C" int vb200_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int64_t cols,
                                int64_t in_stride, int64_t out_stride, void* stream) {
    if ((cols & 7) || (in_stride & 7) || (out_stride & 7) || cols <= 0)
        return vb200_set_error(VB200_EINVAL, "swiglu_fwd: cols and strides must be multiples of 8");
    if (rows <= 0) return VB200_OK;
    swiglu_fwd_kernel<<<ew_grid(rows * (cols >> 3)), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)out, rows, (int)(cols >> 3),
        in_stride, out_stride);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup,
                                int64_t rows, int64_t cols, int64_t in_stride, int64_t dout_stride,
                                int64_t dgrad_stride, void* stream) {
    if ((cols & 7) || (in_stride & 7) || (dout_stride & 7) || (dgrad_stride & 7) || cols <= 0)
        return vb200_set_error(VB200_EINVAL, "swiglu_bwd: cols and strides must be multiples of 8");
    if (rows <= 0) return VB200_OK;
    swiglu_bwd_kernel<<<ew_grid(rows * (cols >> 3)), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)dgate,
        (__nv_bfloat16*)dup, rows, (int)(cols >> 3), in_stride, dout_stride, dgrad_stride);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
