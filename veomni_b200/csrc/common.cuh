// Shared device helpers for the veomni_b200 sm_100a kernels.
// Everything here is header-only; each .cu is compiled separately by veomni_b200/build.py
// with `-gencode arch=compute_100a,code=sm_100a -lineinfo`.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/veomni_b200.h"

#define VB_HOST_CHECK_LAUNCH()                                      \
    do {                                                            \
        cudaError_t e__ = cudaGetLastError();                       \
        if (e__ != cudaSuccess) return vb200_set_cuda_error(e__);   \
    } while (0)

#define VB_CUDA_TRY(expr)                                           \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return vb200_set_cuda_error(e__);   \
    } while (0)

// Defined in runtime.cu: records the last CUDA error string for vb200_last_error().
extern "C" int vb200_set_cuda_error(cudaError_t e);
extern "C" int vb200_set_error(int code, const char* msg);
extern "C" void vb200_count_launch(int n);

namespace vb {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v) {  // reduce over WIDTH consecutive lanes
#pragma unroll
    for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte streaming global load / store (no L1 allocation: every byte is touched once).
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}
// Plain (coherent) 16-byte load: used for peer-mapped addresses written by other GPUs.
__device__ __forceinline__ uint4 ldg_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p)
                 : "memory");
    return r;
}
__device__ __forceinline__ void stg_v4(void* p, const uint4& v) {
    asm volatile("st.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}

// Packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2): one issue slot for two lanes of the softmax arithmetic.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rc = *reinterpret_cast<uint64_t*>(&c), rd;
    asm("fma.rn.ftz.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rd;
    asm("add.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
    uint64_t ra = *reinterpret_cast<uint64_t*>(&a), rb = *reinterpret_cast<uint64_t*>(&b), rd;
    asm("mul.rn.ftz.f32x2 %0, %1, %2;" : "=l"(rd) : "l"(ra), "l"(rb));
    return *reinterpret_cast<float2*>(&rd);
}

__device__ __forceinline__ float2 bf2_to_f2(uint32_t u) {
    __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    float2 a = bf2_to_f2(u.x), b = bf2_to_f2(u.y), c = bf2_to_f2(u.z), d = bf2_to_f2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 u;
    u.x = f2_to_bf2(f[0], f[1]); u.y = f2_to_bf2(f[2], f[3]);
    u.z = f2_to_bf2(f[4], f[5]); u.w = f2_to_bf2(f[6], f[7]);
    return u;
}
// bf16x2 * bf16x2 -> bf16x2, one rounding (the product of two bf16 values is exact in fp32, so this equals the
// reference's fp32 multiply followed by `.to(bfloat16)`); SASS HMUL2.BF16.
__device__ __forceinline__ uint32_t bf2_mul(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// Sum of squares of the 8 bf16 values of one 16-byte vector into two packed accumulators (FFMA2). Every RMSNorm
// forward that must agree bit-for-bit (plain and fused-add) uses this one accumulation order.
__device__ __forceinline__ void sumsq8(const uint4& u, float2& acc0, float2& acc1) {
    const float2 a = bf2_to_f2(u.x), b = bf2_to_f2(u.y), c = bf2_to_f2(u.z), d = bf2_to_f2(u.w);
    acc0 = ffma2(a, a, acc0);
    acc1 = ffma2(b, b, acc1);
    acc0 = ffma2(c, c, acc0);
    acc1 = ffma2(d, d, acc1);
}
// y = w * bf16(x * rs) on one vector of 8 (two roundings, as Qwen3RMSNorm): FMUL2 + F2FP pack + HMUL2.BF16.
__device__ __forceinline__ uint4 norm_scale8(const uint4& u, const uint4& wv, float2 rs2) {
    uint4 o;
    float2 t;
    t = fmul2(bf2_to_f2(u.x), rs2); o.x = bf2_mul(wv.x, f2_to_bf2(t.x, t.y));
    t = fmul2(bf2_to_f2(u.y), rs2); o.y = bf2_mul(wv.y, f2_to_bf2(t.x, t.y));
    t = fmul2(bf2_to_f2(u.z), rs2); o.z = bf2_mul(wv.z, f2_to_bf2(t.x, t.y));
    t = fmul2(bf2_to_f2(u.w), rs2); o.w = bf2_mul(wv.w, f2_to_bf2(t.x, t.y));
    return o;
}
// Programmatic dependent launch: the dependent grid may start once every CTA of the primary has executed
// launch_dependents (or exited); its reads of the primary's results must come after griddep_wait().
__device__ __forceinline__ void griddep_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// Round an fp32 value to bf16 precision and back (mirrors `.to(bfloat16)` in the reference).
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ---- mbarrier / bulk-async (TMA 1-D) primitives --------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (the launch reports an error) instead of hanging the box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 24)) __trap();
    }
}
// global -> shared bulk copy, completion counted on `bar` (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
// shared -> global bulk copy (bulk async-group completion).
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
template <int N>
__device__ __forceinline__ void bulk_wait_all() {
    asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy smem writes -> visible to the async proxy (needed before bulk_s2g)
__device__ __forceinline__ void fence_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

}  // namespace vb
