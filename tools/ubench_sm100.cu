// Pipe-throughput microbenchmarks for the softmax stage of the attention kernels (B200, sm_100a).
//
// The ncu captures of the round-1 backward kernels (profiles/r01_attn_bwd_*_ts_ncu.txt) show ~2000 clk per 128x64 tile
// with no pipe saturated, and three restructurings (TMEM operands, ping-pong warpgroups, packed f32x2 math) left the
// time unchanged. These loops measure, per SM and for 1/2/4 warps per scheduler, what the building blocks of that stage
// cost in isolation: MUFU.EX2, F2FP (bf16x2 pack), FFMA2, tcgen05.ld / tcgen05.st, and the whole per-tile sequence
// (tcgen05.ld S,dP -> exp2 -> dS -> pack -> tcgen05.st) without any MMA or barrier — the floor the real kernel can reach.
//
// Build + run: python tools/ubench.py   (nvcc -gencode arch=compute_100a,code=sm_100a, one JSON line per measurement)
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../veomni_b200/csrc/common.cuh"
#include "../veomni_b200/csrc/umma.cuh"

using namespace vb;

constexpr int ITERS = 512;

// --- MUFU.EX2: 8 independent chains per thread ---------------------------------------------------------------
__global__ void k_ex2(float* out, long long* cyc) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = exp2f(a[i]) - 1.0f;  // 1 MUFU + 1 FADD
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// --- F2FP: pack two fp32 into bf16x2 -------------------------------------------------------------------------
__global__ void k_f2fp(uint32_t* out, long long* cyc) {
    float a[8];
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 1.0f + 0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const uint32_t p = f2_to_bf2(a[i], a[i + 1]);
            acc ^= p;
            a[i] = __uint_as_float((p << 16) | 0x3f800000u & 0x7fffffffu);  // keep a dependence on the packed value
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// --- FFMA2 (packed fp32 pairs) --------------------------------------------------------------------------------
__global__ void k_ffma2(float* out, long long* cyc) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(0.5f + threadIdx.x * 1e-4f, 0.25f + i);
    const float2 b = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ffma2(a[i], b, c);
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// --- tcgen05.ld / tcgen05.st: every warp streams its own lane quadrant ------------------------------------------
template <bool STORE>
__global__ void k_tmem(uint32_t* out, long long* cyc) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) & 3) * 128;  // own 128 columns
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = lane + i;
    tmem_st32(base, v);
    tmem_st32(base + 32, v);
    tmem_st32(base + 64, v);
    tmem_st32(base + 96, v);
    tmem_wait_st();
    __syncthreads();
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        if (STORE) {
            tmem_st32(base + (it & 3) * 32, v);
            tmem_wait_st();
        } else {
            tmem_ld32(base + (it & 3) * 32, v);  // includes tcgen05.wait::ld
            acc += v[it & 31];
        }
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + v[0];
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// --- the whole softmax stage of one 128x64 backward tile, no MMA, no barriers -------------------------------------
// 8 warps (two per lane quadrant, 32 columns each) as in attn_bwd_dq_tc_kernel: ld S, ld dP, 32x (FFMA, EX2, FADD, FMUL),
// pack, st dS to TMEM. Reports cycles per tile for the CTA.
__global__ void k_softmax_tile(uint32_t* out, long long* cyc, int packed) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc(&slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    const int q = warp & 3, cw = (warp >> 2) & 1;
    const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
    uint32_t z[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(0.01f * ((lane + i) & 15));
    tmem_st32(lane_base + cw * 32, z);
    tmem_st32(lane_base + 128 + cw * 32, z);
    tmem_wait_st();
    __syncthreads();
    const float sl2 = 0.1275f, lse2 = 0.5f, dl = 0.125f;
    uint32_t acc = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 256; ++it) {
        uint32_t sv[32], dv[32], pk[16];
        tmem_ld32_nowait(lane_base + cw * 32, sv);
        tmem_ld32_nowait(lane_base + 128 + cw * 32, dv);
        tmem_wait_ld();
        if (packed) {
            const float2 sl2v = make_float2(sl2, sl2), nlse = make_float2(-lse2, -lse2), ndl = make_float2(-dl, -dl);
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float2 t = ffma2(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), sl2v, nlse);
                const float2 u = fadd2(make_float2(__uint_as_float(dv[i]), __uint_as_float(dv[i + 1])), ndl);
                const float2 w = fmul2(make_float2(exp2f(t.x), exp2f(t.y)), u);
                pk[i >> 1] = f2_to_bf2(w.x, w.y);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float p0 = exp2f(__uint_as_float(sv[i]) * sl2 - lse2), p1 = exp2f(__uint_as_float(sv[i + 1]) * sl2 - lse2);
                pk[i >> 1] = f2_to_bf2(p0 * (__uint_as_float(dv[i]) - dl), p1 * (__uint_as_float(dv[i + 1]) - dl));
            }
        }
        tmem_st16(lane_base + 256 + cw * 16, pk);
        tmem_wait_st();
        acc ^= pk[it & 15];
    }
    const long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = (t1 - t0) / 256;
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

// --- the same softmax stage while one thread streams the tile's 20 MMAs (16x M128 N64 K16 + 4x M128 N128 K16) ------
// Columns 0..255 belong to the softmax warps (as in k_softmax_tile), the MMAs accumulate into 256..511: any slowdown of
// either side against its solo run is TMEM-port / issue contention, not a data dependence.
__global__ void __launch_bounds__(288, 1)
k_softmax_vs_mma(uint32_t* out, long long* cyc_soft, long long* cyc_mma, int run_mma, int run_soft) {
    extern __shared__ __align__(1024) uint8_t smem[];  // 64 KB of zeros: A [128][128] bf16 + B [128][128] bf16, SWIZZLE_128B boxes
    __shared__ uint32_t slot;
    __shared__ uint64_t bar[2];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 65536 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc(&slot, 512);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    constexpr int TILES = 128;
    if (warp == 0) {
        if (run_mma) {
            constexpr uint32_t idesc64 = umma_idesc(0, 0, 128, 64), idesc128 = umma_idesc(0, 1, 128, 128);
            const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 32768) >> 4;
            const long long t0 = clock64();
            for (int t = 0; t < TILES; ++t) {
                if (t >= 2) mbar_wait(&bar[t & 1], (uint32_t)((t >> 1) - 1) & 1u);  // at most two tiles of MMAs in flight
                if (elect_one_sync()) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t off = (uint32_t)(k >> 2) * (128 * 128) + (uint32_t)(k & 3) * 32;
                        umma_f16_bo(tmem + 256, a16, off, 16, 1024, b16, off, 16, 1024, idesc64, k ? 1u : 0u);
                        umma_f16_bo(tmem + 320, a16, off, 16, 1024, b16, off, 16, 1024, idesc64, k ? 1u : 0u);
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_bo(tmem + 384, a16, k * 32, 16, 1024, b16, k * 16 * 128, 64 * 128, 1024, idesc128, 1u);
                    umma_commit(&bar[t & 1]);
                }
                __syncwarp();
            }
            mbar_wait(&bar[(TILES - 1) & 1], (uint32_t)((TILES - 1) >> 1) & 1u);
            const long long t1 = clock64();
            if (lane == 0) cyc_mma[blockIdx.x] = (t1 - t0) / TILES;
        }
    } else if (run_soft) {
        const int sw = warp - 1, q = sw & 3, cw = (sw >> 2) & 1;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t z[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) z[i] = __float_as_uint(0.01f * ((lane + i) & 15));
        tmem_st32(lane_base + cw * 32, z);
        tmem_st32(lane_base + 128 + cw * 32, z);
        tmem_wait_st();
        const float sl2 = 0.1275f, lse2 = 0.5f, dl = 0.125f;
        uint32_t acc = 0;
        const long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < TILES; ++it) {
            uint32_t sv[32], dv[32], pk[16];
            tmem_ld32_nowait(lane_base + cw * 32, sv);
            tmem_ld32_nowait(lane_base + 128 + cw * 32, dv);
            tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
                const float p0 = exp2f(__uint_as_float(sv[i]) * sl2 - lse2), p1 = exp2f(__uint_as_float(sv[i + 1]) * sl2 - lse2);
                pk[i >> 1] = f2_to_bf2(p0 * (__uint_as_float(dv[i]) - dl), p1 * (__uint_as_float(dv[i + 1]) - dl));
            }
            tmem_st16(lane_base + 64 + cw * 16, pk);
            tmem_wait_st();
            acc ^= pk[it & 15];
        }
        const long long t1 = clock64();
        out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
        if (sw == 0 && lane == 0) cyc_soft[blockIdx.x] = (t1 - t0) / TILES;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc(tmem, 512);
    }
}

static double median_cycles(long long* d_cyc, int blocks) {
    long long* h = (long long*)malloc(sizeof(long long) * blocks);
    cudaMemcpy(h, d_cyc, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
    for (int i = 0; i < blocks; ++i)
        for (int j = i + 1; j < blocks; ++j)
            if (h[j] < h[i]) { long long t = h[i]; h[i] = h[j]; h[j] = t; }
    const double m = (double)h[blocks / 2];
    free(h);
    return m;
}

int main() {
    const int blocks = 148;
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * blocks * 1024);
    cudaMalloc(&cyc, sizeof(long long) * blocks);
    for (int warps = 4; warps <= 32; warps *= 2) {
        const int thr = warps * 32;
        for (int rep = 0; rep < 2; ++rep) k_ex2<<<blocks, thr>>>(out, cyc);
        cudaDeviceSynchronize();
        double c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"mufu_ex2\", \"warps_per_sm\": %d, \"lane_ops_per_clk_per_sm\": %.2f}\n", warps, (double)ITERS * 8 * thr / c);
        for (int rep = 0; rep < 2; ++rep) k_f2fp<<<blocks, thr>>>((uint32_t*)out, cyc);
        cudaDeviceSynchronize();
        c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"f2fp_bf16x2\", \"warps_per_sm\": %d, \"lane_ops_per_clk_per_sm\": %.2f}\n", warps, (double)ITERS * 4 * thr / c);
        for (int rep = 0; rep < 2; ++rep) k_ffma2<<<blocks, thr>>>(out, cyc);
        cudaDeviceSynchronize();
        c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"ffma2\", \"warps_per_sm\": %d, \"lane_pair_ops_per_clk_per_sm\": %.2f}\n", warps, (double)ITERS * 8 * thr / c);
    }
    for (int warps = 4; warps <= 16; warps *= 2) {
        const int thr = warps * 32;
        for (int rep = 0; rep < 2; ++rep) k_tmem<false><<<blocks, thr>>>((uint32_t*)out, cyc);
        cudaDeviceSynchronize();
        double c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"tcgen05_ld_32x32b_x32 (+wait)\", \"warps_per_sm\": %d, \"bytes_per_clk_per_sm\": %.1f, \"clk_per_ld\": %.1f}\n", warps,
               (double)ITERS * 4096 * warps / c, c / ITERS);
        for (int rep = 0; rep < 2; ++rep) k_tmem<true><<<blocks, thr>>>((uint32_t*)out, cyc);
        cudaDeviceSynchronize();
        c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"tcgen05_st_32x32b_x32 (+wait)\", \"warps_per_sm\": %d, \"bytes_per_clk_per_sm\": %.1f, \"clk_per_st\": %.1f}\n", warps,
               (double)ITERS * 4096 * warps / c, c / ITERS);
    }
    for (int packed = 0; packed < 2; ++packed) {
        for (int rep = 0; rep < 2; ++rep) k_softmax_tile<<<blocks, 256>>>((uint32_t*)out, cyc, packed);
        cudaDeviceSynchronize();
        const double c = median_cycles(cyc, blocks);
        printf("{\"bench\": \"bwd_softmax_stage_128x64_tile (8 warps, no MMA)\", \"packed_f32x2\": %d, \"clk_per_tile\": %.0f}\n", packed, c);
    }
    {
        long long* cyc2;
        cudaMalloc(&cyc2, sizeof(long long) * blocks);
        cudaFuncSetAttribute(k_softmax_vs_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
        const int modes[3][2] = {{1, 0}, {0, 1}, {1, 1}};
        for (int m = 0; m < 3; ++m) {
            cudaMemset(cyc, 0, sizeof(long long) * blocks);
            cudaMemset(cyc2, 0, sizeof(long long) * blocks);
            for (int rep = 0; rep < 2; ++rep) k_softmax_vs_mma<<<blocks, 288, 65536>>>((uint32_t*)out, cyc, cyc2, modes[m][0], modes[m][1]);
            cudaDeviceSynchronize();
            printf("{\"bench\": \"bwd tile: softmax stage vs 20-MMA stream\", \"mma\": %d, \"softmax\": %d, \"softmax_clk_per_tile\": %.0f, "
                   "\"mma_clk_per_tile\": %.0f}\n", modes[m][0], modes[m][1], median_cycles(cyc, blocks), median_cycles(cyc2, blocks));
        }
        cudaFree(cyc2);
    }
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("{\"error\": \"%s\"}\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
