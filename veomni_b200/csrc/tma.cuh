// TMA (cp.async.bulk.tensor) helpers: host-side tensor-map encoding through the driver entry
// point (no link-time dependency on libcuda) and device-side tile loads.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace vb {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        fn = (PFN_encodeTiled)p;
    }
    return fn;
}

// bf16 tensor addressed as [rows2][rows1][inner] (inner contiguous); strides in elements.
// Box = {64 (128 B, SWIZZLE_128B), 1, box_rows}: one call loads a [box_rows][64] swizzled sub-tile.
inline int make_tmap_3d(CUtensorMap* m, const void* base, uint64_t inner, uint64_t dim1, uint64_t dim2,
                        uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {inner, dim1, dim2};
    cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {64, 1, box_rows};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled failed");
    return 0;
}

// bf16 tensor [dim2][dim1][inner] with a 2-D box {box_inner (<= 64), box_rows} inside one dim2 slice
// (expert weight matrices [G][rows][inner]).
inline int make_tmap_3d_box(CUtensorMap* m, const void* base, uint64_t inner, uint64_t dim1, uint64_t dim2,
                            uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box_inner, uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[3] = {inner, dim1, dim2};
    cuuint64_t strides[2] = {stride1_elems * 2, stride2_elems * 2};
    cuuint32_t box[3] = {box_inner, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled failed");
    return 0;
}

// 2-D bf16 matrix [rows][cols] (cols contiguous, row stride in elements), box = {64, box_rows}.
inline int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t row_stride_elems,
                        uint32_t box_rows) {
    PFN_encodeTiled enc = get_encode_fn();
    if (!enc) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled entry point unavailable");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {row_stride_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return vb200_set_error(VB200_ECUDA, "cuTensorMapEncodeTiled failed");
    return 0;
}

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, int c2,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tm, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}

// Byte address of 16-byte chunk `chunk` (0..HD/8-1) of row `r` in a [rows][HD] bf16 tile stored as HD/64
// sub-tiles of [rows][64] with the TMA 128-byte swizzle (tile base 1024-byte aligned, rows % 8 == 0).
__device__ __forceinline__ uint32_t swz_addr(uint32_t base, int rows, int r, int chunk) {
    return base + (uint32_t)(chunk >> 3) * (uint32_t)rows * 128u + (uint32_t)r * 128u +
           (uint32_t)(((chunk & 7) ^ (r & 7)) << 4);
}

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// acc[NB][4] += A[16 x HD] * B[NB*8 x HD]^T, both operands row-major tiles in swizzled smem.
template <int HD, int NB>
__device__ __forceinline__ void gemm_nt(float (&acc)[NB][4], uint32_t a_base, int a_rows, int a_r0,
                                        uint32_t b_base, int b_rows, int b_r0, int lane) {
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t a[4];
        ldsm_x4(a, swz_addr(a_base, a_rows, a_r0 + (lane & 15), ks * 2 + (lane >> 4)));
#pragma unroll
        for (int nb2 = 0; nb2 < NB / 2; ++nb2) {
            uint32_t b[4];
            const int row = b_r0 + nb2 * 16 + (lane & 7) + ((lane >> 4) << 3);
            ldsm_x4(b, swz_addr(b_base, b_rows, row, ks * 2 + ((lane >> 3) & 1)));
            mma_bf16(acc[2 * nb2], a, b[0], b[1]);
            mma_bf16(acc[2 * nb2 + 1], a, b[2], b[3]);
        }
    }
}

// acc[HD/8][4] += P[16 x KN] * Z[KN x HD]; P as packed bf16 A fragments, Z a row-major swizzled tile.
template <int HD, int KN>
__device__ __forceinline__ void gemm_rt(float (&acc)[HD / 8][4], const uint32_t (&a)[KN / 16][4], uint32_t z_base,
                                        int z_rows, int z_r0, int lane) {
#pragma unroll
    for (int kk = 0; kk < KN / 16; ++kk) {
#pragma unroll
        for (int nb2 = 0; nb2 < HD / 16; ++nb2) {
            uint32_t b[4];
            const int row = z_r0 + kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
            ldsm_x4_t(b, swz_addr(z_base, z_rows, row, nb2 * 2 + (lane >> 4)));
            mma_bf16(acc[2 * nb2], a[kk], b[0], b[1]);
            mma_bf16(acc[2 * nb2 + 1], a[kk], b[2], b[3]);
        }
    }
}

}  // namespace vb
