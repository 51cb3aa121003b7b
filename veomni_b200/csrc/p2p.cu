// NVLink / NVSwitch peer-memory runtime and collectives for sm_100a.
//
// One process per GPU. Every rank owns a *symmetric* data region and a small signal pad, both
// cudaMalloc'd here and exported with CUDA IPC so each peer maps them into its own address space
// (vb200_symm_* / vb200_ipc_*). A `vb200_comm` handle then holds the N mapped base pointers.
// Collectives are single kernels that read peers' memory directly with coalesced 16-byte loads
// over NVLink (pull model) and synchronise with flags in the peers' signal pads
// (st.release.sys / ld.acquire.sys) instead of NCCL's channel/proxy machinery:
//
//   vb200_allgather      FSDP2 unit all-gather   (torch/_fsdp_collectives.py:81-95, DefaultAllGather)
//   vb200_reduce_scatter FSDP2 unit reduce-scatter (:116-131, DefaultReduceScatter) incl. the divide
//   vb200_all_to_all     Ulysses seq<->head exchange (veomni/distributed/sequence_parallel/ulysses.py:86-122)
//                        and any chunked pull (EP dispatch/combine row blocks, moe/comm.py:36-42)
//   vb200_comm_barrier   flag barrier (tests, allocation-lifetime fences)
//
// Protocol of every collective kernel on channel c with epoch e (= previous epoch + 1, kept in
// local device memory, so no host state is needed):
//   1. ready:  thread p of CTA 0 stores e into peer p's pad[c][READY][my rank]   (release.sys)
//   2. every CTA waits until its own pad[c][READY][p] >= e for all p             (acquire.sys)
//   3. pull:   coalesced 16 B loads from peers' regions, local stores
//   4. done:   the last CTA to finish stores e into every peer's pad[c][DONE][my rank], waits for
//              all peers' DONE (so when the kernel retires, peers no longer read our region and
//              it may be overwritten) and publishes epoch[c] = e.
// Waits are bounded (~30 s of clock64) and report VB200_ETIMEOUT through the comm's error word
// instead of hanging the GPU.
#include <cstring>

#include <cstdlib>

#include "common.cuh"

namespace vb {

constexpr int kMaxWorld = 8;
constexpr int kMaxChannels = 32;
constexpr int kPadWords = kMaxChannels * 2 * kMaxWorld;  // uint64 flag words per rank (4 KB)
constexpr long long kSpinLimit = 60000000000LL;          // clock64 ticks (~30 s)

struct CommDev {
    int rank, world;
    uint8_t* data[kMaxWorld];   // peers' symmetric data regions, mapped locally
    uint64_t* sig[kMaxWorld];   // peers' signal pads
    uint32_t* state;            // local: epoch[kMaxChannels], counter[kMaxChannels], error (+3 pad), counter2[kMaxChannels]
};
constexpr int kStateError = 2 * kMaxChannels;
constexpr int kStateCounter2 = 2 * kMaxChannels + 4;
constexpr int kStateWords = 3 * kMaxChannels + 4;

struct CommHost {
    CommDev dev;
    int64_t data_bytes;
};

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ int pad_idx(int ch, int kind, int src) { return (ch * 2 + kind) * kMaxWorld + src; }

// A flag word is {payload:32 (high), epoch:32 (low)}. READY flags carry, as payload, the byte offset
// >> 8 of the buffer this rank exposes for the collective inside its own symmetric region, so buffer
// offsets need not be equal across ranks (each rank runs its own allocator).
// thread t < world: tell peer t that `kind` of epoch e has happened on this rank
__device__ __forceinline__ void signal_peers(const CommDev& c, int ch, int kind, uint32_t e, int64_t my_off = 0) {
    if ((int)threadIdx.x < c.world)
        st_release_sys(c.sig[threadIdx.x] + pad_idx(ch, kind, c.rank), ((uint64_t)(my_off >> 8) << 32) | e);
}
// thread t < world: wait until peer t has signalled `kind` of epoch >= e; publishes the peers' payload
// offsets in shared memory (`peer_off`, may be null) and ends with a block-wide barrier
__device__ __forceinline__ void wait_peers(const CommDev& c, int ch, int kind, uint32_t e, int64_t* peer_off = nullptr) {
    if ((int)threadIdx.x < c.world) {
        const uint64_t* f = c.sig[c.rank] + pad_idx(ch, kind, threadIdx.x);
        const long long t0 = clock64();
        uint64_t v;
        while ((int32_t)((uint32_t)(v = ld_acquire_sys(f)) - e) < 0) {
            __nanosleep(40);
            if (clock64() - t0 > kSpinLimit) {
                atomicExch(c.state + kStateError, 1u);
                break;
            }
        }
        if (peer_off) peer_off[threadIdx.x] = (int64_t)(v >> 32) << 8;
    }
    __syncthreads();
}
// Called by every CTA after its part of the work. Returns true in the last CTA to arrive.
__device__ __forceinline__ bool grid_arrive_last(const CommDev& c, int ch) {
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(c.state + kMaxChannels + ch, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    return is_last != 0;
}
__device__ __forceinline__ void finish_epoch(const CommDev& c, int ch, uint32_t e) {
    // last CTA only
    signal_peers(c, ch, 1, e);
    wait_peers(c, ch, 1, e);
    if (threadIdx.x == 0) {
        c.state[kMaxChannels + ch] = 0;  // counter
        __threadfence();
        c.state[ch] = e;  // epoch
    }
}

// One warp that performs the flag exchange a collective starts with (optionally sending this rank's flag first) and exits.
// Launched in front of a many-CTA collective kernel on the same stream: the kernel then finds every peer ready and holds
// its SMs only for the transfer, instead of parking 16-32 SMs' worth of CTAs in a spin loop until the slowest peer has
// reached the same point of its step (1.2 ms of the 1.85 ms an in-step all-gather launch took at N=8 was that wait, with
// the GEMMs next to it running on 116 SMs). The kernel repeats the exchange — same epoch, same values: idempotent.
__global__ void __launch_bounds__(32)
comm_gate_kernel(CommDev c, int ch, int kind, int64_t off, int do_signal) {
    const uint32_t e = c.state[ch] + 1;
    if (do_signal) signal_peers(c, ch, kind, e, off);
    wait_peers(c, ch, kind, e);
}

// Copy `bytes` from src to dst (same alignment modulo 16) with the whole grid; 2-byte granularity.
__device__ __forceinline__ void grid_copy(uint8_t* dst, const uint8_t* src, int64_t bytes) {
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    int64_t head = (16 - ((uintptr_t)src & 15)) & 15;
    if (head > bytes) head = bytes;
    if (((uintptr_t)src & 15) != ((uintptr_t)dst & 15)) head = bytes;  // fall back to 2-byte copies
    for (int64_t i = tid * 2; i < head; i += nthr * 2)
        *reinterpret_cast<uint16_t*>(dst + i) = *reinterpret_cast<const volatile uint16_t*>(src + i);
    const int64_t nvec = (bytes - head) >> 4;
    const uint4* s4 = reinterpret_cast<const uint4*>(src + head);
    uint4* d4 = reinterpret_cast<uint4*>(dst + head);
    int64_t i = tid;
    for (; i + 7 * nthr < nvec; i += 8 * nthr) {  // 8 independent 16 B peer loads in flight per thread
        uint4 r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) r[u] = ldg_v4(s4 + i + u * nthr);
#pragma unroll
        for (int u = 0; u < 8; ++u) d4[i + u * nthr] = r[u];
    }
    for (; i < nvec; i += nthr) d4[i] = ldg_v4(s4 + i);
    const int64_t done = head + (nvec << 4);
    for (int64_t j = done + tid * 2; j < bytes; j += nthr * 2)
        *reinterpret_cast<uint16_t*>(dst + j) = *reinterpret_cast<const volatile uint16_t*>(src + j);
}

__global__ void __launch_bounds__(512)
barrier_kernel(CommDev c, int ch) {
    const uint32_t e = c.state[ch] + 1;
    signal_peers(c, ch, 0, e);
    wait_peers(c, ch, 0, e);
    finish_epoch(c, ch, e);
}

// All-gather in place inside the symmetric region: rank p's shard already sits at
// region_p[off_p + p*shard_bytes] of rank p; every rank fills the other N-1 slots from the owners.
__global__ void __launch_bounds__(512)
allgather_kernel(CommDev c, int ch, int64_t off, int64_t shard_bytes) {
    __shared__ int64_t peer_off[kMaxWorld];
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    wait_peers(c, ch, 0, e, peer_off);
    for (int q = 1; q < c.world; ++q) {
        const int p = (c.rank + q) % c.world;  // staggered so sources are hit evenly
        grid_copy(c.data[c.rank] + off + (int64_t)p * shard_bytes, c.data[p] + peer_off[p] + (int64_t)p * shard_bytes,
                  shard_bytes);
    }
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// Reduce-scatter: every rank holds N chunks of `chunk` fp32 values at its buffer; rank r sums chunk r of
// all ranks in rank order 0..N-1 (deterministic), scales, writes `out` (any memory).
// W = compile-time world size (0 = generic up to 8), UN = vectors per thread per iteration: UN*W = 16
// independent 16-byte loads are in flight per thread (one of the W sources is local; a loaded NVSwitch round trip
// measured ~7 us, so 8 in flight left 32 CTAs at half the link rate).
template <int W, int UN>
__global__ void __launch_bounds__(512)
reduce_scatter_f32_kernel(CommDev c, int ch, int64_t off, int64_t chunk, float scale, float* __restrict__ out) {
    __shared__ int64_t peer_off[kMaxWorld];
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    wait_peers(c, ch, 0, e, peer_off);
    constexpr int NP = W ? W : kMaxWorld;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t rank_base = (int64_t)c.rank * chunk * 4;
    // All loops over peers are fully unrolled with a predicate so the peer pointers stay in registers.
#define VB_SRC(p) (c.data[p] + peer_off[p] + rank_base)
    int64_t head = ((16 - ((uintptr_t)VB_SRC(0) & 15)) & 15) >> 2;  // floats until 16-byte alignment
    if (head > chunk) head = chunk;
    const int64_t nvec = (chunk - head) >> 2;
    const int64_t tail0 = head + (nvec << 2);
    // scalar head [0, head) and tail [tail0, chunk)
    for (int64_t j = tid; j < head + (chunk - tail0); j += nthr) {
        const int64_t i = j < head ? j : tail0 + (j - head);
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (p < c.world) {
                float v;
                asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(VB_SRC(p) + i * 4) : "memory");
                a += v;
            }
        out[i] = a * scale;
    }
    const bool out_aligned = (((uintptr_t)(out + head)) & 15) == 0;
    for (int64_t v0 = tid; v0 < nvec; v0 += nthr * UN) {
        uint4 r[UN][NP];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t v = v0 + (int64_t)u * nthr;
#pragma unroll
            for (int p = 0; p < NP; ++p)
                if (p < c.world && v < nvec) r[u][p] = ldg_v4(VB_SRC(p) + (head << 2) + (v << 4));
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t v = v0 + (int64_t)u * nthr;
            if (v < nvec) {
                float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    if (p < c.world) {
                        a.x += __uint_as_float(r[u][p].x); a.y += __uint_as_float(r[u][p].y);
                        a.z += __uint_as_float(r[u][p].z); a.w += __uint_as_float(r[u][p].w);
                    }
                a.x *= scale; a.y *= scale; a.z *= scale; a.w *= scale;
                float* o = out + head + v * 4;
                if (out_aligned) {
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w) : "memory");
                } else {
                    asm volatile("st.global.f32 [%0], %1;" ::"l"(o), "f"(a.x) : "memory");
                    asm volatile("st.global.f32 [%0], %1;" ::"l"(o + 1), "f"(a.y) : "memory");
                    asm volatile("st.global.f32 [%0], %1;" ::"l"(o + 2), "f"(a.z) : "memory");
                    asm volatile("st.global.f32 [%0], %1;" ::"l"(o + 3), "f"(a.w) : "memory");
                }
            }
        }
    }
#undef VB_SRC
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// bf16-input variant: the gradients of a bf16 FSDP unit are exact bf16 values, so pulling them as bf16 and
// accumulating in fp32 in rank order is bit-identical to the fp32 reduce-scatter of their fp32 copies — at half
// the NVLink bytes and without the fp32 staging pass. Input layout: N chunks of `chunk` bf16 at `off`.
template <int W, int UN>
__global__ void __launch_bounds__(512)
reduce_scatter_bf16_kernel(CommDev c, int ch, int64_t off, int64_t chunk, float scale, float* __restrict__ out) {
    __shared__ int64_t peer_off[kMaxWorld];
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    wait_peers(c, ch, 0, e, peer_off);
    constexpr int NP = W ? W : kMaxWorld;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t rank_base = (int64_t)c.rank * chunk * 2;
#define VB_SRC(p) (c.data[p] + peer_off[p] + rank_base)
    int64_t head = ((16 - ((uintptr_t)VB_SRC(0) & 15)) & 15) >> 1;  // bf16 elements until 16-byte alignment
    if (head > chunk) head = chunk;
    const int64_t nvec = (chunk - head) >> 3;
    const int64_t tail0 = head + (nvec << 3);
    for (int64_t j = tid; j < head + (chunk - tail0); j += nthr) {
        const int64_t i = j < head ? j : tail0 + (j - head);
        float a = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (p < c.world) {
                uint16_t v;
                asm volatile("ld.global.u16 %0, [%1];" : "=h"(v) : "l"(VB_SRC(p) + i * 2) : "memory");
                a += __uint_as_float((uint32_t)v << 16);
            }
        out[i] = a * scale;
    }
    const bool out_aligned = (((uintptr_t)(out + head)) & 15) == 0;
    for (int64_t v0 = tid; v0 < nvec; v0 += nthr * UN) {
        uint4 r[UN][NP];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t v = v0 + (int64_t)u * nthr;
#pragma unroll
            for (int p = 0; p < NP; ++p)
                if (p < c.world && v < nvec) r[u][p] = ldg_v4(VB_SRC(p) + (head << 1) + (v << 4));
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
            const int64_t v = v0 + (int64_t)u * nthr;
            if (v < nvec) {
                float a[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) a[i] = 0.f;
#pragma unroll
                for (int p = 0; p < NP; ++p)
                    if (p < c.world) {
                        const uint32_t w[4] = {r[u][p].x, r[u][p].y, r[u][p].z, r[u][p].w};
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            a[2 * i] += __uint_as_float(w[i] << 16);
                            a[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
                        }
                    }
                float* o = out + head + v * 8;
                if (out_aligned) {
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(a[0] * scale), "f"(a[1] * scale), "f"(a[2] * scale), "f"(a[3] * scale) : "memory");
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4), "f"(a[4] * scale), "f"(a[5] * scale), "f"(a[6] * scale), "f"(a[7] * scale) : "memory");
                } else {
#pragma unroll
                    for (int i = 0; i < 8; ++i) asm volatile("st.global.f32 [%0], %1;" ::"l"(o + i), "f"(a[i] * scale) : "memory");
                }
            }
        }
    }
#undef VB_SRC
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// FSDP2 reduce-scatter copy-in without the dtype conversion: packs the unsharded bf16 gradients of one unit into
// the [world, S] chunk-major layout torch._chunk_cat produces (chunk r of every parameter, dim-0 zero-padded to a
// multiple of world, concatenated), keeping bf16. One launch: grid.y = parameter, grid.x blocks stride over it.
struct PackEntry {
    const void* src;
    int64_t numel;   // real elements of the parameter
    int64_t chunk;   // elements per rank chunk = ceil(dim0 / world) * inner
    int64_t off;     // element offset of this parameter inside a rank's row of the output
};
constexpr int kPackMax = 24;
struct PackArgs {
    int n;
    int world;
    int64_t row;  // S: elements per rank row
    PackEntry e[kPackMax];
};

__device__ __forceinline__ __nv_bfloat16 pack_elem(const __nv_bfloat16* p, int64_t i) { return p[i]; }
__device__ __forceinline__ __nv_bfloat16 pack_elem(const float* p, int64_t i) { return __float2bfloat16_rn(p[i]); }
__device__ __forceinline__ uint4 pack_vec8(const __nv_bfloat16* p, int64_t v) { return ldg_stream(p + (v << 3)); }
__device__ __forceinline__ uint4 pack_vec8(const float* p, int64_t v) {  // 8 fp32 -> 8 bf16 (round to nearest even)
    const uint4 a = ldg_stream(p + (v << 3)), b = ldg_stream(p + (v << 3) + 4);
    uint4 r;
    r.x = f2_to_bf2(__uint_as_float(a.x), __uint_as_float(a.y));
    r.y = f2_to_bf2(__uint_as_float(a.z), __uint_as_float(a.w));
    r.z = f2_to_bf2(__uint_as_float(b.x), __uint_as_float(b.y));
    r.w = f2_to_bf2(__uint_as_float(b.z), __uint_as_float(b.w));
    return r;
}

// SrcT = bf16: the reduce-scatter copy-in; SrcT = float: also the all-gather copy-in (world = 1, the fp32 master
// shards cast to the bf16 all-gather input in one pass).
template <typename SrcT, bool VEC>
__global__ void __launch_bounds__(512)
fsdp_pack_bf16_kernel(PackArgs a, __nv_bfloat16* __restrict__ out) {
    const PackEntry& e = a.e[blockIdx.y];
    const SrcT* src = reinterpret_cast<const SrcT*>(e.src);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t padded = e.chunk * a.world;
    if (VEC) {  // chunk, off, row multiples of 8 and 16-byte aligned pointers: 8 elements per thread step
        const int64_t nvec = padded >> 3;
        const int64_t full = e.numel >> 3;  // vectors entirely inside the real data
        for (int64_t v0 = tid; v0 < nvec; v0 += 4 * nthr) {
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t v = v0 + u * nthr;
                if (v < full) r[u] = pack_vec8(src, v);
                else if (v < nvec) {
                    __align__(16) __nv_bfloat16 t[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) t[i] = ((v << 3) + i < e.numel) ? pack_elem(src, (v << 3) + i) : __float2bfloat16_rn(0.f);
                    r[u] = *reinterpret_cast<uint4*>(t);
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int64_t v = v0 + u * nthr;
                if (v < nvec) {
                    const int64_t el = v << 3;
                    const int64_t rk = el / e.chunk;
                    stg_v4(out + rk * a.row + e.off + (el - rk * e.chunk), r[u]);
                }
            }
        }
    } else {
        for (int64_t el = tid; el < padded; el += nthr) {
            const int64_t rk = el / e.chunk;
            out[rk * a.row + e.off + (el - rk * e.chunk)] = el < e.numel ? pack_elem(src, el) : __float2bfloat16_rn(0.f);
        }
    }
}

// Chunked pull ("all-to-all"): for each descriptor d and each peer p, copy `rows` segments of
// seg_bytes from peer p's exposed buffer at  src_off + rank*src_rank_stride + row*src_row_stride
// to local memory at               dst + p*dst_peer_stride + row*dst_row_stride.
struct A2ADesc {
    int64_t src_off, src_rank_stride, src_row_stride;
    uint8_t* dst;
    int64_t dst_peer_stride, dst_row_stride;
    int64_t rows, seg_bytes;
};
struct A2AArgs {
    int n;
    A2ADesc d[4];
};

__global__ void __launch_bounds__(512)
all_to_all_kernel(CommDev c, int ch, int64_t off, A2AArgs args) {
    __shared__ int64_t peer_off[kMaxWorld];
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    wait_peers(c, ch, 0, e, peer_off);
    const int lane = threadIdx.x & 31;
    const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t GW = ((int64_t)gridDim.x * blockDim.x) >> 5;
    constexpr int R = 8;  // row segments in flight per warp (one 16 B load per lane each)
    // Loop order: descriptor, peer, rows. Peer and descriptor are uniform per sweep, so a row costs two adds — the first
    // version decoded (peer, row) from a flat item index with a 64-bit division per item and was instruction-bound on the
    // 512-byte k/v segments of the Ulysses exchange (228-304 GB/s).
    for (int di = 0; di < args.n; ++di) {
        const A2ADesc& d = args.d[di];
        const int vec_per_seg = (int)(d.seg_bytes >> 4);
        for (int q = 0; q < c.world; ++q) {
            const int p = (c.rank + q) % c.world;  // staggered over sources
            const uint8_t* sbase = c.data[p] + peer_off[p] + d.src_off + (int64_t)c.rank * d.src_rank_stride;
            uint8_t* dbase = d.dst + (int64_t)p * d.dst_peer_stride;
            for (int64_t row0 = gw * R; row0 < d.rows; row0 += GW * R) {
                const int nr = (int)((d.rows - row0) < R ? (d.rows - row0) : R);
                for (int v = lane; v < vec_per_seg; v += 32) {
                    uint4 val[R];
#pragma unroll
                    for (int u = 0; u < R; ++u)
                        if (u < nr) val[u] = ldg_v4(sbase + (row0 + u) * d.src_row_stride + ((int64_t)v << 4));
#pragma unroll
                    for (int u = 0; u < R; ++u)
                        if (u < nr) stg_v4(dbase + (row0 + u) * d.dst_row_stride + ((int64_t)v << 4), val[u]);
                }
            }
        }
    }
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// largest i in [0, n) with prefix[i] <= v (prefix[0] = 0, prefix[n] > v; zero-sized entries are skipped)
__device__ __forceinline__ int find_segment(const int64_t* prefix, int n, int64_t v) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (prefix[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// Variable-size chunk pull (EP dispatch / combine): a device-side list of up to `nchunks` copies
// {peer, src_off (bytes in the peer's region), dst_off (bytes from dst), bytes}; chunk sizes are
// multiples of 16 bytes. One CTA-group sweeps each chunk; the list itself lives in device memory so
// it can be produced by the routing kernel without a host round trip.
struct ChunkDesc {
    int64_t src_off, dst_off, bytes;
    int32_t peer, pad;
};

constexpr int kChunkSmem = 1024;  // chunk lists up to this length are swept as ONE flattened index space

__global__ void __launch_bounds__(512)
chunk_pull_kernel(CommDev c, int ch, int64_t off, const ChunkDesc* __restrict__ chunks, int nchunks,
                  uint8_t* __restrict__ dst) {
    __shared__ int64_t peer_off[kMaxWorld];
    __shared__ int64_t s_prefix[kChunkSmem + 1];  // 16-byte vectors before chunk k
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    const bool flat = nchunks <= kChunkSmem;
    if (flat) {
        // exclusive prefix of the chunk sizes (block-wide: 2 chunks per thread, then a serial pass over 32 warp totals)
        __shared__ int64_t warp_tot[16];
        const int t = threadIdx.x;
        const int64_t a0 = 2 * t < nchunks ? (chunks[2 * t].bytes >> 4) : 0;
        const int64_t a1 = 2 * t + 1 < nchunks ? (chunks[2 * t + 1].bytes >> 4) : 0;
        int64_t v = a0 + a1;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t u = __shfl_up_sync(0xffffffffu, v, o);
            if ((t & 31) >= o) v += u;
        }
        if ((t & 31) == 31) warp_tot[t >> 5] = v;
        __syncthreads();
        int64_t base = 0;
        for (int w = 0; w < (t >> 5); ++w) base += warp_tot[w];
        const int64_t incl = base + v;  // inclusive over pairs
        if (2 * t < kChunkSmem) {
            s_prefix[2 * t] = incl - a0 - a1;
            s_prefix[2 * t + 1] = incl - a1;
        }
        if (t == 511) s_prefix[kChunkSmem] = incl;
    }
    wait_peers(c, ch, 0, e, peer_off);
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    if (flat) {
        // every thread keeps 8 independent 16-byte peer loads in flight regardless of how small the chunks are (an EP
        // exchange is E/EP x EP blocks of a few hundred KB: swept one by one they never fill the unrolled loop)
        const int64_t vtot = s_prefix[nchunks < kChunkSmem ? nchunks : kChunkSmem];
        constexpr int UN = 8;
        for (int64_t v0 = tid; v0 < vtot; v0 += nthr * UN) {
            uint4 r[UN];
            uint4* d[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int64_t v = v0 + (int64_t)u * nthr;
                d[u] = nullptr;
                if (v < vtot) {
                    const int k = find_segment(s_prefix, nchunks, v);
                    const ChunkDesc cd = chunks[k];
                    const int64_t o = (v - s_prefix[k]) << 4;
                    r[u] = ldg_v4(c.data[cd.peer] + peer_off[cd.peer] + cd.src_off + o);
                    d[u] = reinterpret_cast<uint4*>(dst + cd.dst_off + o);
                }
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (d[u]) stg_v4(d[u], r[u]);
        }
    } else {
        for (int k = 0; k < nchunks; ++k) {
            const ChunkDesc d = chunks[k];
            const int64_t nvec = d.bytes >> 4;
            const uint4* s = reinterpret_cast<const uint4*>(c.data[d.peer] + peer_off[d.peer] + d.src_off);
            uint4* o = reinterpret_cast<uint4*>(dst + d.dst_off);
            int64_t i = tid;
            for (; i + 7 * nthr < nvec; i += 8 * nthr) {
                uint4 r[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) r[u] = ldg_v4(s + i + u * nthr);
#pragma unroll
                for (int u = 0; u < 8; ++u) o[i + u * nthr] = r[u];
            }
            for (; i < nvec; i += nthr) o[i] = ldg_v4(s + i);
        }
    }
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// ---- FSDP2 unit all-gather with the copy-out fused in ----------------------------------------------------------
// Replaces DefaultAllGather (torch/_fsdp_collectives.py:81-95) *and* foreach_all_gather_copy_out's
// fsdp::split_with_sizes_copy (:196-212, :346-412): every rank's shard row (the concatenation of its shards of the
// unit's parameters, written by the copy-in) is pulled from its owner and stored straight into the per-parameter
// unsharded tensors, rank p's piece of parameter i at dst_i + p * bytes_i. The [world, row] all-gather output buffer
// is never filled and the extra read + write pass over the whole unit disappears.
constexpr int kScatterMax = 64;
struct ScatterArgs {
    int n;
    int vec_ok;                          // every offset / size / destination is a multiple of 16 bytes
    int64_t off[kScatterMax];            // byte offset of parameter i inside a shard row
    int64_t bytes[kScatterMax];          // bytes of one rank's shard of parameter i
    uint8_t* dst[kScatterMax];           // the parameter's unsharded tensor (local memory, world * bytes)
};

constexpr int kTileVecs = 8;  // 16-byte vectors per thread per tile: a tile is blockDim * 8 * 16 B = 64 KB of one parameter

__global__ void __launch_bounds__(1024)
allgather_scatter_kernel(CommDev c, int ch, int64_t off, int64_t shard_bytes, ScatterArgs a) {
    __shared__ int64_t peer_off[kMaxWorld];
    __shared__ int s_tprefix[kScatterMax + 1];  // 64 KB tiles before parameter i inside one rank's row
    __shared__ int64_t s_off[kScatterMax], s_bytes[kScatterMax];
    __shared__ uint8_t* s_dst[kScatterMax];
    __shared__ const uint8_t* s_base[kMaxWorld];
    const uint32_t e = c.state[ch] + 1;
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);
    const int64_t tile_bytes = (int64_t)blockDim.x * kTileVecs * 16;
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int i = 0; i < a.n; ++i) {
            s_tprefix[i] = acc;
            acc += (int)((a.bytes[i] + tile_bytes - 1) / tile_bytes);
        }
        s_tprefix[a.n] = acc;
    }
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        s_off[i] = a.off[i];
        s_bytes[i] = a.bytes[i];
        s_dst[i] = a.dst[i];
    }
    wait_peers(c, ch, 0, e, peer_off);  // ends with a block-wide barrier
    if ((int)threadIdx.x < c.world) s_base[threadIdx.x] = c.data[threadIdx.x] + peer_off[threadIdx.x] + (int64_t)threadIdx.x * shard_bytes;
    __syncthreads();
    if (a.vec_ok) {
        // Work = (peer q, 64 KB tile of one parameter). A CTA walks the tiles of a peer with stride gridDim, so the
        // parameter index only ever moves forward (no per-vector search: the copy itself is ~3 instructions per
        // 16 bytes and must stay that cheap for 32 CTAs to keep the links busy); every thread has 8 independent
        // 16-byte peer loads in flight, then stores them.
        const int tiles_per_peer = s_tprefix[a.n];
        for (int q = 0; q < c.world; ++q) {
            const int p = (c.rank + q) % c.world;  // q = 0: this rank's own shard; sources staggered over the peers
            const uint8_t* src_row = s_base[p];
            int i = 0;
            for (int t = blockIdx.x; t < tiles_per_peer; t += gridDim.x) {
                while (t >= s_tprefix[i + 1]) ++i;
                const int64_t base = (int64_t)(t - s_tprefix[i]) * tile_bytes;
                const int64_t left = s_bytes[i] - base;  // > 0
                const uint8_t* src = src_row + s_off[i] + base;
                uint8_t* dst = s_dst[i] + (int64_t)p * s_bytes[i] + base;
                uint4 r[kTileVecs];
#pragma unroll
                for (int u = 0; u < kTileVecs; ++u) {
                    const int64_t o = ((int64_t)u * blockDim.x + threadIdx.x) << 4;
                    if (o < left) r[u] = ldg_v4(src + o);
                }
#pragma unroll
                for (int u = 0; u < kTileVecs; ++u) {
                    const int64_t o = ((int64_t)u * blockDim.x + threadIdx.x) << 4;
                    if (o < left) stg_v4(dst + o, r[u]);
                }
            }
        }
    } else {
        for (int q = 0; q < c.world; ++q) {
            const int p = (c.rank + q) % c.world;
            for (int i = 0; i < a.n; ++i) grid_copy(s_dst[i] + (int64_t)p * s_bytes[i], s_base[p] + s_off[i], s_bytes[i]);
        }
    }
    if (grid_arrive_last(c, ch)) finish_epoch(c, ch, e);
}

// ---- FSDP2 unit reduce-scatter with the copy-in fused in -----------------------------------------------------------
// Replaces foreach_reduce_scatter_copy_in (torch._chunk_cat, _fsdp_collectives.py:667-675) + DefaultReduceScatter
// (:116-131) + the divide (:701-759). Phase A reads the unit's bf16 gradients *in place* through a pointer table and
// pushes chunk p of every parameter (dim-0 zero-padded to a multiple of world, as chunk_cat lays it out) to peer p's
// staging buffer, slot = this rank, with 16-byte NVLink stores (posted writes: no round-trip latency to cover, so few
// CTAs keep the links busy). Phase B, after every peer's "pushed" flag has arrived, sums the world slots of the local
// staging buffer in rank order 0..N-1 in fp32 (deterministic, bit-identical to the fp32 reduce-scatter of the same
// gradients because bf16 -> fp32 is exact), scales, and writes the fp32 shard.
constexpr int kPushMax = 64;
struct PushArgs {
    int n;
    int world;
    int vec_ok;    // chunks are multiples of 8 elements and gradient pointers are 16-byte aligned
    int64_t row;   // elements per rank row = sum of chunks
    const __nv_bfloat16* src[kPushMax];
    int64_t numel[kPushMax];
    int64_t chunk[kPushMax];  // off[i] = sum of chunk[0..i)
};

__device__ __forceinline__ uint4 ldg_cg_v4(const void* p) {  // L2-coherent load (data written by peers over NVLink)
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

template <int W>
__global__ void __launch_bounds__(512)
reduce_scatter_push_bf16_kernel(CommDev c, int ch, int64_t off, PushArgs a, float scale, float* __restrict__ out, int phase) {
    // phase 3: the whole collective in one launch. phase 1 (push) + phase 2 (reduce) as two launches with a gate kernel
    // between them: the wait for the slowest peer's pushes then costs one warp instead of the whole grid.
    __shared__ int64_t peer_off[kMaxWorld];
    __shared__ int64_t s_prefix[kPushMax + 1];  // element offset of parameter i inside a row
    __shared__ const __nv_bfloat16* s_src[kPushMax];
    __shared__ int64_t s_numel[kPushMax], s_chunk[kPushMax];
    __shared__ int s_tprefix[kPushMax + 1];  // 64 KB tiles before parameter i inside one row
    __shared__ int s_last;
    const uint32_t e = c.state[ch] + 1;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthr = (int64_t)gridDim.x * blockDim.x;
    const int64_t row = a.row;
    if (phase & 1) {
    if (blockIdx.x == 0) signal_peers(c, ch, 0, e, off);  // "my staging buffer (at off) is free for epoch e"
    if (threadIdx.x == 0) {
        int64_t acc = 0;
        int tacc = 0;
        const int64_t tile_elems = (int64_t)blockDim.x * kTileVecs * 8;
        for (int i = 0; i < a.n; ++i) {
            s_prefix[i] = acc;
            s_tprefix[i] = tacc;
            acc += a.chunk[i];
            tacc += (int)((a.chunk[i] + tile_elems - 1) / tile_elems);
        }
        s_prefix[a.n] = acc;
        s_tprefix[a.n] = tacc;
    }
    for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
        s_src[i] = a.src[i];
        s_numel[i] = a.numel[i];
        s_chunk[i] = a.chunk[i];
    }
    wait_peers(c, ch, 0, e, peer_off);
    // ---- phase A: push -------------------------------------------------------------------------------------------
    // Work = (peer q, 64 KB tile of one parameter's chunk), walked like the all-gather's tiles: no per-vector search.
    if (a.vec_ok) {
        const int64_t tile_elems = (int64_t)blockDim.x * kTileVecs * 8;
        const int tiles_per_peer = s_tprefix[a.n];
        for (int q = 0; q < c.world; ++q) {
            const int p = (c.rank + q) % c.world;  // q = 0: own chunk into the local staging slot
            __nv_bfloat16* dst_row = reinterpret_cast<__nv_bfloat16*>(c.data[p] + peer_off[p]) + (int64_t)c.rank * row;
            int i = 0;
            for (int t = blockIdx.x; t < tiles_per_peer; t += gridDim.x) {
                while (t >= s_tprefix[i + 1]) ++i;
                const int64_t base = (int64_t)(t - s_tprefix[i]) * tile_elems;  // element inside the chunk
                const int64_t left = s_chunk[i] - base;                          // elements of the chunk from here on
                const int64_t g0 = (int64_t)p * s_chunk[i] + base;               // element of gradient i
                const __nv_bfloat16* src = s_src[i];
                const int64_t numel = s_numel[i];
                __nv_bfloat16* dst = dst_row + s_prefix[i] + base;
                uint4 r[kTileVecs];
#pragma unroll
                for (int u = 0; u < kTileVecs; ++u) {
                    const int64_t x = ((int64_t)u * blockDim.x + threadIdx.x) << 3;
                    if (x < left) {
                        const int64_t g = g0 + x;
                        if (g + 8 <= numel) {
                            r[u] = ldg_stream(src + g);
                        } else {  // dim-0 padding: zeros
                            __align__(16) __nv_bfloat16 tv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) tv[j] = g + j < numel ? src[g + j] : __float2bfloat16_rn(0.f);
                            r[u] = *reinterpret_cast<uint4*>(tv);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < kTileVecs; ++u) {
                    const int64_t x = ((int64_t)u * blockDim.x + threadIdx.x) << 3;
                    if (x < left) stg_v4(dst + x, r[u]);
                }
            }
        }
    } else {
        for (int q = 0; q < c.world; ++q) {
            const int p = (c.rank + q) % c.world;
            __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(c.data[p] + peer_off[p]) + (int64_t)c.rank * row;
            for (int64_t x = tid; x < row; x += nthr) {
                const int i = find_segment(s_prefix, a.n, x);
                const int64_t g = (int64_t)p * s_chunk[i] + (x - s_prefix[i]);
                dst[x] = g < s_numel[i] ? s_src[i][g] : __float2bfloat16_rn(0.f);
            }
        }
    }
    // every store of this CTA is performed system-wide before the CTA counts itself as done
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(c.state + kMaxChannels + ch, 1u);
        s_last = (prev == gridDim.x - 1);
        __threadfence();
    }
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        signal_peers(c, ch, 1, e);  // "everything I owe you for epoch e has been written"
        if (phase == 1 && threadIdx.x == 0) c.state[kMaxChannels + ch] = 0;  // every CTA has counted itself already
    }
    }  // phase & 1
    if (!(phase & 2)) return;
    wait_peers(c, ch, 1, e);
    // ---- phase B: reduce the world slots of the local staging buffer in rank order -------------------------------
    const __nv_bfloat16* stage = reinterpret_cast<const __nv_bfloat16*>(c.data[c.rank] + off);
    constexpr int NP = W ? W : kMaxWorld;
    constexpr int UN2 = W == 0 ? 2 : (W <= 2 ? 8 : 4);
    if (a.vec_ok && ((((uintptr_t)out) & 15) == 0)) {
        const int64_t nvec = row >> 3;
        for (int64_t v0 = tid; v0 < nvec; v0 += nthr * UN2) {
            uint4 r[UN2][NP];
#pragma unroll
            for (int u = 0; u < UN2; ++u) {
                const int64_t v = v0 + (int64_t)u * nthr;
#pragma unroll
                for (int s = 0; s < NP; ++s)
                    if (s < c.world && v < nvec) r[u][s] = ldg_cg_v4(stage + (int64_t)s * row + (v << 3));
            }
#pragma unroll
            for (int u = 0; u < UN2; ++u) {
                const int64_t v = v0 + (int64_t)u * nthr;
                if (v < nvec) {
                    float acc[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
                    for (int s = 0; s < NP; ++s)
                        if (s < c.world) {
                            const uint32_t w[4] = {r[u][s].x, r[u][s].y, r[u][s].z, r[u][s].w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                acc[2 * j] += __uint_as_float(w[j] << 16);
                                acc[2 * j + 1] += __uint_as_float(w[j] & 0xffff0000u);
                            }
                        }
                    float* o = out + (v << 3);
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o), "f"(acc[0] * scale), "f"(acc[1] * scale), "f"(acc[2] * scale), "f"(acc[3] * scale) : "memory");
                    asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(o + 4), "f"(acc[4] * scale), "f"(acc[5] * scale), "f"(acc[6] * scale), "f"(acc[7] * scale) : "memory");
                }
            }
        }
    } else {
        for (int64_t x = tid; x < row; x += nthr) {
            float acc = 0.f;
            for (int s = 0; s < c.world; ++s) {
                uint16_t v;
                asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(v) : "l"(stage + (int64_t)s * row + x) : "memory");
                acc += __uint_as_float((uint32_t)v << 16);
            }
            out[x] = acc * scale;
        }
    }
    // completion: no peer handshake is needed (peers write this staging buffer again only after the READY flag of this
    // rank's *next* reduce-scatter kernel, which is stream-ordered after this one); the last CTA resets the counters
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned prev = atomicAdd(c.state + kStateCounter2 + ch, 1u);
        if (prev == gridDim.x - 1) {
            c.state[kMaxChannels + ch] = 0;
            c.state[kStateCounter2 + ch] = 0;
            __threadfence();
            c.state[ch] = e;
        }
    }
}

}  // namespace vb

using namespace vb;

// ---- allocation / IPC -------------------------------------------------------------------------
extern "C" int vb200_symm_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return vb200_set_error(VB200_EINVAL, "symm_alloc: bad arguments");
    VB_CUDA_TRY(cudaMalloc(ptr, (size_t)bytes));
    VB_CUDA_TRY(cudaMemset(*ptr, 0, (size_t)bytes));
    VB_CUDA_TRY(cudaDeviceSynchronize());
    return VB200_OK;
}
extern "C" int vb200_symm_free(void* ptr) {
    VB_CUDA_TRY(cudaFree(ptr));
    return VB200_OK;
}
extern "C" int vb200_ipc_get_handle(const void* ptr, void* handle64) {
    cudaIpcMemHandle_t h;
    VB_CUDA_TRY(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return VB200_OK;
}
extern "C" int vb200_ipc_open_handle(const void* handle64, void** ptr) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    VB_CUDA_TRY(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return VB200_OK;
}
extern "C" int vb200_ipc_close_handle(void* ptr) {
    VB_CUDA_TRY(cudaIpcCloseMemHandle(ptr));
    return VB200_OK;
}

// ---- comm handle ------------------------------------------------------------------------------
extern "C" int64_t vb200_comm_signal_bytes(void) { return (int64_t)kPadWords * 8; }

extern "C" int vb200_comm_create(void** comm, int32_t rank, int32_t world, void* const* peer_data,
                                 void* const* peer_signal, int64_t data_bytes) {
    if (!comm || world < 1 || world > kMaxWorld || rank < 0 || rank >= world)
        return vb200_set_error(VB200_EINVAL, "comm_create: world must be in [1,8] and 0 <= rank < world");
    CommHost* h = new CommHost();
    h->dev.rank = rank;
    h->dev.world = world;
    for (int p = 0; p < kMaxWorld; ++p) {
        h->dev.data[p] = (uint8_t*)(p < world ? peer_data[p] : peer_data[rank]);
        h->dev.sig[p] = (uint64_t*)(p < world ? peer_signal[p] : peer_signal[rank]);
    }
    h->data_bytes = data_bytes;
    const size_t state_bytes = sizeof(uint32_t) * kStateWords;
    cudaError_t e = cudaMalloc((void**)&h->dev.state, state_bytes);
    if (e == cudaSuccess) e = cudaMemset(h->dev.state, 0, state_bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        delete h;
        return vb200_set_cuda_error(e);
    }
    *comm = h;
    return VB200_OK;
}
extern "C" int vb200_comm_destroy(void* comm) {
    if (!comm) return VB200_OK;
    CommHost* h = (CommHost*)comm;
    cudaFree(h->dev.state);
    delete h;
    return VB200_OK;
}
// Returns VB200_ETIMEOUT if any kernel on this comm ever timed out waiting for a peer (synchronises).
extern "C" int vb200_comm_check(void* comm) {
    CommHost* h = (CommHost*)comm;
    uint32_t err = 0;
    VB_CUDA_TRY(cudaMemcpy(&err, h->dev.state + kStateError, 4, cudaMemcpyDeviceToHost));
    if (err) return vb200_set_error(VB200_ETIMEOUT, "a peer signal wait timed out");
    return VB200_OK;
}

static int check_ch(int ch) {
    if (ch < 0 || ch >= kMaxChannels) return vb200_set_error(VB200_EINVAL, "channel out of range");
    return 0;
}
static bool comm_cluster_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VB200_COMM_CLUSTER");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
static bool comm_gate_enabled() {  // VB200_COMM_GATE=0: no gate kernels (the collectives spin on all their CTAs; A-B runs)
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("VB200_COMM_GATE");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}
static int clamp_ctas(int n) { return n < 1 ? 1 : (n > 2 * kNumSMs ? 2 * kNumSMs : n); }

extern "C" int vb200_comm_barrier(void* comm, int32_t channel, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->dev, channel);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_allgather(void* comm, int32_t channel, int64_t region_offset, int64_t shard_bytes,
                               int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || shard_bytes < 0 || (shard_bytes & 1) ||
        region_offset + shard_bytes * h->dev.world > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "allgather: offset must be 256-byte aligned and inside the region");
    allgather_kernel<<<clamp_ctas(num_ctas), 512, 0, (cudaStream_t)stream>>>(h->dev, channel, region_offset, shard_bytes);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_reduce_scatter_f32(void* comm, int32_t channel, int64_t region_offset, int64_t chunk_elems,
                                        float scale, float* out, int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || chunk_elems < 0 ||
        region_offset + chunk_elems * 4 * h->dev.world > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "reduce_scatter: offset must be 256-byte aligned and inside the region");
    const int g = clamp_ctas(num_ctas);
    cudaStream_t st = (cudaStream_t)stream;
    // VB200_RS_GENERIC=1 forces the world-size-generic instantiation (the one 8 GPUs use) for testing on fewer GPUs
    const char* gen = getenv("VB200_RS_GENERIC");
    switch ((gen && gen[0] == '1') ? 0 : h->dev.world) {
        case 1: reduce_scatter_f32_kernel<1, 8><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        case 2: reduce_scatter_f32_kernel<2, 8><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        case 4: reduce_scatter_f32_kernel<4, 4><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        default: reduce_scatter_f32_kernel<0, 2><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
    }
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_reduce_scatter_bf16(void* comm, int32_t channel, int64_t region_offset, int64_t chunk_elems,
                                         float scale, float* out, int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || chunk_elems < 0 ||
        region_offset + chunk_elems * 2 * h->dev.world > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "reduce_scatter_bf16: offset must be 256-byte aligned and inside the region");
    const int g = clamp_ctas(num_ctas);
    cudaStream_t st = (cudaStream_t)stream;
    // VB200_RS_GENERIC=1 forces the world-size-generic instantiation (the one 8 GPUs use) for testing on fewer GPUs
    const char* gen = getenv("VB200_RS_GENERIC");
    switch ((gen && gen[0] == '1') ? 0 : h->dev.world) {
        case 1: reduce_scatter_bf16_kernel<1, 8><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        case 2: reduce_scatter_bf16_kernel<2, 8><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        case 4: reduce_scatter_bf16_kernel<4, 4><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
        default: reduce_scatter_bf16_kernel<0, 2><<<g, 512, 0, st>>>(h->dev, channel, region_offset, chunk_elems, scale, out); break;
    }
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

// desc: n x 4 int64 (src pointer, numel, chunk elements, row offset); out: [world, row_elems] bf16
extern "C" int vb200_fsdp_pack_bf16(const int64_t* desc, int32_t n, int32_t world, int64_t row_elems, void* out,
                                    int32_t src_dtype, void* stream) {
    if (n < 0 || world <= 0 || row_elems < 0 || (n > 0 && (!desc || !out)) || (src_dtype != 0 && src_dtype != 1))
        return vb200_set_error(VB200_EINVAL, "fsdp_pack_bf16: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    for (int base = 0; base < n; base += kPackMax) {
        PackArgs a;
        a.n = n - base < kPackMax ? n - base : kPackMax;
        a.world = world;
        a.row = row_elems;
        bool vec = (row_elems % 8 == 0) && (((uintptr_t)out & 15) == 0);
        int64_t biggest = 0;
        for (int i = 0; i < a.n; ++i) {
            const int64_t* d = desc + (int64_t)(base + i) * 4;
            a.e[i].src = (const void*)(uintptr_t)d[0];
            a.e[i].numel = d[1];
            a.e[i].chunk = d[2];
            a.e[i].off = d[3];
            if (d[1] < 0 || d[2] <= 0 || d[3] < 0 || d[3] + d[2] > row_elems || d[1] > d[2] * world)
                return vb200_set_error(VB200_EINVAL, "fsdp_pack_bf16: inconsistent descriptor");
            vec = vec && d[2] % 8 == 0 && d[3] % 8 == 0 && ((uintptr_t)d[0] & 15) == 0;
            if (d[2] * world > biggest) biggest = d[2] * world;
        }
        int64_t blocks = (biggest / 8 + 512 * 4 - 1) / (512 * 4);
        if (blocks < 1) blocks = 1;
        if (blocks > 148 * 2) blocks = 148 * 2;
        dim3 grid((unsigned)blocks, (unsigned)a.n);
        if (src_dtype == 0) {
            if (vec) fsdp_pack_bf16_kernel<__nv_bfloat16, true><<<grid, 512, 0, st>>>(a, (__nv_bfloat16*)out);
            else fsdp_pack_bf16_kernel<__nv_bfloat16, false><<<grid, 512, 0, st>>>(a, (__nv_bfloat16*)out);
        } else {
            if (vec) fsdp_pack_bf16_kernel<float, true><<<grid, 512, 0, st>>>(a, (__nv_bfloat16*)out);
            else fsdp_pack_bf16_kernel<float, false><<<grid, 512, 0, st>>>(a, (__nv_bfloat16*)out);
        }
        vb200_count_launch(1);
        VB_HOST_CHECK_LAUNCH();
    }
    return VB200_OK;
}


// table: n x 3 int64 = {byte offset of the parameter inside a shard row, bytes of one rank's shard, destination pointer}
extern "C" int vb200_allgather_scatter(void* comm, int32_t channel, int64_t region_offset, int64_t shard_bytes,
                                       const int64_t* table, int32_t n, int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || shard_bytes < 0 || (shard_bytes & 1) || n < 0 || (n > 0 && !table) ||
        region_offset + shard_bytes * h->dev.world > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "allgather_scatter: offset must be 256-byte aligned and inside the region");
    if (n > kScatterMax) return vb200_set_error(VB200_EINVAL, "allgather_scatter: at most 64 parameters per call");
    ScatterArgs a;
    a.n = n;
    a.vec_ok = (shard_bytes & 15) == 0;
    for (int i = 0; i < n; ++i) {
        a.off[i] = table[3 * i];
        a.bytes[i] = table[3 * i + 1];
        a.dst[i] = (uint8_t*)(uintptr_t)table[3 * i + 2];
        if (a.off[i] < 0 || a.bytes[i] < 0 || (a.bytes[i] & 1) || a.off[i] + a.bytes[i] > shard_bytes || (a.bytes[i] > 0 && !a.dst[i]))
            return vb200_set_error(VB200_EINVAL, "allgather_scatter: inconsistent parameter table");
        if ((a.off[i] | a.bytes[i] | (int64_t)(uintptr_t)a.dst[i]) & 15) a.vec_ok = 0;
    }
    // num_ctas counts 512-thread units. Launched as half as many 1024-thread CTAs in clusters of two: the same threads
    // and bytes in flight, but on a quarter of the TPCs — a communication CTA on an SM keeps the GEMMs' CTA pairs (cuBLAS
    // 2-SM tcgen05 kernels need a whole TPC, all of its registers and shared memory) off that TPC for as long as it runs.
    const bool pair = comm_cluster_enabled();  // VB200_COMM_CLUSTER=0: plain 512-thread CTAs (debugging / A-B)
    int ctas = pair ? (clamp_ctas(num_ctas) + 1) / 2 : clamp_ctas(num_ctas);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(pair ? 1024 : 512);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = (cudaStream_t)stream;
    cfg.numAttrs = 0;
    if (pair && ctas >= 2) {
        ctas &= ~1;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cfg.gridDim = dim3(ctas);
    const bool gate = comm_gate_enabled() && h->dev.world > 1;
    if (gate) comm_gate_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->dev, (int)channel, 0, region_offset, 1);
    VB_CUDA_TRY(cudaLaunchKernelEx(&cfg, allgather_scatter_kernel, h->dev, (int)channel, region_offset, shard_bytes, a));
    vb200_count_launch(gate ? 2 : 1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

// desc: n x 3 int64 = {gradient pointer (bf16, contiguous), numel, chunk elements}; the staging buffer
// [world, row_elems] bf16 sits at region_offset of every rank's region; out: row_elems fp32
extern "C" int vb200_reduce_scatter_push_bf16(void* comm, int32_t channel, int64_t region_offset, const int64_t* desc,
                                              int32_t n, int64_t row_elems, float scale, float* out, int32_t num_ctas,
                                              void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || row_elems < 0 || n < 0 || (n > 0 && (!desc || !out)) ||
        region_offset + row_elems * 2 * h->dev.world > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "reduce_scatter_push: staging must be 256-byte aligned and inside the region");
    if (n > kPushMax) return vb200_set_error(VB200_EINVAL, "reduce_scatter_push: at most 64 parameters per call");
    PushArgs a;
    a.n = n;
    a.world = h->dev.world;
    a.row = row_elems;
    a.vec_ok = 1;
    int64_t acc = 0;
    for (int i = 0; i < n; ++i) {
        a.src[i] = (const __nv_bfloat16*)(uintptr_t)desc[3 * i];
        a.numel[i] = desc[3 * i + 1];
        a.chunk[i] = desc[3 * i + 2];
        if (a.numel[i] < 0 || a.chunk[i] < 0 || a.numel[i] > a.chunk[i] * h->dev.world || (a.numel[i] > 0 && !a.src[i]))
            return vb200_set_error(VB200_EINVAL, "reduce_scatter_push: inconsistent descriptor");
        if ((a.chunk[i] & 7) || ((uintptr_t)a.src[i] & 15)) a.vec_ok = 0;
        acc += a.chunk[i];
    }
    if (acc != row_elems) return vb200_set_error(VB200_EINVAL, "reduce_scatter_push: chunks do not add up to the row");
    int g = clamp_ctas(num_ctas);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.blockDim = dim3(512);
    cfg.stream = (cudaStream_t)stream;
    if (comm_cluster_enabled() && g >= 2) {  // CTA pairs share a TPC (see vb200_allgather_scatter)
        g &= ~1;
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
    }
    cfg.gridDim = dim3(g);
    const char* gen = getenv("VB200_RS_GENERIC");
    const int ch32 = (int)channel;
    const int w = (gen && gen[0] == '1') ? 0 : h->dev.world;
    const bool gate = comm_gate_enabled() && h->dev.world > 1;
    cudaStream_t st = (cudaStream_t)stream;
    auto launch = [&](int phase) -> cudaError_t {
        switch (w) {
            case 1: return cudaLaunchKernelEx(&cfg, reduce_scatter_push_bf16_kernel<1>, h->dev, ch32, region_offset, a, scale, out, phase);
            case 2: return cudaLaunchKernelEx(&cfg, reduce_scatter_push_bf16_kernel<2>, h->dev, ch32, region_offset, a, scale, out, phase);
            case 4: return cudaLaunchKernelEx(&cfg, reduce_scatter_push_bf16_kernel<4>, h->dev, ch32, region_offset, a, scale, out, phase);
            default: return cudaLaunchKernelEx(&cfg, reduce_scatter_push_bf16_kernel<0>, h->dev, ch32, region_offset, a, scale, out, phase);
        }
    };
    if (gate) {
        comm_gate_kernel<<<1, 32, 0, st>>>(h->dev, ch32, 0, region_offset, 1);  // staging free on every peer
        VB_CUDA_TRY(launch(1));                                                    // push
        comm_gate_kernel<<<1, 32, 0, st>>>(h->dev, ch32, 1, region_offset, 0);  // every peer's pushes have landed here
        VB_CUDA_TRY(launch(2));                                                    // local rank-order reduction
        vb200_count_launch(4);
    } else {
        VB_CUDA_TRY(launch(3));
        vb200_count_launch(1);
    }
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_all_to_all(void* comm, int32_t channel, int64_t region_offset, int32_t n_desc,
                                const int64_t* desc /* n x 8 */, int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    if (n_desc < 1 || n_desc > 4) return vb200_set_error(VB200_EINVAL, "all_to_all: 1..4 descriptors");
    CommHost* h = (CommHost*)comm;
    A2AArgs a;
    a.n = n_desc;
    for (int i = 0; i < n_desc; ++i) {
        const int64_t* d = desc + 8 * i;
        a.d[i].src_off = d[0]; a.d[i].src_rank_stride = d[1]; a.d[i].src_row_stride = d[2];
        a.d[i].dst = (uint8_t*)(uintptr_t)d[3]; a.d[i].dst_peer_stride = d[4]; a.d[i].dst_row_stride = d[5];
        a.d[i].rows = d[6]; a.d[i].seg_bytes = d[7];
        if ((d[0] | d[1] | d[2] | d[3] | d[4] | d[5] | d[7]) & 15)
            return vb200_set_error(VB200_EINVAL, "all_to_all: offsets, strides and segments must be 16-byte multiples");
        const int64_t last = region_offset + d[0] + (h->dev.world - 1) * d[1] + (d[6] > 0 ? (d[6] - 1) * d[2] : 0) + d[7];
        if (d[0] < 0 || region_offset < 0 || (region_offset & 255) || last > h->data_bytes)
            return vb200_set_error(VB200_EINVAL, "all_to_all: source range outside the symmetric region");
    }
    all_to_all_kernel<<<clamp_ctas(num_ctas), 512, 0, (cudaStream_t)stream>>>(h->dev, channel, region_offset, a);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}

extern "C" int vb200_chunk_pull(void* comm, int32_t channel, int64_t region_offset, const void* chunks,
                                int32_t nchunks, void* dst, int32_t num_ctas, void* stream) {
    if (check_ch(channel)) return VB200_EINVAL;
    CommHost* h = (CommHost*)comm;
    if (region_offset < 0 || (region_offset & 255) || region_offset > h->data_bytes)
        return vb200_set_error(VB200_EINVAL, "chunk_pull: offset must be 256-byte aligned and inside the region");
    chunk_pull_kernel<<<clamp_ctas(num_ctas), 512, 0, (cudaStream_t)stream>>>(
        h->dev, channel, region_offset, (const ChunkDesc*)chunks, nchunks, (uint8_t*)dst);
    vb200_count_launch(1);
    VB_HOST_CHECK_LAUNCH();
    return VB200_OK;
}
