// Gradient all-reduce over NVLink peer memory: ONE kernel per arena bucket, launched like any other kernel of the step, so it
// can sit inside the step's CUDA graph on a forked stream and overlap the rest of backward.
//
// Replaces the implicit gradient exchange of the reference's data-parallel wrapper (nn.DataParallel at
// diff_cifar/main.py:235-238; DDP in pdearena's Lightning trainer): every replica ends the step with the SUM of all
// replicas' gradients (the 1/world mean is folded into the optimiser kernel's grad_scale).
//
// Every rank owns a "symmetric" allocation of the same size (cudaMalloc here, exported with cudaIpcGetMemHandle, opened by the
// other ranks of the box with cudaIpcOpenMemHandle): [flag area | fp32 gradient arena].  On B200 boxes all peers sit behind
// NVSwitch, so a plain ld/st to a peer-mapped address travels at NVLink speed.
//
// Two-shot, in place, deterministic:  rank r owns the r-th 1/world slice of the bucket.  After a start barrier (every rank's
// bucket is complete) it reads its slice from every rank in rank order, adds, and writes the sum into the slice of EVERY
// rank; an end barrier makes the result visible before anyone consumes it.  A slice is read only by its owner and each
// element is read before the same thread overwrites it, so no staging buffer is needed, and every replica receives bitwise
// the same sum (the replicas cannot drift apart).  Per rank and bucket of B bytes: (world-1)/world * B read from and
// written to peers, against 2*(world-1)/world * B for a ring.
//
// Barriers are per CTA (CTA b of every rank meets CTA b of every other rank) through monotonically increasing epochs kept in
// device memory: nothing is reset between launches, so the kernel replays from a CUDA graph.  Waits are bounded (~30 s): a
// lost peer traps instead of hanging the GPU.
#include <string.h>

#include "common.cuh"

namespace {
using namespace ub;

constexpr int kMaxWorld = 8;
// CTAs per launch: identical on every rank and launch (the barrier pairs CTAs by index); UB200_P2P_BLOCKS overrides the default
// (same value on every rank).  A CTA is 256 threads at <= 64 registers and no shared memory, so it fits NEXT TO the one-per-SM
// CTAs of the persistent conv kernels (96-160 registers x 320 threads): an overlapped bucket takes issue slots, not SMs.
constexpr int kArMaxBlocks = 148;
constexpr int kArThreads = 256;
// flag area (uint32 words): start[kArMaxBlocks][kMaxWorld] | end[kArMaxBlocks][kMaxWorld] | epoch[kArMaxBlocks] (local only)
constexpr int kFlagWords = 2 * kArMaxBlocks * kMaxWorld + kArMaxBlocks;

struct Peers {
    float *buf[kMaxWorld];             // gradient arena of every rank (own pointer at [rank])
    uint32_t *flags[kMaxWorld];        // flag area of every rank
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint64_t evict_first_policy() {
    uint64_t pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// peer data: never from a stale local cache line (system scope).  EF: L2 evict-first, so that a bucket streaming through does
// not push the activations the concurrent backward kernels re-read out of L2.
template <bool EF>
__device__ __forceinline__ float4 ld_peer(const float4 *p) {
    float4 v;
    if constexpr (EF)
        asm volatile("ld.relaxed.sys.global.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(evict_first_policy()) : "memory");
    else
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
template <bool EF>
__device__ __forceinline__ void st_peer(float4 *p, const float4 &v) {
    if constexpr (EF)
        asm volatile("st.relaxed.sys.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;"
                     ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(evict_first_policy()) : "memory");
    else
        asm volatile("st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// CTA b of every rank meets CTA b of every other rank.  which = 0 start, 1 end.
__device__ __forceinline__ void cta_barrier(const Peers &P, int rank, int world, int which, uint32_t epoch) {
    __syncthreads();                                   // every thread of this CTA has issued its stores
    if ((int)threadIdx.x < world) {
        const int peer = threadIdx.x;
        const int word = (which * kArMaxBlocks + blockIdx.x) * kMaxWorld;
        __threadfence_system();
        st_release_sys(P.flags[peer] + word + rank, epoch);             // "rank has arrived" in the peer's flag area
        const uint32_t *mine = P.flags[rank] + word + peer;
        const long long t0 = clock64();
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
            if (clock64() - t0 > 60000000000LL) __trap();                // ~30 s at 1.9 GHz: a peer is gone
        }
    }
    __syncthreads();
}

// WORLD is a template parameter so that the loads of all ranks for U = 8 / WORLD consecutive turns are in flight together:
// eight 16-byte loads per thread whatever the world size (two ranks are otherwise latency-bound: 0.6 MB on the wire).
template <int WORLD, bool EF>
__global__ void __launch_bounds__(kArThreads, 4) p2p_allreduce_kernel(const Peers P, int rank, int64_t off4, int64_t n4) {
    constexpr int U = 8 / WORLD;
    uint32_t *epoch_slot = P.flags[rank] + 2 * kArMaxBlocks * kMaxWorld + blockIdx.x;
    const uint32_t epoch = *epoch_slot + 1u;           // same sequence of launches on every rank
    cta_barrier(P, rank, WORLD, 0, epoch);
    // this rank's slice of the bucket, in float4 units
    const int64_t per = (n4 + WORLD - 1) / WORLD;
    const int64_t lo = off4 + (int64_t)rank * per;
    int64_t hi = lo + per;
    if (hi > off4 + n4) hi = off4 + n4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += U * stride) {
        float4 a[U][WORLD];
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int p = 0; p < WORLD; ++p)
                if (i + u * stride < hi) a[u][p] = ld_peer<EF>(reinterpret_cast<const float4 *>(P.buf[p]) + i + u * stride);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u * stride < hi) {
                float4 s = a[u][0];
#pragma unroll
                for (int p = 1; p < WORLD; ++p) { s.x += a[u][p].x; s.y += a[u][p].y; s.z += a[u][p].z; s.w += a[u][p].w; }   // rank order
#pragma unroll
                for (int p = 0; p < WORLD; ++p) st_peer<EF>(reinterpret_cast<float4 *>(P.buf[p]) + i + u * stride, s);
            }
        }
    }
    cta_barrier(P, rank, WORLD, 1, epoch);
    if (threadIdx.x == 0) *epoch_slot = epoch;
}

}  // namespace

extern "C" {

size_t ub200_p2p_flag_bytes(void) { return ((size_t)kFlagWords * 4 + 1023) & ~(size_t)1023; }

int ub200_p2p_alloc(size_t bytes, void **ptr, unsigned char *handle64) {
    UB_REQUIRE(ptr && handle64 && bytes > 0, UB200_E_BADARG);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return (int)e; }
    memcpy(handle64, &h, 64);
    *ptr = p;
    return UB200_OK;
}

int ub200_p2p_free(void *ptr) { return ptr ? (int)cudaFree(ptr) : UB200_OK; }

int ub200_p2p_open(const unsigned char *handle64, void **peer_ptr) {
    UB_REQUIRE(handle64 && peer_ptr, UB200_E_BADARG);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    return (int)cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess);
}

int ub200_p2p_close(void *peer_ptr) { return peer_ptr ? (int)cudaIpcCloseMemHandle(peer_ptr) : UB200_OK; }

int ub200_p2p_allreduce_sum_f32(void *const *bases, int rank, int world, int64_t offset_floats, int64_t count, int ctas,
                                void *stream) {
    UB_REQUIRE(bases && world >= 2 && world <= kMaxWorld && rank >= 0 && rank < world, UB200_E_BADARG);
    UB_REQUIRE(offset_floats >= 0 && count > 0, UB200_E_BADARG);
    UB_REQUIRE(offset_floats % 4 == 0 && count % 4 == 0, UB200_E_UNSUPPORTED);      // float4 granularity (arena slices are)
    Peers P{};
    const size_t fb = ub200_p2p_flag_bytes();
    for (int p = 0; p < world; ++p) {
        UB_REQUIRE(bases[p] && ub::aligned16(bases[p]), UB200_E_BADARG);
        P.flags[p] = reinterpret_cast<uint32_t *>(bases[p]);
        P.buf[p] = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(bases[p]) + fb);
    }
    static const int def_blocks = [] {
        const char *e = getenv("UB200_P2P_BLOCKS");
        const int b = e ? atoi(e) : kArMaxBlocks;
        return b < 1 ? 1 : (b > kArMaxBlocks ? kArMaxBlocks : b);
    }();
    const int blocks = ctas <= 0 ? def_blocks : (ctas > kArMaxBlocks ? kArMaxBlocks : ctas);
    cudaStream_t st = ub::as_stream(stream);
    // diagnostics (A/B of what an overlapped bucket costs the kernels it runs next to): 1 = no launch at all, 2 = barriers only
    static const int dbg = [] { const char *e = getenv("UB200_P2P_DEBUG"); return e ? atoi(e) : 0; }();
    if (dbg == 1) return UB200_OK;
    const int64_t o4 = offset_floats / 4, n4 = dbg == 2 ? 0 : count / 4;
    static const int ef = [] { const char *e = getenv("UB200_P2P_EVICT_FIRST"); return e ? atoi(e) : 0; }();
    // (same shared-memory carve-out as the conv kernels, or a bucket could not be resident next to them: common.cuh)
#define P2P_LAUNCH(W)                                                                                    \
    if (ef) {                                                                                            \
        prefer_max_smem_carveout(reinterpret_cast<const void *>(p2p_allreduce_kernel<W, true>));         \
        p2p_allreduce_kernel<W, true><<<blocks, kArThreads, 0, st>>>(P, rank, o4, n4);                   \
    } else {                                                                                             \
        prefer_max_smem_carveout(reinterpret_cast<const void *>(p2p_allreduce_kernel<W, false>));        \
        p2p_allreduce_kernel<W, false><<<blocks, kArThreads, 0, st>>>(P, rank, o4, n4);                  \
    }                                                                                                    \
    break
    switch (world) {                       // one instantiation per world size
        case 2: P2P_LAUNCH(2);
        case 3: P2P_LAUNCH(3);
        case 4: P2P_LAUNCH(4);
        case 5: P2P_LAUNCH(5);
        case 6: P2P_LAUNCH(6);
        case 7: P2P_LAUNCH(7);
        default: P2P_LAUNCH(8);
    }
#undef P2P_LAUNCH
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
