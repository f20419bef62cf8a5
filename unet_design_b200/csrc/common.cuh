// Shared helpers for the sm_100a kernels of the U-Net hot path.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_set>

#include "unet_b200.h"

#define UB_LAUNCH_CHECK()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

#define UB_REQUIRE(cond, code) \
    do {                       \
        if (!(cond)) return (code); \
    } while (0)

namespace ub {

constexpr int kSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline cudaStream_t as_stream(void *s) { return reinterpret_cast<cudaStream_t>(s); }

// Blocks for a grid-stride kernel: enough to cover `work` items, capped at `waves` full waves of
// `per_sm` resident CTAs on every SM.
inline int grid_for(int64_t work, int threads, int per_sm, int waves = 4) {
    int64_t need = (work + threads - 1) / threads;
    int64_t cap = (int64_t)kSMs * per_sm * waves;
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

// Programmatic dependent launch (PDL).  A kernel launched with the programmatic-stream-serialization attribute may start
// while its predecessor in the stream is still draining: its CTAs become resident as SMs free up and run their prologue
// (barrier init, TMEM allocation, descriptor prefetch), then block in pdl_wait() until the predecessor has completed and
// its memory is visible.  pdl_wait() must precede the first read of upstream data AND the first global write (the
// predecessor may still be reading a buffer this kernel overwrites).  pdl_trigger() lets the NEXT kernel do the same.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// launch attribute for the kernels above; UB200_PDL=0 switches it off (A/B)
inline bool pdl_enabled() {
    static const bool on = [] { const char *e = getenv("UB200_PDL"); return !(e && e[0] == '0'); }();
    return on;
}
inline void pdl_attr(cudaLaunchAttribute &a) {
    a.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    a.val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
}

// Shared-memory carve-out.  Two kernels are resident on one SM together only when they agree on the SM's L1 / shared-memory
// split; a kernel that wants another split waits until the SM has drained.  The persistent conv kernels need the largest
// shared-memory configuration, so every kernel that is meant to run NEXT TO them on another stream (weight-gradient side
// stream, column sums, all-reduce buckets, the streaming GroupNorm kernels) asks for the same one.  None of them lives off L1:
// their global accesses are streaming.  UB200_MAX_CARVEOUT=0 switches the request off (A/B).
inline void prefer_max_smem_carveout(const void *kernel) {
    static const bool on = [] { const char *e = getenv("UB200_MAX_CARVEOUT"); return !(e && e[0] == '0'); }();
    if (!on) return;
    static std::mutex mu;
    static std::unordered_set<const void *> done;
    std::lock_guard<std::mutex> lk(mu);
    if (done.insert(kernel).second)
        cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
}

// kernel<<<grid, block, smem, stream>>>(args...) with the PDL attribute (the kernel must call pdl_trigger / pdl_wait)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args &&...args) {
    prefer_max_smem_carveout(reinterpret_cast<const void *>(kernel));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute attr[1];
    pdl_attr(attr[0]);
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// Streaming (read-once / write-once) 128-bit accesses: keep them out of L1.
__device__ __forceinline__ float4 ld_stream(const float4 *p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream(float4 *p, const float4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ld_stream_u4(const uint4 *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream_u4(uint4 *p, const uint4 &v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ float bf16lo(uint32_t v) { return __uint_as_float(v << 16); }
__device__ __forceinline__ float bf16hi(uint32_t v) { return __uint_as_float(v & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t *>(&t);
}
__device__ __forceinline__ void unpack8(const uint4 &v, float (&f)[8]) {
    f[0] = bf16lo(v.x); f[1] = bf16hi(v.x); f[2] = bf16lo(v.y); f[3] = bf16hi(v.y);
    f[4] = bf16lo(v.z); f[5] = bf16hi(v.z); f[6] = bf16lo(v.w); f[7] = bf16hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ub
