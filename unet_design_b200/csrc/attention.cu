// Attention core of AttnBlock (diff_cifar/model.py:100-119: softmax(Q K^T / sqrt(C)) V, one head, T = H*W tokens) and its
// backward on tcgen05 tensor cores, as SIX calls of one batched GEMM kernel with fused epilogues:
//
//   forward    P  = softmax_rows(alpha * Q K^T)        (K-major A, K-major B, epilogue SOFTMAX)      P is kept for backward
//              O  = P V                                (K-major A, MN-major B)
//   backward   dS = P o (dO V^T - rowsum(P o dO V^T)) * scale      (K-major A, K-major B, epilogue SOFTMAX_BWD)
//              dV = P^T dO                             (MN-major A, MN-major B)
//              dQ = dS K                               (K-major A, MN-major B)
//              dK = dS^T Q                             (MN-major A, MN-major B)
//
// Every operand is a row-major bf16 matrix [R = samples * T rows, ld] (q, k, v are channel slices of the fused qkv
// projection, the gradients slices of one dqkv buffer), so "transposed" operands cost nothing: a TMA box of 64 rows x 64
// columns lands in shared memory as the MN-major swizzled layout tcgen05 consumes in place (the conv_wgrad trick), the
// plain ones as the K-major layout (conv_fprop).  One CTA owns a group of 256 consecutive rows = 256 / T whole samples:
// M = 256 (two 128-row accumulators), N <= 256, K <= 256, so the score tile of a group fits TMEM exactly (2 x 256
// columns) and a softmax row is one thread's registers.  Samples shorter than 256 tokens share a group; scores between
// different samples are masked to zero probability (block-diagonal), which makes every product above exact per sample.
//
// Replaces F.scaled_dot_product_attention (PyTorch's sm80-style flash kernels) at unet_design_b200/diff_cifar/model.py.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "tc_common.cuh"

namespace {
using namespace ub;
using namespace ub::tc;

constexpr int kAttThreads = 320;                 // TMA warp, MMA warp, 8 epilogue warps (2 M-tiles x 4 lane quadrants)
constexpr uint32_t kHalfStage = 32 * 1024;       // one operand's k-chunk: 256 (rows | MN) x 64 (K | rows) bf16
constexpr uint32_t kStageBytes = 2 * kHalfStage;
constexpr int kStages = 3;

struct BgemmParams {
    int groups, N, K;                            // 256-row groups, output columns, reduction extent
    int a_mn, b_mn;                              // operand majors
    int epilogue, T;
    float alpha;
    const __nv_bfloat16 *p; int64_t ld_p;        // SOFTMAX_BWD: P [R, 256]
};

__device__ __forceinline__ void sts_u4(uint32_t addr, const uint4 &v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 lds_u4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

__global__ void __launch_bounds__(kAttThreads, 1) bgemm256_kernel(const __grid_constant__ CUtensorMap tm_a,
                                                                const __grid_constant__ CUtensorMap tm_b,
                                                                const __grid_constant__ CUtensorMap tm_out,
                                                                const BgemmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + kStages * kStageBytes);
    uint64_t *empty = full + kStages;
    uint64_t *acc_full = empty + kStages;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(acc_full + 1);
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int g = blockIdx.x;
    const int kchunks = p.K >> 6, nchunks = p.N >> 6;
    pdl_trigger();
    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm_a); prefetch_tmap(&tm_b); prefetch_tmap(&tm_out);
        for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
        mbar_init(acc_full, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    pdl_wait();
    const int row0 = g * 256;

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform loop, lane 0 issues) =====================
        const bool leader = lane == 0;
        const uint32_t base = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
        // bytes per stage: A is always 256 x 64 elements (32 KB); B is N x 64 (K-major: one box of N rows; MN-major: N/64 boxes)
        const uint32_t tx = kHalfStage + (uint32_t)p.N * 128u;
        int s = 0; uint32_t ph = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
            const uint32_t fb = full0 + 8u * s, sa = base + (uint32_t)s * kStageBytes, sb = sa + kHalfStage;
            mbar_wait_a(empty0 + 8u * s, ph ^ 1u);
            if (leader) {
                mbar_arrive_expect_tx_a(fb, tx);
                if (!p.a_mn) {            // A[m, k]: rows = m, two boxes of 128 rows x 64 k
                    tma_load_2d_a(sa, &tm_a, fb, 64 * kc, row0);
                    tma_load_2d_a(sa + 16384, &tm_a, fb, 64 * kc, row0 + 128);
                } else {                  // A stored [k rows][m contiguous]: four boxes of 64 k-rows x 64 m
                    for (int j = 0; j < 4; ++j) tma_load_2d_a(sa + 8192u * j, &tm_a, fb, 64 * j, row0 + 64 * kc);
                }
                if (!p.b_mn) {            // B[n, k]: rows = n (the group's 256 keys), one box of N rows x 64 k
                    tma_load_2d_a(sb, &tm_b, fb, 64 * kc, row0);
                } else {                  // B stored [k rows][n contiguous]: N/64 boxes of 64 k-rows x 64 n
                    for (int j = 0; j < nchunks; ++j) tma_load_2d_a(sb + 8192u * j, &tm_b, fb, 64 * j, row0 + 64 * kc);
                }
            }
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        const bool leader = lane == 0;
        const uint32_t idesc = make_idesc(128, p.N, p.a_mn, p.b_mn);
        const uint32_t base = smem_u32(smem), full0 = smem_u32(full), empty0 = smem_u32(empty);
        int s = 0; uint32_t ph = 0;
        for (int kc = 0; kc < kchunks; ++kc) {
            mbar_wait_a(full0 + 8u * s, ph);
            tc_fence_after();
            const uint32_t sa = base + (uint32_t)s * kStageBytes, sb = sa + kHalfStage;
            if (leader) {
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    // K-major: 128 rows x 128 bytes, 8-row groups 1024 bytes apart, k-step = 32 bytes inside the row
                    // MN-major: 64-element chunks 8 KB apart (LBO), 8 k-row groups 1024 bytes apart, k-step = 16 rows = 2 KB
                    const uint64_t da = p.a_mn ? make_smem_desc(sa + 16384u * mt, 8192, 1024, 2u) : make_smem_desc(sa + 16384u * mt, 16, 1024, 2u);
                    const uint64_t db = p.b_mn ? make_smem_desc(sb, 8192, 1024, 2u) : make_smem_desc(sb, 16, 1024, 2u);
                    const uint32_t ka = p.a_mn ? 128u : 2u, kb = p.b_mn ? 128u : 2u;       // descriptor units of 16 bytes
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_bf16(tmem_base + 256u * mt, da + (uint64_t)(k * ka), db + (uint64_t)(k * kb), idesc, (kc | k) ? 1u : 0u);
                }
                umma_commit_a(empty0 + 8u * s);
            }
            if (++s == kStages) { s = 0; ph ^= 1u; }
        }
        if (leader) umma_commit(acc_full);
    } else {
        // ===================== epilogue: thread = one row of the 256 x N tile =====================
        const int mt = (warp - 2) >> 2, qd = warp & 3;
        const int r = mt * 128 + qd * 32 + lane;                 // row inside the group
        const uint32_t trow = tmem_base + ((uint32_t)(qd * 32) << 16) + 256u * mt;
        mbar_wait(acc_full, 0);
        tc_fence_after();
        // the smem ring is free now (every MMA has completed): reuse it as the output staging, one [128 rows x 128 B]
        // swizzled box per (M-tile, 64-column chunk) = 16 KB each, 8 boxes
        const uint32_t stage0 = smem_u32(smem);
        const uint32_t rl = (uint32_t)(qd * 32 + lane);          // row inside the M-tile
        const uint32_t row_off = rl * 128u, sw = rl & 7u;
        float scale_row = p.alpha, delta = 0.f, mx = -INFINITY;
        const int blk = r / p.T;                                 // sample inside the group (block-diagonal mask)
        const int c_lo = blk * p.T, c_hi = c_lo + p.T;
        if (p.epilogue == UB200_BGEMM_SOFTMAX) {
            for (int ch = 0; ch < nchunks; ++ch) {
                float v[64];
                tmem_ld64(trow + 64u * ch, v);
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const int c = 64 * ch + i;
                    if (c >= c_lo && c < c_hi) mx = fmaxf(mx, v[i]);
                }
            }
            mx *= p.alpha;                                       // alpha > 0 (scale * log2 e)
        } else if (p.epilogue == UB200_BGEMM_SOFTMAX_BWD) {
            const __nv_bfloat16 *prow = p.p + (int64_t)(row0 + r) * p.ld_p;
            for (int ch = 0; ch < nchunks; ++ch) {
                float v[64];
                tmem_ld64(trow + 64u * ch, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f[8];
                    unpack8(__ldg(reinterpret_cast<const uint4 *>(prow + 64 * ch) + j), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) delta = fmaf(f[i], v[8 * j + i], delta);
                }
            }
        }
        float sum = 0.f;
        for (int ch = 0; ch < nchunks; ++ch) {
            float v[64];
            tmem_ld64(trow + 64u * ch, v);
            if (p.epilogue == UB200_BGEMM_SOFTMAX) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const int c = 64 * ch + i;
                    const float e = (c >= c_lo && c < c_hi) ? ex2(fmaf(v[i], p.alpha, -mx)) : 0.f;
                    sum += e;
                    v[i] = e;
                }
            } else if (p.epilogue == UB200_BGEMM_SOFTMAX_BWD) {
                const __nv_bfloat16 *prow = p.p + (int64_t)(row0 + r) * p.ld_p;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f[8];
                    unpack8(__ldg(reinterpret_cast<const uint4 *>(prow + 64 * ch) + j), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[8 * j + i] = f[i] * (v[8 * j + i] - delta) * p.alpha;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) v[i] *= scale_row;
            }
            const uint32_t box = stage0 + (uint32_t)(mt * 4 + ch) * 16384u;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float f[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = v[8 * j + i];
                sts_u4(box + row_off + (((uint32_t)j ^ sw) << 4), pack8(f));
            }
        }
        if (p.epilogue == UB200_BGEMM_SOFTMAX) {
            // normalise in place: the row sits in shared memory as unnormalised bf16 exponentials
            const float inv = 1.0f / sum;
            for (int ch = 0; ch < nchunks; ++ch) {
                const uint32_t box = stage0 + (uint32_t)(mt * 4 + ch) * 16384u;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const uint32_t a = box + row_off + (((uint32_t)j ^ sw) << 4);
                    float f[8];
                    unpack8(lds_u4(a), f);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[i] *= inv;
                    sts_u4(a, pack8(f));
                }
            }
        }
        tc_fence_before();
        fence_proxy_async();
        named_barrier_sync(1, 256);                              // all eight epilogue warps have written their rows
        if (warp == 2 && lane == 0) {
            for (int m2 = 0; m2 < 2; ++m2)
                for (int ch = 0; ch < nchunks; ++ch)
                    tma_store_2d(&tm_out, stage0 + (uint32_t)(m2 * 4 + ch) * 16384u, 64 * ch, row0 + 128 * m2);
            tma_store_commit();
            tma_store_wait_all();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

#undef UB_REQUIRE
#define UB_REQUIRE(cond, code)                                                                              \
    do {                                                                                                    \
        if (!(cond)) {                                                                                      \
            if (getenv("UB200_DEBUG")) fprintf(stderr, "ub200_bgemm256: requirement failed: %s\n", #cond);  \
            return (code);                                                                                  \
        }                                                                                                   \
    } while (0)

extern "C" int ub200_bgemm256(const ub200_bgemm_args *a, void *stream) {
    UB_REQUIRE(a && a->a && a->b && a->out && a->groups > 0, UB200_E_BADARG);
    UB_REQUIRE(a->N > 0 && a->N <= 256 && a->N % 64 == 0 && a->K > 0 && a->K <= 256 && a->K % 64 == 0, UB200_E_UNSUPPORTED);
    UB_REQUIRE(a->epilogue >= UB200_BGEMM_PLAIN && a->epilogue <= UB200_BGEMM_SOFTMAX_BWD, UB200_E_BADARG);
    UB_REQUIRE(a->T > 0 && a->T <= 256 && 256 % a->T == 0, UB200_E_UNSUPPORTED);
    UB_REQUIRE(a->epilogue == UB200_BGEMM_PLAIN || a->N == 256, UB200_E_UNSUPPORTED);          // softmax rows span the group's 256 keys
    UB_REQUIRE(a->epilogue != UB200_BGEMM_SOFTMAX_BWD || (a->p && a->ld_p >= 256 && a->ld_p % 8 == 0 && ub::aligned16(a->p)),
               UB200_E_BADARG);
    UB_REQUIRE(a->epilogue != UB200_BGEMM_SOFTMAX || a->alpha > 0.f, UB200_E_BADARG);
    // MN-major operands span the group's 256 rows in K; a K-major B holds the group's 256 keys as its N rows
    UB_REQUIRE((!a->a_mn_major || a->K == 256) && (!a->b_mn_major || a->K == 256) && (a->b_mn_major || a->N == 256), UB200_E_UNSUPPORTED);
    UB_REQUIRE(a->ld_a % 8 == 0 && a->ld_b % 8 == 0 && a->ld_out % 8 == 0 && ub::aligned16(a->a) && ub::aligned16(a->b) &&
                   ub::aligned16(a->out) && a->groups < (1 << 22),
               UB200_E_UNSUPPORTED);
    const int64_t R = a->groups * 256;
    // columns each operand exposes: K-major A: K; MN-major A: 256 (M); K-major B: K; MN-major B: N
    const int64_t cols_a = a->a_mn_major ? 256 : a->K, cols_b = a->b_mn_major ? a->N : a->K;
    UB_REQUIRE(a->ld_a >= cols_a && a->ld_b >= cols_b && a->ld_out >= a->N, UB200_E_BADARG);
    CUtensorMap ta, tb, to;
    {
        const int64_t dims[2] = {cols_a, R};
        const int64_t str[1] = {a->ld_a};
        const int box[2] = {64, a->a_mn_major ? 64 : 128};
        int rc = encode_bf16_tensor_map(&ta, a->a, 2, dims, str, box);
        UB_REQUIRE(rc == 0, rc);
    }
    {
        const int64_t dims[2] = {cols_b, R};
        const int64_t str[1] = {a->ld_b};
        const int box[2] = {64, a->b_mn_major ? 64 : (int)a->N};
        int rc = encode_bf16_tensor_map(&tb, a->b, 2, dims, str, box);
        UB_REQUIRE(rc == 0, rc);
    }
    {
        const int64_t dims[2] = {a->N, R};
        const int64_t str[1] = {a->ld_out};
        const int box[2] = {64, 128};
        int rc = encode_bf16_tensor_map(&to, a->out, 2, dims, str, box);
        UB_REQUIRE(rc == 0, rc);
    }
    BgemmParams p{};
    p.groups = (int)a->groups; p.N = (int)a->N; p.K = (int)a->K; p.a_mn = a->a_mn_major ? 1 : 0; p.b_mn = a->b_mn_major ? 1 : 0;
    p.epilogue = a->epilogue; p.T = a->T; p.alpha = a->alpha;
    p.p = reinterpret_cast<const __nv_bfloat16 *>(a->p); p.ld_p = a->ld_p;
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(bgemm256_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    });
    if (attr_err != cudaSuccess) return (int)attr_err;
    const size_t smem = 1024 + (size_t)kStages * kStageBytes + (2 * kStages + 1) * sizeof(uint64_t) + 16;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)a->groups, 1, 1);
    cfg.blockDim = dim3(kAttThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ub::as_stream(stream);
    cudaLaunchAttribute attr[1];
    ub::pdl_attr(attr[0]);
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, bgemm256_kernel, ta, tb, to, p);
    if (le != cudaSuccess) return (int)le;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}
