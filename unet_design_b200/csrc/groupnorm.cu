// GroupNorm statistics, fused GroupNorm-apply + activation (+ scale/shift, + dropout), and backward.
// NHWC bf16 activations, fp32 maths.  Memory-bound: every pass streams whole pixels (all channels,
// 16 bytes per thread) so global accesses are fully coalesced whatever the group width.
//
// Replaces nn.GroupNorm + Swish/SiLU/GELU (+ nn.Dropout) of the reference's conv blocks:
// diff_cifar/model.py:130-132,:139-142,:393-395; diff_mnist/.../layers.py:284-288,:330-334;
// pdearena/pdearena/modules/twod_unetbase.py:30-31 (GroupNorm(1, C) + GELU).
//
// stats layout: float [N, G, 2] = (sum, sum of squares) over the (C/G)*HW slab; consumers derive
// mean / rstd (biased variance, eps) themselves, so a producer (this file's stats kernel, or the conv
// epilogue) only ever accumulates.
#include <cooperative_groups.h>

#include <cstdlib>
#include <mutex>
#include <unordered_set>

#include "tc_common.cuh"

namespace cg = cooperative_groups;

namespace {
using namespace ub;

// ---- Philox4x32-10 (counter-based; the same (seed, offset, index) regenerates the mask in backward)
__device__ __forceinline__ uint4 philox4(uint64_t seed, uint64_t ctr) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)ctr, c1 = (uint32_t)(ctr >> 32), c2 = 0x9E3779B9u, c3 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        c0 = hi1 ^ c1 ^ k0; c1 = lo1; c2 = hi0 ^ c3 ^ k1; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}

// keep-mask for the 8 channels starting at element index `elem` (multiple of 8): ONE Philox call per 8 elements, 16 random
// bits each (the keep probability is quantised to 1/65536: p = 0.1 -> 0.100006).  These kernels are instruction-issue bound
// and the generator is ~60 of their ~115 instructions per 16-byte chunk, so two calls per chunk (32 bits per element) cost
// a quarter of the dropout layers' time for resolution nobody needs.
__device__ __forceinline__ void dropout_mask8(uint64_t seed, uint64_t offset, int64_t elem, float p, float (&m)[8]) {
    const uint32_t thr = (uint32_t)fminf(p * 65536.0f, 65535.0f);
    const float keep_scale = 1.0f / (1.0f - p);
    const uint4 r0 = philox4(seed, offset + (uint64_t)(elem >> 3));
    const uint32_t r[8] = {r0.x & 0xffffu, r0.x >> 16, r0.y & 0xffffu, r0.y >> 16, r0.z & 0xffffu, r0.z >> 16, r0.w & 0xffffu, r0.w >> 16};
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = r[u] >= thr ? keep_scale : 0.f;
}

// The activation takes the PRE-SCALED argument h = kPre<ACT> * z: the factor is folded into the per-channel affine
// coefficients (z = x*A + B is one fma per element either way), which removes the `0.5 * z` every SiLU evaluation starts
// with.  SiLU through ONE special-function op (tanh.approx.f32, relative error 2^-11: far below the bf16 rounding of every
// value this feeds), with h = z/2 and t = tanh(h):
//     silu(z)  = z * (0.5 t + 0.5)              = h * t + h                       (1 MUFU + 1 FMA)
//     silu'(z) = s + z s (1 - s),  s = 0.5t+0.5 = 0.5 * (1 + t + h (1 - t^2))     (1 MUFU + 3 FMA)
// These kernels are instruction-issue bound, not bandwidth bound, on B200: every instruction per element counts.
template <int ACT>
__device__ __forceinline__ constexpr float act_pre_scale() { return ACT == UB200_ACT_SILU ? 0.5f : 1.0f; }

__device__ __forceinline__ float tanh_approx(float h) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return t;
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float h) {
    if constexpr (ACT == UB200_ACT_SILU) return fmaf(h, tanh_approx(h), h);
    else if constexpr (ACT == UB200_ACT_GELU) return 0.5f * h * (1.0f + erff(h * 0.70710678118654752f));
    else if constexpr (ACT == UB200_ACT_RELU) return fmaxf(h, 0.f);
    else return h;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float h) {   // d act / d z at z = h / kPre
    if constexpr (ACT == UB200_ACT_SILU) {
        const float t = tanh_approx(h);
        return fmaf(fmaf(h, fmaf(-t, t, 1.0f), t), 0.5f, 0.5f);
    } else if constexpr (ACT == UB200_ACT_GELU) {
        return 0.5f * (1.0f + erff(h * 0.70710678118654752f)) + h * 0.3989422804014327f * __expf(-0.5f * h * h);
    } else if constexpr (ACT == UB200_ACT_RELU) {
        return h > 0.f ? 1.0f : 0.f;               // torch: relu'(0) = 0
    } else return 1.0f;
}

struct Shape {
    int64_t HW; int C, G, cpg, chunks, rows;   // rows = pixels handled per pass by one CTA
    int64_t pix_per_cta;
};

// ---------------------------------------------------------------------------------------------
// statistics: grid (splits, N); each CTA streams its pixel range, every thread owns one 8-channel chunk
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gn_stats_kernel(const __nv_bfloat16 *__restrict__ x, int64_t ld, Shape sh,
                                                      float *__restrict__ stats) {
    extern __shared__ float sg[];   // [G][2]
    const int64_t n = blockIdx.y;
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) sg[i] = 0.f;
    __syncthreads();
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    if (r < sh.rows) {
        float s[8], ss[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) s[u] = ss[u] = 0.f;
        const int64_t p0 = (int64_t)blockIdx.x * sh.pix_per_cta;
        int64_t p1 = p0 + sh.pix_per_cta;
        if (p1 > sh.HW) p1 = sh.HW;
        const __nv_bfloat16 *base = x + n * sh.HW * ld + 8 * q;
        for (int64_t p = p0 + r; p < p1; p += sh.rows) {
            float f[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(base + p * ld)), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { s[u] += f[u]; ss[u] += f[u] * f[u]; }
        }
        // channels of one chunk fall into at most 8 groups; merge equal neighbours before the smem atomics
        int g_prev = (8 * q) / sh.cpg;
        float a = 0.f, b = 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int g = (8 * q + u) / sh.cpg;
            if (g != g_prev) { atomicAdd(&sg[2 * g_prev], a); atomicAdd(&sg[2 * g_prev + 1], b); a = b = 0.f; g_prev = g; }
            a += s[u]; b += ss[u];
        }
        atomicAdd(&sg[2 * g_prev], a); atomicAdd(&sg[2 * g_prev + 1], b);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) atomicAdd(stats + n * 2 * sh.G + i, sg[i]);
}

// per-thread affine y = act(x*A + B) coefficients for its 8 channels
struct Coef { float A[8], B[8]; };

__device__ __forceinline__ void mean_rstd(const float *stats, int64_t n, int G, int g, float inv_cnt, float eps,
                                          float &mean, float &rstd) {
    if (!stats) { mean = 0.f; rstd = 1.f; return; }      // no normalisation: plain activation (norm=False blocks)
    const float s = __ldg(stats + (n * G + g) * 2), ss = __ldg(stats + (n * G + g) * 2 + 1);
    mean = s * inv_cnt;
    const float var = fmaxf(ss * inv_cnt - mean * mean, 0.f);
    rstd = rsqrtf(var + eps);
}

__device__ __forceinline__ Coef make_coef(const Shape &sh, int64_t n, int q, const float *stats, const float *gamma,
                                          const float *beta, const float *scale, const float *shift, float eps,
                                          float pre) {
    Coef k;
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = 8 * q + u;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f, sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        k.A[u] = pre * (rstd * ga * sc);
        k.B[u] = pre * ((be - mean * rstd * ga) * sc + sf);
    }
    return k;
}

template <int ACT, bool DROP>
__global__ void __launch_bounds__(256) gn_act_fwd_kernel(const __nv_bfloat16 *__restrict__ x, int64_t ld_x, Shape sh,
                                                        const float *__restrict__ stats, const float *__restrict__ gamma,
                                                        const float *__restrict__ beta, const float *__restrict__ scale,
                                                        const float *__restrict__ shift, float eps, float p_drop,
                                                        uint64_t seed, uint64_t offset, const uint64_t *__restrict__ off_dev,
                                                        const __nv_bfloat16 *__restrict__ addend, int64_t ld_add,
                                                        __nv_bfloat16 *__restrict__ y,
                                                        int64_t ld_y) {
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    if (r >= sh.rows) return;
    const Coef k = make_coef(sh, n, q, stats, gamma, beta, scale, shift, eps, act_pre_scale<ACT>());
    if (DROP && off_dev) offset += __ldg(off_dev);   // device-resident counter: CUDA-graph replays draw fresh masks
    const int64_t p0 = (int64_t)blockIdx.x * sh.pix_per_cta;
    int64_t p1 = p0 + sh.pix_per_cta;
    if (p1 > sh.HW) p1 = sh.HW;
    for (int64_t p = p0 + r; p < p1; p += sh.rows) {
        const int64_t pix = n * sh.HW + p;
        float f[8];
        unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(x + pix * ld_x + 8 * q)), f);
        float m[8];
        if (DROP) dropout_mask8(seed, offset, pix * sh.C + 8 * q, p_drop, m);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            f[u] = act_fwd<ACT>(fmaf(f[u], k.A[u], k.B[u]));
            if (DROP) f[u] *= m[u];
        }
        if (addend) {
            float r8[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(addend + pix * ld_add + 8 * q)), r8);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += r8[u];
        }
        *reinterpret_cast<uint4 *>(y + pix * ld_y + 8 * q) = pack8(f);
    }
}

// backward pass 1: dz = gy * mask * act'(z) is computed ONCE, stored (bf16) in the gx buffer, and reduced to
// Q[n,c] = (sum_p dz, sum_p dz*x).  Pass 2 then needs no transcendental and no dropout mask.
template <int ACT, bool DROP>
__global__ void __launch_bounds__(256, 3) gn_act_bwd_reduce(const __nv_bfloat16 *__restrict__ gy, int64_t ld_gy,
                                                           const __nv_bfloat16 *__restrict__ x, int64_t ld_x, Shape sh,
                                                           const float *__restrict__ stats, const float *__restrict__ gamma,
                                                           const float *__restrict__ beta, const float *__restrict__ scale,
                                                           const float *__restrict__ shift, float eps, float p_drop,
                                                           uint64_t seed, uint64_t offset, const uint64_t *__restrict__ off_dev,
                                                           __nv_bfloat16 *__restrict__ dzbuf, int64_t ld_dz,
                                                           float *__restrict__ Q) {
    extern __shared__ float part[];   // [rows][chunks][16]: per-thread partial sums, combined without atomics
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    if (r < sh.rows) {
        const Coef k = make_coef(sh, n, q, stats, gamma, beta, scale, shift, eps, act_pre_scale<ACT>());
        if (DROP && off_dev) offset += __ldg(off_dev);   // device-resident counter: CUDA-graph replays draw fresh masks
        float q1[8], q2[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) q1[u] = q2[u] = 0.f;
        const int64_t p0 = (int64_t)blockIdx.x * sh.pix_per_cta;
        int64_t p1 = p0 + sh.pix_per_cta;
        if (p1 > sh.HW) p1 = sh.HW;
        const __nv_bfloat16 *xb = x + n * sh.HW * ld_x + 8 * q;
        const __nv_bfloat16 *gb = gy + n * sh.HW * ld_gy + 8 * q;
        __nv_bfloat16 *db = dzbuf + n * sh.HW * ld_dz + 8 * q;
        int64_t p = p0 + r;
        // two pixels per iteration: four independent 16-byte loads in flight per thread
        for (; p + sh.rows < p1; p += 2 * sh.rows) {
            const uint4 xa = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x));
            const uint4 ga = ld_stream_u4(reinterpret_cast<const uint4 *>(gb + p * ld_gy));
            const uint4 xc = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + (p + sh.rows) * ld_x));
            const uint4 gc = ld_stream_u4(reinterpret_cast<const uint4 *>(gb + (p + sh.rows) * ld_gy));
            float f[8], g[8], m[8];
            unpack8(xa, f); unpack8(ga, g);
            if (DROP) dropout_mask8(seed, offset, (n * sh.HW + p) * sh.C + 8 * q, p_drop, m);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float dz = g[u] * act_bwd<ACT>(fmaf(f[u], k.A[u], k.B[u]));
                if (DROP) dz *= m[u];
                q1[u] += dz; q2[u] = fmaf(dz, f[u], q2[u]);
                g[u] = dz;
            }
            *reinterpret_cast<uint4 *>(db + p * ld_dz) = pack8(g);
            unpack8(xc, f); unpack8(gc, g);
            if (DROP) dropout_mask8(seed, offset, (n * sh.HW + p + sh.rows) * sh.C + 8 * q, p_drop, m);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float dz = g[u] * act_bwd<ACT>(fmaf(f[u], k.A[u], k.B[u]));
                if (DROP) dz *= m[u];
                q1[u] += dz; q2[u] = fmaf(dz, f[u], q2[u]);
                g[u] = dz;
            }
            *reinterpret_cast<uint4 *>(db + (p + sh.rows) * ld_dz) = pack8(g);
        }
        for (; p < p1; p += sh.rows) {
            float f[8], g[8], m[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x)), f);
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(gb + p * ld_gy)), g);
            if (DROP) dropout_mask8(seed, offset, (n * sh.HW + p) * sh.C + 8 * q, p_drop, m);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float dz = g[u] * act_bwd<ACT>(fmaf(f[u], k.A[u], k.B[u]));
                if (DROP) dz *= m[u];
                q1[u] += dz; q2[u] = fmaf(dz, f[u], q2[u]);
                g[u] = dz;
            }
            *reinterpret_cast<uint4 *>(db + p * ld_dz) = pack8(g);
        }
        float4 *dst = reinterpret_cast<float4 *>(part + (size_t)threadIdx.x * 16);
        dst[0] = make_float4(q1[0], q1[1], q1[2], q1[3]); dst[1] = make_float4(q1[4], q1[5], q1[6], q1[7]);
        dst[2] = make_float4(q2[0], q2[1], q2[2], q2[3]); dst[3] = make_float4(q2[4], q2[5], q2[6], q2[7]);
    }
    __syncthreads();
    // channel c = 8q+u: sum over the `rows` threads that own chunk q, then one global atomic per (c, kind)
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) {
        const int c = i >> 1, kind = i & 1, cq = c >> 3, u = c & 7;
        float acc = 0.f;
        for (int rr = 0; rr < sh.rows; ++rr) acc += part[(size_t)(rr * sh.chunks + cq) * 16 + kind * 8 + u];
        atomicAdd(Q + (n * sh.C + c) * 2 + kind, acc);
    }
}

// backward pass 2.  With Q1 = sum dz, Q2x = sum dz*x per (n,c):  sum dz*xhat = rstd*(Q2x - mean*Q1), and
//   dx = rstd*(dz*gs - m1 - xhat*m2) = dz*P + x*Qc + R     (P, Qc, R per channel; gs = gamma*(1+scale))
// dz is read back from the gx buffer (written by pass 1) and overwritten in place: three fmas per element.
__global__ void __launch_bounds__(256, 4) gn_act_bwd_apply(const __nv_bfloat16 *__restrict__ x, int64_t ld_x, Shape sh,
                                                          const float *__restrict__ stats, const float *__restrict__ gamma,
                                                          const float *__restrict__ beta, const float *__restrict__ scale,
                                                          float eps, const float *__restrict__ Q,
                                                          __nv_bfloat16 *__restrict__ gx, int64_t ld_gx,
                                                          float *__restrict__ dgamma, float *__restrict__ dbeta,
                                                          float *__restrict__ dscale, float *__restrict__ dshift) {
    extern __shared__ float sg[];   // [G][2]: sum_c gs*Q1, sum_c gs*(sum dz*xhat)
    const int64_t n = blockIdx.y;
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) sg[i] = 0.f;
    __syncthreads();
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float Q1 = __ldg(Q + (n * sh.C + c) * 2);
        const float Q2 = rstd * (__ldg(Q + (n * sh.C + c) * 2 + 1) - mean * Q1);      // sum dz * xhat
        if (stats) {      // without normalisation the statistics do not depend on x: no mean-subtraction terms
            atomicAdd(&sg[2 * (c / sh.cpg)], ga * sc * Q1);
            atomicAdd(&sg[2 * (c / sh.cpg) + 1], ga * sc * Q2);
        }
        if (blockIdx.x == 0) {   // parameter gradients: once per sample
            if (dgamma) atomicAdd(dgamma + c, sc * Q2);
            if (dbeta) atomicAdd(dbeta + c, sc * Q1);
            if (dscale) dscale[n * sh.C + c] = ga * Q2 + be * Q1;
            if (dshift) dshift[n * sh.C + c] = Q1;
        }
    }
    __syncthreads();
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    if (r >= sh.rows) return;
    float P[8], Qc[8], R[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = 8 * q + u, g = c / sh.cpg;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, g, inv_cnt, eps, mean, rstd);
        const float m1 = sg[2 * g] * inv_cnt, m2 = sg[2 * g + 1] * inv_cnt;
        const float gs = (gamma ? __ldg(gamma + c) : 1.f) * (scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f);
        P[u] = rstd * gs;
        Qc[u] = -rstd * rstd * m2;
        R[u] = -rstd * m1 - mean * Qc[u];
    }
    const int64_t p0 = (int64_t)blockIdx.x * sh.pix_per_cta;
    int64_t p1 = p0 + sh.pix_per_cta;
    if (p1 > sh.HW) p1 = sh.HW;
    const __nv_bfloat16 *xb = x + n * sh.HW * ld_x + 8 * q;
    __nv_bfloat16 *gb = gx + n * sh.HW * ld_gx + 8 * q;
    int64_t p = p0 + r;
    for (; p + sh.rows < p1; p += 2 * sh.rows) {
        const uint4 xa = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x));
        const uint4 da = *reinterpret_cast<const uint4 *>(gb + p * ld_gx);
        const uint4 xc = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + (p + sh.rows) * ld_x));
        const uint4 dc = *reinterpret_cast<const uint4 *>(gb + (p + sh.rows) * ld_gx);
        float f[8], d[8];
        unpack8(xa, f); unpack8(da, d);
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u]));
        *reinterpret_cast<uint4 *>(gb + p * ld_gx) = pack8(d);
        unpack8(xc, f); unpack8(dc, d);
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u]));
        *reinterpret_cast<uint4 *>(gb + (p + sh.rows) * ld_gx) = pack8(d);
    }
    for (; p < p1; p += sh.rows) {
        float f[8], d[8];
        unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x)), f);
        unpack8(*reinterpret_cast<const uint4 *>(gb + p * ld_gx), d);
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u]));
        *reinterpret_cast<uint4 *>(gb + p * ld_gx) = pack8(d);
    }
}

// ---------------------------------------------------------------------------------------------
// Fused single-pass variants: one thread-block CLUSTER per sample keeps the sample's slab in (distributed) shared
// memory, so every tensor crosses HBM once.  forward: read x, write y (2 passes instead of 3) and emit the raw
// statistics for the backward; backward: read x and gy, write gx (3 passes instead of 6), one launch instead of
// memset + two kernels.  The slab arrives through bulk async copies (cp.async.bulk -> mbarrier): the whole per-CTA
// slab is in flight at once, which is what these latency-bound kernels were missing (one 16-byte load per thread
// per iteration reached ~3 TB/s).  Sums are combined without floating-point shared atomics (those are CAS loops):
// per-thread partials are folded through an 8 KB staging buffer, then across the cluster through DSMEM.
// Used when the slab fits one cluster; otherwise the multi-pass kernels above run.
// ---------------------------------------------------------------------------------------------
constexpr int kFusedThreads = 256;
constexpr int kFusedFixedBytes = 16 + 8192;      // mbarrier + fold staging (128 x 16 floats)
constexpr uint32_t kBulkPiece = 16384;           // dense slabs are fetched in pieces of this many bytes

struct FusedShape {
    int64_t HW; int C, G, cpg, chunks, rows, cs; int64_t pix_per_cta;
};

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

// warp 0 fetches `npx` pixels x C channels (bf16) into a dense smem slab; completion is counted on `bar`
__device__ __forceinline__ void fetch_slab(uint32_t dst, const __nv_bfloat16 *src, int64_t ld, int C, int npx, uint32_t bar) {
    const int lane = threadIdx.x;
    if (ld == C) {                                   // contiguous in HBM: few large pieces
        const uint32_t total = (uint32_t)npx * (uint32_t)C * 2u;
        for (uint32_t o = lane * kBulkPiece; o < total; o += 32 * kBulkPiece)
            bulk_g2s(dst + o, reinterpret_cast<const uint8_t *>(src) + o, min(kBulkPiece, total - o), bar);
    } else {                                         // a channel slice of a wider buffer: one piece per pixel
        const uint32_t row = (uint32_t)C * 2u;
        for (int p = lane; p < npx; p += 32) bulk_g2s(dst + p * row, src + (int64_t)p * ld, row, bar);
    }
}

// Sum 16 per-thread values over the threads that share chunk q (all r), for every q: out[2*c + kind], c = 8q+u,
// value index kind*8+u.  stage: float4[4][128].  Every thread of the CTA must call this (it synchronises).
__device__ __forceinline__ void cta_channel_sums(const float (&v)[16], float4 *stage, float *out, const FusedShape &sh,
                                                 int q, int r) {
    const int half = (sh.rows + 1) >> 1;
    float w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = v[i];
    if (r >= half && r < sh.rows) {
        const int idx = (r - half) * sh.chunks + q;
#pragma unroll
        for (int k = 0; k < 4; ++k) stage[k * 128 + idx] = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
    __syncthreads();
    if (r < half && r + half < sh.rows) {
        const int idx = r * sh.chunks + q;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 t = stage[k * 128 + idx];
            w[4 * k] += t.x; w[4 * k + 1] += t.y; w[4 * k + 2] += t.z; w[4 * k + 3] += t.w;
        }
    }
    __syncthreads();
    if (r < half) {
        const int idx = r * sh.chunks + q;
#pragma unroll
        for (int k = 0; k < 4; ++k) stage[k * 128 + idx] = make_float4(w[4 * k], w[4 * k + 1], w[4 * k + 2], w[4 * k + 3]);
    }
    __syncthreads();
    const float *sf = reinterpret_cast<const float *>(stage);
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) {     // i = (k*chunks + cq)*4 + j: conflict-free reads
        const int j = i & 3, cq = (i >> 2) % sh.chunks, k = (i >> 2) / sh.chunks;
        float acc = 0.f;
        for (int rr = 0; rr < half; ++rr) acc += sf[(k * 128 + rr * sh.chunks + cq) * 4 + j];
        const int val = 4 * k + j;
        out[2 * (8 * cq + (val & 7)) + (val >> 3)] = acc;
    }
    __syncthreads();
}

template <int ACT, bool DROP>
__global__ void __launch_bounds__(kFusedThreads) gn_fused_fwd_kernel(
    const __nv_bfloat16 *__restrict__ x, int64_t ld_x, FusedShape sh, float *__restrict__ stats_out, float eps,
    const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ scale,
    const float *__restrict__ shift, float p_drop, uint64_t seed, uint64_t offset, const uint64_t *__restrict__ off_dev,
    const __nv_bfloat16 *__restrict__ addend, int64_t ld_add, __nv_bfloat16 *__restrict__ y, int64_t ld_y) {
    extern __shared__ __align__(128) uint8_t fsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int64_t n = blockIdx.x / sh.cs;
    const int gpad = (2 * sh.G + 3) & ~3;
    float4 *stage = reinterpret_cast<float4 *>(fsm + 16);
    float *chan = reinterpret_cast<float *>(fsm + kFusedFixedBytes);       // [C][2] this CTA
    float *cta_gs = chan + 2 * sh.C;                                       // [G][2] this CTA
    float *tot_gs = cta_gs + gpad;                                         // [G][2] whole sample
    uint4 *xs = reinterpret_cast<uint4 *>(tot_gs + gpad);                  // [pix][chunks]
    const uint32_t bar = tc::smem_u32(fsm);
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    const int64_t p0 = (int64_t)rank * sh.pix_per_cta;
    const int npx = (int)(min(p0 + sh.pix_per_cta, sh.HW) - p0);
    if (threadIdx.x == 0) {
        tc::mbar_init(reinterpret_cast<uint64_t *>(fsm), 1);
        tc::fence_barrier_init();
        tc::mbar_arrive_expect_tx_a(bar, (uint32_t)npx * (uint32_t)sh.C * 2u);
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();
    if (threadIdx.x < 32) fetch_slab(tc::smem_u32(xs), x + (n * sh.HW + p0) * ld_x, ld_x, sh.C, npx, bar);
    const bool active = r < sh.rows;
    if (DROP && off_dev) offset += __ldg(off_dev);
    tc::mbar_wait_a(bar, 0);
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = 0.f;
    const int step = sh.rows * sh.chunks;                                   // slab entries between a thread's pixels
    if (active) {
        const uint4 *xp = xs + r * sh.chunks + q;
        for (int p = r; p < npx; p += sh.rows, xp += step) {
            float f[8];
            unpack8(*xp, f);
#pragma unroll
            for (int u = 0; u < 8; ++u) { v[u] += f[u]; v[8 + u] = fmaf(f[u], f[u], v[8 + u]); }
        }
    }
    cta_channel_sums(v, stage, chan, sh, q, r);
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) {
        const int g = i >> 1, kind = i & 1;
        float acc = 0.f;
        for (int c = g * sh.cpg; c < (g + 1) * sh.cpg; ++c) acc += chan[2 * c + kind];
        cta_gs[i] = acc;
    }
    cluster.sync();
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) {
        float acc = 0.f;
        for (int rk = 0; rk < sh.cs; ++rk) acc += cluster.map_shared_rank(cta_gs, rk)[i];
        tot_gs[i] = acc;
        if (rank == 0) stats_out[n * 2 * sh.G + i] = acc;                   // raw sums: what the backward consumes
    }
    cluster.sync();                                                         // peers are done reading our cta_gs
    // y = act(x * A[c] + B[c]): the per-channel coefficients are computed ONCE per CTA (one channel per thread, one
    // integer division each) into the `chan` table instead of eight channels redundantly in every thread
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const int g = c / sh.cpg;
        const float mean = tot_gs[2 * g] * inv_cnt;
        const float var = fmaxf(tot_gs[2 * g + 1] * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + eps);
        const float g0 = gamma ? __ldg(gamma + c) : 1.f, b0 = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f;
        const float sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        const float a = rstd * g0 * sc;
        chan[c] = act_pre_scale<ACT>() * a;
        chan[sh.C + c] = act_pre_scale<ACT>() * fmaf(-mean, a, fmaf(b0, sc, sf));
    }
    __syncthreads();
    if (!active) return;
    float A[8], B[8];
    {
        const float4 a0 = *reinterpret_cast<const float4 *>(chan + 8 * q), a1 = *reinterpret_cast<const float4 *>(chan + 8 * q + 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(chan + sh.C + 8 * q), b1 = *reinterpret_cast<const float4 *>(chan + sh.C + 8 * q + 4);
        A[0] = a0.x; A[1] = a0.y; A[2] = a0.z; A[3] = a0.w; A[4] = a1.x; A[5] = a1.y; A[6] = a1.z; A[7] = a1.w;
        B[0] = b0.x; B[1] = b0.y; B[2] = b0.z; B[3] = b0.w; B[4] = b1.x; B[5] = b1.y; B[6] = b1.z; B[7] = b1.w;
    }
    __nv_bfloat16 *yp = y + (n * sh.HW + p0 + r) * ld_y + 8 * q;
    const __nv_bfloat16 *ap = addend ? addend + (n * sh.HW + p0 + r) * ld_add + 8 * q : nullptr;
    const int64_t ystep = (int64_t)sh.rows * ld_y, astep = (int64_t)sh.rows * ld_add;
    const uint4 *xp = xs + r * sh.chunks + q;
    int64_t elem = (n * sh.HW + p0 + r) * sh.C + 8 * q;                     // dropout counter index
    const int64_t estep = (int64_t)sh.rows * sh.C;
#pragma unroll 2
    for (int p = r; p < npx; p += sh.rows, xp += step, yp += ystep, elem += estep) {
        float f[8];
        unpack8(*xp, f);
        float m[8];
        if (DROP) dropout_mask8(seed, offset, elem, p_drop, m);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            f[u] = act_fwd<ACT>(fmaf(f[u], A[u], B[u]));
            if (DROP) f[u] *= m[u];
        }
        if (ap) {
            float r8[8];
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(ap)), r8);
            ap += astep;
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] += r8[u];
        }
        *reinterpret_cast<uint4 *>(yp) = pack8(f);
    }
}

// XSLAB: x is kept in shared memory next to dz (small slabs: shortest latency chain).  Otherwise only gy -> dz lives in
// shared memory and x is streamed from HBM in the sums pass and read again (L2-resident: the same CTA touched it
// microseconds earlier) in the apply pass, which halves the footprint: large layers keep 3 CTAs per SM.
template <int ACT, bool DROP, bool XSLAB>
__global__ void __launch_bounds__(kFusedThreads, 3) gn_fused_bwd_kernel(
    const __nv_bfloat16 *__restrict__ gy, int64_t ld_gy, const __nv_bfloat16 *__restrict__ x, int64_t ld_x, FusedShape sh,
    const float *__restrict__ stats, float eps, const float *__restrict__ gamma, const float *__restrict__ beta,
    const float *__restrict__ scale, const float *__restrict__ shift, float p_drop, uint64_t seed, uint64_t offset,
    const uint64_t *__restrict__ off_dev, __nv_bfloat16 *__restrict__ gx, int64_t ld_gx, float *__restrict__ dgamma,
    float *__restrict__ dbeta, float *__restrict__ dscale, float *__restrict__ dshift,
    const __nv_bfloat16 *__restrict__ gadd, int64_t ld_gadd) {
    extern __shared__ __align__(128) uint8_t fsm[];
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int64_t n = blockIdx.x / sh.cs;
    float4 *stage = reinterpret_cast<float4 *>(fsm + 16);
    float *cta_q = reinterpret_cast<float *>(fsm + kFusedFixedBytes);      // [C][2] this CTA: sum dz, sum dz*x
    float *tot_q = cta_q + 2 * sh.C;                                       // [C][2] whole sample
    float *sg = tot_q + 2 * sh.C;                                          // [G][2]
    uint4 *slab = reinterpret_cast<uint4 *>(sg + ((2 * sh.G + 3) & ~3));
    uint4 *ds = slab;                                                      // [pix][chunks] gy, then dz in place
    uint4 *xs = slab + (size_t)sh.pix_per_cta * sh.chunks;                 // [pix][chunks] x (XSLAB only)
    const uint32_t bar = tc::smem_u32(fsm);
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    const bool active = r < sh.rows;
    const int64_t p0 = (int64_t)rank * sh.pix_per_cta;
    const int npx = (int)(min(p0 + sh.pix_per_cta, sh.HW) - p0);
    if (threadIdx.x == 0) {
        tc::mbar_init(reinterpret_cast<uint64_t *>(fsm), 1);
        tc::fence_barrier_init();
        tc::mbar_arrive_expect_tx_a(bar, (XSLAB ? 2u : 1u) * (uint32_t)npx * (uint32_t)sh.C * 2u);
    }
    pdl_trigger();
    __syncthreads();
    pdl_wait();
    if (threadIdx.x < 32) {
        fetch_slab(tc::smem_u32(ds), gy + (n * sh.HW + p0) * ld_gy, ld_gy, sh.C, npx, bar);
        if (XSLAB) fetch_slab(tc::smem_u32(xs), x + (n * sh.HW + p0) * ld_x, ld_x, sh.C, npx, bar);
    }
    const __nv_bfloat16 *xb = x + (n * sh.HW + p0) * ld_x + 8 * q;
    const int step = sh.rows;
    const int sstep = sh.rows * sh.chunks;                                 // slab entries between a thread's pixels
    const int64_t xstep = (int64_t)step * ld_x;
    uint4 xa = make_uint4(0, 0, 0, 0), xc = xa;
    if (!XSLAB && active) {                                                // first two pixels travel with the bulk copy
        if (r < npx) xa = *reinterpret_cast<const uint4 *>(xb + (int64_t)r * ld_x);
        if (r + step < npx) xc = *reinterpret_cast<const uint4 *>(xb + (int64_t)(r + step) * ld_x);
    }
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    // z = x * A[c] + B[c]: per-channel coefficients computed once per CTA (one channel per thread) into the tot_q area,
    // which is not needed before the cluster reduction
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f, sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        const float a = rstd * ga * sc;
        tot_q[c] = act_pre_scale<ACT>() * a;
        tot_q[sh.C + c] = act_pre_scale<ACT>() * ((be - mean * rstd * ga) * sc + sf);
    }
    __syncthreads();
    Coef k;
    if (active) {
        const float4 a0 = *reinterpret_cast<const float4 *>(tot_q + 8 * q), a1 = *reinterpret_cast<const float4 *>(tot_q + 8 * q + 4);
        const float4 b0 = *reinterpret_cast<const float4 *>(tot_q + sh.C + 8 * q), b1 = *reinterpret_cast<const float4 *>(tot_q + sh.C + 8 * q + 4);
        k.A[0] = a0.x; k.A[1] = a0.y; k.A[2] = a0.z; k.A[3] = a0.w; k.A[4] = a1.x; k.A[5] = a1.y; k.A[6] = a1.z; k.A[7] = a1.w;
        k.B[0] = b0.x; k.B[1] = b0.y; k.B[2] = b0.z; k.B[3] = b0.w; k.B[4] = b1.x; k.B[5] = b1.y; k.B[6] = b1.z; k.B[7] = b1.w;
    }
    if (DROP && off_dev) offset += __ldg(off_dev);
    tc::mbar_wait_a(bar, 0);
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = 0.f;
    const int64_t ebase = (n * sh.HW + p0) * sh.C + 8 * q;                  // dropout counter index of pixel 0
    auto one = [&](int p, uint4 *dp, const uint4 &xv) {
        float f[8], g[8], m[8];
        unpack8(xv, f);
        unpack8(*dp, g);
        if (DROP) dropout_mask8(seed, offset, ebase + (int64_t)p * sh.C, p_drop, m);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float dz = g[u] * act_bwd<ACT>(fmaf(f[u], k.A[u], k.B[u]));
            if (DROP) dz *= m[u];
            v[u] += dz; v[8 + u] = fmaf(dz, f[u], v[8 + u]);
            g[u] = dz;
        }
        *dp = pack8(g);
    };
    if (active) {
        uint4 *dp = ds + r * sh.chunks + q;
        if (XSLAB) {
            const uint4 *xp = xs + r * sh.chunks + q;
#pragma unroll 2
            for (int p = r; p < npx; p += step, dp += sstep, xp += sstep) one(p, dp, *xp);
        } else {
            const __nv_bfloat16 *xn = xb + (int64_t)(r + 2 * step) * ld_x;  // next pair to prefetch
            for (int p = r; p < npx; p += 2 * step, dp += 2 * sstep, xn += 2 * xstep) {   // two pixels per turn, the next two in flight
                uint4 na = make_uint4(0, 0, 0, 0), nc = na;
                if (p + 2 * step < npx) na = *reinterpret_cast<const uint4 *>(xn);
                if (p + 3 * step < npx) nc = *reinterpret_cast<const uint4 *>(xn + xstep);
                one(p, dp, xa);
                if (p + step < npx) one(p + step, dp + sstep, xc);
                xa = na; xc = nc;
            }
        }
    }
    cta_channel_sums(v, stage, cta_q, sh, q, r);
    cluster.sync();
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) {
        float acc = 0.f;
        for (int rk = 0; rk < sh.cs; ++rk) acc += cluster.map_shared_rank(cta_q, rk)[i];
        tot_q[i] = acc;
    }
    cluster.sync();                                                         // peers are done reading our cta_q: reuse it
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float Q1 = tot_q[2 * c];
        const float Q2 = rstd * (tot_q[2 * c + 1] - mean * Q1);                  // sum dz * xhat
        cta_q[2 * c] = ga * sc * Q1;
        cta_q[2 * c + 1] = ga * sc * Q2;
        if (rank == 0) {                                                          // parameter gradients: once per sample
            if (dgamma) atomicAdd(dgamma + c, sc * Q2);
            if (dbeta) atomicAdd(dbeta + c, sc * Q1);
            if (dscale) dscale[n * sh.C + c] = ga * Q2 + be * Q1;
            if (dshift) dshift[n * sh.C + c] = Q1;
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) {
        const int g = i >> 1, kind = i & 1;
        float acc = 0.f;
        // without normalisation the statistics do not depend on x: no mean-subtraction terms
        if (stats) for (int c = g * sh.cpg; c < (g + 1) * sh.cpg; ++c) acc += cta_q[2 * c + kind];
        sg[i] = acc;
    }
    __syncthreads();
    // dx = dz * P[c] + x * Qc[c] + R[c]: tables again, one channel per thread (tot_q and cta_q are free now)
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const int g = c / sh.cpg;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, g, inv_cnt, eps, mean, rstd);
        const float m1 = sg[2 * g] * inv_cnt, m2 = sg[2 * g + 1] * inv_cnt;
        const float gs = (gamma ? __ldg(gamma + c) : 1.f) * (scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f);
        const float qc = -rstd * rstd * m2;
        tot_q[c] = rstd * gs;
        tot_q[sh.C + c] = qc;
        cta_q[c] = -rstd * m1 - mean * qc;
    }
    __syncthreads();
    if (!active) return;
    float P[8], Qc[8], R[8];
    {
        const float4 *tp = reinterpret_cast<const float4 *>(tot_q + 8 * q), *tq = reinterpret_cast<const float4 *>(tot_q + sh.C + 8 * q);
        const float4 *tr = reinterpret_cast<const float4 *>(cta_q + 8 * q);
        const float4 p0v = tp[0], p1v = tp[1], q0v = tq[0], q1v = tq[1], r0v = tr[0], r1v = tr[1];
        P[0] = p0v.x; P[1] = p0v.y; P[2] = p0v.z; P[3] = p0v.w; P[4] = p1v.x; P[5] = p1v.y; P[6] = p1v.z; P[7] = p1v.w;
        Qc[0] = q0v.x; Qc[1] = q0v.y; Qc[2] = q0v.z; Qc[3] = q0v.w; Qc[4] = q1v.x; Qc[5] = q1v.y; Qc[6] = q1v.z; Qc[7] = q1v.w;
        R[0] = r0v.x; R[1] = r0v.y; R[2] = r0v.z; R[3] = r0v.w; R[4] = r1v.x; R[5] = r1v.y; R[6] = r1v.z; R[7] = r1v.w;
    }
    // gadd: a second gradient of x (the ResBlock's shortcut / residual branch), summed here instead of by a separate add
    const bool has_add = gadd != nullptr;
    auto fin = [&](const uint4 *dp, __nv_bfloat16 *op, const uint4 &xv, const uint4 &av) {
        float f[8], d[8];
        unpack8(xv, f);
        unpack8(*dp, d);
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u]));
        if (has_add) {
            float a[8];
            unpack8(av, a);
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] += a[u];
        }
        *reinterpret_cast<uint4 *>(op) = pack8(d);
    };
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    const uint4 *dp = ds + r * sh.chunks + q;
    __nv_bfloat16 *op = gx + (n * sh.HW + p0 + r) * ld_gx + 8 * q;
    const __nv_bfloat16 *ap = has_add ? gadd + (n * sh.HW + p0 + r) * ld_gadd + 8 * q : nullptr;
    const int64_t ostep = (int64_t)step * ld_gx, astep = (int64_t)step * ld_gadd;
    if (XSLAB) {
        const uint4 *xp = xs + r * sh.chunks + q;
        for (int p = r; p < npx; p += 2 * step, dp += 2 * sstep, xp += 2 * sstep, op += 2 * ostep) {
            const bool two = p + step < npx;
            uint4 a0 = zero4, a1 = zero4;
            if (has_add) {
                a0 = ld_stream_u4(reinterpret_cast<const uint4 *>(ap));
                if (two) a1 = ld_stream_u4(reinterpret_cast<const uint4 *>(ap + astep));
                ap += 2 * astep;
            }
            fin(dp, op, *xp, a0);
            if (two) fin(dp + sstep, op + ostep, xp[sstep], a1);
        }
    } else {
        const __nv_bfloat16 *xp = xb + (int64_t)r * ld_x;
        for (int p = r; p < npx; p += 4 * step, dp += 4 * sstep, xp += 4 * xstep, op += 4 * ostep) {   // four L2 reads in flight
            uint4 t[4], a[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (p + j * step < npx) {
                    t[j] = *reinterpret_cast<const uint4 *>(xp + j * xstep);
                    a[j] = has_add ? ld_stream_u4(reinterpret_cast<const uint4 *>(ap + j * astep)) : zero4;
                }
            if (has_add) ap += 4 * astep;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (p + j * step < npx) fin(dp + j * sstep, op + j * ostep, t[j], a[j]);
        }
    }
}

// gx += gadd (the multi-pass fallback of the fused backward's second-gradient input)
__global__ void __launch_bounds__(256) add_rows_kernel(__nv_bfloat16 *__restrict__ gx, int64_t ld_gx,
                                                      const __nv_bfloat16 *__restrict__ gadd, int64_t ld_gadd, int64_t pixels,
                                                      int chunks) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pixels * chunks; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t p = i / chunks;
        const int q = (int)(i - p * chunks);
        float a[8], b[8];
        uint4 *dst = reinterpret_cast<uint4 *>(gx + p * ld_gx + 8 * q);
        unpack8(*dst, a);
        unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(gadd + p * ld_gadd + 8 * q)), b);
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] += b[u];
        *dst = pack8(a);
    }
}

// ---------------------------------------------------------------------------------------------
// Streaming two-launch variants.  The cluster kernels above read every tensor once, but each CTA walks a chain of ~8
// dependent phases (bulk copy -> sums -> fold -> cluster barrier -> DSMEM -> table -> apply) on a 32-64 KB slab: ~10 us of
// latency per CTA, which caps them at 2.3-3.2 TB/s forward and 1.5-1.8 TB/s backward on the 32x32 layers (cold cache).
// Here the chain is cut at a kernel boundary instead: a sums kernel and an apply kernel, both pure streams with four
// 16-byte loads in flight per thread and no barrier inside the loop.  The second kernel's reads are L2 hits (the first
// kernel touched the same lines microseconds earlier; a layer's tensors are 17-100 MB against 126 MB of L2), so HBM
// still sees each tensor once.  Sums travel between the two launches as per-CTA partials [N][splits][...] that every
// apply CTA reduces itself in a fixed order: no atomics, no memset, bit-reproducible.
//   forward : gn_stream_stats_kernel (partials [N][splits][2G]) -> gn_stream_fwd_kernel (writes stats [N,G,2] too)
//   backward: gn_stream_bwd_reduce   (dz parked in gx, partials [N][splits][2C]) -> gn_stream_bwd_apply
// The forward apply kernel with splits = 1 and part = stats is the "statistics already known" entry
// (ub200_gn_act_fwd_nhwc_bf16; a conv epilogue that accumulates ub200_conv_args.gn_partial feeds it directly).
// ---------------------------------------------------------------------------------------------
constexpr int kStreamThreads = 256;
constexpr int kStreamMaxSplits = 64;

struct StreamShape {
    int64_t HW; int C, G, cpg, chunks, rows, splits, iters;   // one CTA: `iters` turns of `rows` pixels
};

// per-thread 16 values (8 channels x 2 kinds) -> chan[2*c + kind] summed over the threads that share a chunk
__device__ __forceinline__ void stream_channel_sums(const float (&v)[16], float *sp, float *chan, const StreamShape &sh, bool active) {
    if (active) {
        float4 *dst = reinterpret_cast<float4 *>(sp + (size_t)threadIdx.x * 16);
        dst[0] = make_float4(v[0], v[1], v[2], v[3]); dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        dst[2] = make_float4(v[8], v[9], v[10], v[11]); dst[3] = make_float4(v[12], v[13], v[14], v[15]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) {
        const int c = i >> 1, kind = i & 1, cq = c >> 3, u = c & 7;
        float acc = 0.f;
        for (int rr = 0; rr < sh.rows; ++rr) acc += sp[(size_t)(rr * sh.chunks + cq) * 16 + kind * 8 + u];
        chan[i] = acc;
    }
    __syncthreads();
}

// out[2*g + kind] = sum over the channels of group g of chan[2*c + kind]; a warp per item for wide groups
__device__ __forceinline__ void stream_group_sums(const float *chan, const StreamShape &sh, float *out_smem, float *out_gmem) {
    if (sh.cpg >= 64) {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int i = warp; i < 2 * sh.G; i += blockDim.x >> 5) {
            const int g = i >> 1, kind = i & 1;
            float acc = 0.f;
            for (int k = lane; k < sh.cpg; k += 32) acc += chan[2 * (g * sh.cpg + k) + kind];
            acc = warp_sum(acc);
            if (lane == 0) { if (out_smem) out_smem[i] = acc; if (out_gmem) out_gmem[i] = acc; }
        }
    } else {
        for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) {
            const int g = i >> 1, kind = i & 1;
            float acc = 0.f;
            for (int k = 0; k < sh.cpg; ++k) acc += chan[2 * (g * sh.cpg + k) + kind];
            if (out_smem) out_smem[i] = acc;
            if (out_gmem) out_gmem[i] = acc;
        }
    }
}

__device__ __forceinline__ void load_tab8(const float *tab, int q, float (&o)[8]) {
    const float4 a = *reinterpret_cast<const float4 *>(tab + 8 * q), b = *reinterpret_cast<const float4 *>(tab + 8 * q + 4);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

__global__ void __launch_bounds__(kStreamThreads, 4) gn_stream_stats_kernel(const __nv_bfloat16 *__restrict__ x, int64_t ld,
                                                                            StreamShape sh, float *__restrict__ part) {
    extern __shared__ __align__(16) float ssm[];      // sp [256][16] | chan [2C]
    float *sp = ssm, *chan = ssm + kStreamThreads * 16;
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    const bool active = r < sh.rows;
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = 0.f;
    pdl_trigger();
    pdl_wait();
    if (active) {
        const int64_t p0 = (int64_t)blockIdx.x * sh.iters * sh.rows;
        const int64_t p1 = min(p0 + (int64_t)sh.iters * sh.rows, sh.HW);
        const __nv_bfloat16 *base = x + n * sh.HW * ld + 8 * q;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        for (int64_t p = p0 + r; p < p1; p += 4 * sh.rows) {
            uint4 t[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int64_t pj = p + (int64_t)j * sh.rows;
                t[j] = pj < p1 ? ld_stream_u4(reinterpret_cast<const uint4 *>(base + pj * ld)) : zero4;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float f[8];
                unpack8(t[j], f);
#pragma unroll
                for (int u = 0; u < 8; ++u) { v[u] += f[u]; v[8 + u] = fmaf(f[u], f[u], v[8 + u]); }
            }
        }
    }
    stream_channel_sums(v, sp, chan, sh, active);
    stream_group_sums(chan, sh, nullptr, part + (n * sh.splits + blockIdx.x) * 2 * sh.G);
}

// y = addend + dropout(act(x * A[n,c] + B[n,c])) with the statistics given as `nparts` partial sums per (sample, group)
template <int ACT, bool DROP>
__global__ void __launch_bounds__(kStreamThreads, DROP ? 3 : 4) gn_stream_fwd_kernel(
    const __nv_bfloat16 *__restrict__ x, int64_t ld_x, StreamShape sh, const float *__restrict__ part, int nparts,
    float *__restrict__ stats_out, float eps, const float *__restrict__ gamma, const float *__restrict__ beta,
    const float *__restrict__ scale, const float *__restrict__ shift, float p_drop, uint64_t seed, uint64_t offset,
    const uint64_t *__restrict__ off_dev, const __nv_bfloat16 *__restrict__ addend, int64_t ld_add,
    __nv_bfloat16 *__restrict__ y, int64_t ld_y) {
    extern __shared__ __align__(16) float ssm[];      // gs [2G padded] | tab [2C]
    float *gs = ssm, *tab = ssm + ((2 * sh.G + 3) & ~3);
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    pdl_trigger();
    pdl_wait();
    if (part) {
        for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) {
            const float *src = part + n * nparts * 2 * sh.G + i;
            float acc = 0.f;
            for (int s = 0; s < nparts; ++s) acc += __ldg(src + (int64_t)s * 2 * sh.G);
            gs[i] = acc;
            if (stats_out && blockIdx.x == 0) stats_out[n * 2 * sh.G + i] = acc;
        }
    }
    __syncthreads();
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        float mean = 0.f, rstd = 1.f;
        if (part) {
            const int g = c / sh.cpg;
            mean = gs[2 * g] * inv_cnt;
            rstd = rsqrtf(fmaxf(gs[2 * g + 1] * inv_cnt - mean * mean, 0.f) + eps);
        }
        const float g0 = gamma ? __ldg(gamma + c) : 1.f, b0 = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f, sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        const float a = rstd * g0 * sc;
        tab[c] = act_pre_scale<ACT>() * a;
        tab[sh.C + c] = act_pre_scale<ACT>() * fmaf(-mean, a, fmaf(b0, sc, sf));
    }
    __syncthreads();
    if (r >= sh.rows) return;
    float A[8], B[8];
    load_tab8(tab, q, A); load_tab8(tab + sh.C, q, B);
    if (DROP && off_dev) offset += __ldg(off_dev);   // device-resident counter: CUDA-graph replays draw fresh masks
    const int64_t p0 = (int64_t)blockIdx.x * sh.iters * sh.rows;
    const int64_t p1 = min(p0 + (int64_t)sh.iters * sh.rows, sh.HW);
    const __nv_bfloat16 *xb = x + n * sh.HW * ld_x + 8 * q;
    const __nv_bfloat16 *ab = addend ? addend + n * sh.HW * ld_add + 8 * q : nullptr;
    __nv_bfloat16 *yb = y + n * sh.HW * ld_y + 8 * q;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (int64_t p = p0 + r; p < p1; p += 4 * sh.rows) {
        uint4 t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t pj = p + (int64_t)j * sh.rows;
            t[j] = pj < p1 ? ld_stream_u4(reinterpret_cast<const uint4 *>(xb + pj * ld_x)) : zero4;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t pj = p + (int64_t)j * sh.rows;
            if (pj >= p1) break;
            float f[8], m[8];
            unpack8(t[j], f);
            if (DROP) dropout_mask8(seed, offset, (n * sh.HW + pj) * sh.C + 8 * q, p_drop, m);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                f[u] = act_fwd<ACT>(fmaf(f[u], A[u], B[u]));
                if (DROP) f[u] *= m[u];
            }
            if (ab) {
                float r8[8];
                unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(ab + pj * ld_add)), r8);   // post-norm blocks only: not prefetched
#pragma unroll
                for (int u = 0; u < 8; ++u) f[u] += r8[u];
            }
            *reinterpret_cast<uint4 *>(yb + pj * ld_y) = pack8(f);
        }
    }
}

// backward launch 1: dz = gy * mask * act'(z) computed once, parked (bf16) in the gx buffer, and reduced to per-CTA
// partials part[n][split][c] = (sum dz, sum dz*x)
template <int ACT, bool DROP>
__global__ void __launch_bounds__(kStreamThreads, DROP ? 2 : 3) gn_stream_bwd_reduce(
    const __nv_bfloat16 *__restrict__ gy, int64_t ld_gy, const __nv_bfloat16 *__restrict__ x, int64_t ld_x, StreamShape sh,
    const float *__restrict__ stats, float eps, const float *__restrict__ gamma, const float *__restrict__ beta,
    const float *__restrict__ scale, const float *__restrict__ shift, float p_drop, uint64_t seed, uint64_t offset,
    const uint64_t *__restrict__ off_dev, __nv_bfloat16 *__restrict__ dzbuf, int64_t ld_dz, float *__restrict__ part) {
    extern __shared__ __align__(16) float ssm[];      // sp [256][16] | chan [2C] (first the A/B table, then the sums)
    float *sp = ssm, *chan = ssm + kStreamThreads * 16;
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    const bool active = r < sh.rows;
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    pdl_trigger();
    pdl_wait();
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f, sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        const float a = rstd * ga * sc;
        chan[c] = act_pre_scale<ACT>() * a;
        chan[sh.C + c] = act_pre_scale<ACT>() * ((be - mean * rstd * ga) * sc + sf);
    }
    __syncthreads();
    float v[16];
#pragma unroll
    for (int u = 0; u < 16; ++u) v[u] = 0.f;
    if (active) {
        float A[8], B[8];
        load_tab8(chan, q, A); load_tab8(chan + sh.C, q, B);
        if (DROP && off_dev) offset += __ldg(off_dev);
        const int64_t p0 = (int64_t)blockIdx.x * sh.iters * sh.rows;
        const int64_t p1 = min(p0 + (int64_t)sh.iters * sh.rows, sh.HW);
        const __nv_bfloat16 *xb = x + n * sh.HW * ld_x + 8 * q;
        const __nv_bfloat16 *gb = gy + n * sh.HW * ld_gy + 8 * q;
        __nv_bfloat16 *db = dzbuf + n * sh.HW * ld_dz + 8 * q;
        const uint4 zero4 = make_uint4(0, 0, 0, 0);
        for (int64_t p = p0 + r; p < p1; p += 2 * sh.rows) {          // two pixels per turn: four 16-byte loads in flight
            const int64_t pb = p + sh.rows;
            const bool two = pb < p1;
            const uint4 xa = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x));
            const uint4 ga = ld_stream_u4(reinterpret_cast<const uint4 *>(gb + p * ld_gy));
            const uint4 xc = two ? ld_stream_u4(reinterpret_cast<const uint4 *>(xb + pb * ld_x)) : zero4;
            const uint4 gc = two ? ld_stream_u4(reinterpret_cast<const uint4 *>(gb + pb * ld_gy)) : zero4;
            float f[8], g[8], m[8];
            unpack8(xa, f); unpack8(ga, g);
            if (DROP) dropout_mask8(seed, offset, (n * sh.HW + p) * sh.C + 8 * q, p_drop, m);
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                float dz = g[u] * act_bwd<ACT>(fmaf(f[u], A[u], B[u]));
                if (DROP) dz *= m[u];
                v[u] += dz; v[8 + u] = fmaf(dz, f[u], v[8 + u]);
                g[u] = dz;
            }
            *reinterpret_cast<uint4 *>(db + p * ld_dz) = pack8(g);
            if (two) {
                unpack8(xc, f); unpack8(gc, g);
                if (DROP) dropout_mask8(seed, offset, (n * sh.HW + pb) * sh.C + 8 * q, p_drop, m);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    float dz = g[u] * act_bwd<ACT>(fmaf(f[u], A[u], B[u]));
                    if (DROP) dz *= m[u];
                    v[u] += dz; v[8 + u] = fmaf(dz, f[u], v[8 + u]);
                    g[u] = dz;
                }
                *reinterpret_cast<uint4 *>(db + pb * ld_dz) = pack8(g);
            }
        }
    }
    __syncthreads();                                                   // every thread is done with the A/B table in chan
    stream_channel_sums(v, sp, chan, sh, active);
    float *dst = part + (n * sh.splits + blockIdx.x) * 2 * sh.C;
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) dst[i] = chan[i];
}

// backward launch 2.  With Q1 = sum dz, Q2x = sum dz*x per (n,c):  sum dz*xhat = rstd*(Q2x - mean*Q1), and
//   dx = rstd*(dz*gs - m1 - xhat*m2) = dz*P + x*Qc + R     (P, Qc, R per channel; gs = gamma*(1+scale))
// dz is read back from the gx buffer and overwritten in place: three fmas per element.
__global__ void __launch_bounds__(kStreamThreads, 4) gn_stream_bwd_apply(
    const __nv_bfloat16 *__restrict__ x, int64_t ld_x, StreamShape sh, const float *__restrict__ stats,
    const float *__restrict__ gamma, const float *__restrict__ beta, const float *__restrict__ scale, float eps,
    const float *__restrict__ part, __nv_bfloat16 *__restrict__ gx, int64_t ld_gx, float *__restrict__ dgamma,
    float *__restrict__ dbeta, float *__restrict__ dscale, float *__restrict__ dshift,
    const __nv_bfloat16 *__restrict__ gadd, int64_t ld_gadd) {
    extern __shared__ __align__(16) float ssm[];      // chan [2C] | tab [3C] | sg [2G]
    float *chan = ssm, *tab = ssm + 2 * sh.C, *sg = tab + 3 * sh.C;
    const int64_t n = blockIdx.y;
    const int q = threadIdx.x % sh.chunks, r = threadIdx.x / sh.chunks;
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
    pdl_trigger();
    pdl_wait();
    // Q[c] = sum of the per-CTA partials (fixed order), folded with gamma*(1+scale) for the group sums
    for (int i = threadIdx.x; i < 2 * sh.C; i += blockDim.x) {
        const float *src = part + n * sh.splits * 2 * sh.C + i;
        float acc = 0.f;
        for (int s = 0; s < sh.splits; ++s) acc += __ldg(src + (int64_t)s * 2 * sh.C);
        tab[i] = acc;                                                  // raw (Q1, Q2x) parked in tab[0 .. 2C)
    }
    __syncthreads();
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float Q1 = tab[2 * c];
        const float Q2 = rstd * (tab[2 * c + 1] - mean * Q1);          // sum dz * xhat
        chan[2 * c] = ga * sc * Q1;
        chan[2 * c + 1] = ga * sc * Q2;
        if (blockIdx.x == 0) {                                         // parameter gradients: once per sample
            if (dgamma) atomicAdd(dgamma + c, sc * Q2);
            if (dbeta) atomicAdd(dbeta + c, sc * Q1);
            if (dscale) dscale[n * sh.C + c] = ga * Q2 + be * Q1;
            if (dshift) dshift[n * sh.C + c] = Q1;
        }
    }
    __syncthreads();
    if (stats) stream_group_sums(chan, sh, sg, nullptr);
    else for (int i = threadIdx.x; i < 2 * sh.G; i += blockDim.x) sg[i] = 0.f;   // no normalisation: no mean-subtraction terms
    __syncthreads();
    for (int c = threadIdx.x; c < sh.C; c += blockDim.x) {
        const int g = c / sh.cpg;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, g, inv_cnt, eps, mean, rstd);
        const float m1 = sg[2 * g] * inv_cnt, m2 = sg[2 * g + 1] * inv_cnt;
        const float gs = (gamma ? __ldg(gamma + c) : 1.f) * (scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f);
        const float qc = -rstd * rstd * m2;
        tab[c] = rstd * gs;
        tab[sh.C + c] = qc;
        tab[2 * sh.C + c] = -rstd * m1 - mean * qc;
    }
    __syncthreads();
    if (r >= sh.rows) return;
    float P[8], Qc[8], R[8];
    load_tab8(tab, q, P); load_tab8(tab + sh.C, q, Qc); load_tab8(tab + 2 * sh.C, q, R);
    const int64_t p0 = (int64_t)blockIdx.x * sh.iters * sh.rows;
    const int64_t p1 = min(p0 + (int64_t)sh.iters * sh.rows, sh.HW);
    const __nv_bfloat16 *xb = x + n * sh.HW * ld_x + 8 * q;
    const __nv_bfloat16 *ab = gadd ? gadd + n * sh.HW * ld_gadd + 8 * q : nullptr;
    __nv_bfloat16 *gb = gx + n * sh.HW * ld_gx + 8 * q;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    for (int64_t p = p0 + r; p < p1; p += 2 * sh.rows) {
        const int64_t pb = p + sh.rows;
        const bool two = pb < p1;
        const uint4 xa = ld_stream_u4(reinterpret_cast<const uint4 *>(xb + p * ld_x));
        const uint4 da = *reinterpret_cast<const uint4 *>(gb + p * ld_gx);
        const uint4 xc = two ? ld_stream_u4(reinterpret_cast<const uint4 *>(xb + pb * ld_x)) : zero4;
        const uint4 dc = two ? *reinterpret_cast<const uint4 *>(gb + pb * ld_gx) : zero4;
        uint4 aa = zero4, ac = zero4;
        if (ab) {
            aa = ld_stream_u4(reinterpret_cast<const uint4 *>(ab + p * ld_gadd));
            if (two) ac = ld_stream_u4(reinterpret_cast<const uint4 *>(ab + pb * ld_gadd));
        }
        float f[8], d[8], a[8];
        unpack8(xa, f); unpack8(da, d); unpack8(aa, a);
#pragma unroll
        for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u])) + a[u];
        *reinterpret_cast<uint4 *>(gb + p * ld_gx) = pack8(d);
        if (two) {
            unpack8(xc, f); unpack8(dc, d); unpack8(ac, a);
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] = fmaf(d[u], P[u], fmaf(f[u], Qc[u], R[u])) + a[u];
            *reinterpret_cast<uint4 *>(gb + pb * ld_gx) = pack8(d);
        }
    }
}

// CTA plan of the streaming kernels; false when the shape is not eligible
bool plan_stream(int64_t N, int64_t HW, int64_t C, int G, StreamShape &sh) {
    if (N <= 0 || HW <= 0 || C <= 0 || G <= 0 || C % 8 != 0 || C % G != 0 || C > 8 * kStreamThreads || N > 65535) return false;
    sh.HW = HW; sh.C = (int)C; sh.G = G; sh.cpg = (int)(C / G); sh.chunks = (int)(C / 8);
    sh.rows = kStreamThreads / sh.chunks;
    const int64_t turns = (HW + sh.rows - 1) / sh.rows;              // row-turns per sample
    // about one wave of 4 resident CTAs per SM, at least 4 turns per CTA; among [want, 2*want] the split that wastes least
    int64_t want = (ub::kSMs * 4 + N - 1) / N;
    const int64_t most = turns / 4 > 1 ? turns / 4 : 1;
    if (want > most) want = most;
    if (want > kStreamMaxSplits) want = kStreamMaxSplits;
    int64_t best = want, best_waste = -1;
    for (int64_t s = want; s <= 2 * want && s <= most && s <= kStreamMaxSplits; ++s) {
        const int64_t it = (turns + s - 1) / s, waste = it * s - turns;
        if ((turns + it - 1) / it != s) continue;                     // would leave an empty CTA
        if (best_waste < 0 || waste < best_waste) { best = s; best_waste = waste; }
    }
    sh.iters = (int)((turns + best - 1) / best);
    sh.splits = (int)((turns + sh.iters - 1) / sh.iters);
    return true;
}

// policy: the cluster kernels win on small slabs (one or two CTAs hold a sample: short chain, one launch)
bool stream_preferred(int64_t HW, int64_t C, bool backward) {
    static const int mode = [] { const char *e = getenv("UB200_GN_STREAM"); return e ? atoi(e) : -1; }();   // A/B switch
    if (mode == 0) return false;
    if (mode == 1) return true;
    static const long min_slab = [] { const char *e = getenv("UB200_GN_STREAM_MIN_SLAB"); return e ? atol(e) : 0L; }();
    static const long max_slab = [] { const char *e = getenv("UB200_GN_STREAM_MAX_SLAB"); return e ? atol(e) : (1L << 40); }();
    // 1 forward, 2 backward, 3 both.  Measured in the config-2 step graph (B200, ms/step): neither 6.42, both 6.32,
    // backward only 6.25, forward only 6.48 -- the one-launch cluster forward stays, the backward streams.
    static const int dir = [] { const char *e = getenv("UB200_GN_STREAM_DIR"); return e ? atoi(e) : 2; }();
    if (!(dir & (backward ? 2 : 1))) return false;
    return HW * C * 2 >= min_slab && HW * C * 2 <= max_slab;
}

// cluster size / smem plan for the fused kernels; returns false when the slab does not fit
bool plan_fused(int64_t N, int64_t HW, int64_t C, int G, int tensors, size_t extra_floats, FusedShape &sh, size_t &smem,
                int only_single_cta = 0) {
    static const bool enabled = [] { const char *e = getenv("UB200_GN_FUSED"); return !(e && e[0] == '0'); }();   // A/B switch
    if (!enabled) return false;
    if (C % 8 != 0 || C % G != 0 || C > 4 * kFusedThreads || N * 8 > 2147483647LL) return false;
    sh.HW = HW; sh.C = (int)C; sh.G = G; sh.cpg = (int)(C / G); sh.chunks = (int)(C / 8);
    sh.rows = (kFusedThreads / sh.chunks) & ~1;      // even, >= 2: the fold staging holds rows/2 * chunks <= 128 entries
    const size_t fixed = (size_t)kFusedFixedBytes + extra_floats * 4;
    if (only_single_cta) {
        const size_t bytes = fixed + (size_t)tensors * HW * C * 2;
        if (bytes > 73 * 1024) return false;
        sh.cs = 1; sh.pix_per_cta = HW; smem = bytes;
        return true;
    }
    // (larger clusters were measured: 16x16x256 forward 18.4 us at the smallest fitting cluster vs 24.6 us at the largest,
    // and non-portable 16-CTA clusters are slower still -- cluster barriers and DSMEM reads cost more than balance gains)
    // smallest cluster whose per-CTA slab leaves room for 3, else 2 CTAs per SM; at 8 CTAs accept one per SM
    static const size_t limits[3] = {73 * 1024, 110 * 1024, 220 * 1024};
    for (int pass = 0; pass < 3; ++pass) {
        for (int cs = (pass == 2 ? 8 : 1); cs <= 8; cs *= 2) {
            const int64_t ppc = (HW + cs - 1) / cs;
            if (cs > 1 && (int64_t)(cs - 1) * ppc >= HW) continue;                 // would leave an empty CTA
            const size_t bytes = fixed + (size_t)tensors * ppc * C * 2;
            if (bytes <= limits[pass]) { sh.cs = cs; sh.pix_per_cta = ppc; smem = bytes; return true; }
        }
    }
    return false;
}

std::mutex g_attr_mu;
std::unordered_set<const void *> g_attr_done;

template <typename K, typename... Args>
int launch_cluster(K kernel, int grid, int cs, size_t smem, cudaStream_t s, Args... args) {
    cudaError_t e = cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_attr_mu);
        if (!g_attr_done.count(reinterpret_cast<const void *>(kernel))) {
            e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e == cudaSuccess) g_attr_done.insert(reinterpret_cast<const void *>(kernel));
        }
    }
    if (e != cudaSuccess) return (int)e;
    prefer_max_smem_carveout(reinterpret_cast<const void *>(kernel));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid, 1, 1);
    cfg.blockDim = dim3(kFusedThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    pdl_attr(attr[1]);
    cfg.attrs = attr; cfg.numAttrs = 2;
    e = cudaLaunchKernelEx(&cfg, kernel, args...);
    return e == cudaSuccess ? (int)cudaPeekAtLastError() : (int)e;
}

int make_shape(int64_t N, int64_t HW, int64_t C, int G, Shape &sh, dim3 &grid) {
    if (N <= 0 || HW <= 0 || C <= 0 || G <= 0) return UB200_E_BADARG;
    if (C % 8 != 0 || C % G != 0 || C > 2048 || G > 1024 || N > 65535) return UB200_E_UNSUPPORTED;
    sh.HW = HW; sh.C = (int)C; sh.G = G; sh.cpg = (int)(C / G); sh.chunks = (int)(C / 8);
    sh.rows = 256 / sh.chunks;
    // aim at ~4 waves of 148 SMs x 8 CTAs; at least `rows` pixels per CTA
    int64_t splits = (148 * 8 * 2 + N - 1) / N;
    int64_t max_splits = (HW + sh.rows - 1) / sh.rows;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    sh.pix_per_cta = (HW + splits - 1) / splits;
    splits = (HW + sh.pix_per_cta - 1) / sh.pix_per_cta;
    grid = dim3((unsigned)splits, (unsigned)N, 1);
    return UB200_OK;
}

#define DISPATCH_ACT_DROP(KERNEL, act, drop, ...)                                                    \
    do {                                                                                             \
        if (drop) {                                                                                  \
            if (act == UB200_ACT_SILU) KERNEL<UB200_ACT_SILU, true> __VA_ARGS__;                     \
            else if (act == UB200_ACT_GELU) KERNEL<UB200_ACT_GELU, true> __VA_ARGS__;                \
            else if (act == UB200_ACT_RELU) KERNEL<UB200_ACT_RELU, true> __VA_ARGS__;                \
            else KERNEL<UB200_ACT_NONE, true> __VA_ARGS__;                                           \
        } else {                                                                                     \
            if (act == UB200_ACT_SILU) KERNEL<UB200_ACT_SILU, false> __VA_ARGS__;                    \
            else if (act == UB200_ACT_GELU) KERNEL<UB200_ACT_GELU, false> __VA_ARGS__;               \
            else if (act == UB200_ACT_RELU) KERNEL<UB200_ACT_RELU, false> __VA_ARGS__;               \
            else KERNEL<UB200_ACT_NONE, false> __VA_ARGS__;                                          \
        }                                                                                            \
    } while (0)

#define STREAM_ACT_DROP(KERNEL, act, drop, ...)                                                                   \
    ((drop) ? ((act) == UB200_ACT_SILU   ? launch_pdl(KERNEL<UB200_ACT_SILU, true>, __VA_ARGS__)                  \
               : (act) == UB200_ACT_GELU ? launch_pdl(KERNEL<UB200_ACT_GELU, true>, __VA_ARGS__)                  \
               : (act) == UB200_ACT_RELU ? launch_pdl(KERNEL<UB200_ACT_RELU, true>, __VA_ARGS__)                  \
                                         : launch_pdl(KERNEL<UB200_ACT_NONE, true>, __VA_ARGS__))                 \
            : ((act) == UB200_ACT_SILU   ? launch_pdl(KERNEL<UB200_ACT_SILU, false>, __VA_ARGS__)                 \
               : (act) == UB200_ACT_GELU ? launch_pdl(KERNEL<UB200_ACT_GELU, false>, __VA_ARGS__)                 \
               : (act) == UB200_ACT_RELU ? launch_pdl(KERNEL<UB200_ACT_RELU, false>, __VA_ARGS__)                 \
                                         : launch_pdl(KERNEL<UB200_ACT_NONE, false>, __VA_ARGS__)))

}  // namespace

extern "C" {

int ub200_gn_stats_nhwc_bf16(const void *x, int64_t ld, int64_t N, int64_t HW, int64_t C, int G, float *stats,
                             void *stream) {
    UB_REQUIRE(x && stats, UB200_E_BADARG);
    Shape sh; dim3 grid;
    int rc = make_shape(N, HW, C, G, sh, grid);
    if (rc) return rc;
    UB_REQUIRE(ld % 8 == 0 && ld >= C && ub::aligned16(x), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(stats, 0, sizeof(float) * 2 * N * G, s);
    if (e != cudaSuccess) return (int)e;
    gn_stats_kernel<<<grid, 256, 2 * G * sizeof(float), s>>>(reinterpret_cast<const __nv_bfloat16 *>(x), ld, sh, stats);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_gn_act_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G, const float *stats,
                               float eps, const float *gamma, const float *beta, const float *scale, const float *shift,
                               int act, float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                               const void *addend, int64_t ld_add, void *y, int64_t ld_y, void *stream) {
    UB_REQUIRE(x && y && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    Shape sh; dim3 grid;
    int rc = make_shape(N, HW, C, G, sh, grid);
    if (rc) return rc;
    UB_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0 && ld_x >= C && ld_y >= C && ub::aligned16(x) && ub::aligned16(y),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(!addend || (ld_add % 8 == 0 && ld_add >= C && ub::aligned16(addend)), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const bool drop = dropout_p > 0.f;
    StreamShape st;
    if (plan_stream(N, HW, C, G, st)) {      // per-CTA coefficient table + four loads in flight; part = stats, one partial
        cudaError_t e = STREAM_ACT_DROP(gn_stream_fwd_kernel, act, drop, dim3((unsigned)st.splits, (unsigned)N, 1),
                                        dim3(kStreamThreads), (size_t)(((2 * G + 3) & ~3) + 2 * C) * 4, s,
                                        reinterpret_cast<const __nv_bfloat16 *>(x), ld_x, st, stats, 1, (float *)nullptr, eps, gamma,
                                        beta, scale, shift, dropout_p, seed, offset, offset_dev,
                                        reinterpret_cast<const __nv_bfloat16 *>(addend), ld_add,
                                        reinterpret_cast<__nv_bfloat16 *>(y), ld_y);
        if (e != cudaSuccess) return (int)e;
        UB_LAUNCH_CHECK();
        return UB200_OK;
    }
    DISPATCH_ACT_DROP(gn_act_fwd_kernel, act, drop,
                      <<<grid, 256, 0, s>>>(reinterpret_cast<const __nv_bfloat16 *>(x), ld_x, sh, stats, gamma, beta,
                                            scale, shift, eps, dropout_p, seed, offset, offset_dev,
                                            reinterpret_cast<const __nv_bfloat16 *>(addend), ld_add,
                                            reinterpret_cast<__nv_bfloat16 *>(y), ld_y));
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

size_t ub200_gn_act_bwd_ws_floats(int64_t N, int64_t C, int G) {
    (void)G;
    return (N > 0 && C > 0) ? (size_t)(2 * N * C) : 0;
}

int ub200_gn_act_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x, int64_t N, int64_t HW,
                               int64_t C, int G, const float *stats, float eps, const float *gamma, const float *beta,
                               const float *scale, const float *shift, int act, float dropout_p, uint64_t seed,
                               uint64_t offset, const uint64_t *offset_dev, void *gx, int64_t ld_gx, int accumulate, float *dgamma, float *dbeta,
                               float *dscale, float *dshift, float *ws, void *stream) {
    UB_REQUIRE(gy && x && gx && ws && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    Shape sh; dim3 grid;
    int rc = make_shape(N, HW, C, G, sh, grid);
    if (rc) return rc;
    UB_REQUIRE(ld_x % 8 == 0 && ld_gy % 8 == 0 && ld_gx % 8 == 0 && ld_x >= C && ld_gy >= C && ld_gx >= C &&
                   ub::aligned16(x) && ub::aligned16(gy) && ub::aligned16(gx),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(accumulate == 0, UB200_E_UNSUPPORTED);     // pass 1 parks dz in the gx buffer
    cudaStream_t s = ub::as_stream(stream);
    cudaError_t e = cudaMemsetAsync(ws, 0, sizeof(float) * 2 * N * C, s);
    if (e != cudaSuccess) return (int)e;
    const bool drop = dropout_p > 0.f;
    const __nv_bfloat16 *gyp = reinterpret_cast<const __nv_bfloat16 *>(gy), *xp = reinterpret_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *gxp = reinterpret_cast<__nv_bfloat16 *>(gx);
    DISPATCH_ACT_DROP(gn_act_bwd_reduce, act, drop,
                      <<<grid, 256, 256 * 16 * sizeof(float), s>>>(gyp, ld_gy, xp, ld_x, sh, stats, gamma, beta, scale, shift,
                                                                   eps, dropout_p, seed, offset, offset_dev, gxp, ld_gx, ws));
    UB_LAUNCH_CHECK();
    gn_act_bwd_apply<<<grid, 256, 2 * G * sizeof(float), s>>>(xp, ld_x, sh, stats, gamma, beta, scale, eps, ws, gxp, ld_gx,
                                                             dgamma, dbeta, dscale, dshift);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

#define DISPATCH_FUSED(KERNEL, act, drop, ...)                                                                  \
    ((drop) ? ((act) == UB200_ACT_SILU   ? launch_cluster(KERNEL(UB200_ACT_SILU, true), __VA_ARGS__)            \
               : (act) == UB200_ACT_GELU ? launch_cluster(KERNEL(UB200_ACT_GELU, true), __VA_ARGS__)            \
               : (act) == UB200_ACT_RELU ? launch_cluster(KERNEL(UB200_ACT_RELU, true), __VA_ARGS__)            \
                                         : launch_cluster(KERNEL(UB200_ACT_NONE, true), __VA_ARGS__))           \
            : ((act) == UB200_ACT_SILU   ? launch_cluster(KERNEL(UB200_ACT_SILU, false), __VA_ARGS__)           \
               : (act) == UB200_ACT_GELU ? launch_cluster(KERNEL(UB200_ACT_GELU, false), __VA_ARGS__)           \
               : (act) == UB200_ACT_RELU ? launch_cluster(KERNEL(UB200_ACT_RELU, false), __VA_ARGS__)           \
                                         : launch_cluster(KERNEL(UB200_ACT_NONE, false), __VA_ARGS__)))
#define FUSED_FWD(A, D) gn_fused_fwd_kernel<A, D>
#define FUSED_BWD_X(A, D) gn_fused_bwd_kernel<A, D, true>
#define FUSED_BWD_S(A, D) gn_fused_bwd_kernel<A, D, false>

int ub200_gn_act_fused_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G, float *stats,
                                     float eps, const float *gamma, const float *beta, const float *scale,
                                     const float *shift, int act, float dropout_p, uint64_t seed, uint64_t offset,
                                     const uint64_t *offset_dev, const void *addend, int64_t ld_add, void *y,
                                     int64_t ld_y, void *stream) {
    UB_REQUIRE(x && y && stats && N > 0 && HW > 0 && C > 0 && G > 0 && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    FusedShape sh; size_t smem = 0;
    const size_t gpad = (size_t)((2 * G + 3) & ~3);
    const bool ok = ld_x % 8 == 0 && ld_y % 8 == 0 && ld_x >= C && ld_y >= C && ub::aligned16(x) && ub::aligned16(y) &&
                    (!addend || (ld_add % 8 == 0 && ld_add >= C && ub::aligned16(addend))) &&
                    plan_fused(N, HW, C, G, 1, 2 * (size_t)C + 2 * gpad, sh, smem);
    if (!ok) {      // slab too large for one cluster: statistics pass + apply pass
        int rc = ub200_gn_stats_nhwc_bf16(x, ld_x, N, HW, C, G, stats, stream);
        if (rc) return rc;
        return ub200_gn_act_fwd_nhwc_bf16(x, ld_x, N, HW, C, G, stats, eps, gamma, beta, scale, shift, act, dropout_p, seed,
                                          offset, offset_dev, addend, ld_add, y, ld_y, stream);
    }
    const bool drop = dropout_p > 0.f;
    return DISPATCH_FUSED(FUSED_FWD, act, drop, (int)(N * sh.cs), sh.cs, smem, ub::as_stream(stream),
                          reinterpret_cast<const __nv_bfloat16 *>(x), ld_x, sh, stats, eps, gamma, beta, scale, shift,
                          dropout_p, seed, offset, offset_dev, reinterpret_cast<const __nv_bfloat16 *>(addend), ld_add,
                          reinterpret_cast<__nv_bfloat16 *>(y), ld_y);
}

int ub200_gn_act_fused_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x, int64_t N, int64_t HW,
                                     int64_t C, int G, const float *stats, float eps, const float *gamma,
                                     const float *beta, const float *scale, const float *shift, int act,
                                     float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                     void *gx, int64_t ld_gx, float *dgamma, float *dbeta, float *dscale, float *dshift,
                                     const void *gadd, int64_t ld_gadd, float *ws, void *stream) {
    UB_REQUIRE(!gadd || (ld_gadd % 8 == 0 && ld_gadd >= C && ub::aligned16(gadd)), UB200_E_UNSUPPORTED);
    UB_REQUIRE(gy && x && gx && ws && N > 0 && HW > 0 && C > 0 && G > 0 && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    FusedShape sh; size_t smem = 0;
    const bool fits = ld_x % 8 == 0 && ld_gy % 8 == 0 && ld_gx % 8 == 0 && ld_x >= C && ld_gy >= C && ld_gx >= C &&
                      ub::aligned16(x) && ub::aligned16(gy) && ub::aligned16(gx);
    const size_t extra = (size_t)(4 * C + ((2 * G + 3) & ~3));
    const bool xslab = fits && plan_fused(N, HW, C, G, 2, extra, sh, smem, 1);
    const bool ok = xslab || (fits && plan_fused(N, HW, C, G, 1, extra, sh, smem));
    if (!ok) {
        int rc = ub200_gn_act_bwd_nhwc_bf16(gy, ld_gy, x, ld_x, N, HW, C, G, stats, eps, gamma, beta, scale, shift, act,
                                            dropout_p, seed, offset, offset_dev, gx, ld_gx, 0, dgamma, dbeta, dscale, dshift,
                                            ws, stream);
        if (rc || !gadd) return rc;
        UB_REQUIRE(C % 8 == 0 && ld_gx % 8 == 0 && ub::aligned16(gx), UB200_E_UNSUPPORTED);
        add_rows_kernel<<<148 * 8, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<__nv_bfloat16 *>(gx), ld_gx,
                                                                    reinterpret_cast<const __nv_bfloat16 *>(gadd), ld_gadd,
                                                                    N * HW, (int)(C / 8));
        UB_LAUNCH_CHECK();
        return UB200_OK;
    }
    const bool drop = dropout_p > 0.f;
#define BWD_ARGS                                                                                                      \
    (int)(N * sh.cs), sh.cs, smem, ub::as_stream(stream), reinterpret_cast<const __nv_bfloat16 *>(gy), ld_gy,         \
        reinterpret_cast<const __nv_bfloat16 *>(x), ld_x, sh, stats, eps, gamma, beta, scale, shift, dropout_p, seed, \
        offset, offset_dev, reinterpret_cast<__nv_bfloat16 *>(gx), ld_gx, dgamma, dbeta, dscale, dshift,                 \
        reinterpret_cast<const __nv_bfloat16 *>(gadd), ld_gadd
    if (xslab) return DISPATCH_FUSED(FUSED_BWD_X, act, drop, BWD_ARGS);
    return DISPATCH_FUSED(FUSED_BWD_S, act, drop, BWD_ARGS);
#undef BWD_ARGS
}

/* ---- streaming two-launch variants (see the kernel comment) ---- */
size_t ub200_gn_stream_ws_floats(int64_t N, int64_t HW, int64_t C, int G) {
    StreamShape sh;
    if (!plan_stream(N, HW, C, G, sh)) return 0;
    return (size_t)N * (size_t)sh.splits * 2 * (size_t)C;      // backward partials; the forward uses [N][splits][2G] of it
}

int ub200_gn_stream_preferred(int64_t N, int64_t HW, int64_t C, int G, int backward) {
    StreamShape sh;
    if (!plan_stream(N, HW, C, G, sh)) return 0;
    if (stream_preferred(HW, C, backward != 0)) return 1;
    // A slab no cluster can hold (pdearena / wmh: GroupNorm(1, C) over 128x128 or 200x200 pixels) used to fall back to memset +
    // atomic statistics (gn_stats_kernel: 66 us for 17 MB at batch 8, 0.25 TB/s) + apply; the streaming pair does the same two
    // passes with per-CTA partials at 4x the bandwidth.  UB200_GN_STREAM=0 keeps the old fallback for A/B.
    static const bool never = [] { const char *e = getenv("UB200_GN_STREAM"); return e && atoi(e) == 0; }();
    if (never) return 0;
    FusedShape fs; size_t smem = 0;
    if (!backward) return plan_fused(N, HW, C, G, 1, 2 * (size_t)C + 2 * (size_t)((2 * G + 3) & ~3), fs, smem) ? 0 : 1;
    const size_t extra = (size_t)(4 * C + ((2 * G + 3) & ~3));
    return (plan_fused(N, HW, C, G, 2, extra, fs, smem, 1) || plan_fused(N, HW, C, G, 1, extra, fs, smem)) ? 0 : 1;
}

int ub200_gn_act_stream_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G, float *stats,
                                      float eps, const float *gamma, const float *beta, const float *scale,
                                      const float *shift, int act, float dropout_p, uint64_t seed, uint64_t offset,
                                      const uint64_t *offset_dev, const void *addend, int64_t ld_add, void *y,
                                      int64_t ld_y, float *ws, void *stream) {
    UB_REQUIRE(x && y && stats && ws && N > 0 && HW > 0 && C > 0 && G > 0 && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    StreamShape sh;
    UB_REQUIRE(plan_stream(N, HW, C, G, sh), UB200_E_UNSUPPORTED);
    UB_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0 && ld_x >= C && ld_y >= C && ub::aligned16(x) && ub::aligned16(y) &&
                   (!addend || (ld_add % 8 == 0 && ld_add >= C && ub::aligned16(addend))),
               UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const dim3 grid((unsigned)sh.splits, (unsigned)N, 1);
    const __nv_bfloat16 *xp = reinterpret_cast<const __nv_bfloat16 *>(x);
    cudaError_t e = launch_pdl(gn_stream_stats_kernel, grid, dim3(kStreamThreads), (size_t)(kStreamThreads * 16 + 2 * C) * 4, s,
                               xp, ld_x, sh, ws);
    if (e != cudaSuccess) return (int)e;
    const size_t smem = (size_t)(((2 * G + 3) & ~3) + 2 * C) * 4;
    const bool drop = dropout_p > 0.f;
    e = STREAM_ACT_DROP(gn_stream_fwd_kernel, act, drop, grid, dim3(kStreamThreads), smem, s, xp, ld_x, sh, (const float *)ws,
                        sh.splits, stats, eps, gamma, beta, scale, shift, dropout_p, seed, offset, offset_dev,
                        reinterpret_cast<const __nv_bfloat16 *>(addend), ld_add, reinterpret_cast<__nv_bfloat16 *>(y), ld_y);
    if (e != cudaSuccess) return (int)e;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_gn_act_stream_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x, int64_t N, int64_t HW,
                                      int64_t C, int G, const float *stats, float eps, const float *gamma,
                                      const float *beta, const float *scale, const float *shift, int act,
                                      float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                      void *gx, int64_t ld_gx, float *dgamma, float *dbeta, float *dscale, float *dshift,
                                      const void *gadd, int64_t ld_gadd, float *ws, void *stream) {
    UB_REQUIRE(gy && x && gx && ws && N > 0 && HW > 0 && C > 0 && G > 0 && dropout_p >= 0.f && dropout_p < 1.f, UB200_E_BADARG);
    UB_REQUIRE(act >= UB200_ACT_NONE && act <= UB200_ACT_RELU, UB200_E_BADARG);
    StreamShape sh;
    UB_REQUIRE(plan_stream(N, HW, C, G, sh), UB200_E_UNSUPPORTED);
    UB_REQUIRE(ld_x % 8 == 0 && ld_gy % 8 == 0 && ld_gx % 8 == 0 && ld_x >= C && ld_gy >= C && ld_gx >= C &&
                   ub::aligned16(x) && ub::aligned16(gy) && ub::aligned16(gx) &&
                   (!gadd || (ld_gadd % 8 == 0 && ld_gadd >= C && ub::aligned16(gadd))),
               UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const dim3 grid((unsigned)sh.splits, (unsigned)N, 1);
    const __nv_bfloat16 *gyp = reinterpret_cast<const __nv_bfloat16 *>(gy), *xp = reinterpret_cast<const __nv_bfloat16 *>(x);
    __nv_bfloat16 *gxp = reinterpret_cast<__nv_bfloat16 *>(gx);
    const bool drop = dropout_p > 0.f;
    cudaError_t e = STREAM_ACT_DROP(gn_stream_bwd_reduce, act, drop, grid, dim3(kStreamThreads),
                                    (size_t)(kStreamThreads * 16 + 2 * C) * 4, s, gyp, ld_gy, xp, ld_x, sh, stats, eps, gamma, beta,
                                    scale, shift, dropout_p, seed, offset, offset_dev, gxp, ld_gx, ws);
    if (e != cudaSuccess) return (int)e;
    e = launch_pdl(gn_stream_bwd_apply, grid, dim3(kStreamThreads), (size_t)(5 * C + 2 * G) * 4, s, xp, ld_x, sh, stats, gamma,
                   beta, scale, eps, (const float *)ws, gxp, ld_gx, dgamma, dbeta, dscale, dshift,
                   reinterpret_cast<const __nv_bfloat16 *>(gadd), ld_gadd);
    if (e != cudaSuccess) return (int)e;
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
