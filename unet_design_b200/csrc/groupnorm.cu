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
#include "common.cuh"

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

// keep-mask for the 8 channels starting at element index `elem` (multiple of 8)
__device__ __forceinline__ void dropout_mask8(uint64_t seed, uint64_t offset, int64_t elem, float p, float (&m)[8]) {
    const uint32_t thr = (uint32_t)fminf(p * 4294967296.0f, 4294967295.0f);
    const float keep_scale = 1.0f / (1.0f - p);
    const uint4 r0 = philox4(seed, offset + (uint64_t)(elem >> 2));
    const uint4 r1 = philox4(seed, offset + (uint64_t)(elem >> 2) + 1);
    const uint32_t r[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int u = 0; u < 8; ++u) m[u] = r[u] >= thr ? keep_scale : 0.f;
}

// sigmoid through ONE special-function op: 0.5 * tanh(z/2) + 0.5  (tanh.approx.f32, relative error 2^-11: far below
// the bf16 rounding of every value this feeds).  These kernels are issue-bound, not bandwidth-bound, on B200.
__device__ __forceinline__ float fast_sigmoid(float z) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * z));
    return fmaf(t, 0.5f, 0.5f);
}

template <int ACT>
__device__ __forceinline__ float act_fwd(float z) {
    if constexpr (ACT == UB200_ACT_SILU) return z * fast_sigmoid(z);
    else if constexpr (ACT == UB200_ACT_GELU) return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
    else return z;
}
template <int ACT>
__device__ __forceinline__ float act_bwd(float z) {   // d act / d z
    if constexpr (ACT == UB200_ACT_SILU) {
        const float s = fast_sigmoid(z);
        return fmaf(z, fmaf(-s, s, s), s);          // s + z*s*(1-s)
    } else if constexpr (ACT == UB200_ACT_GELU) {
        return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * __expf(-0.5f * z * z);
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
                                          const float *beta, const float *scale, const float *shift, float eps) {
    Coef k;
    const float inv_cnt = 1.0f / ((float)sh.cpg * (float)sh.HW);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const int c = 8 * q + u;
        float mean, rstd;
        mean_rstd(stats, n, sh.G, c / sh.cpg, inv_cnt, eps, mean, rstd);
        const float ga = gamma ? __ldg(gamma + c) : 1.f, be = beta ? __ldg(beta + c) : 0.f;
        const float sc = scale ? 1.f + __ldg(scale + n * sh.C + c) : 1.f, sf = shift ? __ldg(shift + n * sh.C + c) : 0.f;
        k.A[u] = rstd * ga * sc;
        k.B[u] = (be - mean * rstd * ga) * sc + sf;
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
    const Coef k = make_coef(sh, n, q, stats, gamma, beta, scale, shift, eps);
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
        const Coef k = make_coef(sh, n, q, stats, gamma, beta, scale, shift, eps);
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
            else KERNEL<UB200_ACT_NONE, true> __VA_ARGS__;                                           \
        } else {                                                                                     \
            if (act == UB200_ACT_SILU) KERNEL<UB200_ACT_SILU, false> __VA_ARGS__;                    \
            else if (act == UB200_ACT_GELU) KERNEL<UB200_ACT_GELU, false> __VA_ARGS__;               \
            else KERNEL<UB200_ACT_NONE, false> __VA_ARGS__;                                          \
        }                                                                                            \
    } while (0)

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
    UB_REQUIRE(act == UB200_ACT_NONE || act == UB200_ACT_SILU || act == UB200_ACT_GELU, UB200_E_BADARG);
    Shape sh; dim3 grid;
    int rc = make_shape(N, HW, C, G, sh, grid);
    if (rc) return rc;
    UB_REQUIRE(ld_x % 8 == 0 && ld_y % 8 == 0 && ld_x >= C && ld_y >= C && ub::aligned16(x) && ub::aligned16(y),
               UB200_E_UNSUPPORTED);
    UB_REQUIRE(!addend || (ld_add % 8 == 0 && ld_add >= C && ub::aligned16(addend)), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const bool drop = dropout_p > 0.f;
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
    UB_REQUIRE(act == UB200_ACT_NONE || act == UB200_ACT_SILU || act == UB200_ACT_GELU, UB200_E_BADARG);
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

}  // extern "C"
