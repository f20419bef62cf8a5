// Haar DWT / iDWT / DTWBlock kernels (memory-bound; fp32 NCHW as in the reference).
//
// Replaces pytorch_wavelets.DWTForward/DWTInverse(mode='zero', wave='haar') at the reference
// call sites diff_cifar/model.py:263-267,:310-321; diff_cifar/diffusion.py:63-70;
// pdearena/pdearena/modules/twod_unetbase.py:169-193; wmh/model.py:68-95.
//
// Arithmetic follows the upstream separable form (W axis first, taps s = fl32(1/sqrt 2)) with
// contraction disabled (__fmul_rn/__fadd_rn), so results are bit-identical to oracle/haar_np.py
// and oracle/haar_c.c.  Algorithmic bytes per launch (E = planes*H*W, b = 4):
//   full 4-band level   : read E*b, write E*b            -> 2*E*b
//   LL_J + tile (r=out/C): read E*b, write r*E*b*4^-J    -> E*b*(1 + r*4^-J)
// Every kernel is a grid-stride loop over 128-bit work items, grid = multiple of 148 SMs.
#include "common.cuh"

namespace {

using namespace ub;

__device__ __forceinline__ float mul_s(float v) { return __fmul_rn(0.70710678118654752440f, v); }
__device__ __forceinline__ float add2(float a, float b) { return __fadd_rn(mul_s(a), mul_s(b)); }   // s*a + s*b
__device__ __forceinline__ float sub2(float a, float b) { return __fsub_rn(mul_s(a), mul_s(b)); }   // s*a - s*b

struct Bands { float ll, lh, hl, hh; };

__device__ __forceinline__ Bands analyse(float a, float b, float c, float d) {
    float lo_t = add2(a, b), hi_t = sub2(a, b);
    float lo_b = add2(c, d), hi_b = sub2(c, d);
    return {add2(lo_t, lo_b), sub2(lo_t, lo_b), add2(hi_t, hi_b), sub2(hi_t, hi_b)};
}
__device__ __forceinline__ float analyse_ll(float a, float b, float c, float d) {
    return add2(add2(a, b), add2(c, d));
}

// ---------------------------------------------------------------------------------------------
// One analysis level, fast path: W % 8 == 0, 16-byte aligned base.  One work item = 4 output
// columns of one output row: 2 x float4 from each of two input rows, one float4 store per band.
// ---------------------------------------------------------------------------------------------
template <bool HIGHS>
__global__ void __launch_bounds__(256) haar_dwt_vec4(const float *__restrict__ x, int64_t planes, int H, int W,
                                                    float *__restrict__ ll, float *__restrict__ highs) {
    const int h2 = (H + 1) >> 1, w2 = W >> 1, wq = w2 >> 2;
    const int64_t items = planes * h2 * wq;
    const int64_t band = (int64_t)h2 * w2;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int jq = (int)(it % wq);
        const int64_t t = it / wq;
        const int i = (int)(t % h2);
        const int64_t p = t / h2;
        const float *row0 = x + (p * H + 2 * i) * (int64_t)W + 8 * jq;
        float4 t0 = ld_stream(reinterpret_cast<const float4 *>(row0));
        float4 t1 = ld_stream(reinterpret_cast<const float4 *>(row0) + 1);
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (2 * i + 1 < H) {
            b0 = ld_stream(reinterpret_cast<const float4 *>(row0 + W));
            b1 = ld_stream(reinterpret_cast<const float4 *>(row0 + W) + 1);
        }
        Bands q0 = analyse(t0.x, t0.y, b0.x, b0.y), q1 = analyse(t0.z, t0.w, b0.z, b0.w);
        Bands q2 = analyse(t1.x, t1.y, b1.x, b1.y), q3 = analyse(t1.z, t1.w, b1.z, b1.w);
        const int64_t o = (p * h2 + i) * (int64_t)w2 + 4 * jq;
        st_stream(reinterpret_cast<float4 *>(ll + o), make_float4(q0.ll, q1.ll, q2.ll, q3.ll));
        if (HIGHS) {
            float *hp = highs + p * 3 * band + (int64_t)i * w2 + 4 * jq;
            st_stream(reinterpret_cast<float4 *>(hp), make_float4(q0.lh, q1.lh, q2.lh, q3.lh));
            st_stream(reinterpret_cast<float4 *>(hp + band), make_float4(q0.hl, q1.hl, q2.hl, q3.hl));
            st_stream(reinterpret_cast<float4 *>(hp + 2 * band), make_float4(q0.hh, q1.hh, q2.hh, q3.hh));
        }
    }
}

// Any extents (odd H / W, tiny planes): one output coefficient per work item, zero extension.
template <bool HIGHS>
__global__ void __launch_bounds__(256) haar_dwt_any(const float *__restrict__ x, int64_t planes, int H, int W,
                                                   float *__restrict__ ll, float *__restrict__ highs) {
    const int h2 = (H + 1) >> 1, w2 = (W + 1) >> 1;
    const int64_t items = planes * h2 * w2;
    const int64_t band = (int64_t)h2 * w2;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(it % w2);
        const int64_t t = it / w2;
        const int i = (int)(t % h2);
        const int64_t p = t / h2;
        const float *src = x + p * (int64_t)H * W;
        const bool r1 = 2 * i + 1 < H, c1 = 2 * j + 1 < W;
        float a = __ldg(src + (int64_t)(2 * i) * W + 2 * j);
        float b = c1 ? __ldg(src + (int64_t)(2 * i) * W + 2 * j + 1) : 0.f;
        float c = r1 ? __ldg(src + (int64_t)(2 * i + 1) * W + 2 * j) : 0.f;
        float d = (r1 && c1) ? __ldg(src + (int64_t)(2 * i + 1) * W + 2 * j + 1) : 0.f;
        Bands q = analyse(a, b, c, d);
        ll[it] = q.ll;
        if (HIGHS) {
            float *hp = highs + p * 3 * band + (int64_t)i * w2 + j;
            hp[0] = q.lh; hp[band] = q.hl; hp[2 * band] = q.hh;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// One synthesis level.  Fast path: w2 % 4 == 0, Wout == 2*w2, aligned.  One work item = 4
// coefficient columns of one coefficient row -> 2 output rows x 8 floats.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void synth(float vll, float vlh, float vhl, float vhh, float &o00, float &o01, float &o10, float &o11) {
    float lo_t = add2(vll, vlh), lo_b = sub2(vll, vlh);
    float hi_t = add2(vhl, vhh), hi_b = sub2(vhl, vhh);
    o00 = add2(lo_t, hi_t); o01 = sub2(lo_t, hi_t);
    o10 = add2(lo_b, hi_b); o11 = sub2(lo_b, hi_b);
}

template <bool HIGHS>
__global__ void __launch_bounds__(256) haar_idwt_vec4(const float *__restrict__ ll, const float *__restrict__ highs,
                                                     int64_t planes, int h2, int w2, int Hout, float *__restrict__ out) {
    const int wq = w2 >> 2, W = 2 * w2;
    const int64_t items = planes * h2 * wq;
    const int64_t band = (int64_t)h2 * w2;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int jq = (int)(it % wq);
        const int64_t t = it / wq;
        const int i = (int)(t % h2);
        const int64_t p = t / h2;
        const int64_t o = (p * h2 + i) * (int64_t)w2 + 4 * jq;
        float4 a = ld_stream(reinterpret_cast<const float4 *>(ll + o));
        float4 z = make_float4(0.f, 0.f, 0.f, 0.f), b = z, c = z, d = z;
        if (HIGHS) {
            const float *hp = highs + p * 3 * band + (int64_t)i * w2 + 4 * jq;
            b = ld_stream(reinterpret_cast<const float4 *>(hp));
            c = ld_stream(reinterpret_cast<const float4 *>(hp + band));
            d = ld_stream(reinterpret_cast<const float4 *>(hp + 2 * band));
        }
        float4 r0a, r0b, r1a, r1b;
        synth(a.x, b.x, c.x, d.x, r0a.x, r0a.y, r1a.x, r1a.y);
        synth(a.y, b.y, c.y, d.y, r0a.z, r0a.w, r1a.z, r1a.w);
        synth(a.z, b.z, c.z, d.z, r0b.x, r0b.y, r1b.x, r1b.y);
        synth(a.w, b.w, c.w, d.w, r0b.z, r0b.w, r1b.z, r1b.w);
        float *dst = out + (p * Hout + 2 * i) * (int64_t)W + 8 * jq;
        st_stream(reinterpret_cast<float4 *>(dst), r0a);
        st_stream(reinterpret_cast<float4 *>(dst) + 1, r0b);
        if (2 * i + 1 < Hout) {
            st_stream(reinterpret_cast<float4 *>(dst + W), r1a);
            st_stream(reinterpret_cast<float4 *>(dst + W) + 1, r1b);
        }
    }
}

template <bool HIGHS>
__global__ void __launch_bounds__(256) haar_idwt_any(const float *__restrict__ ll, const float *__restrict__ highs,
                                                    int64_t planes, int h2, int w2, int Hout, int Wout,
                                                    float *__restrict__ out) {
    const int64_t items = planes * h2 * w2;
    const int64_t band = (int64_t)h2 * w2;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(it % w2);
        const int64_t t = it / w2;
        const int i = (int)(t % h2);
        const int64_t p = t / h2;
        float vlh = 0.f, vhl = 0.f, vhh = 0.f;
        if (HIGHS) {
            const float *hp = highs + p * 3 * band + (int64_t)i * w2 + j;
            vlh = __ldg(hp); vhl = __ldg(hp + band); vhh = __ldg(hp + 2 * band);
        }
        float o00, o01, o10, o11;
        synth(__ldg(ll + it), vlh, vhl, vhh, o00, o01, o10, o11);
        float *dst = out + p * (int64_t)Hout * Wout;
        const bool r1 = 2 * i + 1 < Hout, c1 = 2 * j + 1 < Wout;
        dst[(int64_t)(2 * i) * Wout + 2 * j] = o00;
        if (c1) dst[(int64_t)(2 * i) * Wout + 2 * j + 1] = o01;
        if (r1) dst[(int64_t)(2 * i + 1) * Wout + 2 * j] = o10;
        if (r1 && c1) dst[(int64_t)(2 * i + 1) * Wout + 2 * j + 1] = o11;
    }
}

// ---------------------------------------------------------------------------------------------
// DTWBlock: LL_J / 2^J + channel tile.  Level extents are carried so that the zero extension of an
// odd intermediate level is reproduced exactly (ll_at returns 0 outside a level's extent).
// ---------------------------------------------------------------------------------------------
struct Ext { int h[4], w[4]; };   // extent of level 0..3 (level 0 = input)

template <int L>
__device__ __forceinline__ float ll_at(const float *__restrict__ src, const Ext &e, int y, int x) {
    if (y >= e.h[L] || x >= e.w[L]) return 0.f;
    if constexpr (L == 0) {
        return __ldg(src + (int64_t)y * e.w[0] + x);
    } else {
        return analyse_ll(ll_at<L - 1>(src, e, 2 * y, 2 * x), ll_at<L - 1>(src, e, 2 * y, 2 * x + 1),
                          ll_at<L - 1>(src, e, 2 * y + 1, 2 * x), ll_at<L - 1>(src, e, 2 * y + 1, 2 * x + 1));
    }
}

__device__ __forceinline__ float ll_dyn(const float *src, const Ext &e, int J, int y, int x) {
    switch (J) {
        case 0: return ll_at<0>(src, e, y, x);
        case 1: return ll_at<1>(src, e, y, x);
        case 2: return ll_at<2>(src, e, y, x);
        default: return ll_at<3>(src, e, y, x);
    }
}

inline Ext make_ext(int64_t H, int64_t W) {
    Ext e;
    e.h[0] = (int)H; e.w[0] = (int)W;
    for (int l = 1; l < 4; ++l) { e.h[l] = (e.h[l - 1] + 1) / 2; e.w[l] = (e.w[l - 1] + 1) / 2; }
    return e;
}

// generic: one work item per (n, c, i, j); writes every k = c, c+C, ... < out_channels
__global__ void __launch_bounds__(256) dwtblock_any(const float *__restrict__ x, int64_t N, int C, Ext e, int J,
                                                   int out_channels, float scale, float *__restrict__ out) {
    const int ho = e.h[J], wo = e.w[J];
    const int64_t plane_o = (int64_t)ho * wo, plane_i = (int64_t)e.h[0] * e.w[0];
    const int64_t items = N * C * plane_o;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(it % wo);
        int64_t t = it / wo;
        const int i = (int)(t % ho);
        t /= ho;
        const int c = (int)(t % C);
        const int64_t n = t / C;
        const float v = ll_dyn(x + (n * C + c) * plane_i, e, J, i, j) * scale;
        for (int k = c; k < out_channels; k += C) out[(n * out_channels + k) * plane_o + (int64_t)i * wo + j] = v;
    }
}

// J = 1 fast path (W % 8 == 0, aligned): 4 output columns per work item, float4 stores per replica.
__global__ void __launch_bounds__(256) dwtblock_j1_vec4(const float *__restrict__ x, int64_t N, int C, int H, int W,
                                                       int out_channels, float *__restrict__ out) {
    const int h2 = (H + 1) >> 1, w2 = W >> 1, wq = w2 >> 2;
    const int64_t plane_o = (int64_t)h2 * w2;
    const int64_t items = N * C * h2 * wq;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int jq = (int)(it % wq);
        int64_t t = it / wq;
        const int i = (int)(t % h2);
        t /= h2;
        const int c = (int)(t % C);
        const int64_t n = t / C;
        const float *row0 = x + ((n * C + c) * H + 2 * i) * (int64_t)W + 8 * jq;
        float4 t0 = ld_stream(reinterpret_cast<const float4 *>(row0));
        float4 t1 = ld_stream(reinterpret_cast<const float4 *>(row0) + 1);
        float4 b0 = make_float4(0.f, 0.f, 0.f, 0.f), b1 = b0;
        if (2 * i + 1 < H) {
            b0 = ld_stream(reinterpret_cast<const float4 *>(row0 + W));
            b1 = ld_stream(reinterpret_cast<const float4 *>(row0 + W) + 1);
        }
        float4 v = make_float4(analyse_ll(t0.x, t0.y, b0.x, b0.y) * 0.5f, analyse_ll(t0.z, t0.w, b0.z, b0.w) * 0.5f,
                               analyse_ll(t1.x, t1.y, b1.x, b1.y) * 0.5f, analyse_ll(t1.z, t1.w, b1.z, b1.w) * 0.5f);
        for (int k = c; k < out_channels; k += C)
            st_stream(reinterpret_cast<float4 *>(out + (n * out_channels + k) * plane_o + (int64_t)i * w2 + 4 * jq), v);
    }
}

// J = 0 fast path (plane % 4 == 0, aligned): the channel tile is a replicated 128-bit copy.
__global__ void __launch_bounds__(256) tile_vec4(const float *__restrict__ x, int64_t N, int C, int64_t plane4,
                                                int out_channels, float *__restrict__ out) {
    const int64_t items = N * C * plane4;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = it % plane4;
        int64_t t = it / plane4;
        const int c = (int)(t % C);
        const int64_t n = t / C;
        float4 v = ld_stream(reinterpret_cast<const float4 *>(x) + it);
        for (int k = c; k < out_channels; k += C)
            st_stream(reinterpret_cast<float4 *>(out) + (n * out_channels + k) * plane4 + q, v);
    }
}

// backward: one work item per input pixel; folds the replicas, applies LL_J^T / 2^J
__global__ void __launch_bounds__(256) dwtblock_bwd_any(const float *__restrict__ g, int64_t N, int C, Ext e, int J,
                                                       int out_channels, float scale, float *__restrict__ gx) {
    const int H = e.h[0], W = e.w[0], ho = e.h[J], wo = e.w[J];
    const int64_t plane_o = (int64_t)ho * wo;
    const int64_t items = N * C * (int64_t)H * W;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int xw = (int)(it % W);
        int64_t t = it / W;
        const int y = (int)(t % H);
        t /= H;
        const int c = (int)(t % C);
        const int64_t n = t / C;
        float acc = 0.f;
        for (int k = c; k < out_channels; k += C)
            acc = __fadd_rn(acc, __ldg(g + (n * out_channels + k) * plane_o + (int64_t)(y >> J) * wo + (xw >> J)));
        acc *= scale;                                     // exact power of two
        for (int l = 0; l < J; ++l) acc = mul_s(mul_s(acc));   // synthesis with zero high bands, per level
        gx[it] = acc;
    }
}

// backward fast path, J = 1, W % 8 == 0, H even, aligned: one work item = 4 coefficient columns of one coefficient
// row -> 2 input rows x 8 floats.  Every coefficient's gradient is read once (float4) instead of four times.
__global__ void __launch_bounds__(256) dwtblock_bwd_j1_vec4(const float *__restrict__ g, int64_t N, int C, int H, int W,
                                                           int out_channels, float *__restrict__ gx) {
    const int h2 = H >> 1, w2 = W >> 1, wq = w2 >> 2;
    const int64_t plane_o = (int64_t)h2 * w2;
    const int64_t items = N * C * h2 * wq;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int jq = (int)(it % wq);
        int64_t t = it / wq;
        const int i = (int)(t % h2);
        t /= h2;
        const int c = (int)(t % C);
        const int64_t n = t / C;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = c; k < out_channels; k += C) {
            const float4 v = ld_stream(reinterpret_cast<const float4 *>(g + (n * out_channels + k) * plane_o + (int64_t)i * w2 + 4 * jq));
            acc.x = __fadd_rn(acc.x, v.x); acc.y = __fadd_rn(acc.y, v.y); acc.z = __fadd_rn(acc.z, v.z); acc.w = __fadd_rn(acc.w, v.w);
        }
        // same expression as the generic path: (acc * 0.5) then s*(s*.)
        const float a = mul_s(mul_s(acc.x * 0.5f)), b = mul_s(mul_s(acc.y * 0.5f));
        const float cc = mul_s(mul_s(acc.z * 0.5f)), d = mul_s(mul_s(acc.w * 0.5f));
        const float4 lo = make_float4(a, a, b, b), hi = make_float4(cc, cc, d, d);
        float *dst = gx + ((n * C + c) * H + 2 * i) * (int64_t)W + 8 * jq;
        st_stream(reinterpret_cast<float4 *>(dst), lo);
        st_stream(reinterpret_cast<float4 *>(dst) + 1, hi);
        st_stream(reinterpret_cast<float4 *>(dst + W), lo);
        st_stream(reinterpret_cast<float4 *>(dst + W) + 1, hi);
    }
}

// forward into NHWC bf16 (pixel stride ld): one work item = 8 consecutive output channels of one pixel
__global__ void __launch_bounds__(256) dwtblock_nhwc_bf16(const float *__restrict__ x, int64_t N, int C, Ext e, int J,
                                                         int out_channels, float scale, const int *__restrict__ chmap,
                                                         __nv_bfloat16 *__restrict__ out, int64_t ld) {
    const int ho = e.h[J], wo = e.w[J];
    const int chunks = out_channels >> 3;
    const int64_t plane_i = (int64_t)e.h[0] * e.w[0];
    const int64_t items = N * ho * (int64_t)wo * chunks;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % chunks);
        int64_t pix = it / chunks;
        const int j = (int)(pix % wo);
        int64_t t = pix / wo;
        const int i = (int)(t % ho);
        const int64_t n = t / ho;
        float f[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int c = chmap ? __ldg(chmap + 8 * q + u) : (8 * q + u) % C;
            f[u] = ll_dyn(x + (n * C + c) * plane_i, e, J, i, j) * scale;
        }
        *reinterpret_cast<uint4 *>(out + pix * ld + 8 * q) = pack8(f);
    }
}

// J = 0 (channel tiling of an image that is already at the target resolution: how the Multi-ResNet decoder fetches its
// skip tensors from the cached Haar pyramid): grid (row-chunk blocks, image row, sample), no 64-bit div/mod per item.
__global__ void __launch_bounds__(256) dwtblock_nhwc_bf16_j0(const float *__restrict__ x, int C, int H, int W, int out_channels,
                                                            const int *__restrict__ chmap, __nv_bfloat16 *__restrict__ out,
                                                            int64_t ld) {
    pdl_trigger();
    pdl_wait();
    const int chunks = out_channels >> 3;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= W * chunks) return;
    const int q = idx % chunks, j = idx / chunks;
    const int i = blockIdx.y;
    const int64_t n = blockIdx.z;
    int c[8];
    if (chmap) {
        const int4 a = __ldg(reinterpret_cast<const int4 *>(chmap + 8 * q)), b = __ldg(reinterpret_cast<const int4 *>(chmap + 8 * q) + 1);
        c[0] = a.x; c[1] = a.y; c[2] = a.z; c[3] = a.w; c[4] = b.x; c[5] = b.y; c[6] = b.z; c[7] = b.w;
    } else {
#pragma unroll
        for (int u = 0; u < 8; ++u) c[u] = (8 * q + u) % C;
    }
    const int64_t plane = (int64_t)H * W;
    const float *src = x + n * C * plane + (int64_t)i * W + j;
    float f[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) f[u] = __ldg(src + c[u] * plane);
    *reinterpret_cast<uint4 *>(out + ((n * H + i) * (int64_t)W + j) * ld + 8 * q) = pack8(f);
}

// ---------------------------------------------------------------------------------------------
// DWTBlock on NHWC bf16 activations (pdearena / wmh: the head conv is learned, so the block sits inside the
// network and has a backward).  J in {0, 1}; fp32 arithmetic, bf16 storage.  One work item = one 16-byte channel
// chunk of one output (forward) / input (backward) pixel.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) dwtblock_nhwc_fwd(const __nv_bfloat16 *__restrict__ x, int64_t ld_x, int64_t N,
                                                        int H, int W, int C, int J, __nv_bfloat16 *__restrict__ out,
                                                        int64_t ld_o, int Cout) {
    const int ho = J ? (H + 1) >> 1 : H, wo = J ? (W + 1) >> 1 : W, chunks = C >> 3;
    const int64_t items = N * ho * (int64_t)wo * chunks;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % chunks);
        const int64_t pix = it / chunks;
        const int j = (int)(pix % wo);
        int64_t t = pix / wo;
        const int i = (int)(t % ho);
        const int64_t n = t / ho;
        float f[8];
        if (J == 0) {
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(x + pix * ld_x + 8 * q)), f);
        } else {
            float a[8], b[8], c[8], d[8];
            const bool r1 = 2 * i + 1 < H, c1 = 2 * j + 1 < W;
            const __nv_bfloat16 *p00 = x + ((n * H + 2 * i) * (int64_t)W + 2 * j) * ld_x + 8 * q;
            const uint4 z = make_uint4(0, 0, 0, 0);
            unpack8(ld_stream_u4(reinterpret_cast<const uint4 *>(p00)), a);
            unpack8(c1 ? ld_stream_u4(reinterpret_cast<const uint4 *>(p00 + ld_x)) : z, b);
            unpack8(r1 ? ld_stream_u4(reinterpret_cast<const uint4 *>(p00 + (int64_t)W * ld_x)) : z, c);
            unpack8((r1 && c1) ? ld_stream_u4(reinterpret_cast<const uint4 *>(p00 + (int64_t)W * ld_x + ld_x)) : z, d);
#pragma unroll
            for (int u = 0; u < 8; ++u) f[u] = analyse_ll(a[u], b[u], c[u], d[u]) * 0.5f;
        }
        const uint4 v = pack8(f);
        for (int k0 = 8 * q; k0 < Cout; k0 += C) *reinterpret_cast<uint4 *>(out + pix * ld_o + k0) = v;
    }
}

__global__ void __launch_bounds__(256) dwtblock_nhwc_bwd(const __nv_bfloat16 *__restrict__ g, int64_t ld_g, int64_t N,
                                                        int H, int W, int C, int J, int Cout,
                                                        __nv_bfloat16 *__restrict__ gx, int64_t ld_gx) {
    const int wo = J ? (W + 1) >> 1 : W, ho = J ? (H + 1) >> 1 : H, chunks = C >> 3;
    const int64_t items = N * H * (int64_t)W * chunks;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % chunks);
        const int64_t pix = it / chunks;
        const int xw = (int)(pix % W);
        int64_t t = pix / W;
        const int y = (int)(t % H);
        const int64_t n = t / H;
        const int64_t opix = (n * ho + (y >> J)) * wo + (xw >> J);
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
        for (int k0 = 8 * q; k0 < Cout; k0 += C) {
            float f[8];
            unpack8(*reinterpret_cast<const uint4 *>(g + opix * ld_g + k0), f);
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] += f[u];
        }
        if (J) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = mul_s(mul_s(acc[u])) * 0.5f;
        }
        *reinterpret_cast<uint4 *>(gx + pix * ld_gx + 8 * q) = pack8(acc);
    }
}

}  // namespace

extern "C" {

int ub200_dwtblock_nhwc_bf16_fwd(const void *x, int64_t ld_x, int64_t N, int64_t H, int64_t W, int64_t C, int J,
                                 void *out, int64_t ld_out, int64_t out_channels, void *stream) {
    UB_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C > 0 && out_channels > 0, UB200_E_BADARG);
    UB_REQUIRE((J == 0 || J == 1) && C % 8 == 0 && out_channels % 8 == 0 && ld_x % 8 == 0 && ld_out % 8 == 0 && ld_x >= C &&
                   ld_out >= out_channels && ub::aligned16(x) && ub::aligned16(out) && H < (1 << 29) && W < (1 << 29),
               UB200_E_UNSUPPORTED);
    const int64_t ho = J ? (H + 1) / 2 : H, wo = J ? (W + 1) / 2 : W;
    int grid = ub::grid_for(N * ho * wo * (C / 8), 256, 8);
    dwtblock_nhwc_fwd<<<grid, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(x), ld_x, N, (int)H, (int)W,
                                                              (int)C, J, reinterpret_cast<__nv_bfloat16 *>(out), ld_out,
                                                              (int)out_channels);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_dwtblock_nhwc_bf16_bwd(const void *gout, int64_t ld_g, int64_t N, int64_t H, int64_t W, int64_t C, int J,
                                 int64_t out_channels, void *gx, int64_t ld_gx, void *stream) {
    UB_REQUIRE(gout && gx && N > 0 && H > 0 && W > 0 && C > 0 && out_channels > 0, UB200_E_BADARG);
    UB_REQUIRE((J == 0 || J == 1) && C % 8 == 0 && out_channels % 8 == 0 && ld_g % 8 == 0 && ld_gx % 8 == 0 && ld_gx >= C &&
                   ld_g >= out_channels && ub::aligned16(gout) && ub::aligned16(gx) && H < (1 << 29) && W < (1 << 29),
               UB200_E_UNSUPPORTED);
    int grid = ub::grid_for(N * H * W * (C / 8), 256, 8);
    dwtblock_nhwc_bwd<<<grid, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(gout), ld_g, N, (int)H,
                                                              (int)W, (int)C, J, (int)out_channels,
                                                              reinterpret_cast<__nv_bfloat16 *>(gx), ld_gx);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_haar_dwt2d_fwd(const float *x, int64_t planes, int64_t H, int64_t W, float *ll, float *highs, void *stream) {
    UB_REQUIRE(x && ll && planes > 0 && H > 0 && W > 0, UB200_E_BADARG);
    UB_REQUIRE(H < (1 << 30) && W < (1 << 30), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const int64_t h2 = (H + 1) / 2, w2 = (W + 1) / 2;
    const bool fast = (W % 8 == 0) && ub::aligned16(x) && ub::aligned16(ll) && (!highs || ub::aligned16(highs));
    if (fast) {
        int grid = ub::grid_for(planes * h2 * (w2 / 4), 256, 8);
        if (highs) haar_dwt_vec4<true><<<grid, 256, 0, s>>>(x, planes, (int)H, (int)W, ll, highs);
        else haar_dwt_vec4<false><<<grid, 256, 0, s>>>(x, planes, (int)H, (int)W, ll, highs);
    } else {
        int grid = ub::grid_for(planes * h2 * w2, 256, 8);
        if (highs) haar_dwt_any<true><<<grid, 256, 0, s>>>(x, planes, (int)H, (int)W, ll, highs);
        else haar_dwt_any<false><<<grid, 256, 0, s>>>(x, planes, (int)H, (int)W, ll, highs);
    }
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_haar_idwt2d(const float *ll, const float *highs, int64_t planes, int64_t h2, int64_t w2, int64_t Hout,
                      int64_t Wout, float *out, void *stream) {
    UB_REQUIRE(ll && out && planes > 0 && h2 > 0 && w2 > 0, UB200_E_BADARG);
    UB_REQUIRE((Hout == 2 * h2 || Hout == 2 * h2 - 1) && (Wout == 2 * w2 || Wout == 2 * w2 - 1), UB200_E_BADARG);
    UB_REQUIRE(Hout < (1 << 30) && Wout < (1 << 30), UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const bool fast = (w2 % 4 == 0) && Wout == 2 * w2 && ub::aligned16(ll) && ub::aligned16(out) &&
                      (!highs || ub::aligned16(highs));
    if (fast) {
        int grid = ub::grid_for(planes * h2 * (w2 / 4), 256, 8);
        if (highs) haar_idwt_vec4<true><<<grid, 256, 0, s>>>(ll, highs, planes, (int)h2, (int)w2, (int)Hout, out);
        else haar_idwt_vec4<false><<<grid, 256, 0, s>>>(ll, highs, planes, (int)h2, (int)w2, (int)Hout, out);
    } else {
        int grid = ub::grid_for(planes * h2 * w2, 256, 8);
        if (highs) haar_idwt_any<true><<<grid, 256, 0, s>>>(ll, highs, planes, (int)h2, (int)w2, (int)Hout, (int)Wout, out);
        else haar_idwt_any<false><<<grid, 256, 0, s>>>(ll, highs, planes, (int)h2, (int)w2, (int)Hout, (int)Wout, out);
    }
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_dwtblock_fwd(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, int J, int64_t out_channels,
                       float *out, void *stream) {
    UB_REQUIRE(x && out && N > 0 && C > 0 && H > 0 && W > 0 && out_channels > 0, UB200_E_BADARG);
    UB_REQUIRE(J >= 0 && J <= 3 && H < (1 << 30) && W < (1 << 30) && C < (1 << 30) && out_channels < (1 << 30),
               UB200_E_UNSUPPORTED);
    cudaStream_t s = ub::as_stream(stream);
    const bool al = ub::aligned16(x) && ub::aligned16(out);
    if (J == 0 && al && (H * W) % 4 == 0) {
        int grid = ub::grid_for(N * C * (H * W / 4), 256, 8);
        tile_vec4<<<grid, 256, 0, s>>>(x, N, (int)C, H * W / 4, (int)out_channels, out);
    } else if (J == 1 && al && W % 8 == 0) {
        int grid = ub::grid_for(N * C * ((H + 1) / 2) * (W / 8), 256, 8);
        dwtblock_j1_vec4<<<grid, 256, 0, s>>>(x, N, (int)C, (int)H, (int)W, (int)out_channels, out);
    } else {
        Ext e = make_ext(H, W);
        int grid = ub::grid_for(N * C * (int64_t)e.h[J] * e.w[J], 256, 8);
        dwtblock_any<<<grid, 256, 0, s>>>(x, N, (int)C, e, J, (int)out_channels, 1.0f / (float)(1 << J), out);
    }
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_dwtblock_bwd(const float *gout, int64_t N, int64_t C, int64_t H, int64_t W, int J, int64_t out_channels,
                       float *gx, void *stream) {
    UB_REQUIRE(gout && gx && N > 0 && C > 0 && H > 0 && W > 0 && out_channels > 0, UB200_E_BADARG);
    UB_REQUIRE(J >= 0 && J <= 3 && H < (1 << 30) && W < (1 << 30) && C < (1 << 30) && out_channels < (1 << 30),
               UB200_E_UNSUPPORTED);
    if (J == 1 && W % 8 == 0 && H % 2 == 0 && ub::aligned16(gout) && ub::aligned16(gx)) {
        int grid = ub::grid_for(N * C * (H / 2) * (W / 8), 256, 8);
        dwtblock_bwd_j1_vec4<<<grid, 256, 0, ub::as_stream(stream)>>>(gout, N, (int)C, (int)H, (int)W, (int)out_channels, gx);
        UB_LAUNCH_CHECK();
        return UB200_OK;
    }
    Ext e = make_ext(H, W);
    int grid = ub::grid_for(N * C * H * W, 256, 8);
    dwtblock_bwd_any<<<grid, 256, 0, ub::as_stream(stream)>>>(gout, N, (int)C, e, J, (int)out_channels,
                                                             1.0f / (float)(1 << J), gx);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_dwtblock_fwd_nhwc_bf16(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, int J,
                                 int64_t out_channels, const int32_t *chmap, void *out_bf16, int64_t ld_out,
                                 void *stream) {
    UB_REQUIRE(x && out_bf16 && N > 0 && C > 0 && H > 0 && W > 0 && out_channels > 0, UB200_E_BADARG);
    UB_REQUIRE(J >= 0 && J <= 3 && out_channels % 8 == 0 && ld_out % 8 == 0 && ld_out >= out_channels &&
                   ub::aligned16(out_bf16) && H < (1 << 30) && W < (1 << 30),
               UB200_E_UNSUPPORTED);
    Ext e = make_ext(H, W);
    if (J == 0 && H <= 65535 && N <= 65535 && W * (out_channels / 8) < (1 << 30)) {
        const int per_row = (int)(W * (out_channels / 8));
        const int threads = per_row >= 256 ? 256 : ((per_row + 31) / 32) * 32;
        dim3 grid((unsigned)((per_row + threads - 1) / threads), (unsigned)H, (unsigned)N);
        cudaError_t le = ub::launch_pdl(dwtblock_nhwc_bf16_j0, grid, dim3(threads), 0, ub::as_stream(stream), x, (int)C, (int)H, (int)W,
                                        (int)out_channels, chmap, reinterpret_cast<__nv_bfloat16 *>(out_bf16), ld_out);
        if (le != cudaSuccess) return (int)le;
        UB_LAUNCH_CHECK();
        return UB200_OK;
    }
    int grid = ub::grid_for(N * (int64_t)e.h[J] * e.w[J] * (out_channels / 8), 256, 8);
    dwtblock_nhwc_bf16<<<grid, 256, 0, ub::as_stream(stream)>>>(x, N, (int)C, e, J, (int)out_channels,
                                                               1.0f / (float)(1 << J), chmap,
                                                               reinterpret_cast<__nv_bfloat16 *>(out_bf16), ld_out);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
