// Train-step tail as flat-arena kernels: gradient sum of squares, then global-norm clip + Adam + EMA in
// one pass, with no host synchronisation (the clip factor is computed on the device from sumsq[0]).
// Replaces torch.nn.utils.clip_grad_norm_ + torch.optim.Adam.step + ema() at diff_cifar/main.py:425-429,
// :57-77.  Memory-bound: reads p, g, m, v, ema and writes p, m, v, ema once (36 bytes per parameter).
#include "common.cuh"

namespace {
using namespace ub;

__global__ void __launch_bounds__(256) sumsq_kernel(const float *__restrict__ g, int64_t n, float *__restrict__ out) {
    float acc = 0.f;
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = ld_stream(reinterpret_cast<const float4 *>(g) + i);
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) { const float t = g[(n4 << 2) + threadIdx.x]; acc += t * t; }
    __shared__ float part[8];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x < 8) {
        float t = part[threadIdx.x];
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffu, t, o);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

struct AdamArgs {
    float max_norm, grad_scale, lr, beta1, beta2, eps, ema_decay, bc1, bc2;
    float weight_decay;             // decoupled (torch.optim.AdamW): p *= 1 - lr * weight_decay before the Adam update
    long long warmup;               // > 0: lr *= min(step - 1, warmup) / warmup   (LambdaLR of diff_cifar/main.py:90-91)
    const long long *step_dev;      // device-resident 1-based step (CUDA-graph replays); NULL = host value baked in
};

// bf16 "shadow" of the parameter arena, written by the optimiser: conv weights keep the [Cout,kh,kw,Cin] order in
// the arena, so their shadow slice IS the packed fprop / wgrad GEMM operand -- no per-layer re-packing.
__device__ __forceinline__ void store_shadow4(__nv_bfloat16 *shadow, int64_t i4, const float4 &p) {
    if (!shadow) return;
    uint2 v = make_uint2(pack_bf16(p.x, p.y), pack_bf16(p.z, p.w));
    *reinterpret_cast<uint2 *>(shadow + 4 * i4) = v;
}

__device__ __forceinline__ void adam1(float &p, float g, float &m, float &v, float *ema, const AdamArgs &a, float clip) {
    g *= clip;
    m = a.beta1 * m + (1.f - a.beta1) * g;
    v = a.beta2 * v + (1.f - a.beta2) * g * g;
    // torch.optim.Adam: p -= lr / bc1 * m / (sqrt(v) / sqrt(bc2) + eps); AdamW decays p first
    p *= 1.f - a.lr * a.weight_decay;
    p -= (a.lr / a.bc1) * m / (sqrtf(v) / a.bc2 + a.eps);
    if (ema) *ema = a.ema_decay * *ema + (1.f - a.ema_decay) * p;
}

__global__ void __launch_bounds__(256) adam_ema_kernel(float *__restrict__ p, const float *__restrict__ g,
                                                      float *__restrict__ m, float *__restrict__ v,
                                                      float *__restrict__ ema, int64_t n,
                                                      const float *__restrict__ sumsq, AdamArgs a,
                                                      __nv_bfloat16 *__restrict__ shadow) {
    if (a.step_dev) {               // bias corrections and warm-up from the device-side step counter
        const float st = (float)__ldg(a.step_dev);
        a.bc1 = 1.f - powf(a.beta1, st);
        a.bc2 = sqrtf(1.f - powf(a.beta2, st));
        if (a.warmup > 0) a.lr *= fminf(st - 1.f, (float)a.warmup) / (float)a.warmup;
    }
    // clip_grad_norm_: coef = max_norm / (total_norm + 1e-6), clamped to 1; grads arrive scaled by 1/grad_scale
    float clip = a.grad_scale;
    if (sumsq && a.max_norm > 0.f) {
        const float norm = sqrtf(__ldg(sumsq)) * a.grad_scale;
        clip *= fminf(1.f, a.max_norm / (norm + 1e-6f));
    }
    const int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4 *>(p)[i], mm = reinterpret_cast<float4 *>(m)[i], vv = reinterpret_cast<float4 *>(v)[i];
        const float4 gg = ld_stream(reinterpret_cast<const float4 *>(g) + i);
        float4 ee = make_float4(0.f, 0.f, 0.f, 0.f);
        if (ema) ee = reinterpret_cast<float4 *>(ema)[i];
        adam1(pp.x, gg.x, mm.x, vv.x, ema ? &ee.x : nullptr, a, clip);
        adam1(pp.y, gg.y, mm.y, vv.y, ema ? &ee.y : nullptr, a, clip);
        adam1(pp.z, gg.z, mm.z, vv.z, ema ? &ee.z : nullptr, a, clip);
        adam1(pp.w, gg.w, mm.w, vv.w, ema ? &ee.w : nullptr, a, clip);
        reinterpret_cast<float4 *>(p)[i] = pp; reinterpret_cast<float4 *>(m)[i] = mm; reinterpret_cast<float4 *>(v)[i] = vv;
        if (ema) reinterpret_cast<float4 *>(ema)[i] = ee;
        store_shadow4(shadow, i, pp);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (n4 << 2) + threadIdx.x;
        adam1(p[i], g[i], m[i], v[i], ema ? ema + i : nullptr, a, clip);
        if (shadow) shadow[i] = __float2bfloat16_rn(p[i]);
    }
}

// dgrad operands of ALL conv layers in one launch: for conv j (blockIdx.y) with table row
// (src offset in the shadow arena, dst offset, Cout, Cin, k):
//   dst[ci][ky][kx][co] = src[co][k-1-ky][k-1-kx][ci]      (rows ci padded to a multiple of 16 with zeros)
__global__ void __launch_bounds__(256) pack_dgrad_batched_kernel(const __nv_bfloat16 *__restrict__ shadow,
                                                                __nv_bfloat16 *__restrict__ dst_arena,
                                                                const long long *__restrict__ table) {
    const long long *row = table + 5 * blockIdx.y;
    const __nv_bfloat16 *src = shadow + row[0];
    __nv_bfloat16 *dst = dst_arena + row[1];
    const int Cout = (int)row[2], Cin = (int)row[3], k = (int)row[4];
    const int rows_pad = (Cin + 15) / 16 * 16;
    if (Cin % 8 == 0 && Cout % 8 == 0 && ((row[0] | row[1]) & 7) == 0) {
        // Tiled transpose, one filter tap at a time: a [64 co] x [64 ci] tile is read as 128-byte rows of the source, turned in
        // shared memory, and written as 128-byte rows of the destination (16 bytes per thread on both sides; all index
        // arithmetic is per tile).  The element-wise loop below needs ~100 instructions per bf16 and was issue-bound at ~150 us
        // for 53 MB, on a second stream next to the forward pass whose GroupNorm kernels are issue-bound themselves.
        __shared__ uint32_t tile[64][33];                     // [co][ci pair], one pad word per row
        const int kk = k * k, co_tiles = (Cout + 63) / 64, ci_tiles = (rows_pad + 63) / 64;
        const int ntiles = kk * co_tiles * ci_tiles;
        const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int cit = t % ci_tiles, rest = t / ci_tiles, cot = rest % co_tiles, tap = rest / co_tiles;
            const int ky = tap / k, kx = tap - ky * k;
            const int src_tap = (k - 1 - ky) * k + (k - 1 - kx);
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int col = pass * 32 + r, co = cot * 64 + col, ci0 = cit * 64 + q * 8;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (co < Cout && ci0 < Cin)
                    v = *reinterpret_cast<const uint4 *>(src + ((long long)co * kk + src_tap) * Cin + ci0);
                uint32_t *d = &tile[col][q * 4];
                d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
            }
            __syncthreads();
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const int cil = pass * 32 + r, ci = cit * 64 + cil, co0 = cot * 64 + q * 8;
                if (ci < rows_pad && co0 < Cout) {
                    const int w = cil >> 1, sh = (cil & 1) * 16;
                    uint32_t h[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) h[j] = (tile[q * 8 + j][w] >> sh) & 0xffffu;
                    const uint4 o = make_uint4(h[0] | (h[1] << 16), h[2] | (h[3] << 16), h[4] | (h[5] << 16), h[6] | (h[7] << 16));
                    *reinterpret_cast<uint4 *>(dst + ((long long)ci * kk + tap) * Cout + co0) = o;
                }
            }
            __syncthreads();
        }
        return;
    }
    const long long total = (long long)rows_pad * k * k * Cout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        long long t = i / Cout;
        const int kx = (int)(t % k); t /= k;
        const int ky = (int)(t % k);
        const int ci = (int)(t / k);
        __nv_bfloat16 v = __float2bfloat16_rn(0.f);
        if (ci < Cin) v = src[(((long long)co * k + (k - 1 - ky)) * k + (k - 1 - kx)) * Cin + ci];
        dst[i] = v;
    }
}

}  // namespace

extern "C" {

int ub200_sumsq_f32(const float *g, int64_t n, float *sumsq, void *stream) {
    UB_REQUIRE(g && sumsq && n > 0, UB200_E_BADARG);
    UB_REQUIRE(ub::aligned16(g), UB200_E_UNSUPPORTED);
    sumsq_kernel<<<ub::grid_for((n + 3) / 4, 256, 8, 2), 256, 0, ub::as_stream(stream)>>>(g, n, sumsq);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_adam_ema_step_f32(float *p, const float *g, float *m, float *v, float *ema, int64_t n, const float *sumsq,
                            float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                            float ema_decay, int64_t step_host, int64_t warmup_steps, const int64_t *step_dev,
                            void *shadow_bf16, void *stream) {
    return ub200_adamw_ema_step_f32(p, g, m, v, ema, n, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, 0.f, ema_decay,
                                    step_host, warmup_steps, step_dev, shadow_bf16, stream);
}

int ub200_adamw_ema_step_f32(float *p, const float *g, float *m, float *v, float *ema, int64_t n, const float *sumsq,
                             float max_norm, float grad_scale, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float ema_decay, int64_t step_host, int64_t warmup_steps,
                             const int64_t *step_dev, void *shadow_bf16, void *stream) {
    UB_REQUIRE(p && g && m && v && n > 0 && (step_dev || step_host >= 1), UB200_E_BADARG);
    UB_REQUIRE(ub::aligned16(p) && ub::aligned16(g) && ub::aligned16(m) && ub::aligned16(v) && (!ema || ub::aligned16(ema)),
               UB200_E_UNSUPPORTED);
    AdamArgs a{max_norm, grad_scale, lr, beta1, beta2, eps, ema_decay, 1.f, 1.f, weight_decay, (long long)warmup_steps,
               reinterpret_cast<const long long *>(step_dev)};
    if (!step_dev) {
        a.bc1 = (float)(1.0 - pow((double)beta1, (double)step_host));
        a.bc2 = (float)sqrt(1.0 - pow((double)beta2, (double)step_host));
        if (warmup_steps > 0) {
            const double s1 = (double)(step_host - 1);
            a.lr *= (float)((s1 < (double)warmup_steps ? s1 : (double)warmup_steps) / (double)warmup_steps);
        }
    }
    UB_REQUIRE(!shadow_bf16 || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7u) == 0, UB200_E_UNSUPPORTED);
    adam_ema_kernel<<<ub::grid_for((n + 3) / 4, 256, 8, 2), 256, 0, ub::as_stream(stream)>>>(
        p, g, m, v, ema, n, sumsq, a, reinterpret_cast<__nv_bfloat16 *>(shadow_bf16));
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_pack_dgrad_weights_batched(const void *shadow_bf16, void *dgrad_arena_bf16, const int64_t *table_dev,
                                     int n_convs, void *stream) {
    UB_REQUIRE(shadow_bf16 && dgrad_arena_bf16 && table_dev && n_convs > 0 && n_convs <= 65535, UB200_E_BADARG);
    dim3 grid(64, (unsigned)n_convs, 1);
    pack_dgrad_batched_kernel<<<grid, 256, 0, ub::as_stream(stream)>>>(
        reinterpret_cast<const __nv_bfloat16 *>(shadow_bf16), reinterpret_cast<__nv_bfloat16 *>(dgrad_arena_bf16),
        reinterpret_cast<const long long *>(table_dev));
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
