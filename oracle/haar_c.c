/* TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned.
 *
 * Plain-C restatement of the 2-D Haar analysis / synthesis the reference reaches
 * through pytorch_wavelets (call sites: diff_cifar/model.py:263-267, :310-321;
 * diff_cifar/diffusion.py:63-70; pdearena/pdearena/modules/twod_unetbase.py:169-193;
 * wmh/model.py:68-95).  Same arithmetic as oracle/haar_np.py: separable, W axis
 * first, taps s = (float)(1/sqrt 2), one zero appended at the end of an odd axis.
 * Built by oracle/Makefile into oracle/_ref/liboracle_haar.so; used by the tests as
 * a checker and by bench.py as the CPU baseline of the DWT sweep.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

static inline float at(const float *p, int64_t H, int64_t W, int64_t y, int64_t x) {
    return (y < H && x < W) ? p[y * W + x] : 0.0f;
}

/* One level over `planes` = N*C contiguous HxW planes.  Any of lh/hl/hh may be NULL
 * (LL-only).  Outputs are [planes, ceil(H/2), ceil(W/2)]. */
void oracle_haar_dwt2_level(const float *x, int64_t planes, int64_t H, int64_t W,
                            float *ll, float *lh, float *hl, float *hh) {
    const float s = (float)(1.0 / sqrt(2.0));
    const int64_t h2 = (H + 1) / 2, w2 = (W + 1) / 2;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < planes; ++p) {
        const float *src = x + p * H * W;
        for (int64_t i = 0; i < h2; ++i)
            for (int64_t j = 0; j < w2; ++j) {
                float a = at(src, H, W, 2 * i, 2 * j), b = at(src, H, W, 2 * i, 2 * j + 1);
                float c = at(src, H, W, 2 * i + 1, 2 * j), d = at(src, H, W, 2 * i + 1, 2 * j + 1);
                float lo_t = s * a + s * b, hi_t = s * a - s * b;
                float lo_b = s * c + s * d, hi_b = s * c - s * d;
                int64_t o = (p * h2 + i) * w2 + j;
                ll[o] = s * lo_t + s * lo_b;
                if (lh) lh[o] = s * lo_t - s * lo_b;
                if (hl) hl[o] = s * hi_t + s * hi_b;
                if (hh) hh[o] = s * hi_t - s * hi_b;
            }
    }
}

/* One synthesis level: bands [planes, h2, w2] -> out [planes, 2*h2, 2*w2]. */
void oracle_haar_idwt2_level(const float *ll, const float *lh, const float *hl, const float *hh,
                             int64_t planes, int64_t h2, int64_t w2, float *out) {
    const float s = (float)(1.0 / sqrt(2.0));
    const int64_t W = 2 * w2;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < planes; ++p)
        for (int64_t i = 0; i < h2; ++i)
            for (int64_t j = 0; j < w2; ++j) {
                int64_t o = (p * h2 + i) * w2 + j;
                float vll = ll[o], vlh = lh ? lh[o] : 0.f, vhl = hl ? hl[o] : 0.f, vhh = hh ? hh[o] : 0.f;
                float lo_t = s * vll + s * vlh, lo_b = s * vll - s * vlh;
                float hi_t = s * vhl + s * vhh, hi_b = s * vhl - s * vhh;
                float *dst = out + p * 4 * h2 * w2;
                dst[(2 * i) * W + 2 * j] = s * lo_t + s * hi_t;
                dst[(2 * i) * W + 2 * j + 1] = s * lo_t - s * hi_t;
                dst[(2 * i + 1) * W + 2 * j] = s * lo_b + s * hi_b;
                dst[(2 * i + 1) * W + 2 * j + 1] = s * lo_b - s * hi_b;
            }
}

/* DTWBlock forward: LL_J / 2^J then channel tile (out[:, k] = y[:, k % C]).
 * scratch must hold 2 * N*C*ceil(H/2)*ceil(W/2) floats when J > 1 (ping-pong). */
void oracle_dwtblock_fwd(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, int J,
                         int64_t out_channels, float *out, float *scratch) {
    const float *cur = x;
    int64_t h = H, w = W;
    float *buf[2] = {scratch, scratch + N * C * ((H + 1) / 2) * ((W + 1) / 2)};
    for (int j = 0; j < J; ++j) {
        float *dst = buf[j & 1];
        oracle_haar_dwt2_level(cur, N * C, h, w, dst, NULL, NULL, NULL);
        cur = dst;
        h = (h + 1) / 2;
        w = (w + 1) / 2;
    }
    const float inv = 1.0f / (float)(1 << J);
    const int64_t plane = h * w;
#pragma omp parallel for collapse(2) schedule(static)
    for (int64_t n = 0; n < N; ++n)
        for (int64_t k = 0; k < out_channels; ++k) {
            const float *src = cur + (n * C + k % C) * plane;
            float *dst = out + (n * out_channels + k) * plane;
            if (J == 0) memcpy(dst, src, (size_t)plane * sizeof(float));
            else for (int64_t i = 0; i < plane; ++i) dst[i] = src[i] * inv; /* x / 2^J, exact */
        }
}
