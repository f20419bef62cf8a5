// Layout converters (NCHW f32 <-> NHWC bf16) and nearest x2 up-sampling on NHWC bf16.
// Memory-bound; 128-bit accesses where the channel count allows.
// Reference ops replaced: F.interpolate(scale_factor=2, mode='nearest') (diff_cifar/model.py:78-79;
// diff_mnist/torch_ddpm/ddpm/models/unet/layers.py:219; pdearena twod_unetbase.py:243).
#include "common.cuh"

namespace {
using namespace ub;

// tile of 32 channels x 32 pixels through shared memory
__global__ void __launch_bounds__(256) nchw_to_nhwc(const float *__restrict__ x, int C, int64_t HW,
                                                   __nv_bfloat16 *__restrict__ out, int64_t ld) {
    __shared__ float tile[32][33];
    const int64_t n = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int64_t p0 = (int64_t)blockIdx.x * 32; p0 < HW; p0 += (int64_t)gridDim.x * 32) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r;
            const int64_t p = p0 + tx;
            tile[r][tx] = (c < C && p < HW) ? __ldg(x + (n * C + c) * HW + p) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t p = p0 + r;
            const int c = c0 + tx;
            if (p < HW && c < C) out[(n * HW + p) * ld + c] = __float2bfloat16_rn(tile[tx][r]);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) nhwc_to_nchw(const __nv_bfloat16 *__restrict__ x, int64_t ld, int C, int64_t HW,
                                                   float *__restrict__ out) {
    __shared__ float tile[32][33];
    const int64_t n = blockIdx.z;
    const int c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int64_t p0 = (int64_t)blockIdx.x * 32; p0 < HW; p0 += (int64_t)gridDim.x * 32) {
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int64_t p = p0 + r;
            const int c = c0 + tx;
            tile[r][tx] = (p < HW && c < C) ? __bfloat162float(x[(n * HW + p) * ld + c]) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int r = ty; r < 32; r += 8) {
            const int c = c0 + r;
            const int64_t p = p0 + tx;
            if (c < C && p < HW) out[(n * C + c) * HW + p] = tile[tx][r];
        }
        __syncthreads();
    }
}

// one work item = one 16-byte channel chunk of one INPUT pixel -> four output pixels
__global__ void __launch_bounds__(256) upsample2x(const uint4 *__restrict__ x, int64_t ld_in8, int64_t N, int H, int W,
                                                 int chunks, uint4 *__restrict__ out, int64_t ld_out8) {
    const int64_t items = N * H * (int64_t)W * chunks;
    const int W2 = 2 * W;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % chunks);
        const int64_t pix = it / chunks;
        const int xw = (int)(pix % W);
        const int64_t t = pix / W;
        const int y = (int)(t % H);
        const int64_t n = t / H;
        const uint4 v = ld_stream_u4(x + pix * ld_in8 + q);
        const int64_t o = ((n * 2 * H + 2 * y) * W2 + 2 * xw) * ld_out8 + q;
        st_stream_u4(out + o, v);
        st_stream_u4(out + o + ld_out8, v);
        st_stream_u4(out + o + (int64_t)W2 * ld_out8, v);
        st_stream_u4(out + o + (int64_t)W2 * ld_out8 + ld_out8, v);
    }
}

__global__ void __launch_bounds__(256) upsample2x_bwd(const uint4 *__restrict__ g, int64_t ld_g8, int64_t N, int H, int W,
                                                     int chunks, uint4 *__restrict__ gx, int64_t ld_gx8) {
    const int64_t items = N * H * (int64_t)W * chunks;
    const int W2 = 2 * W;
    for (int64_t it = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; it < items; it += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(it % chunks);
        const int64_t pix = it / chunks;
        const int xw = (int)(pix % W);
        const int64_t t = pix / W;
        const int y = (int)(t % H);
        const int64_t n = t / H;
        const int64_t o = ((n * 2 * H + 2 * y) * W2 + 2 * xw) * ld_g8 + q;
        float a[8], b[8], c[8], d[8];
        unpack8(ld_stream_u4(g + o), a);
        unpack8(ld_stream_u4(g + o + ld_g8), b);
        unpack8(ld_stream_u4(g + o + (int64_t)W2 * ld_g8), c);
        unpack8(ld_stream_u4(g + o + (int64_t)W2 * ld_g8 + ld_g8), d);
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = (a[u] + b[u]) + (c[u] + d[u]);
        gx[pix * ld_gx8 + q] = pack8(a);
    }
}

}  // namespace

extern "C" {

int ub200_nchw_f32_to_nhwc_bf16(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, void *out_bf16, int64_t ld,
                                void *stream) {
    UB_REQUIRE(x && out_bf16 && N > 0 && C > 0 && H > 0 && W > 0 && ld >= C, UB200_E_BADARG);
    UB_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, UB200_E_UNSUPPORTED);
    const int64_t HW = H * W;
    int64_t gx = (HW + 31) / 32;
    if (gx > 148 * 16) gx = 148 * 16;
    dim3 grid((unsigned)gx, (unsigned)((C + 31) / 32), (unsigned)N);
    nchw_to_nhwc<<<grid, 256, 0, ub::as_stream(stream)>>>(x, (int)C, HW, reinterpret_cast<__nv_bfloat16 *>(out_bf16), ld);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_nhwc_bf16_to_nchw_f32(const void *x_bf16, int64_t ld, int64_t N, int64_t C, int64_t H, int64_t W, float *out,
                                void *stream) {
    UB_REQUIRE(x_bf16 && out && N > 0 && C > 0 && H > 0 && W > 0 && ld >= C, UB200_E_BADARG);
    UB_REQUIRE(N <= 65535 && (C + 31) / 32 <= 65535, UB200_E_UNSUPPORTED);
    const int64_t HW = H * W;
    int64_t gx = (HW + 31) / 32;
    if (gx > 148 * 16) gx = 148 * 16;
    dim3 grid((unsigned)gx, (unsigned)((C + 31) / 32), (unsigned)N);
    nhwc_to_nchw<<<grid, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<const __nv_bfloat16 *>(x_bf16), ld, (int)C, HW, out);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_upsample2x_nhwc_bf16(const void *x, int64_t ld_in, int64_t N, int64_t H, int64_t W, int64_t C, void *out,
                               int64_t ld_out, void *stream) {
    UB_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && C > 0, UB200_E_BADARG);
    UB_REQUIRE(C % 8 == 0 && ld_in % 8 == 0 && ld_out % 8 == 0 && ld_in >= C && ld_out >= C && ub::aligned16(x) &&
                   ub::aligned16(out) && H < (1 << 29) && W < (1 << 29),
               UB200_E_UNSUPPORTED);
    int grid = ub::grid_for(N * H * W * (C / 8), 256, 8);
    upsample2x<<<grid, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<const uint4 *>(x), ld_in / 8, N, (int)H, (int)W,
                                                       (int)(C / 8), reinterpret_cast<uint4 *>(out), ld_out / 8);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

int ub200_upsample2x_bwd_nhwc_bf16(const void *gout, int64_t ld_g, int64_t N, int64_t H, int64_t W, int64_t C, void *gx,
                                   int64_t ld_gx, void *stream) {
    UB_REQUIRE(gout && gx && N > 0 && H > 0 && W > 0 && C > 0, UB200_E_BADARG);
    UB_REQUIRE(C % 8 == 0 && ld_g % 8 == 0 && ld_gx % 8 == 0 && ld_g >= C && ld_gx >= C && ub::aligned16(gout) &&
                   ub::aligned16(gx) && H < (1 << 29) && W < (1 << 29),
               UB200_E_UNSUPPORTED);
    int grid = ub::grid_for(N * H * W * (C / 8), 256, 8);
    upsample2x_bwd<<<grid, 256, 0, ub::as_stream(stream)>>>(reinterpret_cast<const uint4 *>(gout), ld_g / 8, N, (int)H,
                                                           (int)W, (int)(C / 8), reinterpret_cast<uint4 *>(gx), ld_gx / 8);
    UB_LAUNCH_CHECK();
    return UB200_OK;
}

}  // extern "C"
