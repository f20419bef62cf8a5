/*
 * unet_b200.h -- C ABI of the B200-native (sm_100a) U-Net hot path.
 *
 * The reference (FabianFalck/unet-design) is 100% Python and has no FFI of its own; its
 * "plugin API" for this path is the nn.Module contract of its model classes.
 * This header is the boundary a maintainer of the reference would bind instead of the
 * ATen / pytorch_wavelets calls listed beside each entry point (file:line under the
 * reference root).  INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _host;
 *   - the caller owns all memory (inputs, outputs, workspaces); nothing is allocated here,
 *     so a caching allocator and CUDA-graph capture keep working;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), re-entrant,
 *     and keeps no mutable global state;
 *   - return value: 0 = ok, <0 = UB200_E_* (bad argument / unsupported shape / no device
 *     code for this GPU), >0 = a cudaError_t from the launch.  Nothing throws or exits.
 *   - "NCHW f32"  = contiguous float  [N, C, H, W]   (the reference's layout and dtype)
 *     "NHWC bf16" = __nv_bfloat16 [N, H, W, C] with an explicit pixel stride `ld` (elements
 *                   between consecutive pixels, >= C): a tensor may be a channel slice of a
 *                   wider buffer, which is how torch.cat([h, skip], 1) disappears.
 */
#ifndef UNET_B200_H
#define UNET_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UB200_OK 0
#define UB200_E_BADARG (-1)       /* null pointer, non-positive extent, misaligned pointer */
#define UB200_E_UNSUPPORTED (-2)  /* shape outside what the kernel family covers */
#define UB200_E_NODEVICE (-3)     /* not an sm_100 device */

/* Library identification; safe to call without a GPU. */
const char *ub200_version(void);
int ub200_abi_version(void);
/* Name of a status code returned by any entry point ("ok", "bad argument", cudaGetErrorName...). */
const char *ub200_status_string(int status);

/* ------------------------------------------------------------------------------------------
 * Haar wavelet (replaces pytorch_wavelets.DWTForward / DWTInverse, mode='zero', wave='haar';
 * reference call sites diff_cifar/model.py:263-267,:310-311; diff_cifar/diffusion.py:63-66;
 * pdearena/pdearena/modules/twod_unetbase.py:169-170,:179-180; wmh/model.py:68-69,:81-82).
 * All NCHW f32.  planes = N*C.  h2 = ceil(H/2), w2 = ceil(W/2); an odd extent is extended by
 * one zero at its end.
 * ------------------------------------------------------------------------------------------ */

/* One analysis level.  x [planes,H,W] -> ll [planes,h2,w2], highs [planes,3,h2,w2] in the order
 * LH (W-low,H-high), HL (W-high,H-low), HH.  highs may be NULL (LL only).  Also the adjoint of
 * ub200_haar_idwt2d (backward of the synthesis). */
int ub200_haar_dwt2d_fwd(const float *x, int64_t planes, int64_t H, int64_t W,
                         float *ll, float *highs, void *stream);

/* One synthesis level.  ll [planes,h2,w2], highs [planes,3,h2,w2] (NULL = zeros) ->
 * out [planes,Hout,Wout] with Hout in {2*h2-1, 2*h2}, Wout likewise (the crop undoes the zero
 * extension).  Also the adjoint of ub200_haar_dwt2d_fwd (its backward). */
int ub200_haar_idwt2d(const float *ll, const float *highs, int64_t planes, int64_t h2, int64_t w2,
                      int64_t Hout, int64_t Wout, float *out, void *stream);

/* J analysis levels in ONE pass over x (DWTForward(J=2|3): the call sites above construct it with J > 1 for the
 * multi-resolution targets, diff_cifar/diffusion.py:63-64): x [planes,H,W] -> ll [planes,H/2^J,W/2^J] and
 * highs[j] [planes,3,H/2^(j+1),W/2^(j+1)], j = 0 the finest (`highs` is a HOST array of J device pointers).
 * Algorithmic traffic 2*E*4 bytes instead of E*4*sum_j 2*4^-j level by level.  Supported when
 * H % 2^J == 0, W % 8 == 0 and all pointers are 16-byte aligned; returns UB200_E_UNSUPPORTED otherwise
 * (the caller then goes level by level through ub200_haar_dwt2d_fwd; results are bit-identical). */
int ub200_haar_dwt2d_multi_fwd(const float *x, int64_t planes, int64_t H, int64_t W, int J,
                               float *ll, float *const *highs, void *stream);
/* Its inverse (and adjoint): ll + highs[0..J) -> out [planes,H,W], same support conditions. */
int ub200_haar_idwt2d_multi(const float *ll, const float *const *highs, int64_t planes, int64_t H, int64_t W,
                            int J, float *out, void *stream);

/* Multi-resolution loss of the staged / multi-res training arms (diff_cifar/diffusion.py:52-91, diff_mnist/main.py:375-403,
 * pdearena pdemodel.py:222-229), fused: with t_k = LL_k(noise) / 2^k, k = 0..J (J in 1..3; H % 2^J == 0, W % 8 == 0),
 *     sums[k] += sum (outs[k] - t_k)^2            (fp32; the caller zeroes sums and divides by numel_k for the means)
 *     grads[k] = (2 / numel_k) * (outs[k] - t_k)  (gradient of sum_k mean-squared-error; grads or grads[k] may be NULL)
 * in ONE pass over the noise: the target pyramid never touches memory.  outs / grads: HOST arrays of J+1 device pointers,
 * level 0 = full resolution [planes, H, W], level k = [planes, H/2^k, W/2^k].  UB200_E_UNSUPPORTED for other shapes
 * (the caller then builds the targets level by level with ub200_dwtblock_fwd). */
int ub200_multires_mse_f32(const float *noise, int64_t planes, int64_t H, int64_t W, int J,
                           const float *const *outs, float *const *grads, float *sums, void *stream);

/* DTWBlock / DWTBlock forward (diff_cifar/model.py:270-323; diff_mnist/mnist_diff/models.py:29-82;
 * twod_unetbase.py:173-193; wmh/model.py:72-95), fused:  out[n,k] = LL_J(x[n, k mod C]) / 2^J,
 * k < out_channels, J in [0,3] (J = 0 is the channel tile alone).  x [N,C,H,W] f32 ->
 * out [N,out_channels,ceil(H/2^J),ceil(W/2^J)] f32. */
int ub200_dwtblock_fwd(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, int J,
                       int64_t out_channels, float *out, void *stream);

/* Adjoint of ub200_dwtblock_fwd: gx[n,c] = LL_J^T( sum_{k mod C == c} gout[n,k] ) / 2^J. */
int ub200_dwtblock_bwd(const float *gout, int64_t N, int64_t C, int64_t H, int64_t W, int J,
                       int64_t out_channels, float *gx, void *stream);

/* Same forward, written straight into the conv blocks' layout: out NHWC bf16 with pixel stride
 * ld_out (the skip half of a torch.cat([h, skip], 1) buffer; diff_cifar/model.py:454).
 * chmap (device int32 [out_channels], nullable) gives the source channel of every output channel;
 * NULL means k mod C.  A chain of DTWBlocks composes into one map, e.g. ((k mod 256) mod 128) mod 3,
 * so the whole Haar encoder reads only the 3-channel image pyramid.  out_channels % 8 == 0. */
int ub200_dwtblock_fwd_nhwc_bf16(const float *x, int64_t N, int64_t C, int64_t H, int64_t W, int J,
                                 int64_t out_channels, const int32_t *chmap, void *out_bf16,
                                 int64_t ld_out, void *stream);

/* DWTBlock inside a network whose head is learned (pdearena twod_unetbase.py:173-193,:203;
 * wmh/model.py:72-95): NHWC bf16 in and out, J in {0,1}, fp32 arithmetic.
 * fwd: out[n,i,j,k] = LL_J(x[n,:,:,k mod C])[i,j] / 2^J;  bwd: its adjoint.  H, W are the INPUT extents,
 * the output is ceil(H/2^J) x ceil(W/2^J) (zero extension of an odd extent).  C, out_channels % 8 == 0. */
int ub200_dwtblock_nhwc_bf16_fwd(const void *x, int64_t ld_x, int64_t N, int64_t H, int64_t W, int64_t C, int J,
                                 void *out, int64_t ld_out, int64_t out_channels, void *stream);
int ub200_dwtblock_nhwc_bf16_bwd(const void *gout, int64_t ld_g, int64_t N, int64_t H, int64_t W, int64_t C,
                                 int J, int64_t out_channels, void *gx, int64_t ld_gx, void *stream);

/* ------------------------------------------------------------------------------------------
 * Layout / resampling (memory-bound).
 * ------------------------------------------------------------------------------------------ */

/* NCHW f32 [N,C,H,W] -> NHWC bf16 (pixel stride ld) and back (boundary of the drop-in modules). */
int ub200_nchw_f32_to_nhwc_bf16(const float *x, int64_t N, int64_t C, int64_t H, int64_t W,
                                void *out_bf16, int64_t ld, void *stream);
int ub200_nhwc_bf16_to_nchw_f32(const void *x_bf16, int64_t ld, int64_t N, int64_t C, int64_t H,
                                int64_t W, float *out, void *stream);

/* F.interpolate(scale_factor=2, mode='nearest') (diff_cifar/model.py:78-79; layers.py:219;
 * twod_unetbase.py:243) on NHWC bf16: out[n,y,x,:] = in[n,y>>1,x>>1,:];  and its adjoint
 * (sum of the 2x2 block, fp32 accumulate).  H, W are the extents of the LOW-resolution tensor
 * (x of the forward, gx of the backward); the other side is [N,2H,2W,C].  C % 8 == 0. */
int ub200_upsample2x_nhwc_bf16(const void *x, int64_t ld_in, int64_t N, int64_t H, int64_t W, int64_t C,
                               void *out, int64_t ld_out, void *stream);
int ub200_upsample2x_bwd_nhwc_bf16(const void *gout, int64_t ld_g, int64_t N, int64_t H, int64_t W,
                                   int64_t C, void *gx, int64_t ld_gx, void *stream);

/* ------------------------------------------------------------------------------------------
 * GroupNorm + activation (replaces nn.GroupNorm + Swish/SiLU/GELU [+ Dropout];
 * diff_cifar/model.py:130-132,:139-142; layers.py:284-288,:330-334; twod_unetbase.py:30-31).
 * NHWC bf16 activations, fp32 statistics and affine parameters (gamma/beta may be NULL = 1/0),
 * biased variance.  C % 8 == 0, C <= 2048.
 * ------------------------------------------------------------------------------------------ */
#define UB200_ACT_NONE 0
#define UB200_ACT_SILU 1
#define UB200_ACT_GELU 2   /* exact erf form (pdearena/pdearena/modules/activations.py:3-9) */
#define UB200_ACT_RELU 3   /* nn.ReLU of the same registry (the `Unetbase` docstring lists gelu / relu / silu) */

/* stats[n,g] = (sum, sum of squares) over the (C/G)*HW slab of sample n; float [N,G,2], zeroed
 * here.  Consumers derive mean / rstd; the conv epilogue can accumulate the same layout
 * (ub200_conv_args.gn_partial), which removes this pass. */
int ub200_gn_stats_nhwc_bf16(const void *x, int64_t ld, int64_t N, int64_t HW, int64_t C, int G,
                             float *stats, void *stream);

/* y = addend + dropout( act( (x - mean) * rstd * gamma[c] * (1 + scale[n,c]) + beta[c] * (1 + scale) + shift[n,c] ) )
 * addend (NHWC bf16, nullable) is the residual of the post-norm blocks: h1 + act(GN(conv2(h1)))
 * (pdearena twod_unetbase.py:149-151,:158-161).  stats == NULL means "no normalisation" (mean 0,
 * rstd 1: the norm=False blocks, which are a plain activation); pass G = 1 then.
 * scale/shift (float [N,C]) may be NULL (diff_mnist's use_scale_shift_norm, layers.py:330-334).
 * Dropout: keep-probability 1-p, Philox4x32-10 counter (seed, offset + *offset_dev + element index/4),
 * scaled 1/(1-p); p = 0 disables it.  offset_dev (nullable) is a device-resident counter so that a
 * captured CUDA graph draws a fresh mask on every replay; backward passes the same triple. */
int ub200_gn_act_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G,
                               const float *stats, float eps, const float *gamma, const float *beta,
                               const float *scale, const float *shift, int act,
                               float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                               const void *addend, int64_t ld_add,
                               void *y, int64_t ld_y, void *stream);

/* Backward of the above.  gy, x NHWC bf16 -> gx NHWC bf16 (accumulate must be 0: the first pass parks
 * dz = gy * mask * act'(z) in the gx buffer, the second pass finishes it in place; gx may not alias gy or x),
 * dgamma/dbeta float [C] (accumulated with atomics: zero them first), dscale/dshift float [N,C]
 * (NULL when unused).  ws is a float workspace of ub200_gn_act_bwd_ws_floats(N, C, G) elements. */
size_t ub200_gn_act_bwd_ws_floats(int64_t N, int64_t C, int G);
int ub200_gn_act_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x,
                               int64_t N, int64_t HW, int64_t C, int G,
                               const float *stats, float eps, const float *gamma, const float *beta,
                               const float *scale, const float *shift, int act,
                               float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                               void *gx, int64_t ld_gx, int accumulate,
                               float *dgamma, float *dbeta, float *dscale, float *dshift,
                               float *ws, void *stream);

/* Single-pass variants of the three calls above: one thread-block cluster per sample keeps the sample's
 * slab in distributed shared memory, so x (and gy) are read from HBM once.  forward = statistics + apply
 * (stats [N,G,2] is an OUTPUT here, kept for the backward); backward = both backward passes.  When the
 * slab does not fit one cluster (8 CTAs x ~200 KB) they run the multi-pass kernels instead: same results
 * either way.  stats == NULL ("no normalisation") is only valid for the backward.  gadd (backward, nullable):
 * a second gradient of x, NHWC bf16 -- the ResBlock input also feeds the shortcut / residual branch
 * (diff_cifar/model.py:167), and summing the two gradients here saves a separate pass over the tensor. */
int ub200_gn_act_fused_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G,
                                     float *stats, float eps, const float *gamma, const float *beta,
                                     const float *scale, const float *shift, int act,
                                     float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                     const void *addend, int64_t ld_add, void *y, int64_t ld_y, void *stream);
int ub200_gn_act_fused_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x,
                                     int64_t N, int64_t HW, int64_t C, int G,
                                     const float *stats, float eps, const float *gamma, const float *beta,
                                     const float *scale, const float *shift, int act,
                                     float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                     void *gx, int64_t ld_gx,
                                     float *dgamma, float *dbeta, float *dscale, float *dshift,
                                     const void *gadd, int64_t ld_gadd,   /* nullable: gx = backward(gy) + gadd */
                                     float *ws, void *stream);

/* Streaming variants of the same two calls: a sums launch and an apply launch, both plain streams over the tensor
 * (no clusters, no shared-memory slab); the apply launch re-reads from L2 what the sums launch just touched, so HBM
 * still sees every tensor once.  Faster than the cluster kernels once a sample's slab is larger than what one or two
 * CTAs hold (ub200_gn_stream_preferred: the policy the torch binding follows; UB200_GN_STREAM=0/1 forces it).
 * Same arguments and results; ws = ub200_gn_stream_ws_floats(N, HW, C, G) floats of scratch, no zeroing needed:
 * sums travel between the launches as per-CTA partials that are reduced in a fixed order (bit-reproducible). */
size_t ub200_gn_stream_ws_floats(int64_t N, int64_t HW, int64_t C, int G);
int ub200_gn_stream_preferred(int64_t N, int64_t HW, int64_t C, int G, int backward);
int ub200_gn_act_stream_fwd_nhwc_bf16(const void *x, int64_t ld_x, int64_t N, int64_t HW, int64_t C, int G,
                                      float *stats, float eps, const float *gamma, const float *beta,
                                      const float *scale, const float *shift, int act,
                                      float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                      const void *addend, int64_t ld_add, void *y, int64_t ld_y,
                                      float *ws, void *stream);
int ub200_gn_act_stream_bwd_nhwc_bf16(const void *gy, int64_t ld_gy, const void *x, int64_t ld_x,
                                      int64_t N, int64_t HW, int64_t C, int G,
                                      const float *stats, float eps, const float *gamma, const float *beta,
                                      const float *scale, const float *shift, int act,
                                      float dropout_p, uint64_t seed, uint64_t offset, const uint64_t *offset_dev,
                                      void *gx, int64_t ld_gx,
                                      float *dgamma, float *dbeta, float *dscale, float *dshift,
                                      const void *gadd, int64_t ld_gadd,
                                      float *ws, void *stream);

/* ------------------------------------------------------------------------------------------
 * 3x3 / 1x1 convolution, stride 1, "same" zero padding, as tcgen05/TMEM implicit GEMM fed by TMA
 * (replaces nn.Conv2d fprop / dgrad / wgrad; diff_cifar/model.py:69,:133,:143,:146,:396;
 * layers.py:286,:300,:305-312; twod_unetbase.py:19-24).  Activations NHWC bf16, weights packed
 * bf16 [Cout, kh, kw, Cin] (ub200_pack_conv_weight), fp32 accumulation in TMEM.
 *
 *   out[n,y,x,co] = bias[co] + rowadd[n,co] + residual[n,y,x,co]
 *                 + sum_{ky,kx,ci} a [n,y+ky-p,x+kx-p,ci] * w [co,ky,kx,ci]
 *                 + sum_{ci2}      a2[n,y,x,ci2]          * w2[co,ci2]          (optional 1x1 term:
 *                   the ResBlock shortcut folded in as extra K slices; model.py:145-148,:167)
 *
 * bias / rowadd (the temb projection, model.py:164) / residual / a2+w2 may be NULL.
 * Cin, Cin2 multiples of 16.  Cout is arbitrary for the fp32 NCHW output (the packed weights are
 * padded to a multiple of 16 rows) and a multiple of 8 for the bf16 NHWC output.
 * dgrad is the same entry point called with spatially flipped, transposed packed weights
 * (ub200_pack_conv_weight with transpose_flip = 1).
 * ------------------------------------------------------------------------------------------ */
typedef struct ub200_conv_args {
    const void *a;  int64_t ld_a;  int64_t Cin;      /* NHWC bf16 input, pixel stride, channels      */
    const void *w;                                   /* packed bf16 [Cout, k, k, Cin]                 */
    int ksize;                                       /* 3 or 1                                        */
    const void *a2; int64_t ld_a2; int64_t Cin2;     /* optional 1x1 term                             */
    const void *w2;                                  /* packed bf16 [Cout, Cin2]                      */
    const float *bias;                               /* [Cout] or NULL                                */
    const float *rowadd;                             /* [N, Cout] or NULL                             */
    const void *residual; int64_t ld_res;            /* NHWC bf16 [N,H,W,Cout] or NULL                */
    void *out; int64_t ld_out;                       /* NHWC bf16 output                              */
    float *out_f32_nchw;                             /* optional: also/instead write NCHW f32 output  */
    int64_t N, H, W, Cout;
    float *gn_partial;                               /* optional [N, G, 2] sum / sumsq of the output  */
    int gn_groups;                                   /*   (next layer's GroupNorm statistics) or 0    */
    const float *bias2;                              /* [Cout] or NULL: a second bias (the fused 1x1  */
                                                     /*   shortcut's, diff_cifar/model.py:167)        */
    int stride;                                      /* 0 / 1: stride 1.  2: nn.Conv2d(.., 3, stride=2, padding=1) of the      */
                                                     /*   down-sampling arms (diff_cifar/model.py:52, layers.py:238,           */
                                                     /*   twod_unet.py Downsample): H, W are the INPUT extents, out / residual */
                                                     /*   / a2 have ceil(H/2) x ceil(W/2) pixels; the TMA traversal stride     */
                                                     /*   fetches every second pixel, so no full-resolution conv is computed   */
} ub200_conv_args;

int ub200_conv_fprop(const ub200_conv_args *args, void *stream);

/* dW[co,ky,kx,ci] += sum_{n,y,x} gout[n,y,x,co] * a[n,y+ky-p,x+kx-p,ci]   (fp32, packed layout
 * [Cout,k,k,Cin] = the channels_last memory of the torch weight).  The pixel range is split across
 * CTAs and the partial sums are ADDED into dw by the L2 (TMA reduce-add boxes; per-thread fp32 atomics
 * when Cin is not a multiple of 32), so the call ACCUMULATES: zero dw first for a plain gradient, and
 * the summation order is not fixed.  Cin, Cout multiples of 16; dw 16-byte aligned. */
int ub200_conv_wgrad(const void *gout, int64_t ld_g, const void *a, int64_t ld_a,
                     int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ksize,
                     float *dw, void *stream);
/* Same for a convolution of stride 1 or 2 (padding k/2): H, W are the extents of `a`; gout has ceil(H/stride) x
 * ceil(W/stride) pixels and dW[co,ky,kx,ci] += sum gout[n,y,x,co] * a[n, stride*y+ky-p, stride*x+kx-p, ci]. */
int ub200_conv_wgrad_strided(const void *gout, int64_t ld_g, const void *a, int64_t ld_a,
                             int64_t N, int64_t H, int64_t W, int64_t Cin, int64_t Cout, int ksize, int stride,
                             float *dw, void *stream);

/* Channel sums of an NHWC bf16 tensor: per_sample[N,C] = sum_p x[n,p,c] (overwritten; the
 * time-embedding / scale-shift gradient, and the workspace of the second pass) and, when total is
 * not NULL, total[C] += sum_n per_sample[n,c] (the conv bias gradient; accumulates). */
int ub200_chansum_nhwc_bf16(const void *x, int64_t ld, int64_t N, int64_t HW, int64_t C,
                            float *per_sample, float *total, void *stream);
/* total[C] += sum_n rows[n, c]: the second pass alone (a second bias that shares the same gradient). */
int ub200_colsum_rows_f32(const float *rows, int64_t N, int64_t C, float *total, void *stream);

/* fp32 weight (element (co,ci,ky,kx) at w[co*s_co + ci*s_ci + ky*s_ky + kx*s_kx], so both the
 * contiguous torch layout and channels_last work) -> packed bf16 [rows_pad,k,k,cols]:
 * transpose_flip = 0: rows = Cout, cols = Cin (fprop / wgrad operand);
 * transpose_flip = 1: rows = Cin, cols = Cout, taps rotated by 180 degrees (dgrad operand).
 * rows_pad = rows rounded up to 16, padding rows are zero. */
int ub200_pack_conv_weight(const float *w, int64_t Cout, int64_t Cin, int ksize,
                           int64_t s_co, int64_t s_ci, int64_t s_ky, int64_t s_kx,
                           int transpose_flip, void *out_bf16, void *stream);

/* ------------------------------------------------------------------------------------------
 * Attention core of AttnBlock (diff_cifar/model.py:100-119; diff_mnist layers.py:371-391 with one head): one batched GEMM
 * kernel on tcgen05 (M = 256 rows per CTA, N, K <= 256) with fused epilogues; forward = 2 calls, backward = 4 (see
 * csrc/attention.cu).  All matrices are row-major bf16 [R = 256 * groups rows, ld]; a group of 256 rows holds 256 / T
 * whole samples of T tokens (T divides 256), scores between different samples of a group are masked.
 *   a_mn_major = 0: A[m, k] = a[row0 + m, k]   (K columns)      1: A[m, k] = a[row0 + k, m]   (256 columns, K = 256)
 *   b_mn_major = 0: B[n, k] = b[row0 + n, k]   (N = 256 rows)   1: B[n, k] = b[row0 + k, n]   (N columns, K = 256)
 *   out[row0 + m, n] = epilogue( sum_k A[m, k] * B[n, k] ):
 *     UB200_BGEMM_PLAIN        alpha * acc
 *     UB200_BGEMM_SOFTMAX      row softmax of acc * scale over the sample's own T keys; pass alpha = scale * log2(e)
 *     UB200_BGEMM_SOFTMAX_BWD  p o (acc - rowsum(p o acc)) * alpha     with p [R, 256] the saved softmax
 * ------------------------------------------------------------------------------------------ */
#define UB200_BGEMM_PLAIN 0
#define UB200_BGEMM_SOFTMAX 1
#define UB200_BGEMM_SOFTMAX_BWD 2
typedef struct ub200_bgemm_args {
    const void *a; int64_t ld_a; int a_mn_major;
    const void *b; int64_t ld_b; int b_mn_major;
    void *out; int64_t ld_out;
    int64_t groups, N, K;
    int epilogue; float alpha; int T;
    const void *p; int64_t ld_p;
} ub200_bgemm_args;
int ub200_bgemm256(const ub200_bgemm_args *args, void *stream);

/* ------------------------------------------------------------------------------------------
 * Train-step tail (diff_cifar/main.py:425-429, :57-77): global-norm clip + Adam + EMA over a
 * flat fp32 parameter arena, no host synchronisation.
 * ------------------------------------------------------------------------------------------ */
/* sumsq[0] += sum(g^2) */
int ub200_sumsq_f32(const float *g, int64_t n, float *sumsq, void *stream);
/* g' = g * grad_scale * min(1, max_norm / (sqrt(sumsq[0]) * grad_scale + 1e-6))  (sumsq NULL or
 * max_norm <= 0: no clip); torch.optim.Adam update with bias correction for the 1-based step_host;
 * ema = decay*ema + (1-decay)*p (ema NULL to skip).  warmup_steps > 0 scales lr by
 * min(step-1, warmup)/warmup (the LambdaLR of diff_cifar/main.py:90-91, stepped after the optimiser).
 * step_dev (nullable) is a device-resident step counter that overrides step_host, so that a
 * captured CUDA graph replays with the right bias correction and learning rate. */

int ub200_adam_ema_step_f32(float *p, const float *g, float *m, float *v, float *ema, int64_t n,
                            const float *sumsq, float max_norm, float grad_scale,
                            float lr, float beta1, float beta2, float eps, float ema_decay,
                            int64_t step_host, int64_t warmup_steps, const int64_t *step_dev,
                            void *shadow_bf16, void *stream);
/* Same with decoupled weight decay, p *= 1 - lr_eff * weight_decay before the Adam update: torch.optim.AdamW as
 * pdearena's PDEModel configures it (pdearena/pdearena/models/pdemodel.py, lr 2e-4, weight_decay 1e-5).
 * weight_decay = 0 is ub200_adam_ema_step_f32 exactly. */
int ub200_adamw_ema_step_f32(float *p, const float *g, float *m, float *v, float *ema, int64_t n,
                             const float *sumsq, float max_norm, float grad_scale,
                             float lr, float beta1, float beta2, float eps, float weight_decay, float ema_decay,
                             int64_t step_host, int64_t warmup_steps, const int64_t *step_dev,
                             void *shadow_bf16, void *stream);

/* shadow_bf16 (nullable, n bf16 elements) receives bf16(p) from the optimiser kernel: conv weights keep
 * the [Cout,kh,kw,Cin] order in the arena, so their shadow slice IS the packed fprop / wgrad operand.
 * The dgrad operands ([Cin_pad,kh,kw,Cout], taps rotated) of ALL conv layers are then produced by one
 * launch from a device table of n_convs rows (src offset in the shadow arena, dst offset in
 * dgrad_arena_bf16, Cout, Cin, k), all in elements. */
int ub200_pack_dgrad_weights_batched(const void *shadow_bf16, void *dgrad_arena_bf16,
                                     const int64_t *table_dev, int n_convs, void *stream);

/* ------------------------------------------------------------------------------------------
 * Batched fp32 row linears of the time-embedding path.  Replaces, in ONE launch per direction, the
 * per-ResBlock `temb_proj` = Swish + Linear(tdim, Cout) (diff_cifar/model.py:134-137, used at :164) and the
 * two Linear layers of every level's TimeEmbedding MLP (diff_cifar/model.py:29-36):
 *     y_i[N, cout_i] = act(x_i[N, K]) @ w_i[cout_i, K]^T + bias_i        act = SiLU if `silu` else identity
 * for i < n_items.  Several items may share one x (all ResBlocks of a level share that level's embedding).
 * Backward, per item: gw_i += gy_i^T @ act(x_i), gbias_i += column sums of gy_i (both ACCUMULATE: they are
 * the caller's gradient buffers), and gx = act'(x) * sum over the items sharing that gx pointer of
 * gy_i @ w_i (OVERWRITTEN).  fp32 FMA arithmetic; cout_i % 4 == 0, K % 4 == 0, all pointers 16-byte aligned.
 * Fields not used by a direction may be NULL.
 * ------------------------------------------------------------------------------------------ */
typedef struct ub200_rowlin_item {
    const float *x;      /* [N, K] */
    const float *w;      /* [cout, K] */
    const float *bias;   /* [cout] or NULL                                   (forward) */
    float *y;            /* [N, cout]                                        (forward) */
    const float *gy;     /* [N, cout]                                        (backward) */
    float *gw;           /* [cout, K], accumulated; NULL = not wanted        (backward) */
    float *gbias;        /* [cout], accumulated; NULL = not wanted           (backward) */
    float *gx;           /* [N, K], overwritten with the group's sum; NULL = not wanted (backward) */
    int64_t cout;
} ub200_rowlin_item;

int ub200_rowlin_fwd(const ub200_rowlin_item *items, int n_items, int64_t N, int64_t K, int silu, void *stream);
int ub200_rowlin_bwd(const ub200_rowlin_item *items, int n_items, int64_t N, int64_t K, int silu, void *stream);

/* ------------------------------------------------------------------------------------------
 * Gradient all-reduce over NVLink peer memory (replaces the implicit gradient exchange of the reference's
 * data-parallel wrappers: nn.DataParallel at diff_cifar/main.py:235-238, Lightning DDP in pdearena).
 * One process per GPU.  Each rank allocates one "symmetric" block [flag area | fp32 gradient arena] with
 * ub200_p2p_alloc (cudaMalloc, zeroed; `handle64` = its 64-byte CUDA IPC handle, to be exchanged between the
 * ranks by any host channel), maps every other rank's block with ub200_p2p_open, and then calls
 * ub200_p2p_allreduce_sum_f32 with the table of the `world` base pointers (its own at index `rank`):
 * arena[offset .. offset + count) of EVERY rank becomes the sum over ranks, bitwise identical on all of
 * them (two-shot, summed in rank order).  A plain kernel launch: stream-ordered, capturable in a CUDA
 * graph, no host synchronisation; all ranks must issue the same sequence of calls (same ranges, same `ctas`).
 * world <= 8, offset and count multiples of 4 floats.  `ctas` = thread blocks of the launch (0: default 148; a
 * bucket that overlaps other kernels can be throttled to a few dozen so that it does not disturb them).
 * The arena starts ub200_p2p_flag_bytes() bytes into the block.
 * ------------------------------------------------------------------------------------------ */
size_t ub200_p2p_flag_bytes(void);
int ub200_p2p_alloc(size_t bytes, void **ptr, unsigned char *handle64);
int ub200_p2p_free(void *ptr);
int ub200_p2p_open(const unsigned char *handle64, void **peer_ptr);
int ub200_p2p_close(void *peer_ptr);
int ub200_p2p_allreduce_sum_f32(void *const *bases, int rank, int world, int64_t offset_floats, int64_t count,
                                int ctas, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* UNET_B200_H */
