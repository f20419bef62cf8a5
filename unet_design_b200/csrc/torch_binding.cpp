// PyTorch C++ extension: torch.ops.unet_b200.* over the C ABI of include/unet_b200.h.
// PyTorch is plumbing here (device memory, current stream, autograd glue lives in Python); every op
// validates device / dtype / strides, fetches the current CUDA stream and calls one extern "C" entry
// point.  Errors surface as c10::Error on the Python side; there is no CPU or eager fallback.
//
// NHWC bf16 activations are torch tensors of logical shape [N, H, W, C] whose last stride is 1 and whose
// pixel stride `ld` = stride(2) may exceed C (a channel slice of a wider buffer: torch.cat elided).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <vector>
#include <torch/library.h>
#include <torch/types.h>

#include "unet_b200.h"

namespace {

using at::Tensor;

// fail loudly (no CPU fallback) before touching the device
#define UB_GUARD(t)                                                                                         \
    TORCH_CHECK((t).is_cuda(), "unet_b200: expected a CUDA tensor; this build has no CPU path");            \
    const c10::cuda::CUDAGuard guard((t).device())


void check_rc(int rc, const char *what) {
    TORCH_CHECK(rc == 0, "unet_b200::", what, " failed: ", ub200_status_string(rc), " (", rc, ")");
}

void *cur_stream() { return (void *)at::cuda::getCurrentCUDAStream().stream(); }

struct Nhwc {
    void *ptr; int64_t N, H, W, C, ld;
};

Nhwc nhwc(const Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16 && t.dim() == 4, name, ": expected a CUDA bf16 [N,H,W,C] tensor");
    const int64_t N = t.size(0), H = t.size(1), W = t.size(2), C = t.size(3);
    const int64_t ld = W > 1 ? t.stride(2) : (H > 1 ? t.stride(1) : (N > 1 ? t.stride(0) : C));
    TORCH_CHECK(C == 1 || t.stride(3) == 1, name, ": channel stride must be 1");
    TORCH_CHECK((W == 1 || t.stride(2) == ld) && (H == 1 || t.stride(1) == W * ld) && (N == 1 || t.stride(0) == H * W * ld),
                name, ": not a dense NHWC view (pixel stride ", ld, ")");
    return {t.data_ptr(), N, H, W, C, ld};
}

const float *f32(const Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), name, ": expected a contiguous CUDA fp32 tensor");
    return t.data_ptr<float>();
}
const float *f32_opt(const c10::optional<Tensor> &t, const char *name) { return t.has_value() ? f32(*t, name) : nullptr; }
float *f32_mut(const Tensor &t, const char *name) { return const_cast<float *>(f32(t, name)); }
void *bf16_opt(const c10::optional<Tensor> &t, int64_t n) {
    if (!t.has_value()) return nullptr;
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kBFloat16 && t->is_contiguous() && t->numel() == n,
                "shadow: expected a contiguous CUDA bf16 tensor with one element per parameter");
    return t->data_ptr();
}
const uint64_t *i64_opt(const c10::optional<Tensor> &t) {
    if (!t.has_value()) return nullptr;
    TORCH_CHECK(t->is_cuda() && t->scalar_type() == at::kLong && t->numel() >= 1, "offset_dev: expected a CUDA int64 tensor");
    return reinterpret_cast<const uint64_t *>(t->data_ptr<int64_t>());
}

// ---------------------------------------------------------------- Haar

// fused multi-resolution MSE: outs[k] at level k (0 = finest); returns {sums[J+1], grad_0 .. grad_J} or {} if not eligible
std::vector<Tensor> multires_mse(const Tensor &noise, const std::vector<Tensor> &outs, bool want_grads) {
    UB_GUARD(noise);
    TORCH_CHECK(noise.dim() == 4, "multires_mse: noise must be [N,C,H,W]");
    const int64_t J = (int64_t)outs.size() - 1;
    const int64_t N = noise.size(0), C = noise.size(1), H = noise.size(2), W = noise.size(3);
    if (J < 1 || J > 3 || H % (1 << J) != 0 || W % 8 != 0) return {};
    const float *op[4] = {nullptr, nullptr, nullptr, nullptr};
    float *gp[4] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<Tensor> res;
    res.push_back(at::zeros({J + 1}, noise.options()));
    for (int64_t k = 0; k <= J; ++k) {
        TORCH_CHECK(outs[k].dim() == 4 && outs[k].size(0) == N && outs[k].size(1) == C && outs[k].size(2) == (H >> k) && outs[k].size(3) == (W >> k),
                    "multires_mse: output shape mismatch at level ", k);
        op[k] = f32(outs[k], "out");
        if (want_grads) { res.push_back(at::empty_like(outs[k])); gp[k] = res.back().data_ptr<float>(); }
    }
    const int rc = ub200_multires_mse_f32(f32(noise, "noise"), N * C, H, W, (int)J, op, want_grads ? gp : nullptr,
                                          res[0].data_ptr<float>(), cur_stream());
    if (rc == UB200_E_UNSUPPORTED) return {};
    check_rc(rc, "multires_mse");
    return res;
}

// J-level fused analysis; returns {ll, highs_1 (finest), ..., highs_J}; empty vector if the shape is not eligible
std::vector<Tensor> haar_dwt2d_multi(const Tensor &x, int64_t J) {
    TORCH_CHECK(x.dim() == 4, "haar_dwt2d_multi: expected [N,C,H,W]");
    UB_GUARD(x);
    const float *xp = f32(x, "x");
    const int64_t N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
    if (!(J == 2 || J == 3) || H % (1 << J) != 0 || W % 8 != 0) return {};
    std::vector<Tensor> out;
    out.push_back(at::empty({N, C, H >> J, W >> J}, x.options()));
    float *hp[3] = {nullptr, nullptr, nullptr};
    for (int64_t j = 1; j <= J; ++j) {
        out.push_back(at::empty({N, C, 3, H >> j, W >> j}, x.options()));
        hp[j - 1] = out.back().data_ptr<float>();
    }
    const int rc = ub200_haar_dwt2d_multi_fwd(xp, N * C, H, W, (int)J, out[0].data_ptr<float>(), hp, cur_stream());
    if (rc == UB200_E_UNSUPPORTED) return {};
    check_rc(rc, "haar_dwt2d_multi_fwd");
    return out;
}

// fused synthesis of J levels; highs finest first; returns an undefined tensor if not eligible
Tensor haar_idwt2d_multi(const Tensor &ll, const std::vector<Tensor> &highs) {
    UB_GUARD(ll);
    const int64_t J = (int64_t)highs.size();
    TORCH_CHECK(ll.dim() == 4 && (J == 2 || J == 3), "haar_idwt2d_multi: expected ll [N,C,h,w] and 2 or 3 levels");
    const int64_t N = ll.size(0), C = ll.size(1), H = ll.size(2) << J, W = ll.size(3) << J;
    const float *hp[3] = {nullptr, nullptr, nullptr};
    for (int64_t j = 1; j <= J; ++j) {
        const Tensor &h = highs[j - 1];
        TORCH_CHECK(h.dim() == 5 && h.size(0) == N && h.size(1) == C && h.size(2) == 3 && h.size(3) == (H >> j) && h.size(4) == (W >> j),
                    "haar_idwt2d_multi: band shape mismatch at level ", j);
        hp[j - 1] = f32(h, "highs");
    }
    Tensor out = at::empty({N, C, H, W}, ll.options());
    const int rc = ub200_haar_idwt2d_multi(f32(ll, "ll"), hp, N * C, H, W, (int)J, out.data_ptr<float>(), cur_stream());
    if (rc == UB200_E_UNSUPPORTED) return Tensor();
    check_rc(rc, "haar_idwt2d_multi");
    return out;
}
std::tuple<Tensor, Tensor> haar_dwt2d_fwd(const Tensor &x, bool want_highs) {
    TORCH_CHECK(x.dim() == 4, "haar_dwt2d_fwd: expected [N,C,H,W]");
    UB_GUARD(x);
    const float *xp = f32(x, "x");
    const int64_t N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3), h2 = (H + 1) / 2, w2 = (W + 1) / 2;
    Tensor ll = at::empty({N, C, h2, w2}, x.options());
    Tensor highs = want_highs ? at::empty({N, C, 3, h2, w2}, x.options()) : at::empty({0}, x.options());
    check_rc(ub200_haar_dwt2d_fwd(xp, N * C, H, W, ll.data_ptr<float>(), want_highs ? highs.data_ptr<float>() : nullptr,
                                  cur_stream()), "haar_dwt2d_fwd");
    return {ll, highs};
}

Tensor haar_idwt2d(const Tensor &ll, const c10::optional<Tensor> &highs, int64_t Hout, int64_t Wout) {
    TORCH_CHECK(ll.dim() == 4, "haar_idwt2d: expected ll [N,C,h,w]");
    UB_GUARD(ll);
    const int64_t N = ll.size(0), C = ll.size(1), h2 = ll.size(2), w2 = ll.size(3);
    if (highs.has_value())
        TORCH_CHECK(highs->dim() == 5 && highs->size(0) == N && highs->size(1) == C && highs->size(2) == 3 &&
                    highs->size(3) == h2 && highs->size(4) == w2, "haar_idwt2d: highs must be [N,C,3,h,w]");
    Tensor out = at::empty({N, C, Hout, Wout}, ll.options());
    check_rc(ub200_haar_idwt2d(f32(ll, "ll"), f32_opt(highs, "highs"), N * C, h2, w2, Hout, Wout, out.data_ptr<float>(),
                               cur_stream()), "haar_idwt2d");
    return out;
}

Tensor dwtblock_fwd(const Tensor &x, int64_t J, int64_t out_channels) {
    TORCH_CHECK(x.dim() == 4, "dwtblock_fwd: expected [N,C,H,W]");
    UB_GUARD(x);
    const int64_t N = x.size(0), C = x.size(1), H = x.size(2), W = x.size(3);
    int64_t h = H, w = W;
    for (int j = 0; j < J; ++j) { h = (h + 1) / 2; w = (w + 1) / 2; }
    Tensor out = at::empty({N, out_channels, h, w}, x.options());
    check_rc(ub200_dwtblock_fwd(f32(x, "x"), N, C, H, W, (int)J, out_channels, out.data_ptr<float>(), cur_stream()), "dwtblock_fwd");
    return out;
}

Tensor dwtblock_bwd(const Tensor &gout, int64_t C, int64_t H, int64_t W, int64_t J) {
    TORCH_CHECK(gout.dim() == 4, "dwtblock_bwd: expected [N,K,h,w]");
    UB_GUARD(gout);
    const int64_t N = gout.size(0);
    Tensor gx = at::empty({N, C, H, W}, gout.options());
    check_rc(ub200_dwtblock_bwd(f32(gout, "gout"), N, C, H, W, (int)J, gout.size(1), gx.data_ptr<float>(), cur_stream()), "dwtblock_bwd");
    return gx;
}

void dwtblock_fwd_nhwc(const Tensor &x, int64_t J, const c10::optional<Tensor> &chmap, const Tensor &out) {
    UB_GUARD(x);
    const Nhwc o = nhwc(out, "out");
    const int32_t *mp = nullptr;
    if (chmap.has_value()) {
        TORCH_CHECK(chmap->is_cuda() && chmap->scalar_type() == at::kInt && chmap->is_contiguous() && chmap->numel() == o.C,
                    "dwtblock_fwd_nhwc: chmap must be a CUDA int32 tensor with one entry per output channel");
        mp = chmap->data_ptr<int32_t>();
    }
    check_rc(ub200_dwtblock_fwd_nhwc_bf16(f32(x, "x"), x.size(0), x.size(1), x.size(2), x.size(3), (int)J, o.C, mp, o.ptr, o.ld,
                                          cur_stream()), "dwtblock_fwd_nhwc_bf16");
}

Tensor dwtblock_nhwc_fwd(const Tensor &x, int64_t J, int64_t out_channels) {
    UB_GUARD(x);
    const Nhwc i = nhwc(x, "x");
    const int64_t ho = J ? (i.H + 1) / 2 : i.H, wo = J ? (i.W + 1) / 2 : i.W;
    Tensor out = at::empty({i.N, ho, wo, out_channels}, x.options());
    check_rc(ub200_dwtblock_nhwc_bf16_fwd(i.ptr, i.ld, i.N, i.H, i.W, i.C, (int)J, out.data_ptr(), out_channels, out_channels,
                                          cur_stream()), "dwtblock_nhwc_bf16_fwd");
    return out;
}

Tensor dwtblock_nhwc_bwd(const Tensor &gout, int64_t H, int64_t W, int64_t C, int64_t J) {
    UB_GUARD(gout);
    const Nhwc g = nhwc(gout, "gout");
    Tensor gx = at::empty({g.N, H, W, C}, gout.options());
    check_rc(ub200_dwtblock_nhwc_bf16_bwd(g.ptr, g.ld, g.N, H, W, C, (int)J, g.C, gx.data_ptr(), C, cur_stream()),
             "dwtblock_nhwc_bf16_bwd");
    return gx;
}

// ---------------------------------------------------------------- layout / resampling
void nchw_to_nhwc(const Tensor &x, const Tensor &out) {
    UB_GUARD(x);
    const Nhwc o = nhwc(out, "out");
    TORCH_CHECK(x.dim() == 4 && x.size(0) == o.N && x.size(1) == o.C && x.size(2) == o.H && x.size(3) == o.W, "nchw_to_nhwc: shape mismatch");
    check_rc(ub200_nchw_f32_to_nhwc_bf16(f32(x, "x"), o.N, o.C, o.H, o.W, o.ptr, o.ld, cur_stream()), "nchw_f32_to_nhwc_bf16");
}

Tensor nhwc_to_nchw(const Tensor &x) {
    UB_GUARD(x);
    const Nhwc i = nhwc(x, "x");
    Tensor out = at::empty({i.N, i.C, i.H, i.W}, x.options().dtype(at::kFloat));
    check_rc(ub200_nhwc_bf16_to_nchw_f32(i.ptr, i.ld, i.N, i.C, i.H, i.W, out.data_ptr<float>(), cur_stream()), "nhwc_bf16_to_nchw_f32");
    return out;
}

void upsample2x(const Tensor &x, const Tensor &out) {
    UB_GUARD(x);
    const Nhwc i = nhwc(x, "x"), o = nhwc(out, "out");
    TORCH_CHECK(o.N == i.N && o.H == 2 * i.H && o.W == 2 * i.W && o.C == i.C, "upsample2x: shape mismatch");
    check_rc(ub200_upsample2x_nhwc_bf16(i.ptr, i.ld, i.N, i.H, i.W, i.C, o.ptr, o.ld, cur_stream()), "upsample2x");
}

void upsample2x_bwd(const Tensor &gout, const Tensor &gx) {
    UB_GUARD(gout);
    const Nhwc g = nhwc(gout, "gout"), o = nhwc(gx, "gx");
    TORCH_CHECK(g.N == o.N && g.H == 2 * o.H && g.W == 2 * o.W && g.C == o.C, "upsample2x_bwd: shape mismatch");
    check_rc(ub200_upsample2x_bwd_nhwc_bf16(g.ptr, g.ld, o.N, o.H, o.W, o.C, o.ptr, o.ld, cur_stream()), "upsample2x_bwd");
}

// ---------------------------------------------------------------- GroupNorm + activation
void gn_stats(const Tensor &x, int64_t G, const Tensor &stats) {
    UB_GUARD(x);
    const Nhwc i = nhwc(x, "x");
    TORCH_CHECK(stats.numel() == i.N * G * 2, "gn_stats: stats must hold N*G*2 floats");
    check_rc(ub200_gn_stats_nhwc_bf16(i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_mut(stats, "stats"), cur_stream()), "gn_stats");
}

void gn_act_fwd(const Tensor &x, int64_t G, const c10::optional<Tensor> &stats, double eps, const c10::optional<Tensor> &gamma,
                const c10::optional<Tensor> &beta, const c10::optional<Tensor> &scale, const c10::optional<Tensor> &shift,
                int64_t act, double dropout_p, int64_t seed, int64_t offset, const c10::optional<Tensor> &offset_dev,
                const c10::optional<Tensor> &addend, const Tensor &y, bool compute_stats) {
    UB_GUARD(x);
    const Nhwc i = nhwc(x, "x"), o = nhwc(y, "y");
    TORCH_CHECK(o.N == i.N && o.H == i.H && o.W == i.W && o.C == i.C, "gn_act_fwd: shape mismatch");
    const void *addp = nullptr; int64_t ld_add = 0;
    if (addend.has_value()) {
        const Nhwc a = nhwc(*addend, "addend");
        TORCH_CHECK(a.N == i.N && a.H == i.H && a.W == i.W && a.C == i.C, "gn_act_fwd: addend shape mismatch");
        addp = a.ptr; ld_add = a.ld;
    }
    if (compute_stats) {     // statistics + apply: two streaming launches, or one cluster kernel for small slabs
        TORCH_CHECK(stats.has_value(), "gn_act_fwd: compute_stats needs a stats buffer");
        if (ub200_gn_stream_preferred(i.N, i.H * i.W, i.C, (int)G, 0)) {
            Tensor ws = at::empty({(int64_t)ub200_gn_stream_ws_floats(i.N, i.H * i.W, i.C, (int)G)}, x.options().dtype(at::kFloat));
            check_rc(ub200_gn_act_stream_fwd_nhwc_bf16(i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_mut(*stats, "stats"), (float)eps,
                                                       f32_opt(gamma, "gamma"), f32_opt(beta, "beta"), f32_opt(scale, "scale"),
                                                       f32_opt(shift, "shift"), (int)act, (float)dropout_p, (uint64_t)seed,
                                                       (uint64_t)offset, i64_opt(offset_dev), addp, ld_add, o.ptr, o.ld,
                                                       ws.data_ptr<float>(), cur_stream()),
                     "gn_act_stream_fwd");
            return;
        }
        check_rc(ub200_gn_act_fused_fwd_nhwc_bf16(i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_mut(*stats, "stats"), (float)eps,
                                                  f32_opt(gamma, "gamma"), f32_opt(beta, "beta"), f32_opt(scale, "scale"),
                                                  f32_opt(shift, "shift"), (int)act, (float)dropout_p, (uint64_t)seed,
                                                  (uint64_t)offset, i64_opt(offset_dev), addp, ld_add, o.ptr, o.ld, cur_stream()),
                 "gn_act_fused_fwd");
        return;
    }
    check_rc(ub200_gn_act_fwd_nhwc_bf16(i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_opt(stats, "stats"), (float)eps,
                                        f32_opt(gamma, "gamma"), f32_opt(beta, "beta"), f32_opt(scale, "scale"),
                                        f32_opt(shift, "shift"), (int)act, (float)dropout_p, (uint64_t)seed, (uint64_t)offset,
                                        i64_opt(offset_dev), addp, ld_add, o.ptr, o.ld, cur_stream()), "gn_act_fwd");
}

void gn_act_bwd(const Tensor &gy, const Tensor &x, int64_t G, const c10::optional<Tensor> &stats, double eps,
                const c10::optional<Tensor> &gamma, const c10::optional<Tensor> &beta, const c10::optional<Tensor> &scale,
                const c10::optional<Tensor> &shift, int64_t act, double dropout_p, int64_t seed, int64_t offset,
                const c10::optional<Tensor> &offset_dev, const Tensor &gx, bool accumulate, const c10::optional<Tensor> &dgamma, const c10::optional<Tensor> &dbeta,
                const c10::optional<Tensor> &dscale, const c10::optional<Tensor> &dshift, const c10::optional<Tensor> &gadd) {
    UB_GUARD(x);
    const Nhwc g = nhwc(gy, "gy"), i = nhwc(x, "x"), o = nhwc(gx, "gx");
    const void *gap = nullptr;
    int64_t ld_ga = 0;
    if (gadd.has_value()) {
        const Nhwc ga = nhwc(*gadd, "gadd");
        TORCH_CHECK(ga.N == i.N && ga.H == i.H && ga.W == i.W && ga.C == i.C, "gn_act_bwd: gadd shape mismatch");
        gap = ga.ptr; ld_ga = ga.ld;
    }
    TORCH_CHECK(g.N == i.N && g.H == i.H && g.W == i.W && g.C == i.C && o.N == i.N && o.H == i.H && o.W == i.W && o.C == i.C,
                "gn_act_bwd: shape mismatch");
    auto mut = [](const c10::optional<Tensor> &t, const char *n) -> float * { return t.has_value() ? f32_mut(*t, n) : nullptr; };
    TORCH_CHECK(!accumulate, "gn_act_bwd: accumulate is not supported");
    if (ub200_gn_stream_preferred(i.N, i.H * i.W, i.C, (int)G, 1)) {
        Tensor ws = at::empty({(int64_t)ub200_gn_stream_ws_floats(i.N, i.H * i.W, i.C, (int)G)}, x.options().dtype(at::kFloat));
        check_rc(ub200_gn_act_stream_bwd_nhwc_bf16(g.ptr, g.ld, i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_opt(stats, "stats"),
                                                   (float)eps, f32_opt(gamma, "gamma"), f32_opt(beta, "beta"), f32_opt(scale, "scale"),
                                                   f32_opt(shift, "shift"), (int)act, (float)dropout_p, (uint64_t)seed,
                                                   (uint64_t)offset, i64_opt(offset_dev), o.ptr, o.ld, mut(dgamma, "dgamma"),
                                                   mut(dbeta, "dbeta"), mut(dscale, "dscale"), mut(dshift, "dshift"), gap, ld_ga,
                                                   ws.data_ptr<float>(), cur_stream()),
                 "gn_act_stream_bwd");
        return;
    }
    Tensor ws = at::empty({(int64_t)ub200_gn_act_bwd_ws_floats(i.N, i.C, (int)G)}, x.options().dtype(at::kFloat));
    check_rc(ub200_gn_act_fused_bwd_nhwc_bf16(g.ptr, g.ld, i.ptr, i.ld, i.N, i.H * i.W, i.C, (int)G, f32_opt(stats, "stats"),
                                              (float)eps, f32_opt(gamma, "gamma"), f32_opt(beta, "beta"), f32_opt(scale, "scale"),
                                              f32_opt(shift, "shift"), (int)act, (float)dropout_p, (uint64_t)seed,
                                              (uint64_t)offset, i64_opt(offset_dev), o.ptr, o.ld, mut(dgamma, "dgamma"),
                                              mut(dbeta, "dbeta"), mut(dscale, "dscale"), mut(dshift, "dshift"), gap, ld_ga,
                                              ws.data_ptr<float>(), cur_stream()),
             "gn_act_bwd");
}

// ---------------------------------------------------------------- convolution
// w / w2: packed bf16 weights (flat); out: NHWC bf16 view or None; out_nchw: fp32 [N,Cout,H,W] or None
void conv_fprop(const Tensor &a, const Tensor &w, int64_t ksize, int64_t Cout, const c10::optional<Tensor> &a2,
                const c10::optional<Tensor> &w2, const c10::optional<Tensor> &bias, const c10::optional<Tensor> &rowadd,
                const c10::optional<Tensor> &residual, const c10::optional<Tensor> &out,
                const c10::optional<Tensor> &out_nchw, const c10::optional<Tensor> &bias2, int64_t stride) {
    UB_GUARD(a);
    const Nhwc A = nhwc(a, "a");
    TORCH_CHECK(stride == 1 || stride == 2, "conv_fprop: stride must be 1 or 2");
    const int64_t Ho = (A.H + stride - 1) / stride, Wo = (A.W + stride - 1) / stride;
    ub200_conv_args args{};
    args.stride = (int)stride;
    args.a = A.ptr; args.ld_a = A.ld; args.Cin = A.C;
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kBFloat16 && w.is_contiguous(), "conv_fprop: w must be packed bf16");
    const int64_t cout_pad = (Cout + 15) / 16 * 16;
    TORCH_CHECK(w.numel() == cout_pad * ksize * ksize * A.C, "conv_fprop: packed weight has ", w.numel(), " elements, expected ",
                cout_pad * ksize * ksize * A.C);
    args.w = w.data_ptr(); args.ksize = (int)ksize;
    if (a2.has_value()) {
        const Nhwc A2 = nhwc(*a2, "a2");
        TORCH_CHECK(A2.N == A.N && A2.H == Ho && A2.W == Wo, "conv_fprop: a2 shape mismatch");
        TORCH_CHECK(w2.has_value() && w2->scalar_type() == at::kBFloat16 && w2->is_contiguous() && w2->numel() == cout_pad * A2.C,
                    "conv_fprop: w2 must be packed bf16 [Cout_pad, Cin2]");
        args.a2 = A2.ptr; args.ld_a2 = A2.ld; args.Cin2 = A2.C; args.w2 = w2->data_ptr();
    }
    if (bias.has_value()) { TORCH_CHECK(bias->numel() == Cout, "conv_fprop: bias size"); args.bias = f32(*bias, "bias"); }
    if (bias2.has_value()) { TORCH_CHECK(bias2->numel() == Cout, "conv_fprop: bias2 size"); args.bias2 = f32(*bias2, "bias2"); }
    if (rowadd.has_value()) { TORCH_CHECK(rowadd->numel() == A.N * Cout, "conv_fprop: rowadd size"); args.rowadd = f32(*rowadd, "rowadd"); }
    if (residual.has_value()) {
        const Nhwc R = nhwc(*residual, "residual");
        TORCH_CHECK(R.N == A.N && R.H == Ho && R.W == Wo && R.C == Cout, "conv_fprop: residual shape mismatch");
        args.residual = R.ptr; args.ld_res = R.ld;
    }
    if (out.has_value()) {
        const Nhwc O = nhwc(*out, "out");
        TORCH_CHECK(O.N == A.N && O.H == Ho && O.W == Wo && O.C == Cout, "conv_fprop: out shape mismatch");
        args.out = O.ptr; args.ld_out = O.ld;
    }
    if (out_nchw.has_value()) {
        TORCH_CHECK(out_nchw->numel() == A.N * Cout * Ho * Wo, "conv_fprop: out_nchw size");
        args.out_f32_nchw = f32_mut(*out_nchw, "out_nchw");
    }
    args.N = A.N; args.H = A.H; args.W = A.W; args.Cout = Cout;
    check_rc(ub200_conv_fprop(&args, cur_stream()), "conv_fprop");
}

void conv_wgrad(const Tensor &gout, const Tensor &a, int64_t ksize, const Tensor &dw, int64_t stride) {
    UB_GUARD(a);
    const Nhwc G = nhwc(gout, "gout"), A = nhwc(a, "a");
    TORCH_CHECK(stride == 1 || stride == 2, "conv_wgrad: stride must be 1 or 2");
    TORCH_CHECK(G.N == A.N && G.H == (A.H + stride - 1) / stride && G.W == (A.W + stride - 1) / stride, "conv_wgrad: shape mismatch");
    TORCH_CHECK(dw.is_cuda() && dw.scalar_type() == at::kFloat && dw.numel() == G.C * ksize * ksize * A.C, "conv_wgrad: dw size");
    // dw is the fp32 gradient in [Cout,k,k,Cin] memory order (channels_last view of [Cout,Cin,k,k] accepted)
    TORCH_CHECK(dw.is_contiguous() || dw.is_contiguous(at::MemoryFormat::ChannelsLast), "conv_wgrad: dw must be dense");
    check_rc(ub200_conv_wgrad_strided(G.ptr, G.ld, A.ptr, A.ld, A.N, A.H, A.W, A.C, G.C, (int)ksize, (int)stride,
                                      (float *)dw.data_ptr(), cur_stream()), "conv_wgrad");
}

void chansum(const Tensor &x, const Tensor &per_sample, const c10::optional<Tensor> &total,
             const c10::optional<Tensor> &total2) {
    UB_GUARD(x);
    const Nhwc X = nhwc(x, "x");
    TORCH_CHECK(per_sample.numel() == X.N * X.C, "chansum: per_sample must hold N*C floats");
    float *ps = f32_mut(per_sample, "per_sample");
    float *tt = total.has_value() ? f32_mut(*total, "total") : nullptr;
    check_rc(ub200_chansum_nhwc_bf16(X.ptr, X.ld, X.N, X.H * X.W, X.C, ps, tt, cur_stream()), "chansum");
    if (total2.has_value())      // a second bias with the same gradient (conv bias + fused shortcut bias)
        check_rc(ub200_colsum_rows_f32(ps, X.N, X.C, f32_mut(*total2, "total2"), cur_stream()), "colsum_rows");
}

// w: fp32 [Cout,Cin,k,k] with any dense strides (contiguous or channels_last)
void pack_conv_weight(const Tensor &w, bool transpose_flip, const Tensor &out) {
    UB_GUARD(w);
    TORCH_CHECK(w.is_cuda() && w.scalar_type() == at::kFloat && w.dim() == 4 && w.size(2) == w.size(3), "pack_conv_weight: w must be fp32 [Cout,Cin,k,k]");
    const int64_t Cout = w.size(0), Cin = w.size(1), k = w.size(2);
    const int64_t rows = transpose_flip ? Cin : Cout, cols = transpose_flip ? Cout : Cin;
    TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kBFloat16 && out.is_contiguous() &&
                out.numel() == (rows + 15) / 16 * 16 * k * k * cols, "pack_conv_weight: out size");
    check_rc(ub200_pack_conv_weight(w.data_ptr<float>(), Cout, Cin, (int)k, w.stride(0), w.stride(1), w.stride(2), w.stride(3),
                                    transpose_flip ? 1 : 0, out.data_ptr(), cur_stream()), "pack_conv_weight");
}

// ---------------------------------------------------------------- attention core (batched 256-row GEMM)
struct Mat2 { void *ptr; int64_t rows, cols, ld; };
Mat2 mat2(const Tensor &t, const char *name) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kBFloat16 && t.dim() == 2 && t.stride(1) == 1, name,
                ": expected a CUDA bf16 [rows, cols] matrix with unit column stride");
    return {t.data_ptr(), t.size(0), t.size(1), t.size(0) > 1 ? t.stride(0) : t.size(1)};
}

void bgemm256(const Tensor &a, bool a_mn, const Tensor &b, bool b_mn, const Tensor &out, int64_t K, int64_t epilogue,
              double alpha, int64_t T, const c10::optional<Tensor> &p) {
    UB_GUARD(a);
    const Mat2 A = mat2(a, "a"), B = mat2(b, "b"), O = mat2(out, "out");
    TORCH_CHECK(A.rows == B.rows && A.rows == O.rows && A.rows % 256 == 0, "bgemm256: all matrices need the same multiple of 256 rows");
    ub200_bgemm_args g{};
    g.a = A.ptr; g.ld_a = A.ld; g.a_mn_major = a_mn ? 1 : 0;
    g.b = B.ptr; g.ld_b = B.ld; g.b_mn_major = b_mn ? 1 : 0;
    g.out = O.ptr; g.ld_out = O.ld;
    g.groups = A.rows / 256; g.N = O.cols; g.K = K;
    g.epilogue = (int)epilogue; g.alpha = (float)alpha; g.T = (int)T;
    TORCH_CHECK(A.cols == (a_mn ? 256 : K) && B.cols == (b_mn ? O.cols : K), "bgemm256: operand column counts do not match N / K");
    if (p.has_value()) {
        const Mat2 P = mat2(*p, "p");
        TORCH_CHECK(P.rows == A.rows && P.cols == 256, "bgemm256: p must be [R, 256]");
        g.p = P.ptr; g.ld_p = P.ld;
    }
    check_rc(ub200_bgemm256(&g, cur_stream()), "bgemm256");
}

// ---------------------------------------------------------------- optimiser tail
void sumsq(const Tensor &g, const Tensor &acc) {
    UB_GUARD(g);
    check_rc(ub200_sumsq_f32(f32(g, "g"), g.numel(), f32_mut(acc, "acc"), cur_stream()), "sumsq");
}

void adam_ema_step(const Tensor &p, const Tensor &g, const Tensor &m, const Tensor &v, const c10::optional<Tensor> &ema,
                   const c10::optional<Tensor> &sumsq_t, double max_norm, double grad_scale, double lr, double beta1,
                   double beta2, double eps, double ema_decay, int64_t step, int64_t warmup_steps,
                   const c10::optional<Tensor> &step_dev, const c10::optional<Tensor> &shadow, double weight_decay) {
    UB_GUARD(p);
    const int64_t n = p.numel();
    TORCH_CHECK(g.numel() == n && m.numel() == n && v.numel() == n && (!ema.has_value() || ema->numel() == n), "adam_ema_step: size mismatch");
    check_rc(ub200_adamw_ema_step_f32(f32_mut(p, "p"), f32(g, "g"), f32_mut(m, "m"), f32_mut(v, "v"),
                                     ema.has_value() ? f32_mut(*ema, "ema") : nullptr, n, f32_opt(sumsq_t, "sumsq"),
                                     (float)max_norm, (float)grad_scale, (float)lr, (float)beta1, (float)beta2, (float)eps,
                                     (float)weight_decay, (float)ema_decay, step, warmup_steps,
                                     reinterpret_cast<const int64_t *>(i64_opt(step_dev)), bf16_opt(shadow, n), cur_stream()),
             "adam_ema_step");
}

void pack_dgrad_weights_batched(const Tensor &shadow, const Tensor &dgrad_arena, const Tensor &table) {
    UB_GUARD(shadow);
    TORCH_CHECK(shadow.scalar_type() == at::kBFloat16 && dgrad_arena.scalar_type() == at::kBFloat16 && shadow.is_contiguous() &&
                dgrad_arena.is_contiguous(), "pack_dgrad_weights_batched: bf16 contiguous arenas expected");
    TORCH_CHECK(table.is_cuda() && table.scalar_type() == at::kLong && table.is_contiguous() && table.dim() == 2 && table.size(1) == 5,
                "pack_dgrad_weights_batched: table must be a CUDA int64 [n,5] tensor");
    check_rc(ub200_pack_dgrad_weights_batched(shadow.data_ptr(), dgrad_arena.data_ptr(), table.data_ptr<int64_t>(),
                                              (int)table.size(0), cur_stream()), "pack_dgrad_weights_batched");
}

const float *rl_f32(const Tensor &t, const char *what) {
    TORCH_CHECK(t.is_cuda() && t.scalar_type() == at::kFloat && t.is_contiguous(), "rowlin: ", what, " must be a contiguous CUDA fp32 tensor");
    return t.data_ptr<float>();
}

// y_i = act(x_i) @ w_i^T + bias_i for a list of items, one launch (bias list: same length, undefined tensors = no bias)
using OptList = c10::List<c10::optional<Tensor>>;
const Tensor *opt_at(const OptList &l, size_t i, Tensor &hold) {
    c10::optional<Tensor> o = l.get(i);
    if (!o.has_value() || !o->defined()) return nullptr;
    hold = *o;
    return &hold;
}

void rowlin_fwd(at::TensorList x, at::TensorList w, const OptList &bias, at::TensorList y, bool silu) {
    const size_t n = x.size();
    TORCH_CHECK(n > 0 && w.size() == n && (size_t)bias.size() == n && y.size() == n, "rowlin_fwd: list lengths differ");
    UB_GUARD(x[0]);
    const int64_t N = x[0].size(0), K = x[0].size(1);
    std::vector<ub200_rowlin_item> items(n);
    for (size_t i = 0; i < n; ++i) {
        TORCH_CHECK(x[i].dim() == 2 && x[i].size(0) == N && x[i].size(1) == K && w[i].dim() == 2 && w[i].size(1) == K, "rowlin_fwd: shapes");
        const int64_t cout = w[i].size(0);
        TORCH_CHECK(y[i].dim() == 2 && y[i].size(0) == N && y[i].size(1) == cout, "rowlin_fwd: y shape");
        Tensor hb;
        const Tensor *pb = opt_at(bias, i, hb);
        items[i] = ub200_rowlin_item{rl_f32(x[i], "x"), rl_f32(w[i], "w"), pb ? rl_f32(*pb, "bias") : nullptr,
                                     const_cast<float *>(rl_f32(y[i], "y")), nullptr, nullptr, nullptr, nullptr, cout};
    }
    check_rc(ub200_rowlin_fwd(items.data(), (int)n, N, K, silu ? 1 : 0, cur_stream()), "rowlin_fwd");
}

// undefined tensors in gw / gbias / gx = that gradient is not wanted; items sharing one gx tensor are summed into it
void rowlin_bwd(at::TensorList x, at::TensorList w, at::TensorList gy, const OptList &gw, const OptList &gbias, const OptList &gx,
                bool silu) {
    const size_t n = x.size();
    TORCH_CHECK(n > 0 && w.size() == n && gy.size() == n && (size_t)gw.size() == n && (size_t)gbias.size() == n && (size_t)gx.size() == n,
                "rowlin_bwd: list lengths differ");
    UB_GUARD(x[0]);
    const int64_t N = x[0].size(0), K = x[0].size(1);
    std::vector<ub200_rowlin_item> items(n);
    for (size_t i = 0; i < n; ++i) {
        const int64_t cout = w[i].size(0);
        TORCH_CHECK(x[i].dim() == 2 && x[i].size(0) == N && x[i].size(1) == K && w[i].dim() == 2 && w[i].size(1) == K &&
                        gy[i].dim() == 2 && gy[i].size(0) == N && gy[i].size(1) == cout, "rowlin_bwd: shapes");
        Tensor h1, h2, h3;
        const Tensor *pw = opt_at(gw, i, h1), *pb = opt_at(gbias, i, h2), *px = opt_at(gx, i, h3);
        TORCH_CHECK(!pw || pw->numel() == cout * K, "rowlin_bwd: gw size");
        TORCH_CHECK(!pb || pb->numel() == cout, "rowlin_bwd: gbias size");
        TORCH_CHECK(!px || px->numel() == N * K, "rowlin_bwd: gx size");
        items[i] = ub200_rowlin_item{rl_f32(x[i], "x"), rl_f32(w[i], "w"), nullptr, nullptr, rl_f32(gy[i], "gy"),
                                     pw ? const_cast<float *>(rl_f32(*pw, "gw")) : nullptr,
                                     pb ? const_cast<float *>(rl_f32(*pb, "gbias")) : nullptr,
                                     px ? const_cast<float *>(rl_f32(*px, "gx")) : nullptr, cout};
    }
    check_rc(ub200_rowlin_bwd(items.data(), (int)n, N, K, silu ? 1 : 0, cur_stream()), "rowlin_bwd");
}

// ---------------------------------------------------------------- peer-memory gradient all-reduce
// (arena fp32 [floats] living in a symmetric block, 64-byte IPC handle as a CPU uint8 tensor, base pointer of the block)
std::tuple<Tensor, Tensor, int64_t> p2p_alloc(int64_t floats, int64_t device) {
    TORCH_CHECK(floats > 0 && floats % 4 == 0, "p2p_alloc: the arena must be a positive multiple of 4 floats");
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    const size_t fb = ub200_p2p_flag_bytes();
    void *base = nullptr;
    Tensor handle = at::empty({64}, at::TensorOptions().dtype(at::kByte));
    check_rc(ub200_p2p_alloc(fb + (size_t)floats * 4, &base, handle.data_ptr<uint8_t>()), "p2p_alloc");
    Tensor arena = at::from_blob(reinterpret_cast<uint8_t *>(base) + fb, {floats}, [base](void *) { ub200_p2p_free(base); },
                                 at::TensorOptions().dtype(at::kFloat).device(at::kCUDA, (c10::DeviceIndex)device));
    return {arena, handle, (int64_t)reinterpret_cast<uintptr_t>(base)};
}

int64_t p2p_open(const Tensor &handle, int64_t device) {
    TORCH_CHECK(!handle.is_cuda() && handle.scalar_type() == at::kByte && handle.numel() == 64 && handle.is_contiguous(),
                "p2p_open: the handle is a CPU uint8[64] tensor");
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    void *peer = nullptr;
    check_rc(ub200_p2p_open(handle.data_ptr<uint8_t>(), &peer), "p2p_open");
    return (int64_t)reinterpret_cast<uintptr_t>(peer);
}

void p2p_close(int64_t ptr, int64_t device) {
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device);
    check_rc(ub200_p2p_close(reinterpret_cast<void *>((uintptr_t)ptr)), "p2p_close");
}

void p2p_allreduce(const Tensor &arena, std::vector<int64_t> bases, int64_t rank, int64_t offset, int64_t count, int64_t ctas) {
    UB_GUARD(arena);
    TORCH_CHECK(arena.scalar_type() == at::kFloat && arena.is_contiguous() && offset >= 0 && offset + count <= arena.numel(),
                "p2p_allreduce: range outside the arena");
    TORCH_CHECK(bases.size() >= 2 && bases.size() <= 8 && rank >= 0 && rank < (int64_t)bases.size(), "p2p_allreduce: bad world / rank");
    TORCH_CHECK((uintptr_t)bases[rank] + ub200_p2p_flag_bytes() == (uintptr_t)arena.data_ptr(),
                "p2p_allreduce: `arena` is not the arena of this rank's symmetric block");
    void *ptrs[8];
    for (size_t i = 0; i < bases.size(); ++i) ptrs[i] = reinterpret_cast<void *>((uintptr_t)bases[i]);
    check_rc(ub200_p2p_allreduce_sum_f32(ptrs, (int)rank, (int)bases.size(), offset, count, (int)ctas, cur_stream()), "p2p_allreduce");
}

}  // namespace

TORCH_LIBRARY(unet_b200, m) {
    m.def("haar_dwt2d_fwd", &haar_dwt2d_fwd);
    m.def("haar_idwt2d", &haar_idwt2d);
    m.def("dwtblock_fwd", &dwtblock_fwd);
    m.def("dwtblock_bwd", &dwtblock_bwd);
    m.def("dwtblock_fwd_nhwc", &dwtblock_fwd_nhwc);
    m.def("dwtblock_nhwc_fwd", &dwtblock_nhwc_fwd);
    m.def("dwtblock_nhwc_bwd", &dwtblock_nhwc_bwd);
    m.def("nchw_to_nhwc", &nchw_to_nhwc);
    m.def("nhwc_to_nchw", &nhwc_to_nchw);
    m.def("upsample2x", &upsample2x);
    m.def("upsample2x_bwd", &upsample2x_bwd);
    m.def("gn_stats", &gn_stats);
    m.def("gn_act_fwd", &gn_act_fwd);
    m.def("gn_act_bwd", &gn_act_bwd);
    m.def("conv_fprop(Tensor a, Tensor w, int ksize, int Cout, Tensor? a2, Tensor? w2, Tensor? bias, Tensor? rowadd, "
          "Tensor? residual, Tensor? out, Tensor? out_nchw, Tensor? bias2, int stride=1) -> ()", &conv_fprop);
    m.def("conv_wgrad(Tensor gout, Tensor a, int ksize, Tensor dw, int stride=1) -> ()", &conv_wgrad);
    m.def("haar_dwt2d_multi", &haar_dwt2d_multi);
    m.def("multires_mse", &multires_mse);
    m.def("bgemm256", &bgemm256);
    m.def("haar_idwt2d_multi", &haar_idwt2d_multi);
    m.def("chansum", &chansum);
    m.def("pack_conv_weight", &pack_conv_weight);
    m.def("sumsq", &sumsq);
    m.def("adam_ema_step", &adam_ema_step);
    m.def("pack_dgrad_weights_batched", &pack_dgrad_weights_batched);
    m.def("rowlin_fwd", &rowlin_fwd);
    m.def("rowlin_bwd", &rowlin_bwd);
    m.def("p2p_alloc", &p2p_alloc);
    m.def("p2p_open", &p2p_open);
    m.def("p2p_close", &p2p_close);
    m.def("p2p_allreduce(Tensor(a!) arena, int[] bases, int rank, int offset, int count, int ctas=0) -> ()", &p2p_allreduce);
}
