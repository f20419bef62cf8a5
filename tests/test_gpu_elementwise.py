"""GPU parity of the memory-bound siblings: layout converters, nearest x2, GroupNorm + activation
(+ scale/shift, + dropout) forward and backward, channel sums, Adam/EMA tail.  Reference = plain PyTorch
fp32 of the same op on the bf16-rounded inputs; tolerance = bf16 output rounding (2^-8 relative)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
BF16 = 6e-3   # norm-wise; bf16 has 8 bits of mantissa (unit round-off 3.9e-3)


@pytest.fixture(scope="module")
def ops():
    from unet_design_b200 import ops as o
    return o


@pytest.mark.parametrize("shape", [(2, 3, 32, 32), (2, 128, 16, 16), (1, 12, 25, 13), (3, 70, 5, 7)])
def test_layout_roundtrip(ops, shape):
    torch.manual_seed(0)
    x = torch.randn(*shape, device="cuda")
    y = ops.to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).to(torch.bfloat16))
    z = ops.to_nchw(y)
    assert torch.equal(z, x.to(torch.bfloat16).float())
    p = ops.to_nhwc(x, pad_to=(shape[1] + 15) // 16 * 16)
    assert p.shape[3] % 16 == 0 and torch.equal(p[..., :shape[1]], y)
    if p.shape[3] > shape[1]:
        assert float(p[..., shape[1]:].abs().max()) == 0


@pytest.mark.parametrize("shape", [(2, 4, 4, 256), (3, 16, 16, 64), (1, 13, 13, 16)])
def test_upsample2x(ops, shape):
    torch.manual_seed(1)
    x = torch.randn(*shape, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = ops.upsample2x(x)
    ref = F.interpolate(x.detach().float().permute(0, 3, 1, 2), scale_factor=2, mode="nearest").permute(0, 2, 3, 1)
    assert torch.equal(y.float(), ref)
    g = torch.randn_like(y)
    y.backward(g)
    n, h, w, c = shape
    want = g.float().reshape(n, h, 2, w, 2, c).sum(dim=(2, 4))
    assert rel_err(x.grad, want) < BF16


def _gn_ref(x, G, gamma, beta, scale, shift, act, eps=1e-5):
    xn = F.group_norm(x.float().permute(0, 3, 1, 2), G, gamma, beta, eps)
    if scale is not None:
        xn = xn * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    y = {"silu": F.silu, "gelu": F.gelu, "relu": F.relu, "none": lambda t: t}[act](xn)
    return y.permute(0, 2, 3, 1)


@pytest.mark.parametrize("shape,G", [((4, 8, 8, 128), 32), ((2, 32, 32, 384), 32), ((3, 4, 4, 512), 32), ((2, 16, 16, 64), 32),
                                     ((2, 16, 16, 64), 1), ((2, 25, 13, 16), 1), ((3, 32, 32, 128), 32), ((2, 32, 32, 256), 32), ((2, 31, 33, 48), 3), ((1, 128, 128, 64), 1), ((2, 8, 8, 1024), 1)])
@pytest.mark.parametrize("act", ["silu", "gelu", "relu", "none"])
@pytest.mark.parametrize("scale_shift", [False, True])
def test_gn_act_forward_backward(ops, shape, G, act, scale_shift):
    if scale_shift and (act != "silu" or shape[3] > 128):
        pytest.skip("scale/shift is the diff_mnist path: SiLU, <=128 channels")
    torch.manual_seed(2)
    n, h, w, c = shape
    x = (torch.randn(*shape, device="cuda") * 1.5 + 0.3).to(torch.bfloat16).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(c, device="cuda")).requires_grad_(True)
    beta = (0.2 * torch.randn(c, device="cuda")).requires_grad_(True)
    scale = (0.3 * torch.randn(n, c, device="cuda")).requires_grad_(True) if scale_shift else None
    shift = (0.3 * torch.randn(n, c, device="cuda")).requires_grad_(True) if scale_shift else None
    y = ops.gn_act(x, gamma, beta, G, act=act, scale=scale, shift=shift)
    g = torch.randn(*shape, device="cuda").to(torch.bfloat16)
    y.backward(g)
    xr = x.detach().float().requires_grad_(True)
    leaves = [t.detach().clone().requires_grad_(True) if t is not None else None for t in (gamma, beta, scale, shift)]
    yr = _gn_ref(xr, G, leaves[0], leaves[1], leaves[2], leaves[3], act)
    yr.backward(g.float())
    assert rel_err(y, yr) < BF16
    assert rel_err(x.grad, xr.grad) < 2 * BF16
    for ours, ref in zip((gamma, beta, scale, shift), leaves):
        if ours is not None:
            assert rel_err(ours.grad, ref.grad) < 1e-2


def test_gn_act_into_channel_slice_and_from_slice(ops):
    torch.manual_seed(3)
    buf = torch.randn(2, 8, 8, 192, device="cuda").to(torch.bfloat16)
    x = buf[..., 64:192]                           # a channel slice of a wider buffer (pixel stride 192)
    gamma, beta = torch.ones(128, device="cuda"), torch.zeros(128, device="cuda")
    y = ops.gn_act(x, gamma, beta, 32, act="silu")
    assert rel_err(y, _gn_ref(x, 32, gamma, beta, None, None, "silu")) < BF16


def test_dropout_mask_statistics_and_backward_consistency(ops):
    torch.manual_seed(4)
    ops.seed_dropout(1234)
    x = torch.randn(8, 16, 16, 128, device="cuda").to(torch.bfloat16).requires_grad_(True)
    # act = none and beta = 6 keep every un-dropped output away from zero, so the mask can be read off y
    gamma, beta = torch.ones(128, device="cuda"), torch.full((128,), 6.0, device="cuda")
    y = ops.gn_act(x, gamma, beta, 32, act="none", dropout_p=0.1)
    y0 = ops.gn_act(x.detach(), gamma, beta, 32, act="none")
    assert float(y0.float().abs().min()) > 0.5
    kept = y.float() != 0
    assert abs(float(kept.float().mean()) - 0.9) < 0.005
    assert rel_err(y.float()[kept], y0.float()[kept] / 0.9) < BF16
    # per-channel / per-pixel keep rates are uniform (no structure in the counter mapping)
    assert float((kept.float().mean(dim=(0, 1, 2)) - 0.9).abs().max()) < 0.05
    # a second draw uses a different counter range -> a different mask
    y2 = ops.gn_act(x.detach(), gamma, beta, 32, act="none", dropout_p=0.1)
    assert float(((y2.float() != 0) ^ kept).float().mean()) > 0.1
    # backward regenerates the same mask from (seed, offset): compare with autograd through the explicit mask
    g = torch.randn_like(y)
    y.backward(g)
    xr = x.detach().float().requires_grad_(True)
    (_gn_ref(xr, 32, gamma, beta, None, None, "none") * kept.float() / 0.9).backward(g.float())
    assert rel_err(x.grad, xr.grad) < 3 * BF16
    # SiLU path: keep rate measured on outputs that are clearly non-zero before dropout
    beta0 = torch.zeros(128, device="cuda")
    ys = ops.gn_act(x.detach(), gamma, beta0, 32, act="silu", dropout_p=0.25)
    y0s = ops.gn_act(x.detach(), gamma, beta0, 32, act="silu")
    live = y0s.float().abs() > 1e-2
    assert abs(float(((ys.float() != 0) & live).sum()) / float(live.sum()) - 0.75) < 0.01


def test_chansum(ops):
    torch.manual_seed(5)
    x = torch.randn(4, 16, 16, 256, device="cuda").to(torch.bfloat16)
    per = torch.full((4, 256), 7.0, device="cuda")     # overwritten
    tot = torch.zeros(256, device="cuda")
    from unet_design_b200._lib import ops as raw
    tot2 = torch.full((256,), -1.0, device="cuda")
    raw().chansum(x, per, tot, tot2)
    assert rel_err(per, x.float().sum(dim=(1, 2))) < 1e-5
    assert rel_err(tot, x.float().sum(dim=(0, 1, 2))) < 1e-5
    assert rel_err(tot2 + 1.0, x.float().sum(dim=(0, 1, 2))) < 1e-5      # total2 accumulates too


def test_adam_ema_clip_matches_torch(ops):
    torch.manual_seed(6)
    n = 100003
    p0 = torch.randn(n, device="cuda")
    ref = p0.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref], lr=2e-4)
    ema_ref = p0.clone()
    p, m, v, ema = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda"), p0.clone()
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * (0.5 if step == 2 else 0.001)
        ref.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref], 1.0)
        opt.step()
        ema_ref.mul_(0.9999).add_(ref.detach(), alpha=1 - 0.9999)
        acc = torch.zeros(1, device="cuda")
        ops.sumsq_(g, acc)
        assert abs(float(acc) - float(g.double().pow(2).sum())) < 1e-4 * float(acc)
        ops.adam_ema_step_(p, g, m, v, ema, acc, 1.0, 1.0, 2e-4, 0.9, 0.999, 1e-8, 0.9999, step)
        assert rel_err(p, ref) < 1e-6 and rel_err(ema, ema_ref) < 1e-6


@pytest.mark.parametrize("shape,J,cout", [((2, 16, 16, 64), 1, 128), ((2, 25, 25, 128), 1, 256), ((1, 13, 9, 16), 1, 16),
                                          ((2, 8, 8, 64), 0, 128), ((1, 128, 128, 64), 1, 128)])
def test_dwtblock_on_nhwc_activations(ops, shape, J, cout):
    """In-network DWTBlock (pdearena / wmh): NHWC bf16 in/out against the fp32 numpy oracle, forward and adjoint."""
    import numpy as np
    from oracle import haar_np
    torch.manual_seed(7)
    n, h, w, c = shape
    x = torch.randn(*shape, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y = ops.dwtblock_act(x, J, cout)
    xn = x.detach().float().permute(0, 3, 1, 2).cpu().numpy()
    want = torch.from_numpy(haar_np.dwtblock(xn, J, cout)).permute(0, 2, 3, 1)
    assert y.shape == want.shape and rel_err(y, want) < BF16
    g = torch.randn_like(y)
    y.backward(g)
    gn = g.float().permute(0, 3, 1, 2).cpu().numpy()
    gx = torch.from_numpy(np.ascontiguousarray(haar_np.dwtblock_bwd(gn, (n, c, h, w), J))).permute(0, 2, 3, 1)
    assert rel_err(x.grad, gx) < BF16


def test_gn_act_addend_and_no_norm_modes(ops):
    torch.manual_seed(8)
    x = torch.randn(2, 8, 8, 64, device="cuda").to(torch.bfloat16).requires_grad_(True)
    add = torch.randn(2, 8, 8, 64, device="cuda").to(torch.bfloat16).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(64, device="cuda")).requires_grad_(True)
    beta = (0.2 * torch.randn(64, device="cuda")).requires_grad_(True)
    y = ops.gn_act(x, gamma, beta, 1, act="gelu", addend=add)
    g = torch.randn_like(y)
    y.backward(g)
    xr, ar = x.detach().float().requires_grad_(True), add.detach().float().requires_grad_(True)
    gr, br = gamma.detach().clone().requires_grad_(True), beta.detach().clone().requires_grad_(True)
    yr = _gn_ref(xr, 1, gr, br, None, None, "gelu") + ar
    yr.backward(g.float())
    assert rel_err(y, yr) < BF16 and rel_err(x.grad, xr.grad) < 2 * BF16 and rel_err(add.grad, ar.grad) < 1e-6
    assert rel_err(gamma.grad, gr.grad) < 1e-2 and rel_err(beta.grad, br.grad) < 1e-2
    # groups = 0: plain activation (norm=False blocks)
    x2 = torch.randn(2, 8, 8, 64, device="cuda").to(torch.bfloat16).requires_grad_(True)
    y2 = ops.gn_act(x2, None, None, 0, act="gelu")
    y2.backward(g)
    x2r = x2.detach().float().requires_grad_(True)
    F.gelu(x2r).backward(g.float())
    assert rel_err(y2, F.gelu(x2.detach().float())) < BF16 and rel_err(x2.grad, x2r.grad) < 2 * BF16


@pytest.mark.gpu
@pytest.mark.parametrize("N,K,couts,silu", [(128, 512, [128, 256, 256, 128], True), (5, 20, [68, 4, 132], True),
                                            (64, 128, [512], False), (33, 512, [256] * 30, True)])
def test_rowlin_batch_matches_torch(ops, N, K, couts, silu):
    """Batched Swish + Linear rows (temb_proj / TimeEmbedding) against torch fp32, forward and all three gradients;
    items alternate between two shared inputs so the input gradient is a sum over items (and over kernel chunks)."""
    torch.manual_seed(11)
    dev = "cuda"
    xa = torch.randn(N, K, device=dev, requires_grad=True)
    xb = torch.randn(N, K, device=dev, requires_grad=True)
    ws = [torch.randn(c, K, device=dev, requires_grad=True) for c in couts]
    bs = [torch.randn(c, device=dev, requires_grad=True) if i % 3 != 2 else None for i, c in enumerate(couts)]
    xs = [xa if i % 2 == 0 else xb for i in range(len(couts))]
    ys = ops.rowlin_batch(xs, ws, bs, silu=silu)
    gys = [torch.randn_like(y) for y in ys]
    torch.autograd.backward(ys, gys)
    got = [None if t.grad is None else t.grad.clone() for t in [xa, xb] + ws + [b for b in bs if b is not None]]
    for t in [xa, xb] + ws + [b for b in bs if b is not None]:
        t.grad = None
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        act = torch.nn.functional.silu if silu else (lambda v: v)
        refs = [act(x) @ w.t() + (b if b is not None else 0.0) for x, w, b in zip(xs, ws, bs)]
        torch.autograd.backward(refs, gys)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    want = [t.grad for t in [xa, xb] + ws + [b for b in bs if b is not None]]
    for y, r in zip(ys, refs):
        assert rel_err(y, r.detach()) < 1e-5
    for g, w_ in zip(got, want):
        assert (g is None) == (w_ is None)
        if g is not None:
            assert rel_err(g, w_) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(4, 8, 8, 128), (3, 32, 32, 256), (1, 128, 128, 64), (2, 16, 16, 512)])
def test_gn_act_fork_sums_both_gradients_in_the_kernel(ops, shape):
    """gn_act_fork returns (act(GN(x)), x); the gradient of the second output is added inside the backward kernel
    (single-CTA slab, cluster slab and multi-pass fallback shapes) and must equal autograd's separate add."""
    torch.manual_seed(12)
    c = shape[3]
    x0 = torch.randn(*shape, device="cuda").to(torch.bfloat16)
    gamma = (1 + 0.1 * torch.randn(c, device="cuda")).requires_grad_(True)
    beta = (0.1 * torch.randn(c, device="cuda")).requires_grad_(True)
    g1 = torch.randn(*shape, device="cuda").to(torch.bfloat16)
    g2 = torch.randn(*shape, device="cuda").to(torch.bfloat16)
    xa = x0.clone().requires_grad_(True)
    a, xf = ops.gn_act_fork(xa, gamma, beta, 32)
    torch.autograd.backward([a, xf], [g1, g2])
    got, gg, gb = xa.grad.float(), gamma.grad.clone(), beta.grad.clone()
    gamma.grad = beta.grad = None
    xb = x0.clone().requires_grad_(True)
    y = ops.gn_act(xb, gamma, beta, 32)
    y.backward(g1)
    want = xb.grad.float() + g2.float()
    assert rel_err(a, y) < 1e-3          # the multi-pass fallback sums its statistics with fp32 atomics: order varies
    assert rel_err(got, want) < BF16
    assert rel_err(gg, gamma.grad) < 1e-4 and rel_err(gb, beta.grad) < 1e-4
