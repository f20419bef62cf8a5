"""GPU parity of the tcgen05 implicit-GEMM convolution (fprop, dgrad, wgrad, fused epilogue terms, 1x1
shortcut K-extension) against torch fp32 convolution of the SAME bf16-rounded operands.
Stated tolerance (north_star): <= 1e-2 relative for bf16; observed error is bf16 output rounding."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-2


@pytest.fixture(scope="module")
def ops():
    from unet_design_b200 import ops as o
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def _mk(n, h, w, cin, cout, k, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = torch.randn(n, h, w, cin, device="cuda", generator=g).to(torch.bfloat16)
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5)
    wt = wt.to(torch.bfloat16).float().contiguous(memory_format=torch.channels_last)
    return a, wt


CASES = [
    # n, h, w, cin, cout, k
    (2, 8, 8, 64, 64, 3), (2, 8, 8, 64, 64, 1), (8, 4, 4, 256, 256, 3), (2, 16, 16, 384, 256, 3), (2, 32, 32, 128, 128, 3),
    (2, 32, 32, 256, 128, 3), (1, 32, 32, 128, 16, 3), (3, 7, 9, 32, 48, 3), (2, 25, 13, 16, 32, 3), (1, 13, 13, 144, 80, 3),
    (2, 16, 16, 512, 256, 1), (2, 8, 8, 128, 512, 3), (1, 16, 16, 64, 320, 3), (64, 16, 16, 64, 320, 1), (1, 128, 128, 64, 64, 3), (1, 96, 192, 64, 64, 3), (1, 200, 200, 16, 16, 3), (5, 4, 4, 512, 256, 3),
]


@pytest.mark.parametrize("case", CASES)
def test_fprop_plain(ops, case):
    n, h, w, cin, cout, k = case
    a, wt = _mk(*case)
    y = ops.conv(a, wt)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt, padding=k // 2).permute(0, 2, 3, 1)
    assert y.shape == ref.shape
    assert rel_err(y, ref) < TOL, f"max abs {float((y.float() - ref).abs().max())}"


@pytest.mark.parametrize("case", [(2, 8, 8, 64, 64, 3), (2, 16, 16, 384, 256, 3), (4, 4, 4, 512, 256, 3), (2, 32, 32, 384, 128, 3),
                                  (3, 7, 9, 32, 48, 3)])
def test_fprop_fused_epilogue_and_shortcut(ops, case):
    n, h, w, cin, cout, k = case
    a, wt = _mk(*case)
    torch.manual_seed(1)
    bias = torch.randn(cout, device="cuda")
    rowadd = torch.randn(n, cout, device="cuda")
    res = torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt, bias, padding=1) + rowadd[:, :, None, None]
    y = ops.conv(a, wt, bias, rowadd=rowadd, residual=res)
    assert rel_err(y, (ref + res.float().permute(0, 3, 1, 2)).permute(0, 2, 3, 1)) < TOL
    # 1x1 shortcut of a second tensor as extra K slices
    cin2 = 2 * cout if cout <= 128 else 512
    a2, w2 = _mk(n, h, w, cin2, cout, 1, seed=3)
    y = ops.conv(a, wt, bias, rowadd=rowadd, a2=a2, w2=w2)
    ref2 = ref + F.conv2d(a2.float().permute(0, 3, 1, 2), w2)
    assert rel_err(y, ref2.permute(0, 2, 3, 1)) < TOL


@pytest.mark.parametrize("case", [(2, 32, 32, 128, 3, 3), (2, 8, 8, 64, 1, 1), (1, 25, 13, 32, 3, 3), (2, 16, 16, 256, 3, 3)])
def test_fprop_narrow_tail_to_nchw(ops, case):
    n, h, w, cin, cout, k = case
    a, wt = _mk(*case)
    bias = torch.randn(cout, device="cuda")
    y = ops.conv(a, wt, bias, out_nchw=True)
    ref = F.conv2d(a.float().permute(0, 3, 1, 2), wt, bias, padding=k // 2)
    assert y.dtype == torch.float32 and y.shape == ref.shape
    assert rel_err(y, ref) < 2e-5          # fp32 accumulate, fp32 store: no bf16 output rounding


@pytest.mark.parametrize("case", CASES)
def test_backward_dgrad_wgrad_bias(ops, case):
    n, h, w, cin, cout, k = case
    a, wt = _mk(*case)
    torch.manual_seed(2)
    a = a.requires_grad_(True)
    wt = wt.requires_grad_(True)
    bias = torch.zeros(cout, device="cuda", requires_grad=True)
    rowadd = torch.zeros(n, cout, device="cuda", requires_grad=True)
    y = ops.conv(a, wt, bias, rowadd=rowadd)
    g = torch.randn_like(y)
    y.backward(g)
    ar = a.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().clone().requires_grad_(True)
    br = torch.zeros(cout, device="cuda", requires_grad=True)
    rr = torch.zeros(n, cout, device="cuda", requires_grad=True)
    yr = F.conv2d(ar, wr, br, padding=k // 2) + rr[:, :, None, None]
    yr.backward(g.float().permute(0, 3, 1, 2))
    assert rel_err(a.grad, ar.grad.permute(0, 2, 3, 1)) < TOL
    assert rel_err(wt.grad, wr.grad) < TOL
    assert rel_err(bias.grad, br.grad) < TOL
    assert rel_err(rowadd.grad, rr.grad) < TOL


def test_backward_with_shortcut_residual_and_tail(ops):
    n, h, w, cin, cout = 2, 16, 16, 128, 64
    a, wt = _mk(n, h, w, cin, cout, 3)
    a2, w2 = _mk(n, h, w, 256, cout, 1, seed=5)
    for t in (a, wt, a2, w2):
        t.requires_grad_(True)
    y = ops.conv(a, wt, None, a2=a2, w2=w2)
    g = torch.randn_like(y)
    y.backward(g)
    ar, a2r = (t.detach().float().permute(0, 3, 1, 2).requires_grad_(True) for t in (a, a2))
    wr, w2r = (t.detach().clone().requires_grad_(True) for t in (wt, w2))
    (F.conv2d(ar, wr, padding=1) + F.conv2d(a2r, w2r)).backward(g.float().permute(0, 3, 1, 2))
    assert rel_err(a2.grad, a2r.grad.permute(0, 2, 3, 1)) < TOL and rel_err(w2.grad, w2r.grad) < TOL
    assert rel_err(a.grad, ar.grad.permute(0, 2, 3, 1)) < TOL and rel_err(wt.grad, wr.grad) < TOL
    # fp32 NCHW tail (Cout = 3): gradient arrives as NCHW fp32
    a, wt = _mk(2, 16, 16, 128, 3, 3, seed=7)
    a.requires_grad_(True); wt.requires_grad_(True)
    bias = torch.zeros(3, device="cuda", requires_grad=True)
    y = ops.conv(a, wt, bias, out_nchw=True)
    g = torch.randn_like(y)
    y.backward(g)
    ar = a.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().clone().requires_grad_(True)
    br = torch.zeros(3, device="cuda", requires_grad=True)
    F.conv2d(ar, wr, br, padding=1).backward(g.to(torch.bfloat16).float())
    assert rel_err(a.grad, ar.grad.permute(0, 2, 3, 1)) < TOL
    assert rel_err(wt.grad, wr.grad) < TOL and rel_err(bias.grad, br.grad) < TOL


def test_linearity_at_full_config_size(ops):
    """BASELINE config-2 size (batch 128, 32x32, 384 -> 128): conv(a1 + a2) == conv(a1) + conv(a2) and a
    spot check of 64 output pixels against torch fp32."""
    a1, wt = _mk(128, 32, 32, 384, 128, 3)
    a2, _ = _mk(128, 32, 32, 384, 128, 3, seed=9)
    y1, y2 = ops.conv(a1, wt), ops.conv(a2, wt)
    ysum = ops.conv((a1.float() + a2.float()).to(torch.bfloat16), wt)
    assert rel_err(ysum, y1.float() + y2.float()) < 2e-2
    ref = F.conv2d(a1[:2].float().permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
    assert rel_err(y1[:2], ref) < TOL


STRIDED = [(2, 32, 32, 128, 128, 3), (2, 16, 16, 64, 64, 3), (3, 8, 8, 256, 256, 3), (2, 25, 13, 32, 48, 3), (1, 128, 128, 64, 64, 3),
           (2, 7, 9, 16, 32, 3), (4, 4, 4, 128, 128, 3), (2, 32, 32, 64, 64, 1)]


@pytest.mark.parametrize("case", STRIDED)
def test_stride2_fprop_dgrad_wgrad(ops, case):
    """nn.Conv2d(C, C, 3, stride=2, padding=1) of the down-sampling arms (diff_cifar/model.py:52, layers.py:238,
    twod_unet.py Downsample) through the TMA traversal stride, forward and all gradients, incl. odd extents."""
    n, h, w, cin, cout, k = case
    a, wt = _mk(*case)
    a = a.requires_grad_(True)
    wt = wt.requires_grad_(True)
    bias = torch.zeros(cout, device="cuda", requires_grad=True)
    y = ops.conv(a, wt, bias, stride=2)
    ar = a.detach().float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wt.detach().clone().requires_grad_(True)
    br = torch.zeros(cout, device="cuda", requires_grad=True)
    yr = F.conv2d(ar, wr, br, stride=2, padding=k // 2)
    assert y.shape == yr.permute(0, 2, 3, 1).shape == (n, (h + 1) // 2, (w + 1) // 2, cout)
    assert rel_err(y, yr.permute(0, 2, 3, 1)) < TOL
    g = torch.randn_like(y)
    y.backward(g)
    yr.backward(g.float().permute(0, 3, 1, 2))
    assert rel_err(a.grad, ar.grad.permute(0, 2, 3, 1)) < TOL
    assert rel_err(wt.grad, wr.grad) < TOL
    assert rel_err(bias.grad, br.grad) < TOL


def test_batched_dgrad_weight_repack_is_exact(ops):
    """`pack_dgrad_weights_batched_` (the tiled-transpose kernel that re-packs every conv's dgrad operand from the bf16 shadow
    arena once per step) against the definition dst[ci][ky][kx][co] = src[co][k-1-ky][k-1-kx][ci]: bit-exact."""
    torch.manual_seed(11)
    shapes = [(256, 256, 3), (128, 384, 3), (768, 256, 1), (16, 128, 3), (128, 16, 3), (272, 80, 3), (64, 64, 1)]
    src_off, dst_off, rows, srcs = 0, 0, [], []
    for cout, cin, k in shapes:
        w = torch.randn(cout, k, k, cin, device="cuda").to(torch.bfloat16)
        srcs.append(w)
        rows.append([src_off, dst_off, cout, cin, k])
        src_off += (w.numel() + 7) // 8 * 8
        dst_off += ((cin + 15) // 16 * 16 * k * k * cout + 7) // 8 * 8
    shadow = torch.zeros(src_off, dtype=torch.bfloat16, device="cuda")
    for w, r in zip(srcs, rows):
        shadow[r[0]:r[0] + w.numel()] = w.reshape(-1)
    dst = torch.full((dst_off,), 7.0, dtype=torch.bfloat16, device="cuda")
    ops.pack_dgrad_weights_batched_(shadow, dst, torch.tensor(rows, dtype=torch.int64, device="cuda"))
    for w, (s0, d0, cout, cin, k) in zip(srcs, rows):
        want = w.flip(1, 2).permute(3, 1, 2, 0).contiguous()                      # [Cin, k, k, Cout], taps rotated by 180 degrees
        got = dst[d0:d0 + want.numel()].view(cin, k, k, cout)
        assert torch.equal(got, want), f"re-pack mismatch for Cout={cout} Cin={cin} k={k}"
