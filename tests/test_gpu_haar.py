"""GPU parity of the Haar kernels (through torch.ops.unet_b200 -> C ABI) against the CPU oracle.
Bar: fp32, bit-exact against oracle/haar_np.py (same expression order, no contraction); the north_star's
1e-6 relative bound against the conv-form restatement of pytorch_wavelets."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import haar_np
from oracle.pytorch_wavelets_restated import DWTForward, DWTInverse

pytestmark = pytest.mark.gpu

SHAPES = [(2, 3, 32, 32), (4, 16, 16, 16), (2, 8, 8, 8), (1, 5, 7, 9), (2, 2, 25, 13), (1, 4, 1, 1), (1, 1, 32, 2),
          (3, 2, 200, 6), (2, 16, 200, 200), (1, 3, 96, 192), (1, 128, 25, 25), (8, 64, 64, 64)]


@pytest.fixture(scope="module")
def ops():
    from unet_design_b200 import ops as o
    return o


@pytest.mark.parametrize("shape", SHAPES)
def test_dwt_level_bit_exact_and_roundtrip(ops, shape):
    torch.manual_seed(0)
    x = torch.randn(*shape)
    yl, yh = ops.haar_dwt2d(x.cuda(), 1)
    ll, lh, hl, hh = haar_np.dwt2_level(x.numpy())
    np.testing.assert_array_equal(yl.cpu().numpy(), ll)
    np.testing.assert_array_equal(yh[0].cpu().numpy(), np.stack([lh, hl, hh], axis=2))
    rec = ops.haar_idwt2d(yl, yh)
    np.testing.assert_array_equal(rec.cpu().numpy(), haar_np.idwt2_level(ll, lh, hl, hh))
    h, w = shape[-2:]
    assert rel_err(rec[..., :h, :w], x) < 1e-6


@pytest.mark.parametrize("shape", SHAPES[:8])
@pytest.mark.parametrize("J", [1, 2, 3])
def test_multilevel_matches_conv_restatement(ops, shape, J):
    torch.manual_seed(1)
    x = torch.randn(*shape)
    yl, yh = ops.haar_dwt2d(x.cuda(), J)
    ref_l, ref_h = DWTForward(J=J)(x)
    assert yl.shape == ref_l.shape and rel_err(yl, ref_l) < 1e-6
    for a, b in zip(yh, ref_h):
        assert a.shape == b.shape and rel_err(a, b, floor=1e-7) < 1e-6
    rec = ops.haar_idwt2d(yl, yh)
    ref = DWTInverse()((ref_l, ref_h))
    assert rec.shape == ref.shape and rel_err(rec, ref) < 1e-6
    # the reference's only use of the inverse: an empty list is the identity
    assert ops.haar_idwt2d(yl, []) is yl


@pytest.mark.parametrize("shape,J,out_ch", [((2, 3, 32, 32), 0, 128), ((2, 128, 32, 32), 1, 128), ((2, 128, 16, 16), 0, 256),
                                            ((1, 5, 7, 9), 1, 12), ((1, 3, 16, 16), 2, 7), ((1, 2, 25, 13), 3, 2),
                                            ((2, 16, 200, 200), 1, 16), ((2, 128, 25, 25), 1, 128), ((2, 3, 33, 31), 0, 8),
                                            ((4, 3, 32, 32), 3, 3), ((4, 3, 32, 32), 2, 3)])
def test_dwtblock_forward_backward_bit_exact(ops, shape, J, out_ch):
    torch.manual_seed(2)
    x = torch.randn(*shape)
    xg = x.cuda().requires_grad_(True)
    y = ops.dwtblock(xg, J, out_ch)
    want = haar_np.dwtblock(x.numpy(), J, out_ch)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), want)
    g = torch.randn(*want.shape)
    y.backward(g.cuda())
    np.testing.assert_array_equal(xg.grad.cpu().numpy(), haar_np.dwtblock_bwd(g.numpy(), shape, J))


def test_dwt_backward_is_the_adjoint(ops):
    torch.manual_seed(3)
    for shape in [(2, 3, 16, 16), (1, 2, 7, 9)]:
        x = torch.randn(*shape, device="cuda", requires_grad=True)
        yl, yh = ops.haar_dwt2d(x, 2)
        gl, gh = torch.randn_like(yl), [torch.randn_like(b) for b in yh]
        (yl * gl).sum().add(sum((a * b).sum() for a, b in zip(yh, gh))).backward()
        xr = x.detach().cpu().requires_grad_(True)
        rl, rh = DWTForward(J=2)(xr)
        (rl * gl.cpu()).sum().add(sum((a * b.cpu()).sum() for a, b in zip(rh, gh))).backward()
        assert rel_err(x.grad, xr.grad) < 1e-6
        # synthesis backward
        ylr = yl.detach().clone().requires_grad_(True)
        yhr = [b.detach().clone().requires_grad_(True) for b in yh]
        rec = ops.haar_idwt2d(ylr, yhr)
        gr = torch.randn_like(rec)
        rec.backward(gr)
        cl = yl.detach().cpu().requires_grad_(True)
        ch = [b.detach().cpu().requires_grad_(True) for b in yh]
        DWTInverse()((cl, ch)).backward(gr.cpu())
        assert rel_err(ylr.grad, cl.grad) < 1e-6
        for a, b in zip(yhr, ch):
            assert rel_err(a.grad, b.grad) < 1e-6


def test_dwtblock_nhwc_with_channel_map(ops):
    torch.manual_seed(4)
    x = torch.randn(3, 3, 16, 16)
    chmap = [((k % 256) % 128) % 3 for k in range(256)]
    out = torch.empty(3, 16, 16, 256, dtype=torch.bfloat16, device="cuda")
    ops.dwtblock_nhwc(x.cuda(), 0, out, torch.tensor(chmap, dtype=torch.int32, device="cuda"))
    want = x[:, chmap].permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(out.cpu(), want)
    # J = 1 into a channel slice of a wider buffer (the skip half of a cat buffer)
    buf = torch.zeros(3, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    ops.dwtblock_nhwc(x.cuda(), 1, buf[..., 32:])
    want = torch.from_numpy(haar_np.dwtblock(x.numpy(), 1, 32)).permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(buf[..., 32:].cpu(), want) and float(buf[..., :32].abs().max()) == 0.0


def test_full_size_properties(ops):
    """Size-independent properties at BASELINE sweep sizes (no CPU oracle needed): perfect reconstruction,
    energy preservation, LL/2 == avg_pool2d, linearity."""
    torch.manual_seed(5)
    x = torch.randn(64, 64, 256, 256, device="cuda")          # 1 GiB
    yl, yh = ops.haar_dwt2d(x, 1)
    rec = ops.haar_idwt2d(yl, yh)
    assert float((rec - x).abs().max()) < 5e-6
    e_in = float(x.double().pow(2).sum())
    e_out = float(yl.double().pow(2).sum() + yh[0].double().pow(2).sum())
    assert abs(e_in - e_out) < 1e-5 * e_in
    assert float((yl * 0.5 - torch.nn.functional.avg_pool2d(x, 2)).abs().max()) < 1e-6
    del rec, yh
    z = torch.randn_like(x)
    lhs = ops.haar_dwt2d(x + 2 * z, 1)[0]
    rhs = yl + 2 * ops.haar_dwt2d(z, 1)[0]
    assert float((lhs - rhs).abs().max()) < 2e-5


@pytest.mark.gpu
def test_multires_helpers_on_gpu():
    """§8 a4 call sites (per-step data / mask / PDE target down-sampling) on the fused Haar kernel, against the numpy oracle."""
    import numpy as np
    from oracle import haar_np
    from unet_design_b200 import multires
    torch.manual_seed(0)
    x = torch.randn(3, 2, 25, 18)
    for J in (1, 2):
        got = multires.downsample(x.cuda(), J).cpu()
        assert np.array_equal(got.numpy(), haar_np.dwtblock(x.numpy(), J, 2))          # bit-exact, like the other Haar paths
    xb, yb = torch.randn(2, 3, 4, 16, 16), torch.randn(2, 1, 4, 16, 16)
    gx, ys = multires.dwt_downsample(xb.cuda(), yb.cuda(), 1, n_levels=3, multi_res_loss=True)
    assert gx.shape == (2, 3, 4, 8, 8) and [t.shape[-1] for t in ys] == [4, 8]
    assert np.array_equal(ys[0].cpu().numpy().reshape(2, 4, 4, 4), haar_np.dwtblock(yb.flatten(0, 1).numpy(), 2, 4))


@pytest.mark.parametrize("shape", [(2, 3, 32, 32), (4, 16, 64, 64), (1, 2, 96, 192), (2, 5, 8, 8), (1, 3, 200, 200), (3, 2, 16, 24)])
@pytest.mark.parametrize("J", [2, 3])
def test_fused_multilevel_is_bit_exact_and_one_launch(ops, shape, J):
    """ub200_haar_dwt2d_multi_fwd / ub200_haar_idwt2d_multi: all J levels in one pass, bit-identical to the numpy oracle
    (and therefore to the level-by-level kernels); backward of each is the other (adjoint identity)."""
    torch.manual_seed(4)
    x = torch.randn(*shape)
    h, w = shape[-2:]
    eligible = h % (1 << J) == 0 and w % 8 == 0
    before = ops.launches()
    yl, yh = ops.haar_dwt2d(x.cuda(), J)
    assert ops.launches() - before == (1 if eligible else J)
    ref_l, ref_h = haar_np.dwt2(x.numpy(), J)
    np.testing.assert_array_equal(yl.cpu().numpy(), ref_l)
    for a, b in zip(yh, ref_h):
        np.testing.assert_array_equal(a.cpu().numpy(), b)
    before = ops.launches()
    rec = ops.haar_idwt2d(yl, yh)
    assert ops.launches() - before == (1 if eligible else J)
    np.testing.assert_array_equal(rec.cpu().numpy(), haar_np.idwt2(ref_l, ref_h))
    assert rel_err(rec[..., :h, :w], x) < 1e-6
    # autograd through the fused pair: <DWT x, g> == <x, DWT^T g>
    xg = x.cuda().requires_grad_(True)
    yl, yh = ops.haar_dwt2d(xg, J)
    gl, gh = torch.randn_like(yl), [torch.randn_like(b) for b in yh]
    loss = (yl * gl).sum() + sum((b * g).sum() for b, g in zip(yh, gh))
    loss.backward()
    want = ops.haar_idwt2d(gl, gh)[..., :h, :w]
    assert rel_err(xg.grad, want) < 1e-6


def test_dwtblock_deeper_than_three_levels_composes(ops):
    """DWTForward accepts any J; the fused kernel holds 3 levels, deeper blocks compose."""
    torch.manual_seed(5)
    for shape, J in (((2, 3, 64, 64), 4), ((1, 2, 100, 36), 5), ((1, 3, 32, 32), 5)):
        x = torch.randn(*shape)
        xg = x.cuda().requires_grad_(True)
        y = ops.dwtblock(xg, J, 7)
        ref = haar_np.dwtblock(x.numpy(), J, 7)
        assert y.shape == ref.shape
        assert rel_err(y, torch.from_numpy(ref)) < 1e-6
        g = torch.randn_like(y)
        y.backward(g)
        assert rel_err(xg.grad, torch.from_numpy(haar_np.dwtblock_bwd(g.cpu().numpy(), shape, J))) < 1e-6


def test_pad_at_start_switch_matches_flipped_oracle():
    """The one unpinned convention, flipped in one place on each side: UB200_HAAR_PAD_AT_START=1 (kernels, read once per
    process, hence the subprocess) against oracle.haar_np.PAD_AT_END = False, on odd extents, bit-exact."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import numpy as np, torch, sys
sys.path.insert(0, "@ROOT@")
from oracle import haar_np
from unet_design_b200 import ops
haar_np.PAD_AT_END = False
torch.manual_seed(0)
for shape in [(2, 2, 25, 13), (1, 5, 7, 9), (1, 3, 200, 6), (2, 3, 8, 8)]:
    x = torch.randn(*shape)
    for J in (1, 2, 3):
        yl, yh = ops.haar_dwt2d(x.cuda(), J)
        rl, rh = haar_np.dwt2(x.numpy(), J)
        assert np.array_equal(yl.cpu().numpy(), rl), (shape, J)
        assert all(np.array_equal(a.cpu().numpy(), b) for a, b in zip(yh, rh)), (shape, J)
        xg = x.cuda().requires_grad_(True)
        y = ops.dwtblock(xg, J, 6)
        assert np.array_equal(y.detach().cpu().numpy(), haar_np.dwtblock(x.numpy(), J, 6)), (shape, J)
        g = torch.randn_like(y)
        y.backward(g)
        assert np.allclose(xg.grad.cpu().numpy(), haar_np.dwtblock_bwd(g.cpu().numpy(), shape, J), atol=1e-6), (shape, J)
    a = x.cuda().permute(0, 2, 3, 1).to(torch.bfloat16) if shape[1] % 8 == 0 else None
print("pad-at-start ok")
""".replace("@ROOT@", root)
    env = dict(os.environ, UB200_HAAR_PAD_AT_START="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "pad-at-start ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("shape,J", [((128, 3, 32, 32), 3), ((4, 3, 32, 32), 2), ((2, 5, 16, 24), 1), ((2, 3, 64, 64), 3)])
def test_fused_multires_loss_matches_oracle(ops, shape, J):
    """ub200_multires_mse_f32: target pyramid LL_k(noise)/2^k + per-level MSE + gradients in one kernel, against the
    level-by-level numpy oracle (diff_cifar/diffusion.py:52-91 builds the same targets with DWTForward per level)."""
    torch.manual_seed(6)
    n, c, h, w = shape
    noise = torch.randn(*shape)
    outs = [torch.randn(n, c, h >> k, w >> k) for k in range(J + 1)]
    outs_gpu = [o.cuda().requires_grad_(True) for o in outs]
    before = ops.launches()
    res = ops.multires_mse(noise.cuda(), outs_gpu)
    assert res is not None and ops.launches() - before == 1
    loss, per_level = res
    loss.backward()
    ref_total = 0.0
    for k, o in enumerate(outs):
        t = torch.from_numpy(haar_np.dwtblock(noise.numpy(), k, None)) if k else noise
        ref = float(((o - t) ** 2).mean())
        ref_total += ref
        assert abs(float(per_level[k].detach()) - ref) < 1e-5 * max(1.0, ref)
        assert rel_err(outs_gpu[k].grad, 2.0 * (o - t) / o.numel()) < 1e-6
    assert abs(float(loss.detach()) - ref_total) < 1e-5 * ref_total
    assert ops.multires_mse(noise.cuda()[..., :-1], outs_gpu) is None        # odd extent: caller goes level by level
