"""CPU tests of the Haar oracle: the three restatements agree, and the algebraic
identities of SURVEY.md §8(c)(3) hold (the reference's own tests pin nothing here)."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import haar_np
from oracle.pytorch_wavelets_restated import DWTForward, DWTInverse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHAPES = [(2, 3, 8, 8), (1, 5, 7, 9), (2, 2, 25, 13), (1, 4, 1, 1), (1, 1, 32, 2), (3, 2, 200, 6)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("J", [1, 2, 3])
def test_conv_and_butterfly_restatements_agree(shape, J):
    torch.manual_seed(0)
    x = torch.randn(*shape)
    yl_t, yh_t = DWTForward(J=J)(x)
    yl_n, yh_n = haar_np.dwt2(x.numpy(), J)
    assert yl_t.shape == yl_n.shape
    np.testing.assert_allclose(yl_t.numpy(), yl_n, rtol=0, atol=2e-6)
    for a, b in zip(yh_t, yh_n):
        np.testing.assert_allclose(a.numpy(), b, rtol=0, atol=2e-6)
    rec_t = DWTInverse()((yl_t, yh_t))
    rec_n = haar_np.idwt2(yl_n, yh_n)
    np.testing.assert_allclose(rec_t.numpy(), rec_n, rtol=0, atol=4e-6)


@pytest.mark.parametrize("shape", SHAPES)
def test_perfect_reconstruction_and_energy(shape):
    torch.manual_seed(1)
    x = torch.randn(*shape).numpy()
    for J in (1, 2, 3):
        yl, yh = haar_np.dwt2(x, J)
        rec = haar_np.idwt2(yl, yh)
        h, w = x.shape[-2:]
        np.testing.assert_allclose(rec[..., :h, :w], x, atol=5e-6)
        assert np.abs(rec[..., h:, :]).max(initial=0) < 5e-6 and np.abs(rec[..., :, w:]).max(initial=0) < 5e-6
        energy = float((yl.astype(np.float64) ** 2).sum() + sum((b.astype(np.float64) ** 2).sum() for b in yh))
        assert abs(energy - float((x.astype(np.float64) ** 2).sum())) <= 1e-5 * max(1.0, energy)


def test_signs_and_band_order_on_a_2x2_block():
    # [[a, b], [c, d]] = [[1, 2], [4, 8]]: LL=(a+b+c+d)/2, LH=(a+b-c-d)/2, HL=(a-b+c-d)/2, HH=(a-b-c+d)/2
    x = np.array([[1.0, 2.0], [4.0, 8.0]], dtype=np.float32)[None, None]
    yl, yh = haar_np.dwt2(x, 1)
    np.testing.assert_allclose(yl.ravel(), [7.5], rtol=1e-6)
    np.testing.assert_allclose(yh[0].ravel(), [-4.5, -2.5, 1.5], rtol=1e-6)
    yl_t, yh_t = DWTForward(J=1)(torch.from_numpy(x))
    np.testing.assert_allclose(yh_t[0].numpy().ravel(), [-4.5, -2.5, 1.5], rtol=1e-6)


def test_ll_is_block_mean_and_avgpool():
    torch.manual_seed(2)
    x = torch.randn(2, 3, 16, 24)
    for J in (1, 2, 3):
        y = haar_np.dwtblock(x.numpy(), J, None)
        ref = F.avg_pool2d(x, 2 ** J).numpy()
        np.testing.assert_allclose(y, ref, atol=1e-6)


def test_identity_inverse_with_empty_highs():
    x = torch.randn(1, 2, 4, 4)
    assert DWTInverse()((x, [])) is x


def test_noise_pyramid_variance():
    # LL_k/2^k of N(0,1) noise has variance 4^-k (reference multi-res loss targets, diffusion.py:63-70)
    rng = np.random.default_rng(0)
    x = rng.standard_normal((8, 3, 64, 64)).astype(np.float32)
    for k in (1, 2, 3):
        v = haar_np.dwtblock(x, k, None).var()
        assert abs(v * 4 ** k - 1.0) < 0.1


def test_dwtblock_adjoint():
    rng = np.random.default_rng(3)
    for shape, J, out in (((2, 3, 8, 8), 0, 7), ((2, 3, 8, 8), 1, 8), ((1, 2, 7, 9), 2, 5), ((1, 4, 25, 25), 1, 4)):
        x = rng.standard_normal(shape).astype(np.float32)
        y = haar_np.dwtblock(x, J, out)
        g = rng.standard_normal(y.shape).astype(np.float32)
        gx = haar_np.dwtblock_bwd(g, shape, J)
        lhs = float((y.astype(np.float64) * g).sum())
        rhs = float((x.astype(np.float64) * gx).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))


def _c_lib():
    path = os.path.join(ROOT, "oracle", "_ref", "liboracle_haar.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/liboracle_haar.so not built (run __graft_entry__.build())")
    return ctypes.CDLL(path)


@pytest.mark.parametrize("shape", SHAPES)
def test_c_restatement_matches_numpy(shape):
    lib = _c_lib()
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(4)
    x = rng.standard_normal(shape).astype(np.float32)
    n, c, h, w = shape
    h2, w2 = (h + 1) // 2, (w + 1) // 2
    outs = [np.empty((n, c, h2, w2), np.float32) for _ in range(4)]
    lib.oracle_haar_dwt2_level(x.ctypes.data_as(fp), ctypes.c_int64(n * c), ctypes.c_int64(h), ctypes.c_int64(w),
                               *[o.ctypes.data_as(fp) for o in outs])
    for got, want in zip(outs, haar_np.dwt2_level(x)):
        np.testing.assert_array_equal(got, want)          # same float32 expression order: bit-exact
    rec = np.empty((n, c, 2 * h2, 2 * w2), np.float32)
    lib.oracle_haar_idwt2_level(*[o.ctypes.data_as(fp) for o in outs], ctypes.c_int64(n * c),
                                ctypes.c_int64(h2), ctypes.c_int64(w2), rec.ctypes.data_as(fp))
    np.testing.assert_array_equal(rec, haar_np.idwt2_level(*outs))
    for J, out_ch in ((0, 7), (1, c), (2, 2 * c + 1)):
        want = haar_np.dwtblock(x, J, out_ch)
        got = np.empty(want.shape, np.float32)
        scratch = np.empty(2 * n * c * h2 * w2 + 1, np.float32)
        lib.oracle_dwtblock_fwd(x.ctypes.data_as(fp), ctypes.c_int64(n), ctypes.c_int64(c), ctypes.c_int64(h),
                                ctypes.c_int64(w), ctypes.c_int(J), ctypes.c_int64(out_ch),
                                got.ctypes.data_as(fp), scratch.ctypes.data_as(fp))
        np.testing.assert_array_equal(got, want)


def test_odd_extent_pad_side_is_one_switch(monkeypatch):
    """The unpinned convention: PAD_AT_END moves the single zero of an odd extent.  Even extents do not
    depend on it; odd ones keep the output extent ceil(n/2), perfect reconstruction and the adjoint identity either way."""
    rng = np.random.default_rng(0)
    even, odd = rng.standard_normal((2, 3, 8, 12)).astype(np.float32), rng.standard_normal((2, 3, 25, 13)).astype(np.float32)
    ref_even, ref_odd = haar_np.dwt2(even, 2), haar_np.dwt2(odd, 2)
    monkeypatch.setattr(haar_np, "PAD_AT_END", False)
    yl, yh = haar_np.dwt2(even, 2)
    np.testing.assert_array_equal(yl, ref_even[0])
    yl, yh = haar_np.dwt2(odd, 2)
    assert yl.shape == ref_odd[0].shape == (2, 3, 7, 4) and not np.array_equal(yl, ref_odd[0])
    rec = haar_np.idwt2(yl, yh)
    np.testing.assert_allclose(haar_np._crop(rec, 25, 13), odd, atol=2e-6)
    # first row of the level-1 LL now sees the zero row: LL[0, j] = (0 + 0 + x[0, 2j-1] + x[0, 2j]) / 2
    ll1 = haar_np.dwt2_level(odd)[0]
    np.testing.assert_allclose(ll1[..., 0, 1], (odd[..., 0, 1] + odd[..., 0, 2]) / 2, atol=1e-6)
    # adjoint of the block (what the kernels' backward computes) under the flipped convention
    g = rng.standard_normal((2, 6, 13, 7)).astype(np.float32)
    lhs = float((haar_np.dwtblock(odd, 1, 6).astype(np.float64) * g).sum())
    rhs = float((odd.astype(np.float64) * haar_np.dwtblock_bwd(g, odd.shape, 1)).sum())
    assert abs(lhs - rhs) < 1e-4 * max(1.0, abs(lhs))
