"""the per-step data / target down-sampling helpers against the numpy Haar oracle."""
import numpy as np
import torch

from oracle import haar_np


def _ref(x, J):
    return torch.from_numpy(haar_np.dwtblock(x.numpy(), J, x.shape[1])) if J > 0 else x


def test_downsample_and_mask(emulated_ops):
    from unet_design_b200 import multires
    torch.manual_seed(0)
    x = torch.randn(3, 2, 25, 18)                       # odd extent: zero padding at the end (wmh 25 -> 13)
    for J in (0, 1, 2):
        got = multires.downsample(x, J)
        assert got.shape == _ref(x, J).shape and torch.allclose(got, _ref(x, J), atol=1e-6)
    mask = (torch.rand(3, 1, 25, 18) > 0.5).float()
    img, m = multires.downsample_image_and_mask(x, mask, 1)
    assert torch.allclose(img, _ref(x, 1), atol=1e-6)
    assert m.dtype == torch.int32 and torch.equal(m, (_ref(mask, 1) > 0.5).int())


def test_pde_dwt_downsample_matches_reference_semantics(emulated_ops):
    from unet_design_b200 import multires
    torch.manual_seed(1)
    x, y = torch.randn(2, 3, 4, 16, 16), torch.randn(2, 1, 4, 16, 16)      # [B, T, C, H, W]
    gx, gy = multires.dwt_downsample(x, y, 1)
    assert gx.shape == (2, 3, 4, 8, 8) and gy.shape == (2, 1, 4, 8, 8)
    assert torch.allclose(gx, _ref(x.flatten(0, 1), 1).reshape(2, 3, 4, 8, 8), atol=1e-6)
    assert torch.allclose(gy, _ref(y.flatten(0, 1), 1).reshape(2, 1, 4, 8, 8), atol=1e-6)
    gx, ys = multires.dwt_downsample(x, y, 0, n_levels=3, multi_res_loss=True)
    assert torch.equal(gx, x) and [t.shape[-1] for t in ys] == [4, 8, 16]   # coarsest first, as the decoder emits them
    for t, J in zip(ys, (2, 1, 0)):
        assert torch.allclose(t, _ref(y.flatten(0, 1), J).reshape(2, 1, 4, *t.shape[-2:]), atol=1e-6)
    # LL_J / 2^J == 2^J x 2^J block mean on even extents
    assert torch.allclose(ys[1], torch.nn.functional.avg_pool2d(y.flatten(0, 1), 2).reshape(2, 1, 4, 8, 8), atol=1e-6)
