"""Multi-resolution data / target helpers: the call sites of SURVEY.md §8 row a4, which in the reference rebuild
`DWTForward` / `DWTBlock` modules (and move their filters to the device) on EVERY step:

* diff_cifar/main.py:403-415 and diff_mnist/main.py:323-336 -- `x_0 = DWTInverse((DWTForward(J)(x_0).yl, [])) / 2^J`; with an
  empty high-pass list the inverse returns `yl` unchanged, so this is `LL_J(x_0) / 2^J`;
* wmh/train_pt.py:547-559 -- the same for image and mask, the mask re-binarised at 0.5;
* pdearena/pdearena/models/pdemodel.py:141-180 `dwt_downsample` -- inputs and (lists of) targets of a [B, T, C, H, W] batch.

Every one of them is `DWTBlock(J, out_channels=C)(x)` = `LL_J(x) / 2^J`, ONE fused kernel here (`ops.dwtblock`, levels
composed in registers, zero padding at the end of odd extents), under `no_grad` as in the reference."""
from __future__ import annotations

from typing import List, Tuple, Union

import torch

from . import ops


@torch.no_grad()
def downsample(x: torch.Tensor, n_downsample: int) -> torch.Tensor:
    """`LL_J(x) / 2^J` of a [N, C, H, W] batch (J = n_downsample >= 0); J = 0 returns x itself (main.py:404)."""
    if n_downsample <= 0:
        return x
    return ops.dwtblock(x.float().contiguous(), int(n_downsample), x.shape[1])


@torch.no_grad()
def downsample_image_and_mask(image: torch.Tensor, mask: torch.Tensor, n_downsample: int, threshold: float = 0.5):
    """wmh/train_pt.py:547-559: both tensors down-sampled, the mask binarised again (int tensor of 0 / 1)."""
    if n_downsample <= 0:
        return image, mask
    m = downsample(mask, n_downsample)
    return downsample(image, n_downsample), torch.where(m > threshold, torch.ones_like(m).int(), torch.zeros_like(m).int())


@torch.no_grad()
def dwt_downsample(x: torch.Tensor, y: torch.Tensor, n_downsample: int, n_levels: int = 0, multi_res_loss: bool = False
                   ) -> Tuple[torch.Tensor, Union[torch.Tensor, List[torch.Tensor]]]:
    """pdemodel.py:141-180.  x, y: [B, T, C, H, W] (time and batch are flattened for the transform and restored).
    Without the multi-resolution loss both are down-sampled J = n_downsample times.  With it, x is down-sampled J times and
    y is returned at every level j = n_downsample .. n_levels-1, coarsest first (the order of the decoder's outputs)."""
    def run(t, j):
        b, s = t.shape[0], t.shape[1]
        flat = torch.flatten(t, 0, 1).float().contiguous()
        out = ops.dwtblock(flat, int(j), flat.shape[1]) if j > 0 else flat
        return out.reshape(b, s, *out.shape[1:])

    if not multi_res_loss:
        return run(x, n_downsample), run(y, n_downsample)
    ys = [run(y, j) for j in range(n_downsample, n_levels)]
    ys.reverse()
    return run(x, n_downsample), ys
