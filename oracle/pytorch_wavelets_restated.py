"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Parity unpinned.

CPU restatement of the two ``pytorch_wavelets`` classes the reference calls,
``DWTForward(J, mode='zero', wave='haar')`` and ``DWTInverse(mode='zero',
wave='haar')`` (pytorch-wavelets, PyPI, un-pinned in
/root/reference/requirements_diff_cifar.txt:41; absent from this image).

It follows the package's *published algorithm*, restated in our own words:

* analysis: one strided grouped correlation along W then one along H per level,
  taps ``h0 = [s, s]`` and ``h1 = [s, -s]`` (pywt's ``dec_lo`` / ``dec_hi``
  reversed because ``conv2d`` is a correlation), ``s = float32(1/sqrt(2))``;
  ``mode='zero'`` with an odd extent appends ONE zero at the end of the axis so the
  output extent is ceil(N/2);
* the four sub-bands of a level come out in the order LL, (W-low,H-high),
  (W-high,H-low), HH -- the "LH, HL, HH" order the reference's comments use
  (diff_mnist/mnist_diff/unet.py:570-592; diff_cifar/model.py:283-285);
* synthesis: grouped transposed correlation with ``g0 = [s, s]``, ``g1 = [s, -s]``;
  when the running low band is one larger than the high band of the next level it
  is cropped (that is how odd extents round-trip);
* ``DWTInverse()((Yl, []))`` returns ``Yl`` unchanged -- the only way the reference
  ever calls it (diff_cifar/model.py:311, diff_cifar/diffusion.py:66,
  pdearena/pdearena/modules/twod_unetbase.py:180, wmh/model.py:82).

Call-site contract it must satisfy (reference file:line):
  diff_cifar/model.py:263-267, :310-311      DTWBlock
  diff_cifar/diffusion.py:63-66              multi-resolution noise targets
  pdearena/pdearena/modules/twod_unetbase.py:169-170, :179-180
  wmh/model.py:68-69, :81-82
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_S = float(torch.tensor(1.0 / math.sqrt(2.0), dtype=torch.float32))


def _check(mode: str, wave: str) -> None:
    if mode != "zero" or wave != "haar":
        raise NotImplementedError(
            "restated oracle covers mode='zero', wave='haar' only (the reference's defaults)")


def _analysis_axis(x: torch.Tensor, lo: torch.Tensor, hi: torch.Tensor, axis: int) -> torch.Tensor:
    """One strided 2-tap analysis along `axis` (2 = H, 3 = W) of an NCHW tensor.

    Output channels are interleaved per input channel as (low, high).
    """
    c = x.shape[1]
    n = x.shape[axis]
    if n % 2 == 1:  # zero mode: one zero at the END of an odd axis
        pad = (0, 0, 0, 1) if axis == 2 else (0, 1, 0, 0)
        x = F.pad(x, pad)
    shape = [1, 1, 1, 1]
    shape[axis] = 2
    bank = torch.cat([lo.reshape(shape), hi.reshape(shape)], dim=0).repeat(c, 1, 1, 1)
    stride = (2, 1) if axis == 2 else (1, 2)
    return F.conv2d(x, bank.to(x.dtype), stride=stride, groups=c)


def _synthesis_axis(lo_band: torch.Tensor, hi_band: torch.Tensor, g0: torch.Tensor, g1: torch.Tensor,
                    axis: int) -> torch.Tensor:
    c = lo_band.shape[1]
    shape = [1, 1, 1, 1]
    shape[axis] = 2
    stride = (2, 1) if axis == 2 else (1, 2)
    w0 = g0.reshape(shape).repeat(c, 1, 1, 1).to(lo_band.dtype)
    w1 = g1.reshape(shape).repeat(c, 1, 1, 1).to(lo_band.dtype)
    return (F.conv_transpose2d(lo_band, w0, stride=stride, groups=c)
            + F.conv_transpose2d(hi_band, w1, stride=stride, groups=c))


class DWTForward(nn.Module):
    """`(Yl, [Yh_1 .. Yh_J]) = DWTForward(J)(x)`; Yh_j is [N, C, 3, h_j, w_j], finest first."""

    def __init__(self, J: int = 1, wave: str = "haar", mode: str = "zero"):
        super().__init__()
        _check(mode, wave)
        self.J = J
        self.mode = mode
        # same buffer names as upstream so that state_dict keys line up
        self.register_buffer("h0_col", torch.tensor([_S, _S]).reshape(1, 1, 2, 1))
        self.register_buffer("h1_col", torch.tensor([_S, -_S]).reshape(1, 1, 2, 1))
        self.register_buffer("h0_row", torch.tensor([_S, _S]).reshape(1, 1, 1, 2))
        self.register_buffer("h1_row", torch.tensor([_S, -_S]).reshape(1, 1, 1, 2))

    def forward(self, x: torch.Tensor):
        highs = []
        ll = x
        for _ in range(self.J):
            rows = _analysis_axis(ll, self.h0_row, self.h1_row, axis=3)   # along W first
            y = _analysis_axis(rows, self.h0_col, self.h1_col, axis=2)    # then along H
            n, _, h, w = y.shape
            y = y.reshape(n, -1, 4, h, w)
            ll = y[:, :, 0].contiguous()
            highs.append(y[:, :, 1:].contiguous())
        return ll, highs


class DWTInverse(nn.Module):
    """`x = DWTInverse()((Yl, [Yh_1 .. Yh_J]))`; with an empty list it is the identity."""

    def __init__(self, wave: str = "haar", mode: str = "zero"):
        super().__init__()
        _check(mode, wave)
        self.mode = mode
        self.register_buffer("g0_col", torch.tensor([_S, _S]).reshape(1, 1, 2, 1))
        self.register_buffer("g1_col", torch.tensor([_S, -_S]).reshape(1, 1, 2, 1))
        self.register_buffer("g0_row", torch.tensor([_S, _S]).reshape(1, 1, 1, 2))
        self.register_buffer("g1_row", torch.tensor([_S, -_S]).reshape(1, 1, 1, 2))

    def forward(self, coeffs):
        ll, highs = coeffs
        for band in highs[::-1]:
            if band is None:
                band = torch.zeros(ll.shape[0], ll.shape[1], 3, ll.shape[-2], ll.shape[-1],
                                   dtype=ll.dtype, device=ll.device)
            # an odd finer level makes the running low band one row/col too large
            if ll.shape[-2] > band.shape[-2]:
                ll = ll[..., :-1, :]
            if ll.shape[-1] > band.shape[-1]:
                ll = ll[..., :-1]
            lh, hl, hh = band[:, :, 0], band[:, :, 1], band[:, :, 2]
            lo = _synthesis_axis(ll, lh, self.g0_col, self.g1_col, axis=2)   # undo H
            hi = _synthesis_axis(hl, hh, self.g0_col, self.g1_col, axis=2)
            ll = _synthesis_axis(lo, hi, self.g0_row, self.g1_row, axis=3)   # undo W
        return ll
