"""Drop-in for the residual block of pdearena's "modern U-Net" (pdearena/pdearena/modules/twod_unet.py:16-61):
PRE-norm `conv2(act(norm2(conv1(act(norm1(x)))))) + shortcut(x)` -- the diff_cifar ResBlock pattern without
time embedding or dropout.  The 1x1 shortcut rides as extra K slices of conv2; an identity shortcut is the conv
epilogue's residual.  The `Unet` container (strided-conv down, ConvTranspose up, optional attention) is out of
scope."""
from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ...diff_cifar.model import _conv_param
from .activations import resolve


class ResidualBlock(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, activation: str = "gelu", norm: bool = False, n_groups: int = 1):
        super().__init__()
        self.activation = resolve(activation)
        self.conv1 = _conv_param(nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), padding=(1, 1)))
        self.conv2 = _conv_param(nn.Conv2d(out_channels, out_channels, kernel_size=(3, 3), padding=(1, 1)))
        if in_channels != out_channels:
            self.shortcut = _conv_param(nn.Conv2d(in_channels, out_channels, kernel_size=(1, 1)))
        else:
            self.shortcut = nn.Identity()
        if norm:
            self.norm1 = nn.GroupNorm(n_groups, in_channels)
            self.norm2 = nn.GroupNorm(n_groups, out_channels)
        else:
            self.norm1 = nn.Identity()
            self.norm2 = nn.Identity()

    def _norm_act(self, x, norm):
        if isinstance(norm, nn.GroupNorm):
            return ops.gn_act(x, norm.weight, norm.bias, norm.num_groups, act=self.activation, eps=norm.eps)
        return ops.gn_act(x, None, None, 0, act=self.activation)

    def forward_nhwc(self, x):
        h = ops.conv(self._norm_act(x, self.norm1), self.conv1.weight, self.conv1.bias)
        a2 = self._norm_act(h, self.norm2)
        if isinstance(self.shortcut, nn.Conv2d):
            return ops.conv(a2, self.conv2.weight, self.conv2.bias + self.shortcut.bias, a2=x, w2=self.shortcut.weight)
        return ops.conv(a2, self.conv2.weight, self.conv2.bias, residual=x)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))
