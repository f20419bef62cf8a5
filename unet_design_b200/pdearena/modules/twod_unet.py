"""Drop-in for the residual block of pdearena's "modern U-Net" (pdearena/pdearena/modules/twod_unet.py:16-61):
PRE-norm `conv2(act(norm2(conv1(act(norm1(x)))))) + shortcut(x)` -- the diff_cifar ResBlock pattern without
time embedding or dropout.  The 1x1 shortcut rides as extra K slices of conv2; an identity shortcut is the conv
epilogue's residual.

`Unet` (:389-548, "Unetmod-64") is the container: `Downsample` is a stride-2 3x3 conv (the kernel's TMA traversal
stride), `Upsample` a ConvTranspose2d(4, 2, 1) that stays on PyTorch's channels_last bf16 kernel,
`AttentionBlock` (:125-175, off in every shipped config: is_attn = mid_attn = False) projects through nn.Linear and
uses PyTorch SDPA.  Same class names, constructor / forward signatures, ModuleList layout and state_dict keys."""
from __future__ import annotations

from typing import List, Optional, Tuple, Union

import torch
import torch.nn.functional as F
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
            return ops.conv(a2, self.conv2.weight, self.conv2.bias, a2=x, w2=self.shortcut.weight, bias2=self.shortcut.bias)
        return ops.conv(a2, self.conv2.weight, self.conv2.bias, residual=x)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class AttentionBlock(nn.Module):
    """twod_unet.py:125-175 (multi-head spatial attention; softmax over the QUERY axis `dim=1` as the reference has it)."""

    def __init__(self, n_channels: int, n_heads: int = 1, d_k: Optional[int] = None, n_groups: int = 1):
        super().__init__()
        if d_k is None:
            d_k = n_channels
        self.norm = nn.GroupNorm(n_groups, n_channels)      # constructed but never applied by the reference forward
        self.projection = nn.Linear(n_channels, n_heads * d_k * 3)
        self.output = nn.Linear(n_heads * d_k, n_channels)
        self.scale = d_k ** -0.5
        self.n_heads = n_heads
        self.d_k = d_k

    def forward_nhwc(self, x):
        # fp32 throughout (a few MFLOP; off in every shipped config): the reference's softmax over the query axis is
        # sharp enough that bf16 logits would dominate the block's error
        n, h, w, c = x.shape
        seq = x.reshape(n, h * w, c).float()
        qkv = self.projection(seq).view(n, -1, self.n_heads, 3 * self.d_k)
        q, k, v = torch.chunk(qkv, 3, dim=-1)
        attn = torch.einsum("bihd,bjhd->bijh", q, k) * self.scale
        attn = attn.softmax(dim=1)
        res = torch.einsum("bijh,bjhd->bihd", attn, v).reshape(n, -1, self.n_heads * self.d_k)
        res = (self.output(res) + seq).to(torch.bfloat16)
        return res.reshape(n, h, w, c)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class _ResAttn(nn.Module):
    def forward_nhwc(self, x):
        x = self.res.forward_nhwc(x)
        return x if isinstance(self.attn, nn.Identity) else self.attn.forward_nhwc(x)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class DownBlock(_ResAttn):
    def __init__(self, in_channels: int, out_channels: int, has_attn: bool = False, activation: str = "gelu", norm: bool = False):
        super().__init__()
        self.res = ResidualBlock(in_channels, out_channels, activation=activation, norm=norm)
        self.attn = AttentionBlock(out_channels) if has_attn else nn.Identity()


class UpBlock(_ResAttn):
    def __init__(self, in_channels: int, out_channels: int, has_attn: bool = False, activation: str = "gelu", norm: bool = False):
        super().__init__()
        self.res = ResidualBlock(in_channels + out_channels, out_channels, activation=activation, norm=norm)
        self.attn = AttentionBlock(out_channels) if has_attn else nn.Identity()


class MiddleBlock(nn.Module):
    def __init__(self, n_channels: int, has_attn: bool = False, activation: str = "gelu", norm: bool = False):
        super().__init__()
        self.res1 = ResidualBlock(n_channels, n_channels, activation=activation, norm=norm)
        self.attn = AttentionBlock(n_channels) if has_attn else nn.Identity()
        self.res2 = ResidualBlock(n_channels, n_channels, activation=activation, norm=norm)

    def forward_nhwc(self, x):
        x = self.res1.forward_nhwc(x)
        if not isinstance(self.attn, nn.Identity):
            x = self.attn.forward_nhwc(x)
        return self.res2.forward_nhwc(x)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Upsample(nn.Module):
    def __init__(self, n_channels: int):
        super().__init__()
        self.conv = nn.ConvTranspose2d(n_channels, n_channels, (4, 4), (2, 2), (1, 1))

    def forward_nhwc(self, x):
        y = F.conv_transpose2d(ops._dense_nhwc(x).permute(0, 3, 1, 2), self.conv.weight.to(torch.bfloat16),
                               self.conv.bias.to(torch.bfloat16), stride=2, padding=1)
        return y.permute(0, 2, 3, 1)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Downsample(nn.Module):
    def __init__(self, n_channels):
        super().__init__()
        self.conv = _conv_param(nn.Conv2d(n_channels, n_channels, (3, 3), (2, 2), (1, 1)))

    def forward_nhwc(self, x):
        return ops.conv(x, self.conv.weight, self.conv.bias, stride=2)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Unet(nn.Module):
    """pdearena "modern U-Net" (twod_unet.py:389-548); `Unetmod-64` = hidden 64, ch_mults (1, 2, 2, 4), norm=True."""

    def __init__(self, n_input_scalar_components: int, n_input_vector_components: int, n_output_scalar_components: int,
                 n_output_vector_components: int, time_history: int, time_future: int, hidden_channels: int,
                 activation: str, norm: bool = False, ch_mults: Union[Tuple[int, ...], List[int]] = (1, 2, 2, 4),
                 is_attn: Union[Tuple[bool, ...], List[bool]] = (False, False, False, False), mid_attn: bool = False,
                 n_blocks: int = 2, use1x1: bool = False) -> None:
        super().__init__()
        self.n_input_scalar_components = n_input_scalar_components
        self.n_input_vector_components = n_input_vector_components
        self.n_output_scalar_components = n_output_scalar_components
        self.n_output_vector_components = n_output_vector_components
        self.time_history = time_history
        self.time_future = time_future
        self.hidden_channels = hidden_channels
        self.activation = resolve(activation)
        n_resolutions = len(ch_mults)
        insize = time_history * (n_input_scalar_components + n_input_vector_components * 2)
        n_channels = hidden_channels
        if use1x1:
            self.image_proj = _conv_param(nn.Conv2d(insize, n_channels, kernel_size=1))
        else:
            self.image_proj = _conv_param(nn.Conv2d(insize, n_channels, kernel_size=(3, 3), padding=(1, 1)))
        down = []
        out_channels = in_channels = n_channels
        for i in range(n_resolutions):
            out_channels = in_channels * ch_mults[i]
            for _ in range(n_blocks):
                down.append(DownBlock(in_channels, out_channels, has_attn=is_attn[i], activation=activation, norm=norm))
                in_channels = out_channels
            if i < n_resolutions - 1:
                down.append(Downsample(in_channels))
        self.down = nn.ModuleList(down)
        self.middle = MiddleBlock(out_channels, has_attn=mid_attn, activation=activation, norm=norm)
        up = []
        in_channels = out_channels
        for i in reversed(range(n_resolutions)):
            out_channels = in_channels
            for _ in range(n_blocks):
                up.append(UpBlock(in_channels, out_channels, has_attn=is_attn[i], activation=activation, norm=norm))
            out_channels = in_channels // ch_mults[i]
            up.append(UpBlock(in_channels, out_channels, has_attn=is_attn[i], activation=activation, norm=norm))
            in_channels = out_channels
            if i > 0:
                up.append(Upsample(in_channels))
        self.up = nn.ModuleList(up)
        self.norm = nn.GroupNorm(8, n_channels) if norm else nn.Identity()
        out_channels = time_future * (n_output_scalar_components + n_output_vector_components * 2)
        if use1x1:
            self.final = _conv_param(nn.Conv2d(in_channels, out_channels, kernel_size=1))
        else:
            self.final = _conv_param(nn.Conv2d(in_channels, out_channels, kernel_size=(3, 3), padding=(1, 1)))

    def forward(self, x: torch.Tensor):
        assert x.dim() == 5
        orig_shape = x.shape
        x = x.reshape(x.size(0), -1, *x.shape[3:])          # collapse T, C
        cin = x.shape[1]
        a = ops.to_nhwc(x.float(), (cin + 15) // 16 * 16)
        w = self.image_proj.weight
        if a.shape[3] != cin:
            w = F.pad(w, (0, 0, 0, 0, 0, a.shape[3] - cin))
        x = ops.conv(a, w, self.image_proj.bias)
        h = [x]
        for m in self.down:
            x = m.forward_nhwc(x)
            h.append(x)
        x = self.middle.forward_nhwc(x)
        for m in self.up:
            if isinstance(m, Upsample):
                x = m.forward_nhwc(x)
            else:
                x = m.forward_nhwc(torch.cat((x, h.pop()), dim=3))
        if isinstance(self.norm, nn.GroupNorm):
            x = ops.gn_act(x, self.norm.weight, self.norm.bias, self.norm.num_groups, act=self.activation, eps=self.norm.eps)
        else:
            x = ops.gn_act(x, None, None, 0, act=self.activation)
        out = ops.conv(x, self.final.weight, self.final.bias, out_nchw=True)
        return out.reshape(orig_shape[0], -1, (self.n_output_scalar_components + self.n_output_vector_components * 2),
                           *orig_shape[3:])
