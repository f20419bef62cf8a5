"""Drop-in for pdearena/pdearena/modules/twod_unetbase.py (the `_G` family that hosts Multi-ResNet), backed by
the sm_100a kernels.  Same class names, constructor / forward signatures, attribute names and state_dict
keys as the reference (file:line): ConvBlock :12-32, FullResnetConvBlock :148-151, PartialResnetConvBlock
:154-161, DWTBlock :164-193, Down_G :200-218, Up_G :221-251, Unetbase_G :254-396.

Blocks are POST-norm: `act(GroupNorm(1, C)(conv3x3(x)))` twice; the residual adds of the Partial / Full variants
ride in the fused GroupNorm+activation kernel (`addend`).  Inside a model activations are NHWC bf16; public
`forward`s keep the reference's NCHW fp32 (and [B, T, C, H, W] for the container).

`Unetbase` (:60-141, the 2015-style baseline: Down = MaxPool2d + ConvBlock, Up = ConvTranspose2d + cat + ConvBlock)
keeps its class, signature and state_dict; its ConvBlocks run on the kernels, MaxPool2d / ConvTranspose2d stay on
PyTorch's channels_last bf16 kernels (SURVEY.md §2.3 allows it; they are not on the Multi-ResNet path).
Not fused (raise NotImplementedError): `up_fct='conv'` of the `_G` family, activations tanh / sigmoid.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ...diff_cifar.model import _HaarFilterBuffers, _conv_param
from .activations import resolve


def _conv(x, conv: nn.Conv2d, out_nchw: bool = False):
    """3x3 / 1x1 conv whose input channel count may need zero-padding to a multiple of 16 (the heads)."""
    w = conv.weight
    cin = w.shape[1]
    if x.shape[3] != cin:                                    # activation was padded by to_nhwc(pad_to=...)
        w = F.pad(w, (0, 0, 0, 0, 0, x.shape[3] - cin))
    return ops.conv(x, w, conv.bias, out_nchw=out_nchw)


def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


class ConvBlock(nn.Module):
    def __init__(self, in_channels, out_channels, num_groups=1, norm: bool = True, activation="gelu") -> None:
        super().__init__()
        self.activation = resolve(activation)
        self.conv1 = _conv_param(nn.Conv2d(in_channels, out_channels, kernel_size=3, padding=1))
        self.conv2 = _conv_param(nn.Conv2d(out_channels, out_channels, kernel_size=3, padding=1))
        if norm:
            self.norm1 = nn.GroupNorm(num_groups, out_channels)
            self.norm2 = nn.GroupNorm(num_groups, out_channels)
        else:
            self.norm1 = nn.Identity()
            self.norm2 = nn.Identity()

    def _norm_act(self, x, norm, addend=None):
        if isinstance(norm, nn.GroupNorm):
            return ops.gn_act(x, norm.weight, norm.bias, norm.num_groups, act=self.activation, eps=norm.eps, addend=addend)
        return ops.gn_act(x, None, None, 0, act=self.activation, addend=addend)

    def forward_nhwc(self, x):
        h = self._norm_act(_conv(x, self.conv1), self.norm1)
        return self._norm_act(_conv(h, self.conv2), self.norm2)

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x, _pad16(x.shape[1]))))


def _as_nchw_view(x):
    """NHWC bf16 tensor as its NCHW (channels_last) view, for the two PyTorch ops of the baseline `Unetbase`."""
    return ops._dense_nhwc(x).permute(0, 3, 1, 2)


def _from_nchw_view(y):
    return y.permute(0, 2, 3, 1)              # channels_last NCHW -> dense NHWC (ops re-pack if a backend answered NCHW)


class Down(nn.Module):
    """twod_unetbase.py:35-45: MaxPool2d(2) then ConvBlock."""

    def __init__(self, in_channels, out_channels, num_groups=1, norm: bool = True, activation="gelu") -> None:
        super().__init__()
        self.conv = ConvBlock(in_channels, out_channels, num_groups, norm, activation)
        self.pool = nn.MaxPool2d(2)

    def forward_nhwc(self, x):
        return self.conv.forward_nhwc(_from_nchw_view(self.pool(_as_nchw_view(x))))

    def forward(self, x: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Up(nn.Module):
    """twod_unetbase.py:48-57: ConvTranspose2d(in, in/2, 2, stride 2), cat([skip, h]), ConvBlock."""

    def __init__(self, in_channels, out_channels, num_groups=1, norm: bool = True, activation="gelu") -> None:
        super().__init__()
        self.up = nn.ConvTranspose2d(in_channels, in_channels // 2, kernel_size=2, stride=2)
        self.conv = ConvBlock(in_channels, out_channels, num_groups, norm, activation)

    def forward_nhwc(self, x1, x2):
        h = F.conv_transpose2d(_as_nchw_view(x1), self.up.weight.to(torch.bfloat16), self.up.bias.to(torch.bfloat16), stride=2)
        h = torch.cat([x2, _from_nchw_view(h)], dim=3)
        return self.conv.forward_nhwc(h)

    def forward(self, x1: torch.Tensor, x2: torch.Tensor):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x1), ops.to_nhwc(x2)))


class Unetbase(nn.Module):
    """pdearena `Unetbase` (twod_unetbase.py:60-141): forward(x[B,T,C,H,W]) -> [B, time_future, C_out, H, W]."""

    def __init__(self, n_input_scalar_components: int, n_input_vector_components: int, n_output_scalar_components: int,
                 n_output_vector_components: int, time_history: int, time_future: int, hidden_channels: int,
                 activation="gelu") -> None:
        super().__init__()
        self.n_input_scalar_components = n_input_scalar_components
        self.n_input_vector_components = n_input_vector_components
        self.n_output_scalar_components = n_output_scalar_components
        self.n_output_vector_components = n_output_vector_components
        self.time_history = time_history
        self.time_future = time_future
        self.hidden_channels = hidden_channels
        self.activation = resolve(activation)
        insize = time_history * (n_input_scalar_components + n_input_vector_components * 2)
        n_channels = hidden_channels
        self.image_proj = ConvBlock(insize, n_channels, activation=activation)
        self.down = nn.ModuleList([Down(n_channels * m, n_channels * 2 * m, activation=activation) for m in (1, 2, 4, 8)])
        self.up = nn.ModuleList([Up(n_channels * 2 * m, n_channels * m, activation=activation) for m in (8, 4, 2, 1)])
        out_channels = time_future * (n_output_scalar_components + n_output_vector_components * 2)
        self.final = _conv_param(nn.Conv2d(n_channels, out_channels, kernel_size=(3, 3), padding=(1, 1)))

    def forward(self, x):
        assert x.dim() == 5
        orig_shape = x.shape
        x = x.reshape(x.size(0), -1, *x.shape[3:])
        h = self.image_proj.forward_nhwc(ops.to_nhwc(x.float(), _pad16(x.shape[1])))
        x1 = self.down[0].forward_nhwc(h)
        x2 = self.down[1].forward_nhwc(x1)
        x3 = self.down[2].forward_nhwc(x2)
        x4 = self.down[3].forward_nhwc(x3)
        y = self.up[0].forward_nhwc(x4, x3)
        y = self.up[1].forward_nhwc(y, x2)
        y = self.up[2].forward_nhwc(y, x1)
        y = self.up[3].forward_nhwc(y, h)
        out = _conv(y, self.final, out_nchw=True)
        return out.reshape(orig_shape[0], -1, (self.n_output_scalar_components + self.n_output_vector_components * 2),
                           *orig_shape[3:])


class FullResnetConvBlock(ConvBlock):
    def forward_nhwc(self, x):
        h = self._norm_act(_conv(x, self.conv1), self.norm1)
        return self._norm_act(_conv(h, self.conv2), self.norm2, addend=x)          # h + x


class PartialResnetConvBlock(ConvBlock):
    """Use this module if the in and output channels are not the same."""

    def forward_nhwc(self, x):
        h = self._norm_act(_conv(x, self.conv1), self.norm1)                          # changing the channels
        return self._norm_act(_conv(h, self.conv2), self.norm2, addend=h)            # h + act(norm2(conv2(h)))


class DWTBlock(nn.Module):
    def __init__(self, J, out_channels, mode="zero", wave="haar") -> None:
        super().__init__()
        if mode != "zero" or wave != "haar":
            raise NotImplementedError("the B200 kernels implement mode='zero', wave='haar' (the reference's defaults)")
        self.J = J
        self.out_channels = out_channels
        self.xfm = _HaarFilterBuffers(["h0_col", "h1_col", "h0_row", "h1_row"])
        self.ifm = _HaarFilterBuffers(["g0_col", "g1_col", "g0_row", "g1_row"])

    def forward_nhwc(self, x):
        if self.J not in (0, 1):
            raise NotImplementedError("in-network DWTBlock supports J in {0, 1}")
        return ops.dwtblock_act(x, self.J, self.out_channels)

    def forward(self, x):
        return ops.dwtblock(x.float(), self.J, self.out_channels)


class Down_G(nn.Module):
    def __init__(self, in_channels, out_channels, num_groups=1, norm: bool = True, activation="gelu", dwt_encoder=False,
                 no_down_up=False, dwt_mode="zero", dwt_wave="haar") -> None:
        super().__init__()
        if dwt_encoder:
            self.down = DWTBlock(J=1 if not no_down_up else 0, out_channels=out_channels, mode=dwt_mode, wave=dwt_wave)
        else:
            self.conv = PartialResnetConvBlock(in_channels, out_channels, num_groups, norm, activation)
            self.pool = nn.AvgPool2d(2) if not no_down_up else nn.Identity()
        self.dwt_encoder = dwt_encoder

    def forward_nhwc(self, x, finest_level: bool = False):
        if self.dwt_encoder:
            return self.down.forward_nhwc(x)
        h = x
        if not finest_level and isinstance(self.pool, nn.AvgPool2d):
            n, hh, ww, c = x.shape
            if hh % 2 or ww % 2:                      # AvgPool2d floors: drop the odd row / column first
                x = x[:, : hh - hh % 2, : ww - ww % 2, :].contiguous()
            h = ops.dwtblock_act(x, 1, c)             # LL/2 == 2x2 average
        return self.conv.forward_nhwc(h)

    def forward(self, x: torch.Tensor, finest_level: bool = False):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), finest_level))


class Up_G(nn.Module):
    def __init__(self, in_channels, out_channels, num_groups=1, norm: bool = True, activation="gelu",
                 up_fct="interpolate_nearest", n_extra_resnet_layers=0, no_skip_connection=False, no_down_up=False,
                 dwt_encoder=False, crop_finest=False) -> None:
        super().__init__()
        if up_fct == "conv":
            raise NotImplementedError("up_fct='conv' (ConvTranspose2d) is not part of the B200 hot path")
        elif up_fct == "interpolate_nearest":
            self.up_conv_channel_dim = _conv_param(nn.Conv2d(in_channels, in_channels // 2, kernel_size=3, padding=1))
        self.conv = PartialResnetConvBlock(in_channels, out_channels, num_groups, norm, activation)
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.up_fct = up_fct
        self.resnet_list = nn.ModuleList([FullResnetConvBlock(out_channels, out_channels, num_groups, norm, activation)
                                          for _ in range(n_extra_resnet_layers)])
        self.no_skip_connection = no_skip_connection
        self.no_down_up = no_down_up
        self.dwt_encoder = dwt_encoder
        self._crop_finest = crop_finest        # wmh only (wmh/model.py:146-155)

    def forward_nhwc(self, x1, x2, finest_level: bool = False):
        h = _conv(x1, self.up_conv_channel_dim)
        if not self.no_down_up:
            h = ops.upsample2x(h)
        if self._crop_finest and finest_level:
            if self.dwt_encoder:                       # 13 -> 26 -> crop to 25 (drop the first row / column)
                h = h[:, 1:, 1:, :]
            else:                                      # 12 -> 24 -> replicate-pad to 25 on the left / top
                h = torch.cat([h[:, :1], h], dim=1)
                h = torch.cat([h[:, :, :1], h], dim=2)
        if self.no_skip_connection:
            x2 = torch.zeros_like(x2)
        h = torch.cat([x2, h], dim=3)
        h = self.conv.forward_nhwc(h)
        for resnet in self.resnet_list:
            h = resnet.forward_nhwc(h)
        return h

    def forward(self, x1: torch.Tensor, x2: torch.Tensor, finest_level: bool = False):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x1), ops.to_nhwc(x2), finest_level))


class _UnetbaseGCore(nn.Module):
    """Shared body of pdearena's and wmh's `Unetbase_G` (the wmh file is a modified copy, wmh/model.py:1-3)."""

    def _build(self, insize, n_channels, out_channels, activation, final_sigmoid, crop_finest):
        dwt = self.dwt_encoder
        kw = dict(activation=activation, dwt_encoder=dwt, no_down_up=self.no_down_up, dwt_mode=self.dwt_mode,
                  dwt_wave=self.dwt_wave)
        self.image_proj_list = nn.ModuleList([])
        down_in_channels = [n_channels, n_channels * 2, n_channels * 4, n_channels * 8]
        self.down = nn.ModuleList([Down_G(c, 2 * c, **kw) for c in down_in_channels])
        up_out_channels = [n_channels * 8, n_channels * 4, n_channels * 2, n_channels]
        self.up = nn.ModuleList([
            Up_G(2 * c, c, activation=activation, up_fct=self.up_fct, n_extra_resnet_layers=self.n_extra_resnet_layers,
                 no_skip_connection=self.no_skip_connection, no_down_up=self.no_down_up, dwt_encoder=dwt,
                 crop_finest=crop_finest) for c in up_out_channels])
        for j, down_in_ch in enumerate(down_in_channels):
            if self.multi_res_loss or self.sequ_mode or j == 0:
                self.image_proj_list.append(PartialResnetConvBlock(insize, down_in_ch, activation=activation))
            else:
                self.image_proj_list.append(nn.Identity())
        self.n_levels = len(self.down)
        self.final_list = nn.ModuleList([])
        for j, up_out_ch in enumerate(up_out_channels):
            if self.multi_res_loss or self.sequ_mode or j == self.n_levels - 1:
                conv = _conv_param(nn.Conv2d(up_out_ch, out_channels, kernel_size=(3, 3), padding=(1, 1)))
                self.final_list.append(nn.Sequential(conv, nn.Sigmoid()) if final_sigmoid else conv)
            else:
                self.final_list.append(nn.Identity())
        self._final_sigmoid = final_sigmoid
        self._wmh_finest = crop_finest

    def _final(self, j, h):
        mod = self.final_list[j]
        conv = mod[0] if self._final_sigmoid else mod
        out = _conv(h, conv, out_nchw=True)                # fp32 NCHW straight from TMEM
        return torch.sigmoid(out) if self._final_sigmoid else out

    def _run(self, x, n_levels_used):
        """x: NCHW fp32 [B, insize, H, W] -> list of per-level NCHW fp32 outputs (or the last one)."""
        h = self.image_proj_list[self.n_levels - n_levels_used].forward_nhwc(ops.to_nhwc(x.float(), _pad16(x.shape[1])))
        skip = [h]
        for i in list(range(self.n_levels))[-n_levels_used:]:
            h = self.down[i].forward_nhwc(h)
            if i != self.n_levels - 1:
                skip.append(h)
        outs = []
        for j in range(n_levels_used):
            s = skip.pop()
            h = self.up[j].forward_nhwc(h, s, finest_level=(self._wmh_finest and j == 0))
            if self.multi_res_loss:
                outs.append(self._final(j, h))
        if self.multi_res_loss:
            return outs
        return self._final(n_levels_used - 1, h)


class Unetbase_G(_UnetbaseGCore):
    """pdearena `Unetbase_G` (twod_unetbase.py:254-396): forward(x[B,T,C,H,W], n_levels_used=None)."""

    def __init__(self, n_input_scalar_components: int, n_input_vector_components: int, n_output_scalar_components: int,
                 n_output_vector_components: int, time_history: int, time_future: int, hidden_channels: int,
                 activation="gelu", dwt_encoder=False, up_fct="interpolate_nearest", n_extra_resnet_layers=0,
                 multi_res_loss=False, sequ_mode=False, no_skip_connection=False, no_down_up=False, dwt_mode="zero",
                 dwt_wave="haar") -> None:
        super().__init__()
        self.n_input_scalar_components = n_input_scalar_components
        self.n_input_vector_components = n_input_vector_components
        self.n_output_scalar_components = n_output_scalar_components
        self.n_output_vector_components = n_output_vector_components
        self.time_history = time_history
        self.time_future = time_future
        self.hidden_channels = hidden_channels
        self.activation = resolve(activation)
        self.dwt_encoder, self.up_fct, self.n_extra_resnet_layers = dwt_encoder, up_fct, n_extra_resnet_layers
        self.multi_res_loss, self.sequ_mode, self.no_skip_connection = multi_res_loss, sequ_mode, no_skip_connection
        self.no_down_up, self.dwt_mode, self.dwt_wave = no_down_up, dwt_mode, dwt_wave
        insize = time_history * (n_input_scalar_components + n_input_vector_components * 2)
        out_channels = time_future * (n_output_scalar_components + n_output_vector_components * 2)
        self._build(insize, hidden_channels, out_channels, activation, final_sigmoid=False, crop_finest=False)

    def forward(self, x, n_levels_used=None):
        if n_levels_used is None:
            n_levels_used = self.n_levels
        assert x.dim() == 5
        orig_shape = x.shape
        nc = self.n_output_scalar_components + self.n_output_vector_components * 2
        out = self._run(x.reshape(x.size(0), -1, *x.shape[3:]), n_levels_used)
        if self.multi_res_loss:
            return [o.reshape(o.shape[0], -1, nc, *o.shape[2:]) for o in out]
        return out.reshape(orig_shape[0], -1, nc, *orig_shape[3:])
