"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against the reference's own classes.

fp32 PyTorch restatement of the `_G` U-Net family the reference uses for its PDE surrogates and the WMH segmentation
model: pdearena/pdearena/modules/twod_unetbase.py (ConvBlock :12-32, PartialResnetConvBlock :154-161,
FullResnetConvBlock :148-151, DWTBlock :164-193, Down_G :200-218, Up_G :221-251, Unetbase_G :254-396) and wmh/model.py
(the modified copy: odd 25 -> 13 extents with the decoder crop / replicate pad :140-157, sigmoid head :253).
Written block-functionally (one class for both containers) rather than as a copy of either file.

Pin: tests/test_oracle_golden.py replays tests/golden/pdearena_unetbase_g_*.pt and wmh_unetbase_g_*.pt, which
tools/make_golden.py produced by running the reference's classes, through this module (state_dict keys are the same).
Used as the CPU baseline of bench.py's --config c3 / c4 / c5 legs and as the GPU library comparator there.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .pytorch_wavelets_restated import DWTForward, DWTInverse

_ACT = {"gelu": F.gelu, "silu": F.silu, "relu": F.relu}


class ConvBlock(nn.Module):
    """act(GN1(conv3x3)) twice; `mode`: 'plain' -> h2, 'partial' -> h1 + h2, 'full' -> h2 + x."""

    def __init__(self, cin, cout, activation="gelu", mode="plain"):
        super().__init__()
        self.act, self.mode = _ACT[activation], mode
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.norm1 = nn.GroupNorm(1, cout)
        self.norm2 = nn.GroupNorm(1, cout)

    def forward(self, x):
        h1 = self.act(self.norm1(self.conv1(x)))
        h2 = self.act(self.norm2(self.conv2(h1)))
        return h2 if self.mode == "plain" else (h1 + h2 if self.mode == "partial" else h2 + x)


class DWTBlock(nn.Module):
    def __init__(self, J, out_channels):
        super().__init__()
        self.J, self.out_channels = J, out_channels
        self.xfm, self.ifm = DWTForward(J=J), DWTInverse()

    def forward(self, x):
        if self.J == 0:
            return x.repeat(1, self.out_channels // x.shape[1] + 1, 1, 1)[:, :self.out_channels]
        yl, _ = self.xfm(x)
        out = self.ifm((yl, [])) / (2.0 ** self.J)
        if x.shape[1] != self.out_channels:
            out = out.repeat(1, self.out_channels // out.shape[1] + 1, 1, 1)[:, :self.out_channels]
        return out


class _Down(nn.Module):
    def __init__(self, cin, cout, activation, dwt_encoder):
        super().__init__()
        self.dwt_encoder = dwt_encoder
        if dwt_encoder:
            self.down = DWTBlock(1, cout)
        else:
            self.conv = ConvBlock(cin, cout, activation, "partial")

    def forward(self, x):
        return self.down(x) if self.dwt_encoder else self.conv(F.avg_pool2d(x, 2))


class _Up(nn.Module):
    def __init__(self, cin, cout, activation, n_extra, dwt_encoder, wmh):
        super().__init__()
        self.up_conv_channel_dim = nn.Conv2d(cin, cin // 2, 3, padding=1)
        self.conv = ConvBlock(cin, cout, activation, "partial")
        self.resnet_list = nn.ModuleList([ConvBlock(cout, cout, activation, "full") for _ in range(n_extra)])
        self.dwt_encoder, self.wmh = dwt_encoder, wmh

    def forward(self, x1, x2, finest_level=False):
        h = F.interpolate(self.up_conv_channel_dim(x1), scale_factor=2)
        if self.wmh and finest_level:
            h = h[:, :, 1:, 1:] if self.dwt_encoder else F.pad(h, (1, 0, 1, 0), mode="replicate")
        h = self.conv(torch.cat([x2, h], dim=1))
        for r in self.resnet_list:
            h = r(h)
        return h


class Unetbase_G(nn.Module):
    """`wmh=False`: pdearena (x [B,T,C,H,W]); `wmh=True`: wmh/model.py (x [B,2,H,W], sigmoid head, odd-extent handling)."""

    def __init__(self, insize, out_channels, hidden_channels, activation="gelu", dwt_encoder=False, n_extra_resnet_layers=0,
                 multi_res_loss=False, wmh=False, n_out_components=None):
        super().__init__()
        c = hidden_channels
        self.multi_res_loss, self.wmh, self.n_levels, self.n_out_components = multi_res_loss, wmh, 4, n_out_components
        self.down = nn.ModuleList([_Down(c * m, c * 2 * m, activation, dwt_encoder) for m in (1, 2, 4, 8)])
        self.up = nn.ModuleList([_Up(c * 2 * m, c * m, activation, n_extra_resnet_layers, dwt_encoder, wmh) for m in (8, 4, 2, 1)])
        self.image_proj_list = nn.ModuleList([ConvBlock(insize, c * m, activation, "partial") if (multi_res_loss or j == 0)
                                              else nn.Identity() for j, m in enumerate((1, 2, 4, 8))])
        self.final_list = nn.ModuleList([])
        for j, m in enumerate((8, 4, 2, 1)):
            if multi_res_loss or j == 3:
                conv = nn.Conv2d(c * m, out_channels, 3, padding=1)
                self.final_list.append(nn.Sequential(conv, nn.Sigmoid()) if wmh else conv)
            else:
                self.final_list.append(nn.Identity())

    def forward(self, x, n_levels_used=None):
        n_levels_used = n_levels_used or self.n_levels
        shape5 = x.shape if x.dim() == 5 else None
        if shape5 is not None:
            x = x.reshape(x.size(0), -1, *x.shape[3:])
        h = self.image_proj_list[self.n_levels - n_levels_used](x)
        skip = [h]
        for i in list(range(self.n_levels))[-n_levels_used:]:
            h = self.down[i](h)
            if i != self.n_levels - 1:
                skip.append(h)
        outs = []
        for j in range(n_levels_used):
            h = self.up[j](h, skip.pop(), finest_level=(self.wmh and j == 0))
            if self.multi_res_loss:
                outs.append(self.final_list[j](h))
        if not self.multi_res_loss:
            outs = [self.final_list[n_levels_used - 1](h)]
        if shape5 is not None:
            outs = [o.reshape(o.shape[0], -1, self.n_out_components, *o.shape[2:]) for o in outs]
        return outs if self.multi_res_loss else outs[0]


def from_reference_cfg(cfg: dict, wmh: bool = False) -> Unetbase_G:
    """Build from the keyword arguments of the reference constructor (the `cfg` stored in the golden fixtures)."""
    if wmh:
        return Unetbase_G(2, 1, cfg["hidden_channels"], cfg.get("activation", "gelu"), cfg.get("dwt_encoder", False),
                          cfg.get("n_extra_resnet_layers", 0), cfg.get("multi_res_loss", False), wmh=True)
    nin = cfg["n_input_scalar_components"] + 2 * cfg["n_input_vector_components"]
    nout = cfg["n_output_scalar_components"] + 2 * cfg["n_output_vector_components"]
    return Unetbase_G(cfg["time_history"] * nin, cfg["time_future"] * nout, cfg["hidden_channels"], cfg.get("activation", "gelu"),
                      cfg.get("dwt_encoder", False), cfg.get("n_extra_resnet_layers", 0), cfg.get("multi_res_loss", False),
                      n_out_components=nout)
