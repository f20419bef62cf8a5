"""Drop-in for diff_mnist/torch_ddpm/ddpm/models/unet/layers.py (guided-diffusion style blocks), backed by the
sm_100a kernels.  Same class names, constructor signatures and state_dict keys as the reference (file:line):
SiLU :11-13, GroupNorm32 :16-18, normalization :92-98, timestep_embedding :101-118, TimestepEmbedSequential
:178-190, Upsample :195-222, Downsample :225-247, ResBlock :250-338 (incl. `use_scale_shift_norm`),
AttentionBlock :341-368, QKVAttention :371-391.

Public `forward`s keep NCHW fp32; `forward_nhwc` is the internal NHWC bf16 path the containers use.
`GroupNorm32` computes in fp32 as the reference does (:16-18): statistics and the affine are fp32 in the fused
kernel.  Gradient checkpointing (`use_checkpoint`) is accepted and ignored (the reference never enables it,
mnist_diff/unet.py:15).
"""
from __future__ import annotations

import math
from abc import abstractmethod

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from .. import ops
from ..diff_cifar.model import _conv_param


class SiLU(nn.Module):
    def forward(self, x):
        return x * th.sigmoid(x)


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def conv_nd(dims, *args, **kwargs):
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    elif dims == 2:
        return _conv_param(nn.Conv2d(*args, **kwargs))
    raise ValueError(f"unsupported dimensions: {dims}")


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def zero_module(module, active=True):
    if active:
        for p in module.parameters():
            p.detach().zero_()
    return module


def normalization(channels):
    return GroupNorm32(32, channels)


def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = th.exp(-math.log(max_period) * th.arange(start=0, end=half, dtype=th.float32, device=timesteps.device) / half)   # built on the device: no H2D copy inside a CUDA-graph capture
    args = timesteps[:, None].float() * freqs[None]
    embedding = th.cat([th.cos(args), th.sin(args)], dim=-1)
    if dim % 2:
        embedding = th.cat([embedding, th.zeros_like(embedding[:, :1])], dim=-1)
    return embedding


class TimestepBlock(nn.Module):
    @abstractmethod
    def forward(self, x, emb):
        """Apply the module to `x` given `emb` timestep embeddings."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb):
        for layer in self:
            x = layer(x, emb) if isinstance(layer, TimestepBlock) else layer(x)
        return x

    def forward_nhwc(self, x, emb):
        for layer in self:
            if isinstance(layer, nn.Identity):
                continue
            x = layer.forward_nhwc(x, emb) if isinstance(layer, TimestepBlock) else layer.forward_nhwc(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        assert dims == 2, "the B200 path is 2-D"
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, channels, channels, 3, padding=1)

    def forward_nhwc(self, x):
        assert x.shape[3] == self.channels
        x = ops.upsample2x(x)
        return ops.conv(x, self.conv.weight, self.conv.bias) if self.use_conv else x

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2):
        super().__init__()
        assert dims == 2, "the B200 path is 2-D"
        self.channels, self.use_conv, self.dims = channels, use_conv, dims
        if use_conv:
            self.op = conv_nd(dims, channels, channels, 3, stride=2, padding=1)
        else:
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)

    def forward_nhwc(self, x):
        assert x.shape[3] == self.channels
        if self.use_conv:      # nn.Conv2d(C, C, 3, stride=2, padding=1) (layers.py:238): TMA traversal stride 2
            return ops.conv(x, self.op.weight, self.op.bias, stride=2)
        n, h, w, c = x.shape
        if h % 2 or w % 2:
            x = x[:, : h - h % 2, : w - w % 2, :].contiguous()
        return ops.dwtblock_act(x, 1, c)          # LL/2 == 2x2 average

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False, use_scale_shift_norm=False,
                 dims=2, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.in_layers = nn.Sequential(
            normalization(channels),
            SiLU(),
            conv_nd(dims, channels, self.out_channels, 3, padding=1),
        )
        self.emb_layers = nn.Sequential(
            SiLU(),
            linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels),
        )
        self.out_layers = nn.Sequential(
            normalization(self.out_channels),
            SiLU(),
            nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)),
        )
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward_nhwc(self, x, emb):
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        gn2, drop, conv2 = self.out_layers[0], self.out_layers[2], self.out_layers[3]
        a1 = ops.gn_act(x, gn1.weight, gn1.bias, gn1.num_groups, act="silu", eps=gn1.eps)
        pre = self.__dict__.pop("_emb_pre", None)       # one-shot: set by precompute_emb_layers for this forward only
        emb_out = pre if pre is not None else self.emb_layers(emb).float()
        p = drop.p if self.training else 0.0
        if self.use_scale_shift_norm:
            h = ops.conv(a1, conv1.weight, conv1.bias)
            scale, shift = th.chunk(emb_out, 2, dim=1)
            # SiLU(GN(h) * (1 + scale) + shift) -> dropout, one kernel (layers.py:330-334)
            a2 = ops.gn_act(h, gn2.weight, gn2.bias, gn2.num_groups, act="silu", eps=gn2.eps, dropout_p=p,
                            scale=scale.contiguous(), shift=shift.contiguous())
        else:
            h = ops.conv(a1, conv1.weight, conv1.bias, rowadd=emb_out.contiguous())
            a2 = ops.gn_act(h, gn2.weight, gn2.bias, gn2.num_groups, act="silu", eps=gn2.eps, dropout_p=p)
        skip = self.skip_connection
        if isinstance(skip, nn.Identity):
            return ops.conv(a2, conv2.weight, conv2.bias, residual=x)
        if skip.kernel_size == (1, 1):          # 1x1 skip rides as extra K slices of conv2
            return ops.conv(a2, conv2.weight, conv2.bias + skip.bias, a2=x, w2=skip.weight)
        return ops.conv(a2, conv2.weight, conv2.bias, residual=ops.conv(x, skip.weight, skip.bias))

    def forward(self, x, emb):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), emb))


def precompute_emb_layers(pairs):
    """`emb_layers(emb)` (SiLU + Linear, layers.py:305-312 of the reference) of every ResBlock that is about to run, as ONE
    batched launch (ops.rowlin_batch) instead of four framework kernels per block (sigmoid, mul, GEMM, bias add) and their
    backward.  `pairs`: [(ResBlock, emb)]; each block picks its row up once in forward_nhwc."""
    pairs = [(b, e) for b, e in pairs if isinstance(b, ResBlock)]
    if not pairs:
        return []
    lins = [b.emb_layers[1] for b, _ in pairs]
    outs = ops.rowlin_batch([e for _, e in pairs], [l.weight for l in lins], [l.bias for l in lins], silu=True)
    for (b, _), o in zip(pairs, outs):
        b.__dict__["_emb_pre"] = o
    return [b for b, _ in pairs]


def clear_emb_layers(blocks):
    """Drop rows that were precomputed but not consumed (a forward that raised half-way): a block called on its own later must
    evaluate its own `emb_layers`."""
    for b in blocks or ():
        b.__dict__.pop("_emb_pre", None)


def batched_time_embed(seqs, sin_emb):
    """[seq(sin_emb) for seq in seqs] for `nn.Sequential(linear, SiLU, linear)` time-embedding MLPs: two batched launches."""
    h = ops.rowlin_batch([sin_emb] * len(seqs), [s[0].weight for s in seqs], [s[0].bias for s in seqs], silu=False)
    return ops.rowlin_batch(h, [s[2].weight for s in seqs], [s[2].bias for s in seqs], silu=True)


class QKVAttention(nn.Module):
    """Kept for API parity; the fused block below calls PyTorch SDPA (attention is out of scope)."""

    def forward(self, qkv):
        ch = qkv.shape[1] // 3
        q, k, v = th.split(qkv, ch, dim=1)
        scale = 1 / math.sqrt(math.sqrt(ch))
        weight = th.einsum("bct,bcs->bts", q * scale, k * scale)
        weight = th.softmax(weight.float(), dim=-1).type(weight.dtype)
        return th.einsum("bts,bcs->bct", weight, v)


class AttentionBlock(nn.Module):
    def __init__(self, channels, num_heads=1, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.num_heads = num_heads
        self.use_checkpoint = use_checkpoint
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.attention = QKVAttention()
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def forward_nhwc(self, x):
        n, h, w, c = x.shape
        heads, ch = self.num_heads, c // self.num_heads
        y = ops.gn_act(x, self.norm.weight, self.norm.bias, self.norm.num_groups, act="none", eps=self.norm.eps)
        qkv = ops.conv(y, self.qkv.weight.unsqueeze(-1), self.qkv.bias)            # [N,H,W,3C], channels = (head, q|k|v, ch)
        qkv = qkv.reshape(n, h * w, heads, 3, ch).permute(3, 0, 2, 1, 4)           # [3, N, heads, T, ch]
        o = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2], scale=1.0 / math.sqrt(ch))   # [N, heads, T, ch]
        o = o.permute(0, 2, 1, 3).reshape(n, h, w, c)
        return ops.conv(o, self.proj_out.weight.unsqueeze(-1), self.proj_out.bias, residual=x)

    def forward(self, x):
        b, c, *spatial = x.shape
        x4 = x.reshape(b, c, spatial[0], -1) if len(spatial) != 2 else x
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x4))).reshape(b, c, *spatial)
