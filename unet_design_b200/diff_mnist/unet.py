"""Drop-in for diff_mnist/mnist_diff/unet.py (`get_unet_wavelet` :11-68, `UNet_wavelet` :75-556) and
diff_mnist/mnist_diff/models.py (`DTWBlock` :12-82), backed by the sm_100a kernels.

Same constructor arguments, attribute names (`time_embed_list`, `input_blocks`, `middle_block`, `out_f_list`,
`out_upsample_list`, `out_activation_list`, `out_reduce_channels_list` -- the staged-training freeze code walks them,
diff_mnist/main.py:258-308) and state_dict keys.  `forward(x, t, y=None, n_levels_used=-1, u_net_norm=False)`
returns `(tensor or list coarse->fine, norms or None)` exactly as the reference, including its quirk of forcing
`model_out_passed_on = True` (:457): at every level the state is collapsed to `out_channels` by GN+SiLU + a 1x1
conv, emitted, and channel-tiled back (:498-510).
"""
from __future__ import annotations

import torch
import torch as th
from torch import nn

from .. import ops
from ..diff_cifar.model import DTWBlock, _PyramidView, _conv_param  # DTWBlock: same logic as mnist_diff/models.py:12-82
from .layers import (AttentionBlock, Downsample, ResBlock, SiLU, TimestepEmbedSequential, Upsample, batched_time_embed,
                     clear_emb_layers, linear, normalization, precompute_emb_layers, timestep_embedding)


def compute_norm(tensor):
    """diff_mnist/utils.py:59-68 (layout independent: works on NHWC as well)."""
    dim = 1
    for sub_dim in tensor.shape[1:]:
        dim *= sub_dim
    return torch.linalg.norm(torch.flatten(tensor.float(), start_dim=1), dim=1) * (1 / dim)


def get_unet_wavelet(image_size, image_channels, num_channels=32, dropout=0.0, num_res_blocks=2, dwt_encoder=False,
                     multi_res_loss=False, model_out_passed_on=False, avg_pool_down=False):
    num_heads = 4
    num_heads_upsample = -1
    attention_resolutions = "168"
    channel_mult = {256: (1, 1, 2, 2, 4, 4), 64: (2, 2, 2, 2), 32: (2, 2, 2, 2), 28: (1, 2, 2), 16: (1, 2, 2, 2),
                    8: (1, 2, 2), 4: (1, 1, 1), 2: (1, 2), 1: (1,)}.get(image_size)
    if channel_mult is None:
        raise ValueError(f"unsupported image size: {image_size}")
    attention_ds = [image_size // int(res) for res in attention_resolutions.split(",")]
    return UNet_wavelet(in_channels=image_channels, model_channels=num_channels, out_channels=image_channels,
                        num_res_blocks=num_res_blocks, attention_resolutions=tuple(attention_ds), dropout=dropout,
                        channel_mult=channel_mult, num_classes=None, use_checkpoint=False, num_heads=num_heads,
                        num_heads_upsample=num_heads_upsample, use_scale_shift_norm=True, dwt_encoder=dwt_encoder,
                        multi_res_loss=multi_res_loss, model_out_passed_on=model_out_passed_on,
                        conv_resample=not avg_pool_down)


class _TileToNhwc(torch.autograd.Function):
    """`h.repeat(1, n/C + 1, 1, 1)[:, :n]` of an NCHW fp32 tensor written straight into NHWC bf16 (unet.py:498-510);
    backward folds the replicas back."""

    @staticmethod
    def forward(ctx, x, out_channels):
        n, c, h, w = x.shape
        out = torch.empty((n, h, w, out_channels), dtype=torch.bfloat16, device=x.device)
        ops.dwtblock_nhwc(x.contiguous(), 0, out)
        ctx.meta = (c, h, w)
        return out

    @staticmethod
    def backward(ctx, g):
        from .._lib import ops as raw
        c, h, w = ctx.meta
        g_nchw = raw().nhwc_to_nchw(ops._dense_nhwc(g))
        return raw().dwtblock_bwd(g_nchw, c, h, w, 0), None


class UNet_wavelet(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False, num_heads=1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, dwt_encoder=False, multi_res_loss=False,
                 model_out_passed_on=False):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        if num_classes is not None:
            raise NotImplementedError("class conditioning is not on the hot path (the reference passes num_classes=None)")
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions, self.dropout = num_res_blocks, attention_resolutions, dropout
        self.channel_mult, self.conv_resample, self.num_classes = channel_mult, conv_resample, num_classes
        self.use_checkpoint, self.num_heads, self.num_heads_upsample = use_checkpoint, num_heads, num_heads_upsample
        self.n_levels = len(channel_mult)
        self.dwt_encoder, self.multi_res_loss, self.model_out_passed_on = dwt_encoder, multi_res_loss, model_out_passed_on

        time_embed_dim = model_channels * 4
        self.time_embed_list = nn.ModuleList([nn.Sequential(
            linear(model_channels, time_embed_dim), SiLU(), linear(time_embed_dim, time_embed_dim),
        ) for _ in range(self.n_levels)])

        ch = model_channels * channel_mult[0]
        ds = 1
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(DTWBlock(J=0, out_channels=ch))])
        input_block_chans = [ch]
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                if self.dwt_encoder:
                    ch = int(mult * model_channels)
                    self.input_blocks.append(TimestepEmbedSequential(DTWBlock(J=0, out_channels=ch)))
                else:
                    layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                       use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                    ch = mult * model_channels
                    if ds in attention_resolutions:
                        layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads))
                    self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                if self.dwt_encoder:
                    ch_downsample = int(channel_mult[level + 1] * model_channels)
                    self.input_blocks.append(TimestepEmbedSequential(DTWBlock(J=1, out_channels=ch_downsample)))
                    input_block_chans.append(ch_downsample)
                else:
                    self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims)))
                    input_block_chans.append(ch)
                ds *= 2

        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
            AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
        )

        self.out_f_list = nn.ModuleList([nn.ModuleList() for _ in range(len(channel_mult))])
        self.out_upsample_list = nn.ModuleList([nn.ModuleList() for _ in range(len(channel_mult))])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + input_block_chans.pop(), time_embed_dim, dropout, out_channels=model_channels * mult,
                                   dims=dims, use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads_upsample))
                self.out_f_list[level].append(TimestepEmbedSequential(*layers))
                if i == num_res_blocks:
                    if level:
                        self.out_upsample_list[level].append(TimestepEmbedSequential(Upsample(ch, conv_resample, dims=dims)))
                        ds //= 2
                    else:
                        self.out_upsample_list[level].append(TimestepEmbedSequential(nn.Identity()))

        self.out_reduce_channels_list = nn.ModuleList([
            _conv_param(nn.Conv2d(in_channels=ch, out_channels=out_channels, kernel_size=1, stride=1))
            for _ in range(len(channel_mult))])
        self.out_activation_list = nn.ModuleList([nn.Sequential(normalization(ch), SiLU()) for _ in range(len(channel_mult))])
        self._chmap_cache = {}

    def compute_time_embedding(self, timesteps, level, y=None):
        if level == -1:
            level = 0
        return self.time_embed_list[level](timestep_embedding(timesteps, self.model_channels))

    # ---- internals
    def _materialise(self, pyramid, view: _PyramidView):
        base = pyramid[view.level]
        n, _, h, w = base.shape
        key = (tuple(view.chmap), base.device)
        cm = self._chmap_cache.get(key)
        if cm is None:
            cm = torch.tensor(view.chmap, dtype=torch.int32, device=base.device)
            self._chmap_cache[key] = cm
        out = torch.empty((n, h, w, len(view.chmap)), dtype=torch.bfloat16, device=base.device)
        return ops.dwtblock_nhwc(base, 0, out, cm)

    def forward(self, x, t, y=None, n_levels_used=-1, u_net_norm=False):
        try:
            return self._forward(x, t, y, n_levels_used, u_net_norm)
        finally:
            clear_emb_layers(self.__dict__.pop("_emb_blocks", None))

    def _forward(self, x, t, y=None, n_levels_used=-1, u_net_norm=False):
        if n_levels_used == -1:
            n_levels_used = len(self.channel_mult)
        timesteps = t.squeeze()
        if timesteps.dim() == 0:
            timesteps = timesteps[None]
        assert y is None, "must specify y if and only if the model is class-conditional"
        norms = {"down": {k: [] for k in range(self.n_levels)}, "middle": [],
                 "up": {k: [] for k in range(self.n_levels)}} if u_net_norm else None
        x = x.float().contiguous()
        lowest_level = len(self.channel_mult) - n_levels_used
        if u_net_norm:
            norms["down"][lowest_level].append(compute_norm(x))
        upper_range = n_levels_used * (self.num_res_blocks + 1) - 1
        ins = [self.input_blocks[0]] + list(self.input_blocks[::-1][:upper_range][::-1])
        start_level = self.n_levels - n_levels_used

        # Time-embedding path of the whole forward in three batched launches: the reference re-evaluates
        # `compute_time_embedding` (sin/cos + Linear, SiLU, Linear) before every block (unet.py:440-520) and every ResBlock
        # runs its own SiLU + Linear; here each level's embedding is computed once and all ResBlock rows together.
        up_levels = list(range(len(self.channel_mult)))[::-1][:n_levels_used]
        plan = []                                   # (ResBlock, level) in execution order
        if not self.dwt_encoder:
            for i, module in enumerate(ins):
                if i > 0:
                    plan += [(blk, start_level + int((i - 1) / (self.num_res_blocks + 1))) for blk in module]
        plan += [(blk, self.n_levels - 1) for blk in self.middle_block]
        for level in up_levels:
            for out_block in self.out_f_list[level]:
                plan += [(blk, level) for blk in out_block]
        plan = [(blk, max(lv, 0)) for blk, lv in plan if isinstance(blk, ResBlock)]
        lv_used = sorted({lv for _, lv in plan})
        embs = dict(zip(lv_used, batched_time_embed([self.time_embed_list[lv] for lv in lv_used],
                                                    timestep_embedding(timesteps, self.model_channels))))
        self.__dict__["_emb_blocks"] = precompute_emb_layers([(blk, embs[lv]) for blk, lv in plan])
        time_emb = lambda level: embs.get(max(level, 0))     # noqa: E731  (None where no ResBlock consumes it)

        hs = []
        if self.dwt_encoder:
            # every encoder tensor is a channel tile of the image pyramid LL_l(x)/2^l: compute the pyramid only
            pyramid = {0: x}
            view = _PyramidView(0, list(range(x.shape[1])))
            for i, module in enumerate(ins):
                blk = module[0]
                if blk.J == 1:
                    src = pyramid[view.level]
                    pyramid[view.level + 1] = ops.dwtblock(src, 1, src.shape[1])
                    view = _PyramidView(view.level + 1, view.chmap)
                view = view.tiled(blk.out_channels)
                hs.append(view)
                if u_net_norm:
                    level = start_level + int((i - 1) / (self.num_res_blocks + 1))
                    norms["down"][level].append(compute_norm(self._materialise(pyramid, view)))
            fetch = lambda v: self._materialise(pyramid, v)
            h = fetch(hs[-1])
        else:
            h = None
            for i, module in enumerate(ins):
                level = start_level + int((i - 1) / (self.num_res_blocks + 1))
                emb = time_emb(level)
                if i == 0:
                    blk = module[0]
                    h = ops.dwtblock_nhwc(x, 0, torch.empty((x.shape[0], x.shape[2], x.shape[3], blk.out_channels),
                                                            dtype=torch.bfloat16, device=x.device))
                else:
                    h = module.forward_nhwc(h, emb)
                if u_net_norm:
                    norms["down"][level].append(compute_norm(h))
                hs.append(h)
            fetch = lambda v: v

        h = self.middle_block.forward_nhwc(h, time_emb(self.n_levels - 1))
        if u_net_norm:
            norms["middle"].append(compute_norm(h))

        self.model_out_passed_on = True            # the reference forces this (unet.py:457)
        model_out_list = []
        out = None
        level_inv = 0
        for i, level in enumerate(list(range(len(self.channel_mult)))[::-1][:n_levels_used]):
            for out_block in self.out_f_list[level]:
                cat_in = th.cat([h, fetch(hs.pop())], dim=3)
                h = out_block.forward_nhwc(cat_in, time_emb(level))
                if u_net_norm:
                    norms["up"][level].append(compute_norm(h))
            gn = self.out_activation_list[i][0]
            a = ops.gn_act(h, gn.weight, gn.bias, gn.num_groups, act="silu", eps=gn.eps)
            n_state_channels = a.shape[3]
            red = self.out_reduce_channels_list[i]
            out = ops.conv(a, red.weight, red.bias, out_nchw=True)          # fp32 NCHW [N, out_channels, H, W]
            if self.multi_res_loss:
                model_out_list.append(out)
            if u_net_norm:
                norms["up"][level].append(compute_norm(out))
            last = level_inv == n_levels_used - 1
            if self.multi_res_loss or not last:
                h = _TileToNhwc.apply(out, n_state_channels)
                if u_net_norm:
                    norms["up"][level].append(compute_norm(h))
            if not last:
                h = self.out_upsample_list[level][0].forward_nhwc(h, None)      # Upsample / Identity: no embedding use
                if u_net_norm:
                    norms["up"][level].append(compute_norm(h))
            level_inv += 1
        if self.multi_res_loss:
            return model_out_list, norms
        return out.type(x.dtype), norms


# --------------------------------------------------------------------------------------------------
# The plain guided-diffusion U-Net of diff_mnist (torch_ddpm/ddpm/models/unet/unet.py:14-311) and its factory
# (torch_ddpm/ddpm/models/utils.py:5-53).  Same constructor / forward signatures and state_dict keys
# (`time_embed`, `input_blocks`, `middle_block`, `output_blocks`, `out`, `out_reduce_channels`).
# --------------------------------------------------------------------------------------------------
from .layers import conv_nd  # noqa: E402
import torch.nn.functional as F  # noqa: E402


class UNetModel(nn.Module):
    def __init__(self, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout=0,
                 channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None, use_checkpoint=False, num_heads=1,
                 num_heads_upsample=-1, use_scale_shift_norm=False):
        super().__init__()
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.locals = [in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout, channel_mult,
                       conv_resample, dims, num_classes, use_checkpoint, num_heads, num_heads_upsample, use_scale_shift_norm]
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions, self.dropout = num_res_blocks, attention_resolutions, dropout
        self.channel_mult, self.conv_resample, self.num_classes = channel_mult, conv_resample, num_classes
        self.use_checkpoint, self.num_heads, self.num_heads_upsample = use_checkpoint, num_heads, num_heads_upsample
        self.n_levels = len(channel_mult)
        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), SiLU(), linear(time_embed_dim, time_embed_dim))
        if self.num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, time_embed_dim)
        ch = model_channels * channel_mult[0]
        input_block_chans = [ch]
        ds = 1
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, ch, 3, padding=1))])
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [ResBlock(ch, time_embed_dim, dropout, out_channels=mult * model_channels, dims=dims,
                                   use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                input_block_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims)))
                input_block_chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
            AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads),
            ResBlock(ch, time_embed_dim, dropout, dims=dims, use_checkpoint=use_checkpoint,
                     use_scale_shift_norm=use_scale_shift_norm),
        )
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [ResBlock(ch + input_block_chans.pop(), time_embed_dim, dropout, out_channels=model_channels * mult,
                                   dims=dims, use_checkpoint=use_checkpoint, use_scale_shift_norm=use_scale_shift_norm)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(AttentionBlock(ch, use_checkpoint=use_checkpoint, num_heads=num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
        self.out = nn.Sequential(normalization(ch), SiLU())
        self.out_reduce_channels = _conv_param(nn.Conv2d(in_channels=ch, out_channels=out_channels, kernel_size=1, stride=1))

    def forward(self, x, t, y=None, n_levels_used: int = -1):
        try:
            return self._forward(x, t, y, n_levels_used)
        finally:
            clear_emb_layers(self.__dict__.pop("_emb_blocks", None))

    def _forward(self, x, t, y=None, n_levels_used: int = -1):
        if n_levels_used == -1:
            n_levels_used = len(self.channel_mult)
        timesteps = t.squeeze()
        if timesteps.dim() == 0:
            timesteps = timesteps[None]
        assert (y is not None) == (self.num_classes is not None), \
            "must specify y if and only if the model is class-conditional"
        emb = batched_time_embed([self.time_embed], timestep_embedding(timesteps, self.model_channels))[0]
        if self.num_classes is not None:
            assert y.shape == (x.shape[0],)
            emb = emb + self.label_emb(y)
        n_in, n_out = n_levels_used * (self.num_res_blocks + 1), n_levels_used * (self.num_res_blocks + 1) - 1
        running = [m for m in self.input_blocks[1:n_in]] + [self.middle_block] + list(self.output_blocks[:n_out])
        self.__dict__["_emb_blocks"] = precompute_emb_layers([(blk, emb) for m in running for blk in m])   # one launch
        hs = []
        cin = x.shape[1]
        h = ops.to_nhwc(x.float(), (cin + 15) // 16 * 16)
        # the reference's slicing of the block lists for a reduced level count (unet.py:233, :241)
        for i, module in enumerate(self.input_blocks[:n_levels_used * (self.num_res_blocks + 1) + 1 - 1]):
            if i == 0:
                conv = module[0]
                w = conv.weight
                if h.shape[3] != cin:
                    w = F.pad(w, (0, 0, 0, 0, 0, h.shape[3] - cin))
                h = ops.conv(h, w, conv.bias)
            else:
                h = module.forward_nhwc(h, emb)
            hs.append(h)
        h = self.middle_block.forward_nhwc(h, emb)
        for module in self.output_blocks[:n_levels_used * (self.num_res_blocks + 1) - 1]:
            h = module.forward_nhwc(th.cat([h, hs.pop()], dim=3), emb)
        gn = self.out[0]
        a = ops.gn_act(h, gn.weight, gn.bias, gn.num_groups, act="silu", eps=gn.eps)
        return ops.conv(a, self.out_reduce_channels.weight, self.out_reduce_channels.bias, out_nchw=True).type(x.dtype)


def get_unet(image_size, image_channels, num_channels=32, dropout=0.0, num_res_blocks=2):
    """torch_ddpm/ddpm/models/utils.py:5-53."""
    channel_mult = {256: (1, 1, 2, 2, 4, 4), 64: (1, 2, 3, 4), 32: (2, 2, 2, 2), 28: (1, 2, 2), 16: (1, 2, 2, 2),
                    8: (1, 2, 2), 4: (1, 2), 2: (1, 2), 1: (1,)}.get(image_size)
    if channel_mult is None:
        raise ValueError(f"unsupported image size: {image_size}")
    attention_ds = [image_size // int(res) for res in "168".split(",")]
    return UNetModel(in_channels=image_channels, model_channels=num_channels, out_channels=image_channels,
                     num_res_blocks=num_res_blocks, attention_resolutions=tuple(attention_ds), dropout=dropout,
                     channel_mult=channel_mult, num_classes=None, use_checkpoint=False, num_heads=4,
                     num_heads_upsample=-1, use_scale_shift_norm=True)
