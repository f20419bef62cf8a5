"""Drop-in for the reference's diff_cifar/model.py, backed by the sm_100a kernels.

Same class names, constructor signatures, `forward` signatures / return structure, attribute names,
ModuleList nesting and `state_dict` keys and shapes as the reference (file:line under
/root/reference): Swish :9-11, TimeEmbedding :14-43, DownSample :46-63, UpSample :66-81,
AttnBlock :84-119, ResBlock :122-169, DTWBlock :253-323, UNetWaveletEnc :326-496, and the commented
plain `UNet` :172-246 as a thin alias.

Public `forward`s take and return the reference's NCHW fp32 tensors.  Inside a model every activation is
NHWC bf16 and every hot op is one of our kernels:

    GroupNorm statistics + (normalise, SiLU, dropout)      -> ops.gn_act      (groupnorm.cu)
    3x3 conv + bias + time-embedding row + 1x1 shortcut /
      identity residual, forward, dgrad and wgrad           -> ops.conv        (conv_fprop.cu, conv_wgrad.cu)
    nearest x2                                              -> ops.upsample2x  (layout.cu)
    Haar pyramid + DTWBlock chain                           -> ops.dwtblock*   (haar.cu)

Conv weights live in fp32 with channels_last strides, i.e. the memory order [Cout, kh, kw, Cin] the
tensor-core kernels consume; `state_dict` / `load_state_dict` are unaffected by strides.
AttnBlock (2.2% of the FLOPs, out of scope per SURVEY.md §2.1) runs on PyTorch's SDPA in bf16.
"""
from __future__ import annotations

import contextlib
import os
import math
from typing import List

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn import init

from .. import ops


class Swish(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(x)


def _conv_param(conv: nn.Conv2d) -> nn.Conv2d:
    """Keep the weight in [Cout, kh, kw, Cin] memory order (channels_last strides of [Cout, Cin, kh, kw])."""
    conv.weight.data = conv.weight.data.contiguous(memory_format=torch.channels_last)
    return conv


class _HaarFilterBuffers(nn.Module):
    """State-dict stand-in for pytorch_wavelets.DWTForward / DWTInverse: same buffer names and shapes
    (the reference's checkpoints carry them), no computation -- the taps are constants of haar.cu."""

    def __init__(self, names):
        super().__init__()
        s = 1.0 / math.sqrt(2.0)
        for name in names:
            taps = torch.tensor([s, s] if name[1] == "0" else [s, -s], dtype=torch.float32)
            self.register_buffer(name, taps.reshape(1, 1, 2, 1) if name.endswith("col") else taps.reshape(1, 1, 1, 2))


class TimeEmbedding(nn.Module):
    def __init__(self, T, d_model, dim):
        assert d_model % 2 == 0
        super().__init__()
        emb = torch.arange(0, d_model, step=2) / d_model * math.log(10000)
        emb = torch.exp(-emb)
        pos = torch.arange(T).float()
        emb = pos[:, None] * emb[None, :]
        emb = torch.stack([torch.sin(emb), torch.cos(emb)], dim=-1).view(T, d_model)
        self.timembedding = nn.Sequential(
            nn.Embedding.from_pretrained(emb),
            nn.Linear(d_model, dim),
            Swish(),
            nn.Linear(dim, dim),
        )
        self.initialize()

    def initialize(self):
        for module in self.modules():
            if isinstance(module, nn.Linear):
                init.xavier_uniform_(module.weight)
                init.zeros_(module.bias)

    def forward(self, t):
        return time_embeddings([self], t)[0]


def time_embeddings(modules, t):
    """[m(t) for m in modules]: the embedding gathers, then both Linear layers of ALL modules as two batched launches
    (Linear, then Swish + Linear) instead of four framework ops per module."""
    embs = [m.timembedding[0](t) for m in modules]
    l1 = [m.timembedding[1] for m in modules]
    l2 = [m.timembedding[3] for m in modules]
    h = ops.rowlin_batch(embs, [l.weight for l in l1], [l.bias for l in l1], silu=False)
    return ops.rowlin_batch(h, [l.weight for l in l2], [l.bias for l in l2], silu=True)


class DownSample(nn.Module):
    """Strided 3x3 conv or 2x2 average pool; only the non-Haar U-Net arm uses it (model.py:369)."""

    def __init__(self, in_ch, type="conv"):
        super().__init__()
        self.type = type
        if type == "conv":
            self.main = _conv_param(nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1))
            self.initialize()
        elif type == "avg_pool":
            self.main = nn.AvgPool2d(2)
        else:
            raise NotImplementedError

    def initialize(self):
        init.xavier_uniform_(self.main.weight)
        init.zeros_(self.main.bias)

    def forward_nhwc(self, x, temb=None):
        if self.type == "conv":
            return ops.conv(x, self.main.weight, self.main.bias, stride=2)     # TMA traversal stride 2 (model.py:52)
        n, h, w, c = x.shape
        if h % 2 or w % 2:                            # AvgPool2d floors: drop the odd row / column first
            x = x[:, : h - h % 2, : w - w % 2, :].contiguous()
        return ops.dwtblock_act(x, 1, c)              # LL/2 == 2x2 average (model.py:54), fwd + adjoint kernels

    def forward(self, x, temb):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), temb))


class UpSample(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.main = _conv_param(nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1))
        self.initialize()

    def initialize(self):
        init.xavier_uniform_(self.main.weight)
        init.zeros_(self.main.bias)

    def forward_nhwc(self, x, temb=None, out=None):
        return ops.conv(ops.upsample2x(x), self.main.weight, self.main.bias, out=out)

    def forward(self, x, temb):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), temb))


class AttnBlock(nn.Module):
    def __init__(self, in_ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.proj = nn.Conv2d(in_ch, in_ch, 1, stride=1, padding=0)
        self.initialize()

    def initialize(self):
        for module in [self.proj_q, self.proj_k, self.proj_v, self.proj]:
            init.xavier_uniform_(module.weight)
            init.zeros_(module.bias)
        init.xavier_uniform_(self.proj.weight, gain=1e-5)

    def fused_param_groups(self):
        """Parameters a training arena should lay out contiguously (train.FlatArena): q|k|v weights, q|k|v biases."""
        return ([self.proj_q.weight, self.proj_k.weight, self.proj_v.weight],
                [self.proj_q.bias, self.proj_k.bias, self.proj_v.bias])

    def forward_nhwc(self, x, out=None):
        n, h, w, c = x.shape
        y = ops.gn_act(x, self.group_norm.weight, self.group_norm.bias, 32, act="none", eps=self.group_norm.eps)
        # q, k, v projections as ONE 1x1 conv on the tensor cores (weights concatenated on the fly: the
        # state_dict keeps the reference's three separate convs); softmax(QK^T/sqrt(C))V on PyTorch SDPA
        if ops.fused_conv_registered(self.proj_q.weight):
            # under train.DDPMTrainStep the three weights (and biases) are neighbours in the parameter arena: their
            # concatenation, its packed operands and its gradient are views -- nothing is copied or re-packed
            qkv = ops.fused_conv1x1(y, *self.fused_param_groups())
        else:
            wqkv = torch.cat([self.proj_q.weight, self.proj_k.weight, self.proj_v.weight], dim=0)
            bqkv = torch.cat([self.proj_q.bias, self.proj_k.bias, self.proj_v.bias], dim=0)
            qkv = ops.conv(y, wqkv, bqkv)
        if ops.attention_core_supported(n * h * w, h * w, c):
            # softmax(Q K^T / sqrt(C)) V and its backward on the tcgen05 batched-GEMM kernel (csrc/attention.cu)
            o = ops.attention_core(qkv)
        else:       # shapes the kernel does not tile (token counts that do not divide 256, C not a multiple of 64)
            qkv = qkv.reshape(n, 1, h * w, 3 * c)
            q, k, v = ops.split3(qkv)
            o = F.scaled_dot_product_attention(q, k, v, scale=int(c) ** (-0.5)).reshape(n, h, w, c)
        # x + proj(o): the residual add rides in the conv epilogue
        return ops.conv(o, self.proj.weight, self.proj.bias, residual=x, out=out)

    def forward(self, x):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x)))


class ResBlock(nn.Module):
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False):
        super().__init__()
        self.in_ch = in_ch
        self.out_ch = out_ch
        self.block1 = nn.Sequential(
            nn.GroupNorm(32, in_ch),
            Swish(),
            _conv_param(nn.Conv2d(in_ch, out_ch, 3, stride=1, padding=1)),
        )
        self.temb_proj = nn.Sequential(
            Swish(),
            nn.Linear(tdim, out_ch),
        )
        self.block2 = nn.Sequential(
            nn.GroupNorm(32, out_ch),
            Swish(),
            nn.Dropout(dropout),
            _conv_param(nn.Conv2d(out_ch, out_ch, 3, stride=1, padding=1)),
        )
        if in_ch != out_ch:
            self.shortcut = _conv_param(nn.Conv2d(in_ch, out_ch, 1, stride=1, padding=0))
        else:
            self.shortcut = nn.Identity()
        self.attn = AttnBlock(out_ch) if attn else nn.Identity()
        self.initialize()

    def initialize(self):
        for module in self.modules():
            if isinstance(module, (nn.Conv2d, nn.Linear)):
                init.xavier_uniform_(module.weight)
                init.zeros_(module.bias)
        init.xavier_uniform_(self.block2[-1].weight, gain=1e-5)

    def forward_nhwc(self, x, temb, row=None, out=None):
        """`row` = temb_proj(temb) when the caller already computed it for all blocks at once (UNetWaveletEnc);
        `out` = where the block's result goes (a channel slice of the next block's concat buffer)."""
        gn1, conv1 = self.block1[0], self.block1[2]
        gn2, drop, conv2 = self.block2[0], self.block2[2], self.block2[3]
        # x also feeds the shortcut / residual branch below: its two gradients are summed in the GroupNorm backward
        a1, x = ops.gn_act_fork(x, gn1.weight, gn1.bias, gn1.num_groups, act="silu", eps=gn1.eps)
        if row is None:
            lin = self.temb_proj[1]
            row = ops.rowlin_batch([temb], [lin.weight], [lin.bias], silu=True)[0]
        # h = conv1(a1) + bias + temb_proj(temb)[:, :, None, None]   (model.py:163-164), one kernel
        h = ops.conv(a1, conv1.weight, conv1.bias, rowadd=row)
        p = drop.p if self.training else 0.0
        a2 = ops.gn_act(h, gn2.weight, gn2.bias, gn2.num_groups, act="silu", eps=gn2.eps, dropout_p=p)
        has_attn = isinstance(self.attn, AttnBlock)
        out2 = None if has_attn else out
        if isinstance(self.shortcut, nn.Conv2d):
            # conv2(a2) + shortcut(x): the 1x1 conv rides along as extra K slices of the same GEMM
            h = ops.conv(a2, conv2.weight, conv2.bias, a2=x, w2=self.shortcut.weight, out=out2, bias2=self.shortcut.bias)
        else:
            h = ops.conv(a2, conv2.weight, conv2.bias, residual=x, out=out2)
        if has_attn:
            h = self.attn.forward_nhwc(h, out=out)
        return h

    def forward(self, x, temb):
        return ops.to_nchw(self.forward_nhwc(ops.to_nhwc(x), temb))


class DTWBlock(nn.Module):
    """LL_J(x) / 2^J followed by a channel tile to `out_channels`; J = 0 is the tile alone (version 1,
    the only live branch of the reference)."""

    def __init__(self, J, out_channels, mode="zero", wave="haar") -> None:
        super().__init__()
        if mode != "zero" or wave != "haar":
            raise NotImplementedError("the B200 kernels implement mode='zero', wave='haar' (the reference's defaults)")
        if not 0 <= J <= 3:
            raise NotImplementedError("fused DTWBlock supports J in [0, 3]")
        self.version = 1
        self.J = J
        self.out_channels = out_channels
        self.xfm = _HaarFilterBuffers(["h0_col", "h1_col", "h0_row", "h1_row"])
        self.ifm = _HaarFilterBuffers(["g0_col", "g1_col", "g0_row", "g1_row"])

    def forward(self, x):
        return ops.dwtblock(x.float(), self.J, self.out_channels)


_SKIP_STREAMS = {}


def _skip_stream(device):
    st = _SKIP_STREAMS.get(device)
    if st is None:
        st = _SKIP_STREAMS[device] = torch.cuda.Stream(device=device)
    return st


class _PyramidView:
    """A DTWBlock-chain output that was never materialised: pyramid level + channel map."""

    __slots__ = ("level", "chmap")

    def __init__(self, level: int, chmap: List[int]):
        self.level, self.chmap = level, chmap

    def tiled(self, out_channels: int) -> "_PyramidView":
        c = len(self.chmap)
        return _PyramidView(self.level, [self.chmap[k % c] for k in range(out_channels)])


class UNetWaveletEnc(nn.Module):
    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, dwt_encoder=False, multi_res_loss=False,
                 downsample_type="conv"):
        super().__init__()
        assert all([i < len(ch_mult) for i in attn]), "attn index out of bound"
        tdim = ch * 4
        self.n_levels = len(ch_mult)
        self.dwt_encoder = dwt_encoder
        self.multi_res_loss = multi_res_loss
        self.downsample_type = downsample_type

        self.time_embedding_list = nn.ModuleList(TimeEmbedding(T, ch, tdim) for _ in range(self.n_levels))
        self.head_list = nn.ModuleList([])
        self.downblocks = nn.ModuleList([nn.ModuleList() for _ in range(self.n_levels)])
        chs = [ch]
        now_ch = ch
        for l, mult in enumerate(ch_mult):
            self.head_list.append(DTWBlock(J=0, out_channels=now_ch))
            out_ch = ch * mult
            for _ in range(num_res_blocks):
                if self.dwt_encoder:
                    self.downblocks[l].append(DTWBlock(J=0, out_channels=out_ch))
                else:
                    self.downblocks[l].append(ResBlock(in_ch=now_ch, out_ch=out_ch, tdim=tdim, dropout=dropout,
                                                       attn=(l in attn)))
                now_ch = out_ch
                chs.append(now_ch)
            if l != len(ch_mult) - 1:
                if self.dwt_encoder:
                    self.downblocks[l].append(DTWBlock(J=1, out_channels=now_ch))
                else:
                    self.downblocks[l].append(DownSample(now_ch, type=self.downsample_type))
                chs.append(now_ch)

        self.middleblocks = nn.ModuleList([
            ResBlock(now_ch, now_ch, tdim, dropout, attn=True),
            ResBlock(now_ch, now_ch, tdim, dropout, attn=False),
        ])

        self.upblocks = nn.ModuleList([nn.ModuleList() for _ in range(self.n_levels)])
        for l, mult in reversed(list(enumerate(ch_mult))):
            out_ch = ch * mult
            for j in range(num_res_blocks + 1):
                chs_pop = chs.pop()
                self.upblocks[l].append(ResBlock(in_ch=chs_pop + now_ch, out_ch=out_ch, tdim=tdim, dropout=dropout,
                                                 attn=(l in attn)))
                now_ch = out_ch
            if l != 0:
                self.upblocks[l].append(UpSample(now_ch))
        assert len(chs) == 0

        self.tail_list = nn.ModuleList([nn.Sequential(
            nn.GroupNorm(32, ch * mult),
            Swish(),
            _conv_param(nn.Conv2d(ch * mult, 3, 3, stride=1, padding=1))
        ) for mult in ch_mult])
        self.initialize()
        self._chmap_cache = {}

    def initialize(self):
        for tail in self.tail_list:
            init.xavier_uniform_(tail[-1].weight, gain=1e-5)
            init.zeros_(tail[-1].bias)

    # ---- internals ---------------------------------------------------------------------------
    def _tail(self, level, h):
        gn, conv = self.tail_list[level][0], self.tail_list[level][2]
        a = ops.gn_act(h, gn.weight, gn.bias, gn.num_groups, act="silu", eps=gn.eps)
        return ops.conv(a, conv.weight, conv.bias, out_nchw=True)      # fp32 NCHW [N,3,H,W] straight from TMEM

    def _materialise(self, pyramid, view: _PyramidView, out=None) -> torch.Tensor:
        base = pyramid[view.level]
        n, _, h, w = base.shape
        key = (tuple(view.chmap), base.device)
        cm = self._chmap_cache.get(key)
        if cm is None:
            cm = torch.tensor(view.chmap, dtype=torch.int32, device=base.device)
            self._chmap_cache[key] = cm
        if out is None:
            out = torch.empty((n, h, w, len(view.chmap)), dtype=torch.bfloat16, device=base.device)
        return ops.dwtblock_nhwc(base, 0, out, cm)

    def _encode_haar(self, x, first):
        """Haar encoder: every skip tensor is a channel tile of one level of the image pyramid
        LL_l(x)/2^l, so only the 3-channel pyramid is computed; skips are written once, in NHWC bf16."""
        pyramid = {first: x}
        view = _PyramidView(first, list(range(x.shape[1]))).tiled(self.head_list[first].out_channels)
        views = [view]
        for level in range(first, self.n_levels):
            for layer in self.downblocks[level]:
                if layer.J == 1:
                    src = pyramid[view.level]
                    pyramid[view.level + 1] = ops.dwtblock(src, 1, src.shape[1])      # LL/2 of the 3-channel image
                    view = _PyramidView(view.level + 1, view.chmap)
                elif layer.J != 0:
                    raise NotImplementedError
                view = view.tiled(layer.out_channels)
                views.append(view)
        return pyramid, views

    @staticmethod
    def _run(layer, h, rows, out=None):
        if isinstance(layer, ResBlock):
            return layer.forward_nhwc(h, None, row=rows[id(layer)], out=out)
        if out is not None:
            return layer.forward_nhwc(h, out=out)
        return layer.forward_nhwc(h)

    def late_grad_params(self):
        """Parameters of the batched time path (time embeddings + every ResBlock's temb projection): their gradients come
        from the launches at the very end of backward (train.FlatArena keeps them out of the early all-reduce buckets)."""
        out = [p for te in self.time_embedding_list for p in te.parameters() if p.requires_grad]
        for m in self.modules():
            if isinstance(m, ResBlock):
                out += [p for p in m.temb_proj.parameters() if p.requires_grad]
        return out

    def forward(self, x, t, n_levels_used=-1):
        if n_levels_used == -1:
            n_levels_used = self.n_levels
        first = self.n_levels - n_levels_used            # finest level in use
        x = x.float().contiguous()

        # ---- time embeddings of every level in use and the temb projections of every ResBlock that will run: three
        # batched launches (reference: one TimeEmbedding call per level, model.py:451/:459/:465, and one Swish + Linear
        # per ResBlock, model.py:164)
        levels = list(range(first, self.n_levels))
        tembs = dict(zip(levels, time_embeddings([self.time_embedding_list[l] for l in levels], t)))
        plan = []
        if not self.dwt_encoder:
            plan += [(layer, lv) for lv in levels for layer in self.downblocks[lv] if isinstance(layer, ResBlock)]
        plan += [(layer, self.n_levels - 1) for layer in self.middleblocks if isinstance(layer, ResBlock)]
        plan += [(layer, lv) for lv in reversed(levels) for layer in self.upblocks[lv] if isinstance(layer, ResBlock)]
        projs = [layer.temb_proj[1] for layer, _ in plan]
        rows = dict(zip((id(layer) for layer, _ in plan),
                        ops.rowlin_batch([tembs[lv] for _, lv in plan], [p.weight for p in projs], [p.bias for p in projs],
                                         silu=True)))

        # ---- encoder
        if self.dwt_encoder:
            pyramid, views = self._encode_haar(x, first)
            hs = views
            h = self._materialise(pyramid, views[-1])
            fetch = lambda v: self._materialise(pyramid, v)
        else:
            h = ops.dwtblock_nhwc(x, 0, torch.empty((x.shape[0], x.shape[2], x.shape[3], self.head_list[first].out_channels),
                                                    dtype=torch.bfloat16, device=x.device))
            hs = [h]
            for level in range(first, self.n_levels):
                for layer in self.downblocks[level]:
                    h = self._run(layer, h, rows)
                    hs.append(h)
            fetch = lambda v: v

        # ---- middle + decoder as one layer list.  In the Haar-encoder arm every decoder ResBlock reads
        # cat(h, skip) (reference model.py:462) where the skip is a gradient-free channel tile of the image pyramid: the
        # producer of h writes straight into the first channels of the concat buffer and the tile kernel fills the rest,
        # so no concat copy is made (the residual-encoder arm keeps torch.cat: its skips carry gradients).
        layers = [(layer, None) for layer in self.middleblocks]
        for l in range(self.n_levels - 1, first - 1, -1):
            layers += [(layer, l) for layer in self.upblocks[l] if isinstance(layer, ResBlock) or l != first]
        model_out_list = []
        # Haar arm: all concat buffers are allocated up front and their skip halves (channel tiles of the pyramid, a
        # function of x alone) are written on a second stream while the main stream runs the time embeddings and the
        # middle blocks; the main stream joins before the first decoder block.
        fulls, side = {}, None
        if self.dwt_encoder:
            c, k = h.shape[3], len(hs)
            for i, (layer, l) in enumerate(layers):
                if l is not None and isinstance(layer, ResBlock):
                    k -= 1
                    skip = hs[k]
                    _, _, sh_, sw_ = pyramid[skip.level].shape
                    if i > 0:                              # the producer is layers[i-1]: it writes channels [0, c)
                        fulls[i] = (torch.empty((h.shape[0], sh_, sw_, c + len(skip.chmap)), dtype=torch.bfloat16,
                                                device=h.device), c, skip)
                if isinstance(layer, ResBlock):
                    c = layer.out_ch
            if fulls and h.is_cuda and os.environ.get("UB200_SKIP_STREAM", "1") != "0":
                side = _skip_stream(h.device)
                side.wait_stream(torch.cuda.current_stream(h.device))
            with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                for full, c0, skip in fulls.values():
                    self._materialise(pyramid, skip, out=ops.channel_slice_alias(full, c0, full.shape[3]))
        for i, (layer, l) in enumerate(layers):
            if l is not None and isinstance(layer, ResBlock):
                skip = hs.pop()
                if i in fulls:
                    if side is not None:
                        torch.cuda.current_stream(h.device).wait_stream(side)
                        side = None
                    h = ops.cat_view(h, fulls[i][0])
                else:
                    h = torch.cat([h, fetch(skip)], dim=3)
            elif l is not None and self.multi_res_loss:    # an UpSample: the coarse output leaves before it
                model_out_list.append(self._tail(l, h))
            out = ops.channel_slice_alias(fulls[i + 1][0], 0, fulls[i + 1][1]) if (i + 1) in fulls else None
            h = self._run(layer, h, rows, out)
        model_out_list.append(self._tail(first, h))
        assert len(hs) == 0

        if self.multi_res_loss:
            assert len(model_out_list) == n_levels_used
            return model_out_list
        return model_out_list[-1]


class UNet(UNetWaveletEnc):
    """The plain U-Net the reference keeps commented out (model.py:172-246): `UNet(T, ch, ch_mult, attn,
    num_res_blocks, dropout).forward(x, t)` -- the residual-encoder arm of the merged class."""

    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout):
        super().__init__(T, ch, ch_mult, attn, num_res_blocks, dropout, dwt_encoder=False)

    def forward(self, x, t):
        return super().forward(x, t)
