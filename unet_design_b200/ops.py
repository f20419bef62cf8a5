"""Autograd glue over `torch.ops.unet_b200` (the C++ extension over the C ABI).

Internal activation format ("NHWC bf16"): a torch bf16 tensor of logical shape [N, H, W, C] with channel
stride 1; it may be a channel slice of a wider buffer.  Public modules take and return the reference's
NCHW fp32 tensors and convert at the boundary.

Nothing here has a CPU or eager fallback: every function ends in a hand-written sm_100a kernel.
"""
from __future__ import annotations

import math
import os
from typing import Optional

import weakref

import torch

from ._lib import ops as _ops

ACT_NONE, ACT_SILU, ACT_GELU, ACT_RELU = 0, 1, 2, 3
_ACT = {"none": ACT_NONE, "silu": ACT_SILU, "swish": ACT_SILU, "gelu": ACT_GELU, "relu": ACT_RELU}


# --------------------------------------------------------------------------------------------------
# launch accounting (bench.py reports how many of OUR kernels ran in the timed region)
# --------------------------------------------------------------------------------------------------
class _Counter:
    launches = 0


def launches() -> int:
    return _Counter.launches


def _count(n: int = 1) -> None:
    _Counter.launches += n


# --------------------------------------------------------------------------------------------------
# per-parameter fast paths installed by the training step (train.DDPMTrainStep):
#   * gradient sinks: kernels that already ACCUMULATE (wgrad atomics, GroupNorm dgamma/dbeta atomics, the bias
#     column sum) write straight into the parameter's slice of the flat gradient arena, instead of allocating a
#     zeroed temporary that autograd then adds into .grad (two launches and three passes over memory per parameter);
#   * packed weights: the bf16 GEMM operands of every conv weight ([Cout,kh,kw,Cin] for fprop/wgrad, the
#     transposed + rotated [Cin,kh,kw,Cout] for dgrad) are refreshed once per optimiser step by the optimiser
#     kernel itself, not re-packed by every forward and backward call.
# Without a registered entry every op falls back to its self-contained path (standalone use, tests).
# --------------------------------------------------------------------------------------------------
class _Registry:
    sinks = {}      # id(param) -> (grad view in the arena, on_ready callback or None)
    packed = {}     # id(param) -> (fprop/wgrad operand, dgrad operand)  bf16, flat
    fused = {}      # id(first weight of a fused group) -> dict (see register_fused_conv)
    owner = {}      # id -> weakref of the tensor that registered it: a recycled id must not inherit stale entries


def _own(t: torch.Tensor) -> int:
    k = id(t)
    ref = _Registry.owner.get(k)
    if ref is not None and ref() is not t:          # the id was recycled by a new tensor: drop what the old one left
        _Registry.sinks.pop(k, None); _Registry.packed.pop(k, None); _Registry.fused.pop(k, None)
    _Registry.owner[k] = weakref.ref(t)
    return k


def _valid(t: torch.Tensor) -> bool:
    ref = _Registry.owner.get(id(t))
    return ref is not None and ref() is t


def register_grad_sink(param: torch.Tensor, grad_view: torch.Tensor, on_ready=None) -> None:
    _Registry.sinks[_own(param)] = (grad_view, on_ready)


def register_packed_weight(param: torch.Tensor, fwd: torch.Tensor, bwd: torch.Tensor) -> None:
    _Registry.packed[_own(param)] = (fwd, bwd)


def register_fused_conv(first_weight: torch.Tensor, packed_fwd, packed_bwd, w_sink, bias, b_sink) -> None:
    """Several 1x1 convs of one input (AttnBlock q/k/v) whose parameters lie next to each other in the training arena:
    their concatenation is then a plain view -- packed operands, the fp32 bias and both gradient sinks of the fused conv."""
    _Registry.fused[_own(first_weight)] = dict(fwd=packed_fwd, bwd=packed_bwd, w_sink=w_sink, bias=bias, b_sink=b_sink)


def fused_conv_registered(first_weight: torch.Tensor) -> bool:
    return id(first_weight) in _Registry.fused and _valid(first_weight)


def clear_registry() -> None:
    _Registry.sinks.clear()
    _Registry.packed.clear()
    _Registry.fused.clear()
    _Registry.owner.clear()


_SKIP_CHANSUM = os.environ.get("UB200_DEBUG_SKIP_CHANSUM", "0") == "1"


class _Side:
    """Weight-gradient launches on a second stream.  A wgrad only feeds the optimiser, so it need not sit on the
    backward critical path; at the coarse levels every kernel covers a fraction of the 148 SMs and is bound by
    launch-to-launch latency, so running the wgrad next to the following dgrad / GroupNorm kernels hides it (measured
    -0.1 ms/step with the <= 8x8 layers only, -0.15 ms with all layers).  The caller (train.DDPMTrainStep) enables this and
    joins the stream after backward; operands are kept alive until the join."""
    enabled = False
    chansum = True
    max_pixels = 1 << 30
    stream = None
    keep = []


def enable_side_wgrad(flag: bool, max_pixels: int = 1 << 30) -> None:
    _Side.enabled, _Side.max_pixels = bool(flag), int(max_pixels)


def join_side_stream() -> None:
    if _Side.stream is not None and _Side.keep:
        torch.cuda.current_stream(_Side.keep[0].device).wait_stream(_Side.stream)
    _Side.keep.clear()


def _wgrad(o, g, a, k, dst, stride: int = 1) -> None:
    n, h, w, _ = a.shape
    if _Side.enabled and a.is_cuda and n * h * w <= _Side.max_pixels:
        if _Side.stream is None:
            _Side.stream = torch.cuda.Stream(device=a.device)
        _Side.stream.wait_stream(torch.cuda.current_stream(a.device))
        with torch.cuda.stream(_Side.stream):
            o.conv_wgrad(g, a, k, dst, stride)
        _Side.keep += [g, a]
    else:
        o.conv_wgrad(g, a, k, dst, stride)


def _chansum(o, g, per, tot, tot2, side_ok: bool) -> None:
    """Bias / time-row column sums of a conv's output gradient.  Like the wgrad they only feed the optimiser (and, for
    `per`, the batched time-embedding backward, which waits for the side stream): off the critical path when both
    destinations are arena sinks or fresh buffers nobody reads before the join."""
    if _SKIP_CHANSUM:               # timing diagnostic only (UB200_DEBUG_SKIP_CHANSUM=1): results are wrong
        per.zero_()
        return
    if side_ok and _Side.enabled and _Side.chansum and g.is_cuda:
        if _Side.stream is None:
            _Side.stream = torch.cuda.Stream(device=g.device)
        _Side.stream.wait_stream(torch.cuda.current_stream(g.device))
        with torch.cuda.stream(_Side.stream):
            o.chansum(g, per, tot, tot2)
        _Side.keep += [g, per]
    else:
        o.chansum(g, per, tot, tot2)


def wait_side_stream(device) -> None:
    """The current stream waits for everything queued on the side stream so far (consumers of its results)."""
    if _Side.stream is not None and _Side.keep:
        torch.cuda.current_stream(device).wait_stream(_Side.stream)


def _sink_of(key):
    return _Registry.sinks.get(key) if key is not None else None


def _key(t):
    """Registry key of a parameter; None for anything else (and for a parameter whose id is a recycled one)."""
    if not isinstance(t, torch.nn.Parameter):
        return None
    k = id(t)
    ref = _Registry.owner.get(k)
    return k if (ref is None or ref() is t) else None


def _dense_nhwc(t: torch.Tensor) -> torch.Tensor:
    """Return `t` if it is a valid NHWC view (channel stride 1, uniform pixel stride), else a packed copy."""
    n, h, w, c = t.shape
    ld = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if n > 1 else c))
    ok = (c == 1 or t.stride(3) == 1) and (w == 1 or t.stride(2) == ld) and (h == 1 or t.stride(1) == w * ld) \
        and (n == 1 or t.stride(0) == h * w * ld) and ld % 8 == 0 and t.data_ptr() % 16 == 0
    return t if ok else t.contiguous()


# --------------------------------------------------------------------------------------------------
# dropout RNG state: (seed, host offset, device offset tensor).  The device counter lets a captured CUDA
# graph draw fresh masks on every replay (bumped once per training step by `advance_dropout_state`).
# --------------------------------------------------------------------------------------------------
class _DropoutState:
    seed = 0x5DEECE66D
    host_offset = 0
    dev_offset: Optional[torch.Tensor] = None


def seed_dropout(seed: int) -> None:
    _DropoutState.seed = int(seed) & 0x7FFFFFFFFFFFFFFF
    _DropoutState.host_offset = 0


def dropout_device_counter(device) -> torch.Tensor:
    d = _DropoutState.dev_offset
    if d is None or d.device != torch.device(device):
        _DropoutState.dev_offset = torch.zeros(1, dtype=torch.int64, device=device)
    return _DropoutState.dev_offset


def advance_dropout_state(device, amount: int = 1 << 32) -> None:
    dropout_device_counter(device).add_(amount)


def _next_dropout_offset(numel: int) -> int:
    off = _DropoutState.host_offset
    _DropoutState.host_offset += (numel + 3) // 4 + 1
    return off


# --------------------------------------------------------------------------------------------------
# layout at the module boundary
# --------------------------------------------------------------------------------------------------
class _ToNhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, pad_to):
        n, c, h, w = x.shape
        cp = max(pad_to, c)
        buf = torch.zeros((n, h, w, cp), dtype=torch.bfloat16, device=x.device) if cp != c else \
            torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
        _ops().nchw_to_nhwc(x.contiguous().float(), buf[..., :c] if cp != c else buf)
        _count()
        ctx.c = c
        return buf

    @staticmethod
    def backward(ctx, g):
        g = _dense_nhwc(g)
        _count()
        return _ops().nhwc_to_nchw(g[..., : ctx.c]), None


class _ToNchw(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _count()
        return _ops().nhwc_to_nchw(_dense_nhwc(x))

    @staticmethod
    def backward(ctx, g):
        n, c, h, w = g.shape
        out = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=g.device)
        _ops().nchw_to_nhwc(g.contiguous().float(), out)
        _count()
        return out


def to_nhwc(x: torch.Tensor, pad_to: int = 0) -> torch.Tensor:
    """NCHW fp32 -> NHWC bf16; channels zero-padded up to `pad_to` (conv operands need C % 16 == 0)."""
    return _ToNhwc.apply(x, pad_to)


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    """NHWC bf16 -> NCHW fp32."""
    return _ToNchw.apply(x)


# --------------------------------------------------------------------------------------------------
# Haar
# --------------------------------------------------------------------------------------------------
class _HaarDwtLevel(torch.autograd.Function):
    """One analysis level; backward is the synthesis bank cropped to the input (pytorch_wavelets AFB2D)."""

    @staticmethod
    def forward(ctx, x):
        ctx.hw = x.shape[-2:]
        ll, highs = _ops().haar_dwt2d_fwd(x.contiguous(), True)
        _count()
        return ll, highs

    @staticmethod
    def backward(ctx, gll, ghighs):
        h, w = ctx.hw
        if gll is None:
            gll = torch.zeros(ghighs.shape[:2] + ghighs.shape[3:], dtype=ghighs.dtype, device=ghighs.device)
        _count()
        return _ops().haar_idwt2d(gll.contiguous(), None if ghighs is None else ghighs.contiguous(), h, w)


class _HaarIdwtLevel(torch.autograd.Function):
    """One synthesis level; backward is the analysis bank (zero extension of a cropped gradient)."""

    @staticmethod
    def forward(ctx, ll, highs, hout, wout):
        _count()
        ctx.has_highs = highs is not None
        return _ops().haar_idwt2d(ll.contiguous(), None if highs is None else highs.contiguous(), hout, wout)

    @staticmethod
    def backward(ctx, g):
        gll, ghighs = _ops().haar_dwt2d_fwd(g.contiguous(), True)
        _count()
        return gll, (ghighs if ctx.has_highs else None), None, None


class _HaarDwtMulti(torch.autograd.Function):
    """J = 2 or 3 analysis levels in one pass over x (ub200_haar_dwt2d_multi_fwd); backward is the fused synthesis."""

    @staticmethod
    def forward(ctx, x, J):
        outs = _ops().haar_dwt2d_multi(x.contiguous(), J)
        _count()
        ctx.J = J
        ctx.shapes = [o.shape for o in outs]
        ctx.opts = (x.dtype, x.device)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        gs = [g.contiguous() if g is not None else torch.zeros(shp, dtype=ctx.opts[0], device=ctx.opts[1])
              for g, shp in zip(grads, ctx.shapes)]
        _count()
        return _ops().haar_idwt2d_multi(gs[0], gs[1:]), None


class _HaarIdwtMulti(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ll, *highs):
        ctx.J = len(highs)
        _count()
        return _ops().haar_idwt2d_multi(ll.contiguous(), [h.contiguous() for h in highs])

    @staticmethod
    def backward(ctx, g):
        outs = _ops().haar_dwt2d_multi(g.contiguous(), ctx.J)
        _count()
        return tuple(outs)


def _multi_ok(h: int, w: int, J: int) -> bool:
    return J in (2, 3) and h % (1 << J) == 0 and w % 8 == 0


def haar_dwt2d(x: torch.Tensor, J: int):
    """`(Yl, [Yh_1..Yh_J])`, the contract of `pytorch_wavelets.DWTForward(J, mode='zero', wave='haar')`.
    J = 2 or 3 on extents divisible by 2^J (W by 8) runs as ONE kernel; anything else level by level."""
    if _multi_ok(x.shape[-2], x.shape[-1], J) and x.dtype == torch.float32:
        outs = _HaarDwtMulti.apply(x, J)
        return outs[0], list(outs[1:])
    highs = []
    ll = x
    for _ in range(J):
        ll, hi = _HaarDwtLevel.apply(ll)
        highs.append(hi)
    return ll, highs


def haar_idwt2d(yl: torch.Tensor, highs) -> torch.Tensor:
    """`pytorch_wavelets.DWTInverse(mode='zero', wave='haar')((Yl, Yh))`; identity for an empty list."""
    J = len(highs)
    if J in (2, 3) and all(h is not None for h in highs) and yl.dtype == torch.float32 \
            and _multi_ok(yl.shape[-2] << J, yl.shape[-1] << J, J) \
            and all(tuple(h.shape[-2:]) == (yl.shape[-2] << (J - 1 - j), yl.shape[-1] << (J - 1 - j)) for j, h in enumerate(highs)):
        return _HaarIdwtMulti.apply(yl, *highs)
    ll = yl
    for band in highs[::-1]:
        if band is None:
            band_hw = ll.shape[-2:]
        else:
            band_hw = band.shape[-2:]
        if ll.shape[-2] > band_hw[0]:
            ll = ll[..., :-1, :]
        if ll.shape[-1] > band_hw[1]:
            ll = ll[..., :-1]
        ll = _HaarIdwtLevel.apply(ll, band, 2 * ll.shape[-2], 2 * ll.shape[-1])
    return ll


class _DwtBlock(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, J, out_channels):
        ctx.meta = (x.shape[1], x.shape[2], x.shape[3], J)
        _count()
        return _ops().dwtblock_fwd(x.contiguous(), J, out_channels)

    @staticmethod
    def backward(ctx, g):
        c, h, w, J = ctx.meta
        _count()
        return _ops().dwtblock_bwd(g.contiguous(), c, h, w, J), None, None


def dwtblock(x: torch.Tensor, J: int, out_channels: int) -> torch.Tensor:
    """Fused DTWBlock / DWTBlock: LL_J(x)/2^J then channel tile, NCHW fp32 (diff_cifar/model.py:270-323).
    The kernel fuses up to 3 levels; deeper transforms (DWTForward accepts any J) compose: LL_J/2^J = (LL_3/8) applied
    repeatedly, the zero extension of odd extents composing level by level."""
    while J > 3:
        x = _DwtBlock.apply(x, 3, x.shape[1])
        J -= 3
    return _DwtBlock.apply(x, J, out_channels)


class _MultiResMse(torch.autograd.Function):
    """sum_k mse(out_k, LL_k(noise) / 2^k) and its gradients in ONE kernel (ub200_multires_mse_f32).  `outs` finest first."""

    @staticmethod
    def forward(ctx, noise, *outs):
        want = any(o.requires_grad for o in outs)
        res = _ops().multires_mse(noise.contiguous(), [o.contiguous() for o in outs], want)
        _count()
        sums = res[0]
        numel = torch.tensor([o.numel() for o in outs], dtype=torch.float32, device=noise.device)
        per_level = sums / numel
        ctx.save_for_backward(*res[1:])
        ctx.n = len(outs)
        return (per_level.sum(), per_level)

    @staticmethod
    def backward(ctx, g_total, g_levels):
        grads = ctx.saved_tensors
        outs = []
        for k, g in enumerate(grads):
            scale = g_total if g_levels is None else g_total + g_levels[k]
            outs.append(g * scale)
        return (None, *outs)


def multires_mse(noise: torch.Tensor, outs_fine_first):
    """(sum of the per-level mean-squared errors, per-level losses [J+1], finest first) against the Haar target pyramid of
    `noise`; None when the shapes are not eligible for the fused kernel (odd extents, more than 4 levels)."""
    J = len(outs_fine_first) - 1
    h, w = noise.shape[-2:]
    if J < 1 or J > 3 or h % (1 << J) or w % 8 or noise.dtype != torch.float32:
        return None
    if any(tuple(o.shape) != (noise.shape[0], noise.shape[1], h >> k, w >> k) or o.dtype != torch.float32
           for k, o in enumerate(outs_fine_first)):
        return None
    return _MultiResMse.apply(noise, *outs_fine_first)


def dwtblock_nhwc(x: torch.Tensor, J: int, out: torch.Tensor, chmap: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Same forward written into an NHWC bf16 view (no autograd: the diffusion encoders see data only).
    `chmap` (int32 [C_out]) names the source channel of each output channel; None = k mod C."""
    _ops().dwtblock_fwd_nhwc(x.contiguous(), J, chmap, out)
    _count()
    return out


# --------------------------------------------------------------------------------------------------
# GroupNorm + activation (+ scale/shift, + dropout)
# --------------------------------------------------------------------------------------------------
class _GnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, scale, shift, G, eps, act, p_drop, addend, fork=False):
        x = _dense_nhwc(x)
        n, h, w, c = x.shape
        stats = None
        if G > 0:                      # statistics are produced by the fused forward kernel below
            stats = torch.empty((n, G, 2), dtype=torch.float32, device=x.device)
        else:
            G = 1                      # G <= 0: no normalisation, plain activation (norm=False blocks)
        y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
        seed = off = 0
        dev = None
        if p_drop > 0.0:
            seed, off, dev = _DropoutState.seed, _next_dropout_offset(x.numel()), dropout_device_counter(x.device)
        _ops().gn_act_fwd(x, G, stats, eps, gamma, beta, scale, shift, act, p_drop, seed, off, dev,
                          _dense_nhwc(addend) if addend is not None else None, y, stats is not None)
        _count(1)
        ctx.has_addend = addend is not None
        ctx.keys = (_key(gamma), _key(beta))
        ctx.save_for_backward(x, stats, gamma, beta, scale, shift)
        ctx.cfg = (G, eps, act, p_drop, seed, off, dev)
        ctx.fork = bool(fork)
        if fork:
            # second output: x itself, for the branch that by-passes the normalisation (ResBlock shortcut / residual).
            # Its gradient comes back to THIS node, where the backward kernel adds it to the GroupNorm gradient.
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, gy, gx2=None):
        x, stats, gamma, beta, scale, shift = ctx.saved_tensors
        G, eps, act, p_drop, seed, off, dev = ctx.cfg
        if gy is None:                 # only the by-pass branch carried a gradient
            gy = torch.zeros(x.shape, dtype=torch.bfloat16, device=x.device)
        gy = _dense_nhwc(gy)
        gadd = _dense_nhwc(gx2.to(torch.bfloat16)) if gx2 is not None else None
        gx = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        sg, sb = _sink_of(ctx.keys[0]), _sink_of(ctx.keys[1])
        dgamma = (sg[0] if sg else torch.zeros_like(gamma)) if gamma is not None else None
        dbeta = (sb[0] if sb else torch.zeros_like(beta)) if beta is not None else None
        dscale = torch.empty_like(scale) if scale is not None else None
        dshift = torch.empty_like(shift) if shift is not None else None
        _ops().gn_act_bwd(gy, x, G, stats, eps, gamma, beta, scale, shift, act, p_drop, seed, off, dev, gx, False,
                          dgamma, dbeta, dscale, dshift, gadd)
        _count(1)
        for sink in (sg, sb):
            if sink and sink[1] is not None:
                sink[1]()
        return (gx, None if sg else dgamma, None if sb else dbeta, dscale, dshift, None, None, None, None,
                (gy if ctx.has_addend else None), None)


def gn_act(x, gamma, beta, groups: int, act: str = "silu", eps: float = 1e-5, dropout_p: float = 0.0,
           scale=None, shift=None, addend=None) -> torch.Tensor:
    """addend + dropout(act(GroupNorm(x) * (1 + scale) + shift)); NHWC bf16 in and out.
    `groups = 0` skips the normalisation (gamma / beta still apply if given)."""
    return _GnAct.apply(x, gamma, beta, scale, shift, groups, eps, _ACT[act], float(dropout_p), addend)


def gn_act_fork(x, gamma, beta, groups: int, act: str = "silu", eps: float = 1e-5):
    """(act(GroupNorm(x)), x): the second output is x for a branch that skips the normalisation; both gradients are
    summed inside the GroupNorm backward kernel instead of by a separate add over the tensor."""
    return _GnAct.apply(x, gamma, beta, None, None, groups, eps, _ACT[act], 0.0, None, True)


class _DwtBlockNhwc(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, J, out_channels):
        x = _dense_nhwc(x)
        ctx.meta = (x.shape[1], x.shape[2], x.shape[3], J)
        _count()
        return _ops().dwtblock_nhwc_fwd(x, J, out_channels)

    @staticmethod
    def backward(ctx, g):
        h, w, c, J = ctx.meta
        _count()
        return _ops().dwtblock_nhwc_bwd(_dense_nhwc(g), h, w, c, J), None, None


def dwtblock_act(x: torch.Tensor, J: int, out_channels: int) -> torch.Tensor:
    """DWTBlock on an NHWC bf16 activation (J in {0, 1}) with its adjoint as backward
    (pdearena twod_unetbase.py:173-193; wmh/model.py:72-95)."""
    return _DwtBlockNhwc.apply(x, J, out_channels)


# --------------------------------------------------------------------------------------------------
# nearest x2 up-sampling
# --------------------------------------------------------------------------------------------------
class _Upsample2x(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _dense_nhwc(x)
        n, h, w, c = x.shape
        out = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
        _ops().upsample2x(x, out)
        _count()
        return out

    @staticmethod
    def backward(ctx, g):
        g = _dense_nhwc(g)
        n, h2, w2, c = g.shape
        gx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.bfloat16, device=g.device)
        _ops().upsample2x_bwd(g, gx)
        _count()
        return gx


def upsample2x(x: torch.Tensor) -> torch.Tensor:
    return _Upsample2x.apply(x)


# --------------------------------------------------------------------------------------------------
# convolution (3x3 / 1x1, stride 1, same padding) on tcgen05
# --------------------------------------------------------------------------------------------------
def _pad16(c: int) -> int:
    return (c + 15) // 16 * 16


def pack_weight(w: torch.Tensor, transpose_flip: bool = False) -> torch.Tensor:
    """fp32 [Cout,Cin,k,k] (any dense strides) -> packed bf16 operand of the implicit GEMM."""
    cout, cin, k, _ = w.shape
    rows, cols = (cin, cout) if transpose_flip else (cout, cin)
    out = torch.empty(_pad16(rows) * k * k * cols, dtype=torch.bfloat16, device=w.device)
    _ops().pack_conv_weight(w.detach(), transpose_flip, out)
    _count()
    return out


class _Conv(torch.autograd.Function):
    """out = conv_k(a, w) [+ conv_1(a2, w2)] + bias + rowadd[:, :, None, None] + residual.

    `out_nchw=True` returns fp32 NCHW [N, Cout, H, W] (any Cout: the network tails); otherwise bf16 NHWC.
    """

    @staticmethod
    def forward(ctx, a, w, bias, rowadd, a2, w2, residual, out_nchw, out_box=None, bias2=None, stride=1):
        o = _ops()
        a = _dense_nhwc(a)
        n, hin, win, cin = a.shape
        h, wd = (hin + stride - 1) // stride, (win + stride - 1) // stride          # output extents (padding k // 2)
        cout, k = w.shape[0], w.shape[2]
        assert w.shape[1] == cin, f"conv: weight expects {w.shape[1]} input channels, activation has {cin}"
        pk = _Registry.packed.get(_key(w))
        wp = pk[0] if pk else pack_weight(w)
        a2d = w2p = None
        if a2 is not None:
            a2d = _dense_nhwc(a2)
            pk2 = _Registry.packed.get(_key(w2))
            w2p = pk2[0] if pk2 else pack_weight(w2)
        res = _dense_nhwc(residual) if residual is not None else None
        if out_nchw:
            out = torch.empty((n, cout, h, wd), dtype=torch.float32, device=a.device)
            o.conv_fprop(a, wp, k, cout, a2d, w2p, bias, rowadd, res, None, out, bias2, stride)
        elif out_box is not None:
            # the caller owns the destination: an NHWC view (pixel stride > Cout) inside a wider buffer, e.g. the first
            # channels of a decoder concat buffer.  Passed in a list so autograd does not treat it as an input.
            dst = out_box[0]
            assert dst.shape == (n, h, wd, cout) and dst.dtype == torch.bfloat16
            o.conv_fprop(a, wp, k, cout, a2d, w2p, bias, rowadd, res, dst, None, bias2, stride)
            out = dst.view_as(dst)
        else:
            out = torch.empty((n, h, wd, cout), dtype=torch.bfloat16, device=a.device)
            o.conv_fprop(a, wp, k, cout, a2d, w2p, bias, rowadd, res, out, None, bias2, stride)
        _count()
        ctx.save_for_backward(a, w, a2d, w2)
        ctx.flags = (bias is not None, rowadd is not None, residual is not None, out_nchw)
        ctx.keys = (_key(w), _key(bias), _key(w2))
        ctx.bias2_key = _key(bias2) if bias2 is not None else None
        ctx.has_bias2 = bias2 is not None
        ctx.stride = stride
        return out

    @staticmethod
    def backward(ctx, g):
        o = _ops()
        a, w, a2, w2 = ctx.saved_tensors
        has_bias, has_rowadd, has_res, out_nchw = ctx.flags
        stride = ctx.stride
        n, hin, win, cin = a.shape
        h, wd = (hin + stride - 1) // stride, (win + stride - 1) // stride
        cout, k = w.shape[0], w.shape[2]
        cpad = _pad16(cout)
        if out_nchw:      # fp32 NCHW gradient of a narrow tail: pad the channel dim for the tensor cores
            gp = torch.zeros((n, h, wd, cpad), dtype=torch.bfloat16, device=g.device)
            o.nchw_to_nhwc(g.contiguous().float(), gp[..., :cout])
            _count()
            g_full, g_valid = gp, gp[..., :cout]
        else:
            g_full = g_valid = _dense_nhwc(g)
            if cpad != cout:
                gp = torch.zeros((n, h, wd, cpad), dtype=torch.bfloat16, device=g.device)
                gp[..., :cout].copy_(g_valid)
                g_full = gp
        needs = ctx.needs_input_grad
        ga = gw = gbias = growadd = ga2 = gw2 = gres = gbias2 = None
        want_b2 = ctx.has_bias2 and needs[9]
        kw, kb, kw2 = ctx.keys
        pk, pk2 = _Registry.packed.get(kw), _Registry.packed.get(kw2)
        if needs[0]:
            # dgrad = the same implicit GEMM with transposed, 180-degree-rotated weights
            if pk is not None and cpad == cout:
                wtp = pk[1]
            else:
                if cpad != cout:
                    w_t = torch.zeros((cpad, cin, k, k), dtype=torch.float32, device=w.device)
                    w_t[:cout].copy_(w.detach())
                else:
                    w_t = w
                wtp = pack_weight(w_t, transpose_flip=True)
            ga = torch.empty((n, hin, win, cin), dtype=torch.bfloat16, device=a.device)
            g_in = g_full
            if stride != 1:
                # transposed convolution: the output gradient zero-stuffed onto the input grid, then the stride-1 kernel
                # (the down-sampling arms are one conv per level of the baseline U-Net, not on the Multi-ResNet path)
                g_in = torch.zeros((n, hin, win, cpad), dtype=torch.bfloat16, device=a.device)
                g_in[:, ::stride, ::stride, :].copy_(g_full)
            o.conv_fprop(g_in, wtp, k, cin, None, None, None, None, None, ga, None, None, 1)
            _count()
        if needs[1]:
            sink = _sink_of(kw) if cpad == cout else None
            if sink is not None:                        # accumulate straight into the gradient arena
                _wgrad(o, g_full, a, k, sink[0], stride)
                _count()
                if sink[1] is not None:
                    sink[1]()
            else:
                dw = torch.zeros((cpad, k, k, cin), dtype=torch.float32, device=w.device)
                o.conv_wgrad(g_full, a, k, dw, stride)
                _count(2)
                gw = dw[:cout].permute(0, 3, 1, 2)      # [Cout,Cin,k,k] view with channels_last strides
        if (has_bias and needs[2]) or (has_rowadd and needs[3]) or want_b2:
            bsink = _sink_of(kb) if (has_bias and needs[2] and cout % 8 == 0) else None
            if cout % 8 == 0:
                per = torch.empty((n, cout), dtype=torch.float32, device=g.device)
                if bsink is not None:
                    tot = bsink[0]
                else:
                    tot = torch.zeros((cout,), dtype=torch.float32, device=g.device) if has_bias else None
                tot2 = b2sink = None
                if want_b2:                 # the second bias receives the same column sums
                    b2sink = _sink_of(ctx.bias2_key)
                    tot2 = b2sink[0] if b2sink is not None else torch.zeros((cout,), dtype=torch.float32, device=g.device)
                # on the side stream only if nothing on the main stream reads the results before the join: sinks, or
                # `per` (consumed by the batched time-embedding backward, which waits for the side stream itself)
                side_ok = (bsink is not None or tot is None) and (b2sink is not None or tot2 is None)
                _chansum(o, g_valid, per, tot, tot2, side_ok)
                _count(1 + (tot is not None) + (tot2 is not None))
                growadd, gbias = (per if has_rowadd else None), (None if bsink is not None else tot)
                gbias2 = None if b2sink is not None else tot2
                for sk in (bsink, b2sink):
                    if sk is not None and sk[1] is not None:
                        sk[1]()
            else:
                per = torch.empty((n, cpad), dtype=torch.float32, device=g.device)
                tot = torch.zeros((cpad,), dtype=torch.float32, device=g.device) if has_bias else None
                o.chansum(g_full, per, tot, None)
                _count(2 if tot is not None else 1)
                growadd = per[:, :cout].contiguous() if has_rowadd else None
                gbias = tot[:cout] if has_bias else None
                gbias2 = tot[:cout].clone() if want_b2 else None
        if a2 is not None:
            if needs[4]:
                w2tp = pk2[1] if (pk2 is not None and cpad == cout) else pack_weight(w2, transpose_flip=True)
                ga2 = torch.empty(a2.shape, dtype=torch.bfloat16, device=a.device)
                o.conv_fprop(g_full, w2tp, 1, a2.shape[3], None, None, None, None, None, ga2, None, None)
                _count()
            if needs[5]:
                sink2 = _sink_of(kw2) if cpad == cout else None
                if sink2 is not None:
                    _wgrad(o, g_full, a2, 1, sink2[0])
                    _count()
                    if sink2[1] is not None:
                        sink2[1]()
                else:
                    dw2 = torch.zeros((cpad, 1, 1, a2.shape[3]), dtype=torch.float32, device=w.device)
                    o.conv_wgrad(g_full, a2, 1, dw2)
                    _count(2)
                    gw2 = dw2[:cout].permute(0, 3, 1, 2)
        if has_res and needs[6]:
            gres = g_valid
        return ga, gw, gbias, growadd, ga2, gw2, gres, None, None, gbias2, None


def conv(a: torch.Tensor, w: torch.Tensor, bias=None, rowadd=None, a2=None, w2=None, residual=None,
         out_nchw: bool = False, out: Optional[torch.Tensor] = None, bias2=None, stride: int = 1) -> torch.Tensor:
    """Fused convolution, k = 1 or 3, padding k // 2, stride 1 or 2 (`nn.Conv2d(C, C, 3, stride=2, padding=1)` of the
    down-sampling arms: the TMA traversal stride fetches every second pixel).  `a` NHWC bf16 with C % 16 == 0.
    `out`: optional destination, an NHWC bf16 view [N,Ho,Wo,Cout] (may be a channel slice of a wider buffer)."""
    return _Conv.apply(a, w, bias, rowadd, a2, w2, residual, out_nchw, [out] if out is not None else None, bias2, stride)


class _FusedConv1x1(torch.autograd.Function):
    """out[..., :] = conv1x1(a, cat(w_i)) + cat(b_i) with the concatenations taken as views of the training arena
    (register_fused_conv): no weight / bias concat in forward, no split + accumulate of the gradients in backward."""

    @staticmethod
    def forward(ctx, a, *params):
        reg = _Registry.fused[id(params[0])]
        a = _dense_nhwc(a)
        n, h, wd, cin = a.shape
        cout = reg["bias"].numel()
        out = torch.empty((n, h, wd, cout), dtype=torch.bfloat16, device=a.device)
        _ops().conv_fprop(a, reg["fwd"], 1, cout, None, None, reg["bias"], None, None, out, None, None)
        _count()
        ctx.save_for_backward(a)
        ctx.reg, ctx.nparams = reg, len(params)
        return out

    @staticmethod
    def backward(ctx, g):
        o = _ops()
        (a,), reg = ctx.saved_tensors, ctx.reg
        n, h, wd, cin = a.shape
        g = _dense_nhwc(g)
        cout = g.shape[3]
        ga = None
        if ctx.needs_input_grad[0]:
            ga = torch.empty((n, h, wd, cin), dtype=torch.bfloat16, device=a.device)
            o.conv_fprop(g, reg["bwd"], 1, cin, None, None, None, None, None, ga, None, None)
            _count()
        _wgrad(o, g, a, 1, reg["w_sink"])
        per = torch.empty((n, cout), dtype=torch.float32, device=g.device)
        _chansum(o, g, per, reg["b_sink"], None, True)
        _count(3)
        return (ga,) + (None,) * ctx.nparams


def fused_conv1x1(a: torch.Tensor, weights, biases) -> torch.Tensor:
    return _FusedConv1x1.apply(a, *weights, *biases)


class _CatView(torch.autograd.Function):
    """`full` already holds `head` in its first channels (the producer wrote there) and a gradient-free tensor in the
    rest: returns `full` as the concatenation without copying; backward hands the head's channel slice back."""

    @staticmethod
    def forward(ctx, head, box):
        full = box[0]
        assert head.data_ptr() == full.data_ptr() and head.shape[:3] == full.shape[:3]
        ctx.c = head.shape[3]
        return full.view_as(full)

    @staticmethod
    def backward(ctx, g):
        return g[..., :ctx.c], None


def cat_view(head: torch.Tensor, full: torch.Tensor) -> torch.Tensor:
    return _CatView.apply(head, [full])


def channel_slice_alias(full: torch.Tensor, c0: int, c1: int) -> torch.Tensor:
    """Channels [c0, c1) of an NHWC buffer as an independent tensor object over the same storage (not an autograd view:
    kernels fill disjoint channel ranges of one concat buffer, which view + in-place tracking would reject)."""
    n, h, w, _ = full.shape
    t = torch.empty(0, dtype=full.dtype, device=full.device)
    t.set_(full.untyped_storage(), full.storage_offset() + c0, (n, h, w, c1 - c0), full.stride())
    return t


class _RowLinBatch(torch.autograd.Function):
    """y_i = act(x_i) @ w_i^T + b_i for a list of items in ONE launch (act = SiLU or identity); backward = two launches.
    The time-embedding path of the reference: TimeEmbedding's two Linear layers (diff_cifar/model.py:29-36) and every
    ResBlock's Swish + Linear `temb_proj` (model.py:134-137).  Items may share their x; its gradient is the sum."""

    @staticmethod
    def forward(ctx, silu, n, *tensors):
        xs, ws, bs = list(tensors[:n]), list(tensors[n:2 * n]), list(tensors[2 * n:3 * n])
        xs = [x if (x.dtype == torch.float32 and x.is_contiguous()) else x.float().contiguous() for x in xs]
        ys = [torch.empty((xs[i].shape[0], ws[i].shape[0]), dtype=torch.float32, device=xs[i].device) for i in range(n)]
        _ops().rowlin_fwd(xs, [w.detach() for w in ws], [b.detach() if b is not None else None for b in bs], ys, bool(silu))
        _count()
        ctx.save_for_backward(*xs, *ws)
        ctx.cfg = (bool(silu), n, [_key(w) for w in ws], [_key(b) for b in bs], [b is not None for b in bs])
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gys):
        silu, n, wkeys, bkeys, has_b = ctx.cfg
        saved = ctx.saved_tensors
        if saved[0].is_cuda:
            wait_side_stream(saved[0].device)      # the row gradients are channel sums that may have run there
        xs, ws = list(saved[:n]), list(saved[n:])
        needs = ctx.needs_input_grad            # (silu, n, xs..., ws..., bs...)
        gys = [g.contiguous() if g is not None else torch.zeros((xs[i].shape[0], ws[i].shape[0]), dtype=torch.float32,
                                                                device=xs[i].device) for i, g in enumerate(gys)]
        gws, gbs, gxs = [None] * n, [None] * n, [None] * n
        ret_w, ret_b, ret_x = [None] * n, [None] * n, [None] * n
        done = []
        shared = {}                              # data_ptr of x -> its gradient buffer (one per distinct input)
        for i in range(n):
            if needs[2 + n + i]:
                sink = _sink_of(wkeys[i])
                if sink is not None:
                    gws[i] = sink[0]
                    done.append(sink)
                else:
                    gws[i] = ret_w[i] = torch.zeros_like(ws[i], memory_format=torch.contiguous_format)
            if has_b[i] and needs[2 + 2 * n + i]:
                sink = _sink_of(bkeys[i])
                if sink is not None:
                    gbs[i] = sink[0]
                    done.append(sink)
                else:
                    gbs[i] = ret_b[i] = torch.zeros((ws[i].shape[0],), dtype=torch.float32, device=ws[i].device)
            if needs[2 + i]:
                key = xs[i].data_ptr()
                if key not in shared:
                    shared[key] = torch.empty_like(xs[i])
                    ret_x[i] = shared[key]       # autograd sums the slots of one tensor: hand the total to the first slot only
                gxs[i] = shared[key]
        wd = [w.detach() for w in ws]
        none = [None] * n
        if _Side.enabled and _Side.chansum and xs[0].is_cuda and any(g is not None for g in gws + gbs) \
                and any(g is not None for g in gxs):
            # weight / bias gradients only feed the optimiser: second stream; the input gradient stays on this one
            if _Side.stream is None:
                _Side.stream = torch.cuda.Stream(device=xs[0].device)
            _Side.stream.wait_stream(torch.cuda.current_stream(xs[0].device))
            with torch.cuda.stream(_Side.stream):
                _ops().rowlin_bwd(xs, wd, gys, gws, gbs, none, silu)
            _Side.keep += list(xs) + list(gys)
            _ops().rowlin_bwd(xs, wd, gys, none, none, gxs, silu)
        else:
            _ops().rowlin_bwd(xs, wd, gys, gws, gbs, gxs, silu)
        _count(2)
        for sink in done:
            if sink[1] is not None:
                sink[1]()
        return (None, None, *ret_x, *ret_w, *ret_b)


def rowlin_batch(xs, ws, bs, silu: bool):
    """[act(x_i) @ w_i^T + b_i for i]: fp32 [N, K] inputs (all the same N and K), fp32 nn.Linear weights [cout_i, K]."""
    n = len(xs)
    assert n > 0 and len(ws) == n and len(bs) == n
    return list(_RowLinBatch.apply(silu, n, *xs, *ws, *bs))


# --------------------------------------------------------------------------------------------------
# attention core: softmax(Q K^T / sqrt(C)) V on tcgen05 (csrc/attention.cu), one head
# --------------------------------------------------------------------------------------------------
BGEMM_PLAIN, BGEMM_SOFTMAX, BGEMM_SOFTMAX_BWD = 0, 1, 2
_LOG2E = 1.4426950408889634


def attention_core_supported(n_tokens_total: int, tokens: int, channels: int) -> bool:
    """One CTA owns 256 rows = 256 / T whole samples; channels in 64-element chunks, at most 256."""
    import os
    return (os.environ.get("UB200_ATTN_CORE", "1") != "0" and tokens <= 256 and 256 % tokens == 0
            and n_tokens_total % 256 == 0 and channels % 64 == 0 and channels <= 256)


class _AttnCore(torch.autograd.Function):
    """o = softmax(q k^T * C^-0.5) v from the fused projection qkv [R, 3C] (R = samples * T rows).  Forward: two batched GEMM
    launches (scores + softmax epilogue, P V); backward: four (dS with the softmax-backward epilogue, dV, dQ, dK).  The
    softmax P (bf16 [R, 256]) is saved, so nothing is recomputed."""

    @staticmethod
    def forward(ctx, qkv2d, T):
        o_ = _ops()
        R, c3 = qkv2d.shape
        C = c3 // 3
        q, k, v = qkv2d[:, :C], qkv2d[:, C:2 * C], qkv2d[:, 2 * C:]
        P = torch.empty((R, 256), dtype=torch.bfloat16, device=qkv2d.device)
        o_.bgemm256(q, False, k, False, P, C, BGEMM_SOFTMAX, float(C) ** -0.5 * _LOG2E, T, None)
        out = torch.empty((R, C), dtype=torch.bfloat16, device=qkv2d.device)
        o_.bgemm256(P, False, v, True, out, 256, BGEMM_PLAIN, 1.0, T, None)
        _count(2)
        ctx.save_for_backward(qkv2d, P)
        ctx.T = T
        return out

    @staticmethod
    def backward(ctx, go):
        o_ = _ops()
        qkv2d, P = ctx.saved_tensors
        T = ctx.T
        R, c3 = qkv2d.shape
        C = c3 // 3
        q, k, v = qkv2d[:, :C], qkv2d[:, C:2 * C], qkv2d[:, 2 * C:]
        if go.stride(1) != 1 or go.stride(0) % 8 != 0 or go.data_ptr() % 16 != 0:
            go = go.contiguous()
        dS = torch.empty((R, 256), dtype=torch.bfloat16, device=go.device)
        o_.bgemm256(go, False, v, False, dS, C, BGEMM_SOFTMAX_BWD, float(C) ** -0.5, T, P)
        dqkv = torch.empty((R, c3), dtype=torch.bfloat16, device=go.device)
        o_.bgemm256(P, True, go, True, dqkv[:, 2 * C:], 256, BGEMM_PLAIN, 1.0, T, None)          # dV = P^T dO
        o_.bgemm256(dS, False, k, True, dqkv[:, :C], 256, BGEMM_PLAIN, 1.0, T, None)             # dQ = dS K
        o_.bgemm256(dS, True, q, True, dqkv[:, C:2 * C], 256, BGEMM_PLAIN, 1.0, T, None)          # dK = dS^T Q
        _count(4)
        return dqkv, None


def attention_core(qkv: torch.Tensor) -> torch.Tensor:
    """qkv: NHWC bf16 [N, H, W, 3C] (the fused q|k|v projection) -> o [N, H, W, C]."""
    n, h, w, c3 = qkv.shape
    qkv = _dense_nhwc(qkv)
    out = _AttnCore.apply(qkv.reshape(n * h * w, c3), h * w)
    return out.reshape(n, h, w, c3 // 3)


class _Split3(torch.autograd.Function):
    """Three channel-slice views of a fused q|k|v projection; backward gathers the three gradients into one
    buffer with strided copies (no zero-fill + add chain as plain slicing would record)."""

    @staticmethod
    def forward(ctx, qkv):
        c = qkv.shape[-1] // 3
        ctx.shape = qkv.shape
        return qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:]

    @staticmethod
    def backward(ctx, gq, gk, gv):
        c = ctx.shape[-1] // 3
        if gq is not None and gk is not None and gv is not None:
            return torch.cat([gq, gk, gv], dim=-1)          # one vectorised launch instead of three strided copies
        ref = next(g for g in (gq, gk, gv) if g is not None)
        out = torch.empty(ctx.shape, dtype=ref.dtype, device=ref.device)
        for i, g in enumerate((gq, gk, gv)):
            if g is None:
                out[..., i * c:(i + 1) * c].zero_()
            else:
                out[..., i * c:(i + 1) * c].copy_(g)
        return out


def split3(qkv: torch.Tensor):
    return _Split3.apply(qkv)


# --------------------------------------------------------------------------------------------------
# optimiser tail
# --------------------------------------------------------------------------------------------------
def sumsq_(g: torch.Tensor, acc: torch.Tensor) -> None:
    _ops().sumsq(g, acc)
    _count()


def adam_ema_step_(p, g, m, v, ema, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, ema_decay, step,
                   warmup_steps: int = 0, step_dev=None, shadow=None, weight_decay: float = 0.0) -> None:
    """Global-norm clip + Adam (AdamW when `weight_decay` > 0: decoupled decay) + EMA + bf16 shadow over flat arenas."""
    _ops().adam_ema_step(p, g, m, v, ema, sumsq, max_norm, grad_scale, lr, beta1, beta2, eps, ema_decay, step,
                         warmup_steps, step_dev, shadow, weight_decay)
    _count()


def pack_dgrad_weights_batched_(shadow, dgrad_arena, table) -> None:
    _ops().pack_dgrad_weights_batched(shadow, dgrad_arena, table)
    _count()
