"""TEST INFRASTRUCTURE ONLY.  A pure-PyTorch CPU emulation of the `torch.ops.unet_b200` op contract.

The build container has no GPU, and a GPU round-trip costs minutes of a small budget, so the *Python*
side of the drop-in modules (level bookkeeping, channel maps, autograd glue, state_dict layout, the
data-parallel step) is exercised on CPU by monkeypatching `unet_design_b200._lib._ops` with this object
(see the `emulated_ops` fixture).  It is never importable from the product package and never used on a
GPU box: the `-m gpu` tests call the real sm_100a kernels through the C ABI.
Each function mirrors the argument meaning of csrc/torch_binding.cpp.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

from oracle import haar_np


def _bf(x):
    return x.to(torch.bfloat16)


class EmulatedOps:
    calls = 0

    # ---- Haar
    def haar_dwt2d_fwd(self, x, want_highs):
        ll, lh, hl, hh = haar_np.dwt2_level(x.numpy())
        highs = torch.from_numpy(np.stack([lh, hl, hh], axis=2)) if want_highs else torch.empty(0)
        return torch.from_numpy(np.ascontiguousarray(ll)), highs

    def haar_idwt2d(self, ll, highs, hout, wout):
        z = np.zeros_like(ll.numpy())
        if highs is None:
            out = haar_np.idwt2_level(ll.numpy(), z, z, z)
        else:
            h = highs.numpy()
            out = haar_np.idwt2_level(ll.numpy(), h[:, :, 0], h[:, :, 1], h[:, :, 2])
        return torch.from_numpy(np.ascontiguousarray(out[..., :hout, :wout]))

    def haar_dwt2d_multi(self, x, J):
        h, w = x.shape[-2:]
        if J not in (2, 3) or h % (1 << J) or w % 8:
            return []
        yl, yh = haar_np.dwt2(x.numpy(), J)
        return [torch.from_numpy(np.ascontiguousarray(yl))] + [torch.from_numpy(np.ascontiguousarray(b)) for b in yh]

    def haar_idwt2d_multi(self, ll, highs):
        return torch.from_numpy(np.ascontiguousarray(haar_np.idwt2(ll.numpy(), [h.numpy() for h in highs])))

    def multires_mse(self, noise, outs, want_grads):
        J = len(outs) - 1
        res = [torch.zeros(J + 1)]
        grads = []
        for k, o in enumerate(outs):
            t = torch.from_numpy(haar_np.dwtblock(noise.numpy(), k, None)) if k else noise
            d = o.detach() - t
            res[0][k] = (d * d).sum()
            grads.append(d * (2.0 / d.numel()))
        return res + (grads if want_grads else [])

    def dwtblock_fwd(self, x, J, out_channels):
        return torch.from_numpy(haar_np.dwtblock(x.numpy(), J, out_channels))

    def dwtblock_bwd(self, g, c, h, w, J):
        return torch.from_numpy(np.ascontiguousarray(haar_np.dwtblock_bwd(g.numpy(), (g.shape[0], c, h, w), J)))

    def dwtblock_fwd_nhwc(self, x, J, chmap, out):
        y = torch.from_numpy(haar_np.dwtblock(x.numpy(), J, None))
        idx = chmap.long() if chmap is not None else torch.arange(out.shape[3]) % x.shape[1]
        out.copy_(_bf(y[:, idx].permute(0, 2, 3, 1)))

    # ---- layout
    def nchw_to_nhwc(self, x, out):
        out.copy_(_bf(x.permute(0, 2, 3, 1)))

    def nhwc_to_nchw(self, x):
        return x.float().permute(0, 3, 1, 2).contiguous()

    def upsample2x(self, x, out):
        out.copy_(x.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))

    def upsample2x_bwd(self, g, gx):
        n, h2, w2, c = g.shape
        gx.copy_(_bf(g.float().reshape(n, h2 // 2, 2, w2 // 2, 2, c).sum(dim=(2, 4))))

    # ---- GroupNorm + act
    def gn_stats(self, x, G, stats):
        n, h, w, c = x.shape
        xf = x.float().reshape(n, h * w, G, c // G)
        stats.copy_(torch.stack([xf.sum(dim=(1, 3)), (xf * xf).sum(dim=(1, 3))], dim=-1))

    @staticmethod
    def _mask(shape, p, seed, off):
        if p <= 0:
            return None
        g = torch.Generator().manual_seed((seed + 7919 * off) & 0x7FFFFFFF)
        return (torch.rand(shape, generator=g) >= p).float() / (1 - p)

    @staticmethod
    def _act(z, act):
        if act == 1:
            return z * torch.sigmoid(z)
        if act == 2:
            return F.gelu(z)
        if act == 3:
            return F.relu(z)
        return z

    def _gn_forward_fp32(self, x, G, stats, eps, gamma, beta, scale, shift, act, p, seed, off):
        n, h, w, c = x.shape
        cnt = (c // G) * h * w
        if stats is None:
            stats = torch.tensor([0.0, float(cnt)]).expand(n, G, 2)      # mean 0, var 1 (eps folded below)
            eps = 0.0
        mean = (stats[..., 0] / cnt).repeat_interleave(c // G, dim=1)[:, None, None, :]
        var = (stats[..., 1] / cnt).repeat_interleave(c // G, dim=1)[:, None, None, :] - mean * mean
        rstd = torch.rsqrt(var.clamp_min(0) + eps)
        xhat = (x.float() - mean) * rstd
        ga = gamma if gamma is not None else torch.ones(c)
        be = beta if beta is not None else torch.zeros(c)
        z = xhat * ga + be
        if scale is not None:
            z = z * (1 + scale[:, None, None, :]) + shift[:, None, None, :]
        y = self._act(z, act)
        m = self._mask(x.shape, p, seed, off)
        return y * m if m is not None else y

    def gn_act_fwd(self, x, G, stats, eps, gamma, beta, scale, shift, act, p, seed, off, off_dev, addend, y,
                   compute_stats=False):
        if compute_stats:
            self.gn_stats(x, G, stats)
        out = self._gn_forward_fp32(x, G, stats, eps, gamma, beta, scale, shift, act, p, seed, off)
        if addend is not None:
            out = out + addend.float()
        y.copy_(_bf(out))

    def dwtblock_nhwc_fwd(self, x, J, out_channels):
        xn = x.float().permute(0, 3, 1, 2).contiguous().numpy()
        y = haar_np.dwtblock(xn, J, out_channels)
        return _bf(torch.from_numpy(y).permute(0, 2, 3, 1)).contiguous()

    def dwtblock_nhwc_bwd(self, g, h, w, c, J):
        gn = g.float().permute(0, 3, 1, 2).contiguous().numpy()
        gx = haar_np.dwtblock_bwd(gn, (g.shape[0], c, h, w), J)
        return _bf(torch.from_numpy(np.ascontiguousarray(gx)).permute(0, 2, 3, 1)).contiguous()

    def gn_act_bwd(self, gy, x, G, stats, eps, gamma, beta, scale, shift, act, p, seed, off, off_dev, gx, accumulate,
                   dgamma, dbeta, dscale, dshift, gadd=None):
        with torch.enable_grad():
            xr = x.float().requires_grad_(True)
            leaves = [xr]
            ga = gamma.detach().clone().requires_grad_(True) if gamma is not None else None
            be = beta.detach().clone().requires_grad_(True) if beta is not None else None
            sc = scale.detach().clone().requires_grad_(True) if scale is not None else None
            sf = shift.detach().clone().requires_grad_(True) if shift is not None else None
            n, h, w, c = x.shape
            # statistics are functions of x: recompute them differentiably
            xf = xr.reshape(n, h * w, G, c // G)
            st = torch.stack([xf.sum(dim=(1, 3)), (xf * xf).sum(dim=(1, 3))], dim=-1) if stats is not None else None
            y = self._gn_forward_fp32(xr, G, st, eps, ga, be, sc, sf, act, p, seed, off)
            leaves += [t for t in (ga, be, sc, sf) if t is not None]
            grads = torch.autograd.grad(y, leaves, gy.float())
        g_iter = iter(grads)
        dx = next(g_iter)
        if gadd is not None:
            dx = dx + gadd.float()
        gx.copy_(_bf(gx.float() + dx) if accumulate else _bf(dx))
        for src, dst in ((ga, dgamma), (be, dbeta), (sc, dscale), (sf, dshift)):
            if src is not None:
                val = next(g_iter)
                if dst is not None:
                    if dst is dgamma or dst is dbeta:
                        dst.add_(val)
                    else:
                        dst.copy_(val)

    # ---- conv
    def conv_fprop(self, a, w, k, cout, a2, w2, bias, rowadd, residual, out, out_nchw, bias2=None, stride=1):
        n, h, wd, cin = a.shape
        cpad = (cout + 15) // 16 * 16
        wt = w.float().reshape(cpad, k, k, cin)[:cout].permute(0, 3, 1, 2)
        y = F.conv2d(a.float().permute(0, 3, 1, 2), wt, padding=k // 2, stride=stride)
        if a2 is not None:
            w2t = w2.float().reshape(cpad, a2.shape[3])[:cout, :, None, None]
            y = y + F.conv2d(a2.float().permute(0, 3, 1, 2), w2t)
        if bias is not None:
            y = y + bias[None, :, None, None]
        if bias2 is not None:
            y = y + bias2[None, :, None, None]
        if rowadd is not None:
            y = y + rowadd[:, :, None, None]
        if residual is not None:
            y = y + residual.float().permute(0, 3, 1, 2)
        if out is not None:
            out.copy_(_bf(y.permute(0, 2, 3, 1)))
        if out_nchw is not None:
            out_nchw.copy_(y)

    def conv_wgrad(self, gout, a, k, dw, stride=1):
        n, h, wd, cin = a.shape
        cout = gout.shape[3]
        xg = a.float().permute(0, 3, 1, 2)
        gg = gout.float().permute(0, 3, 1, 2)
        gw = torch.nn.grad.conv2d_weight(xg, (cout, cin, k, k), gg, padding=k // 2, stride=stride)     # [Cout,Cin,k,k]
        if dw.dim() == 4 and tuple(dw.shape) == (cout, cin, k, k):      # a [Cout,Cin,k,k] view with channels_last strides
            dw.add_(gw)
        else:
            dw.view(cout, k, k, cin).add_(gw.permute(0, 2, 3, 1))

    def bgemm256(self, a, a_mn, b, b_mn, out, K, epilogue, alpha, T, p):
        """Contract of ub200_bgemm256 (csrc/attention.cu): per group of 256 rows, out = epilogue(A @ B^T), fp32 math."""
        R, N = out.shape
        for g in range(R // 256):
            rows = slice(256 * g, 256 * g + 256)
            A = a[rows].float().t() if a_mn else a[rows, :K].float()            # [256, K]
            B = b[rows, :N].float().t() if b_mn else b[rows, :K].float()        # [N, K]
            acc = A @ B.t()
            if epilogue == 1:
                idx = torch.arange(256)
                mask = (idx[:, None] // T) == (idx[None, :] // T)
                z = (acc * alpha).masked_fill(~mask, float("-inf"))
                acc = torch.exp2(z - z.max(dim=1, keepdim=True).values)
                acc = acc / acc.sum(dim=1, keepdim=True)
            elif epilogue == 2:
                pp = p[rows].float()
                acc = pp * (acc - (pp * acc).sum(dim=1, keepdim=True)) * alpha
            else:
                acc = acc * alpha
            out[rows] = _bf(acc)

    def chansum(self, x, per_sample, total, total2=None):
        s = x.float().sum(dim=(1, 2))
        per_sample.copy_(s)
        if total is not None:
            total.add_(s.sum(0))
        if total2 is not None:
            total2.add_(s.sum(0))

    def pack_conv_weight(self, w, transpose_flip, out):
        cout, cin, k, _ = w.shape
        if transpose_flip:
            src = w.flip(2, 3).permute(1, 2, 3, 0)       # [Cin,k,k,Cout], taps rotated
        else:
            src = w.permute(0, 2, 3, 1)                  # [Cout,k,k,Cin]
        rows = src.shape[0]
        buf = torch.zeros(((rows + 15) // 16 * 16,) + tuple(src.shape[1:]), dtype=torch.bfloat16)
        buf[:rows] = _bf(src)
        out.copy_(buf.reshape(-1))

    # ---- optimiser tail
    def sumsq(self, g, acc):
        acc.add_((g.double() ** 2).sum().float())

    # ---- batched row linears (time-embedding path)
    def rowlin_fwd(self, xs, ws, bs, ys, silu):
        for x, w, b, y in zip(xs, ws, bs, ys):
            a = F.silu(x) if silu else x
            y.copy_(a @ w.t() + (b if b is not None else 0.0))

    def rowlin_bwd(self, xs, ws, gys, gws, gbs, gxs, silu):
        first = {}
        for x, w, gy, gw, gb, gx in zip(xs, ws, gys, gws, gbs, gxs):
            a = F.silu(x) if silu else x
            if gw is not None:
                gw.add_((gy.t() @ a).view(gw.shape))
            if gb is not None:
                gb.add_(gy.sum(0))
            if gx is not None:
                t = gy @ w
                if silu:
                    s = torch.sigmoid(x)
                    t = t * (s * (1 + x * (1 - s)))
                if id(gx) in first:
                    gx.add_(t)
                else:
                    first[id(gx)] = True
                    gx.copy_(t)

    def pack_dgrad_weights_batched(self, shadow, dgrad_arena, table):
        for src, dst, cout, cin, k in table.tolist():
            w = shadow[src:src + cout * k * k * cin].reshape(cout, k, k, cin)
            t = w.flip(1, 2).permute(3, 1, 2, 0)                                    # [Cin,k,k,Cout], taps rotated
            rows = (cin + 15) // 16 * 16
            buf = torch.zeros((rows, k, k, cout), dtype=torch.bfloat16)
            buf[:cin] = t
            dgrad_arena[dst:dst + buf.numel()] = buf.reshape(-1)

    def adam_ema_step(self, p, g, m, v, ema, sumsq, max_norm, grad_scale, lr, b1, b2, eps, decay, step,
                      warmup_steps=0, step_dev=None, shadow=None, weight_decay=0.0):
        if step_dev is not None:
            step = int(step_dev)
        if warmup_steps > 0:
            lr = lr * min(step - 1, warmup_steps) / warmup_steps
        clip = grad_scale
        if sumsq is not None and max_norm > 0:
            norm = float(sumsq.sqrt()) * grad_scale
            clip *= min(1.0, max_norm / (norm + 1e-6))
        gg = g * clip
        m.mul_(b1).add_(gg, alpha=1 - b1)
        v.mul_(b2).addcmul_(gg, gg, value=1 - b2)
        bc1, bc2 = 1 - b1 ** step, math.sqrt(1 - b2 ** step)
        p.mul_(1 - lr * weight_decay)
        p.sub_((lr / bc1) * m / (v.sqrt() / bc2 + eps))
        if ema is not None:
            ema.mul_(decay).add_(p, alpha=1 - decay)
        if shadow is not None:
            shadow.copy_(p)
