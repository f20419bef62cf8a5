"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Plain PyTorch fp32 restatement of the reference's diff_cifar Multi-ResNet
(`UNetWaveletEnc` and its blocks) and of the DDPM training loss, used

* by the GPU parity tests as the floating-point oracle of the conv blocks
  (/root/reference does not exist on the GPU box), and
* by `bench.py` as the CPU baseline / `--impl reference` arm.

It is pinned: `tools/make_golden.py` imports the reference's own classes from
/root/reference, copies one `state_dict` into both, and stores input/output/gradient
vectors under tests/golden/; `tests/test_oracle_golden.py` replays them through
this file.  `state_dict` keys and shapes equal the reference's, so checkpoints move
between the reference, this oracle and the CUDA drop-in.

Reference (file:line): diff_cifar/model.py:9-11 Swish, :14-43 TimeEmbedding,
:46-63 DownSample, :66-81 UpSample, :84-119 AttnBlock, :122-169 ResBlock,
:253-323 DTWBlock, :326-496 UNetWaveletEnc; diff_cifar/diffusion.py:17-91 trainer.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .pytorch_wavelets_restated import DWTForward, DWTInverse


class Swish(nn.Module):
    def forward(self, x):  # model.py:9-11
        return x * torch.sigmoid(x)


def _xavier(*mods, gain=1.0):
    for m in mods:
        nn.init.xavier_uniform_(m.weight, gain=gain)
        nn.init.zeros_(m.bias)


class TimeEmbedding(nn.Module):  # model.py:14-43
    def __init__(self, T, d_model, dim):
        super().__init__()
        freq = torch.exp(-(torch.arange(0, d_model, 2) / d_model * math.log(10000)))
        ang = torch.arange(T).float()[:, None] * freq[None, :]
        table = torch.stack([ang.sin(), ang.cos()], dim=-1).reshape(T, d_model)
        self.timembedding = nn.Sequential(
            nn.Embedding.from_pretrained(table), nn.Linear(d_model, dim), Swish(), nn.Linear(dim, dim))
        _xavier(self.timembedding[1], self.timembedding[3])

    def forward(self, t):
        return self.timembedding(t)


class DownSample(nn.Module):  # model.py:46-63
    def __init__(self, in_ch, type="conv"):
        super().__init__()
        if type == "conv":
            self.main = nn.Conv2d(in_ch, in_ch, 3, stride=2, padding=1)
            _xavier(self.main)
        elif type == "avg_pool":
            self.main = nn.AvgPool2d(2)
        else:
            raise NotImplementedError

    def forward(self, x, temb):
        return self.main(x)


class UpSample(nn.Module):  # model.py:66-81
    def __init__(self, in_ch):
        super().__init__()
        self.main = nn.Conv2d(in_ch, in_ch, 3, stride=1, padding=1)
        _xavier(self.main)

    def forward(self, x, temb):
        return self.main(F.interpolate(x, scale_factor=2, mode="nearest"))


class AttnBlock(nn.Module):  # model.py:84-119
    def __init__(self, in_ch):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, in_ch)
        self.proj_q = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_k = nn.Conv2d(in_ch, in_ch, 1)
        self.proj_v = nn.Conv2d(in_ch, in_ch, 1)
        self.proj = nn.Conv2d(in_ch, in_ch, 1)
        _xavier(self.proj_q, self.proj_k, self.proj_v)
        _xavier(self.proj, gain=1e-5)

    def forward(self, x):
        b, c, h, w = x.shape
        y = self.group_norm(x)
        q = self.proj_q(y).flatten(2).transpose(1, 2)          # [B, HW, C]
        k = self.proj_k(y).flatten(2)                          # [B, C, HW]
        v = self.proj_v(y).flatten(2).transpose(1, 2)          # [B, HW, C]
        p = F.softmax(torch.bmm(q, k) * (int(c) ** (-0.5)), dim=-1)
        o = torch.bmm(p, v).transpose(1, 2).reshape(b, c, h, w)
        return x + self.proj(o)


class ResBlock(nn.Module):  # model.py:122-169
    def __init__(self, in_ch, out_ch, tdim, dropout, attn=False):
        super().__init__()
        self.in_ch, self.out_ch = in_ch, out_ch
        self.block1 = nn.Sequential(nn.GroupNorm(32, in_ch), Swish(), nn.Conv2d(in_ch, out_ch, 3, padding=1))
        self.temb_proj = nn.Sequential(Swish(), nn.Linear(tdim, out_ch))
        self.block2 = nn.Sequential(nn.GroupNorm(32, out_ch), Swish(), nn.Dropout(dropout),
                                    nn.Conv2d(out_ch, out_ch, 3, padding=1))
        self.shortcut = nn.Conv2d(in_ch, out_ch, 1) if in_ch != out_ch else nn.Identity()
        self.attn = AttnBlock(out_ch) if attn else nn.Identity()
        # model.py:155-160 re-draws EVERY conv/linear below this block, the nested attention's too
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Linear)):
                _xavier(m)
        nn.init.xavier_uniform_(self.block2[-1].weight, gain=1e-5)

    def forward(self, x, temb):
        h = self.block1(x)
        h = h + self.temb_proj(temb)[:, :, None, None]
        h = self.block2(h)
        return self.attn(h + self.shortcut(x))


class DTWBlock(nn.Module):  # model.py:253-323 (version 1, the only live branch)
    def __init__(self, J, out_channels, mode="zero", wave="haar"):
        super().__init__()
        self.J, self.out_channels = J, out_channels
        self.xfm = DWTForward(J=J, mode=mode, wave=wave)
        self.ifm = DWTInverse(mode=mode, wave=wave)

    def forward(self, x):
        if self.J > 0:
            yl, _ = self.xfm(x)
            x = self.ifm((yl, [])) / math.pow(2, self.J)
        reps = int(self.out_channels / x.shape[1]) + 1
        return x.repeat(1, reps, 1, 1)[:, : self.out_channels]


class UNetWaveletEnc(nn.Module):  # model.py:326-496
    def __init__(self, T, ch, ch_mult, attn, num_res_blocks, dropout, dwt_encoder=False,
                 multi_res_loss=False, downsample_type="conv"):
        super().__init__()
        assert all(i < len(ch_mult) for i in attn), "attn index out of bound"
        tdim = ch * 4
        self.n_levels = len(ch_mult)
        self.dwt_encoder, self.multi_res_loss, self.downsample_type = dwt_encoder, multi_res_loss, downsample_type
        self.time_embedding_list = nn.ModuleList(TimeEmbedding(T, ch, tdim) for _ in range(self.n_levels))
        self.head_list = nn.ModuleList()
        self.downblocks = nn.ModuleList(nn.ModuleList() for _ in range(self.n_levels))
        widths, cur = [ch], ch
        for lvl, mult in enumerate(ch_mult):
            self.head_list.append(DTWBlock(J=0, out_channels=cur))
            out_ch = ch * mult
            for _ in range(num_res_blocks):
                self.downblocks[lvl].append(
                    DTWBlock(J=0, out_channels=out_ch) if dwt_encoder
                    else ResBlock(cur, out_ch, tdim, dropout, attn=(lvl in attn)))
                cur = out_ch
                widths.append(cur)
            if lvl != self.n_levels - 1:
                self.downblocks[lvl].append(
                    DTWBlock(J=1, out_channels=cur) if dwt_encoder else DownSample(cur, type=downsample_type))
                widths.append(cur)
        self.middleblocks = nn.ModuleList([ResBlock(cur, cur, tdim, dropout, attn=True),
                                           ResBlock(cur, cur, tdim, dropout, attn=False)])
        self.upblocks = nn.ModuleList(nn.ModuleList() for _ in range(self.n_levels))
        for lvl in reversed(range(self.n_levels)):
            out_ch = ch * ch_mult[lvl]
            for _ in range(num_res_blocks + 1):
                self.upblocks[lvl].append(ResBlock(widths.pop() + cur, out_ch, tdim, dropout, attn=(lvl in attn)))
                cur = out_ch
            if lvl != 0:
                self.upblocks[lvl].append(UpSample(cur))
        assert not widths
        self.tail_list = nn.ModuleList(
            nn.Sequential(nn.GroupNorm(32, ch * m), Swish(), nn.Conv2d(ch * m, 3, 3, padding=1)) for m in ch_mult)
        for tail in self.tail_list:
            _xavier(tail[-1], gain=1e-5)

    def forward(self, x, t, n_levels_used=-1):
        n = self.n_levels if n_levels_used == -1 else n_levels_used
        first = self.n_levels - n                      # finest level in use
        h = self.head_list[-n](x)
        skips = [h]
        for lvl in range(first, self.n_levels):
            temb = self.time_embedding_list[lvl](t)
            for layer in self.downblocks[lvl]:
                h = layer(h) if self.dwt_encoder else layer(h, temb)
                skips.append(h)
        temb = self.time_embedding_list[self.n_levels - 1](t)
        for layer in self.middleblocks:
            h = layer(h, temb)
        outs = []
        for lvl in range(self.n_levels - 1, first - 1, -1):
            for layer in self.upblocks[lvl]:
                if isinstance(layer, ResBlock):
                    h = layer(torch.cat([h, skips.pop()], dim=1), self.time_embedding_list[lvl](t))
                elif lvl != first:                      # UpSample; skipped on the finest level in use
                    if self.multi_res_loss:
                        outs.append(self.tail_list[lvl](h))
                    h = layer(h, None)
        outs.append(self.tail_list[first](h))
        assert not skips
        return outs if self.multi_res_loss else outs[-1]


class GaussianDiffusionTrainer(nn.Module):  # diffusion.py:17-91
    def __init__(self, model, beta_1, beta_T, T, multi_res_loss=False, sequ_train_algo=False, device=None):
        super().__init__()
        self.model, self.T = model, T
        self.multi_res_loss, self.sequ_train_algo, self.device = multi_res_loss, sequ_train_algo, device
        self.register_buffer("betas", torch.linspace(beta_1, beta_T, T).double())
        abar = torch.cumprod(1.0 - self.betas, dim=0)
        self.register_buffer("sqrt_alphas_bar", abar.sqrt())
        self.register_buffer("sqrt_one_minus_alphas_bar", (1.0 - abar).sqrt())

    def q_sample(self, x_0, t, noise):
        a = self.sqrt_alphas_bar[t].float().view(-1, 1, 1, 1)
        b = self.sqrt_one_minus_alphas_bar[t].float().view(-1, 1, 1, 1)
        return a * x_0 + b * noise

    def loss_from(self, x_0, t, noise, n_levels_used=-1, n_downsample=0):
        """Deterministic part of diffusion.py:38-91 (t and noise supplied by the caller)."""
        out = self.model(self.q_sample(x_0, t, noise), t, n_levels_used=n_levels_used)
        if not self.multi_res_loss:
            return F.mse_loss(out, noise, reduction="none").mean(), []
        targets = []
        for k in reversed(range(self.model.n_levels)):
            if self.sequ_train_algo:
                k -= n_downsample
            if k > 0:
                yl, _ = DWTForward(J=k, mode="zero", wave="haar")(noise)
                targets.append(yl / math.pow(2, k))
            elif k == 0:
                targets.append(noise)
        losses = [F.mse_loss(o, n, reduction="none").mean() for o, n in zip(out, targets)]
        return sum(losses), losses

    def forward(self, x_0, n_levels_used=-1, n_downsample=0):
        t = torch.randint(self.T, size=(x_0.shape[0],), device=x_0.device)
        return self.loss_from(x_0, t, torch.randn_like(x_0), n_levels_used, n_downsample)


class GaussianDiffusionSampler(nn.Module):  # diffusion.py:94-222 (mean_type 'epsilon' / 'xstart' / 'xprev')
    def __init__(self, model, beta_1, beta_T, T, mean_type="epsilon", var_type="fixedlarge", multi_res_loss=False):
        super().__init__()
        self.model, self.T, self.mean_type, self.var_type, self.multi_res_loss = model, T, mean_type, var_type, multi_res_loss
        betas = torch.linspace(beta_1, beta_T, T).double()
        alphas = 1.0 - betas
        abar = torch.cumprod(alphas, dim=0)
        abar_prev = torch.cat([torch.ones(1, dtype=torch.float64), abar[:-1]])
        self.betas = betas
        self.recip = (1.0 / abar).sqrt()
        self.recipm1 = (1.0 / abar - 1.0).sqrt()
        self.post_var = betas * (1.0 - abar_prev) / (1.0 - abar)
        self.post_logvar = torch.log(torch.cat([self.post_var[1:2], self.post_var[1:]]))
        self.c1 = abar_prev.sqrt() * betas / (1.0 - abar)
        self.c2 = alphas.sqrt() * (1.0 - abar_prev) / (1.0 - abar)

    @torch.no_grad()
    def forward(self, x_T, n_levels_used, noises):
        """`noises[i]` is the draw of loop iteration i (time step T-1-i); the last iteration (t = 0) adds none."""
        x = x_T
        logvar = {"fixedlarge": torch.log(torch.cat([self.post_var[1:2], self.betas[1:]])),
                  "fixedsmall": self.post_logvar}[self.var_type]
        for i, ts in enumerate(reversed(range(self.T))):
            t = torch.full((x.shape[0],), ts, dtype=torch.long)
            out = self.model(x, t, n_levels_used=n_levels_used)
            out = out[-1] if self.multi_res_loss else out
            if self.mean_type == "xprev":
                mean = out
            else:
                x0 = out if self.mean_type == "xstart" else self.recip[ts].float() * x - self.recipm1[ts].float() * out
                mean = self.c1[ts].float() * x0 + self.c2[ts].float() * x
            noise = noises[i] if ts > 0 else torch.zeros_like(x)
            x = mean + torch.exp(0.5 * logvar[ts].float()) * noise
        return torch.clip(x, -1, 1)
