"""Drop-in for the training half of the reference's diff_cifar/diffusion.py (GaussianDiffusionTrainer,
:17-91): DDPM Algorithm 1 around the B200 model, with the multi-resolution noise targets
(`LL_k(noise) / 2^k`, :52-78) computed by the fused Haar kernel instead of rebuilding
`DWTForward` / `DWTInverse` modules and moving their filters to the device on every step.

The sampler (:94-222) is out of scope."""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

from .. import ops


def extract(v, t, x_shape):
    """Gather per-sample coefficients and reshape to [B, 1, 1, 1] (diffusion.py:8-14)."""
    out = torch.gather(v, index=t, dim=0).float()
    return out.view([t.shape[0]] + [1] * (len(x_shape) - 1))


class GaussianDiffusionTrainer(nn.Module):
    def __init__(self, model, beta_1, beta_T, T, multi_res_loss=False, sequ_train_algo=False, device=None):
        super().__init__()
        self.model = model
        self.T = T
        self.multi_res_loss = multi_res_loss
        self.sequ_train_algo = sequ_train_algo
        self.device = device
        self.register_buffer("betas", torch.linspace(beta_1, beta_T, T).double())
        alphas = 1.0 - self.betas
        alphas_bar = torch.cumprod(alphas, dim=0)
        self.register_buffer("sqrt_alphas_bar", torch.sqrt(alphas_bar))
        self.register_buffer("sqrt_one_minus_alphas_bar", torch.sqrt(1.0 - alphas_bar))

    def loss_from(self, x_0, t, noise, n_levels_used=-1, n_downsample=0):
        """The deterministic part of `forward` (t and noise supplied by the caller)."""
        x_t = (extract(self.sqrt_alphas_bar, t, x_0.shape) * x_0
               + extract(self.sqrt_one_minus_alphas_bar, t, x_0.shape) * noise)
        model_out = self.model(x_t, t, n_levels_used=n_levels_used)
        loss_list = []
        if self.multi_res_loss:
            targets = []
            for k in list(range(0, self.model.n_levels))[::-1]:
                if self.sequ_train_algo:
                    k = k - n_downsample
                if k > 0:
                    targets.append(ops.dwtblock(noise, k, noise.shape[1]))     # LL_k(noise) / 2^k, one kernel
                elif k == 0:
                    targets.append(noise)
            loss = 0.0
            for out, n in zip(model_out, targets):
                loss_res = F.mse_loss(out, n, reduction="none").mean()
                loss = loss + loss_res
                loss_list.append(loss_res)
        else:
            loss = F.mse_loss(model_out, noise, reduction="none").mean()
        return loss, loss_list

    def forward(self, x_0, n_levels_used=-1, n_downsample=0):
        """Algorithm 1."""
        t = torch.randint(self.T, size=(x_0.shape[0],), device=x_0.device)
        noise = torch.randn_like(x_0)
        return self.loss_from(x_0, t, noise, n_levels_used, n_downsample)
