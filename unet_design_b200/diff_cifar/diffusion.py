"""Drop-in for the training half of the reference's diff_cifar/diffusion.py (GaussianDiffusionTrainer,
:17-91): DDPM Algorithm 1 around the B200 model, with the multi-resolution noise targets
(`LL_k(noise) / 2^k`, :52-78) computed by the fused Haar kernel instead of rebuilding
`DWTForward` / `DWTInverse` modules and moving their filters to the device on every step.

`GaussianDiffusionSampler` (:94-222) is the inference side of the same boundary: DDPM Algorithm 2,
T sequential model calls.  One reverse step (time-step vector, model forward, posterior mean, noise) is captured as a
CUDA graph whose time index lives on the device, so the T-step loop is T graph replays with no host synchronisation."""
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
        if self.multi_res_loss and not self.sequ_train_algo and len(model_out) == self.model.n_levels:
            # all levels present (coarse -> fine): target pyramid, the per-level MSEs and their gradients in ONE kernel
            fused = ops.multires_mse(noise, list(model_out)[::-1])
            if fused is not None:
                loss, per_level = fused
                return loss, list(per_level.flip(0).unbind(0))
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


class GaussianDiffusionSampler(nn.Module):
    """Drop-in for diff_cifar/diffusion.py:94-222 (same constructor, buffers and methods).

    Reference quirk kept: the reference asserts `mean_type in ['xprev' 'xstart', 'epsilon']` (missing
    comma), so its default `'eps'` is rejected and main.py passes `'epsilon'`; `'eps'` is rejected here too, the three
    spelled-out modes are accepted."""

    def __init__(self, model, beta_1, beta_T, T, img_size=32, mean_type='eps', var_type='fixedlarge', multi_res_loss=False):
        assert mean_type in ['xprev', 'xstart', 'epsilon', 'xprevxstart']
        assert var_type in ['fixedlarge', 'fixedsmall']
        super().__init__()
        self.model = model
        self.T = T
        self.img_size = img_size
        self.mean_type = mean_type
        self.var_type = var_type
        self.multi_res_loss = multi_res_loss
        self.register_buffer('betas', torch.linspace(beta_1, beta_T, T).double())
        alphas = 1. - self.betas
        alphas_bar = torch.cumprod(alphas, dim=0)
        alphas_bar_prev = F.pad(alphas_bar, [1, 0], value=1)[:T]
        self.register_buffer('sqrt_recip_alphas_bar', torch.sqrt(1. / alphas_bar))
        self.register_buffer('sqrt_recipm1_alphas_bar', torch.sqrt(1. / alphas_bar - 1))
        self.register_buffer('posterior_var', self.betas * (1. - alphas_bar_prev) / (1. - alphas_bar))
        self.register_buffer('posterior_log_var_clipped',
                             torch.log(torch.cat([self.posterior_var[1:2], self.posterior_var[1:]])))
        self.register_buffer('posterior_mean_coef1', torch.sqrt(alphas_bar_prev) * self.betas / (1. - alphas_bar))
        self.register_buffer('posterior_mean_coef2', torch.sqrt(alphas) * (1. - alphas_bar_prev) / (1. - alphas_bar))
        self.use_cuda_graph = True
        self._graphs = {}

    # ---- the reference's helper methods (:139-200)
    def q_mean_variance(self, x_0, x_t, t):
        assert x_0.shape == x_t.shape
        mean = (extract(self.posterior_mean_coef1, t, x_t.shape) * x_0
                + extract(self.posterior_mean_coef2, t, x_t.shape) * x_t)
        return mean, extract(self.posterior_log_var_clipped, t, x_t.shape)

    def predict_xstart_from_eps(self, x_t, t, eps):
        assert x_t.shape == eps.shape
        return (extract(self.sqrt_recip_alphas_bar, t, x_t.shape) * x_t
                - extract(self.sqrt_recipm1_alphas_bar, t, x_t.shape) * eps)

    def predict_xstart_from_xprev(self, x_t, t, xprev):
        assert x_t.shape == xprev.shape
        return (extract(1. / self.posterior_mean_coef1, t, x_t.shape) * xprev
                - extract(self.posterior_mean_coef2 / self.posterior_mean_coef1, t, x_t.shape) * x_t)

    def _model_out(self, x_t, t, n_levels_used):
        out = self.model(x_t, t, n_levels_used=n_levels_used)
        return out[-1] if self.multi_res_loss else out          # the finest output (:178, :185, :192)

    def p_mean_variance(self, x_t, t, n_levels_used):
        log_var = {'fixedlarge': torch.log(torch.cat([self.posterior_var[1:2], self.betas[1:]])),
                   'fixedsmall': self.posterior_log_var_clipped}[self.var_type]
        log_var = extract(log_var, t, x_t.shape)
        out = self._model_out(x_t, t, n_levels_used)
        if self.mean_type == 'xprev':
            mean = out
        elif self.mean_type == 'xstart':
            mean, _ = self.q_mean_variance(out, x_t, t)
        elif self.mean_type == 'epsilon':
            mean, _ = self.q_mean_variance(self.predict_xstart_from_eps(x_t, t, eps=out), x_t, t)
        else:
            raise NotImplementedError(self.mean_type)
        return mean, log_var        # the reference clips x_0 AFTER the mean is formed (:199): no effect on the result

    def _step(self, x_t, t, n_levels_used, noise):
        mean, log_var = self.p_mean_variance(x_t, t, n_levels_used)
        return mean + torch.exp(0.5 * log_var) * noise

    @torch.no_grad()
    def forward(self, x_T, n_levels_used, noises=None):
        """Algorithm 2 (:207-222).  `noises` (optional, tests): the T-1 Gaussian draws in the order the loop uses them."""
        if noises is None and self.use_cuda_graph and x_T.is_cuda and not self.training:
            return self._forward_graphed(x_T, n_levels_used)
        x_t = x_T
        for i, time_step in enumerate(reversed(range(self.T))):
            t = x_t.new_ones([x_T.shape[0], ], dtype=torch.long) * time_step
            if time_step > 0:
                noise = noises[i] if noises is not None else torch.randn_like(x_t)
            else:
                noise = torch.zeros_like(x_t)
            x_t = self._step(x_t, t, n_levels_used, noise)
        return torch.clip(x_t, -1, 1)

    def _forward_graphed(self, x_T, n_levels_used):
        key = (tuple(x_T.shape), n_levels_used, x_T.device)
        g = self._graphs.get(key)
        if g is None:
            g = self._capture(x_T, n_levels_used)
            self._graphs[key] = g
        graph, x_buf, t_dev = g
        x_buf.copy_(x_T)
        t_dev.fill_(self.T - 1)
        for _ in range(self.T):
            graph.replay()
        return torch.clip(x_buf, -1, 1).clone()

    def _capture(self, x_T, n_levels_used):
        dev = x_T.device
        x_buf = x_T.clone()
        t_dev = torch.full((1,), self.T - 1, dtype=torch.long, device=dev)

        def one_step():
            t = t_dev.expand(x_buf.shape[0])
            noise = torch.randn_like(x_buf) * (t_dev > 0).to(x_buf.dtype)        # no noise at t == 0 (:217-220)
            x_buf.copy_(self._step(x_buf, t, n_levels_used, noise))
            t_dev.sub_(1).clamp_(min=0)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):                        # warm-up outside capture: allocator, tensor maps, cuBLAS
            for _ in range(2):
                one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            one_step()
        return graph, x_buf, t_dev
