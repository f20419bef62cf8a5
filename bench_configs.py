"""BASELINE.json configs 1, 3, 4, 5 for bench.py --config (config 2, the headline, lives in bench.py itself).

Each entry builds (a) the drop-in model + loss closure for `unet_design_b200.train.TrainStep`, (b) the synthetic batch
(SURVEY.md 8d shapes, seeds, optimisers) and, where an fp32 restatement exists in oracle/, (c) the same step on the reference
algorithm for the CPU baseline and the GPU library (cuDNN) comparator.  FLOPs per sample are the fwd+bwd totals PyTorch's
FlopCounterMode reports for the reference models."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _pde_cfg(time_history, dwt, extra=0):
    return dict(n_input_scalar_components=1, n_input_vector_components=1, n_output_scalar_components=1,
                n_output_vector_components=1, time_history=time_history, time_future=1, hidden_channels=64,
                activation="gelu", dwt_encoder=dwt, n_extra_resnet_layers=extra)


def custom_mse(pred, target):
    """pdearena/pdearena/modules/loss.py `custommse_loss`: squared error summed over time and fields, mean over the rest."""
    return ((pred - target) ** 2).sum(dim=(1, 2)).mean()


def dice_loss(y_true, y_pred, smooth=1.0):
    """wmh/train_pt.py:102-112."""
    inter = torch.sum(torch.flatten(y_true) * torch.flatten(y_pred))
    return 1.0 - (2.0 * inter + smooth) / (torch.sum(y_true) + torch.sum(y_pred) + smooth)


class _MnistDiffusion:
    """The forward process of diff_mnist (torch_ddpm/ddpm/diffusion.py:43-100): Diffusion(0.1, 20, N=30)."""

    def __init__(self, dev, n=30, beta_min=0.1, beta_max=20.0):
        betas = torch.linspace(beta_min / n, beta_max / n, n)
        acp = torch.cumprod(1.0 - betas, dim=0)
        self.n, self.a, self.b = n, acp.sqrt().to(dev), (1.0 - acp).sqrt().to(dev)

    def loss(self, model, x0):
        t = torch.randint(self.n, (x0.shape[0],), device=x0.device)
        noise = torch.randn_like(x0)
        x_t = self.a[t].view(-1, 1, 1, 1) * x0 + self.b[t].view(-1, 1, 1, 1) * noise
        out, _ = model(x_t, t.unsqueeze(-1))
        return torch.mean(torch.mean(torch.square(out - noise).reshape(x0.shape[0], -1), dim=-1))


CONFIGS = {
    # name: (description, batch per GPU, unit, GFLOP per sample fwd+bwd)
    "c1": ("diff_mnist Multi-ResNet (Haar encoder) DDPM train step, synthetic 1x32x32, batch 64", 64, "images/s", 3.23),
    "c3": ("pdearena Navier-Stokes 2D Multi-ResNet surrogate (Unetbase-64_G, Haar encoder), synthetic 4x3x128x128 -> 1x3x128x128, "
           "batch 8 per GPU", 8, "samples/s", 55.0),
    "c3u": ("pdearena Navier-Stokes 2D residual U-Net arm (Unetbase-64_G, dwt_encoder=False), synthetic 4x3x128x128, batch 8 per GPU",
            8, "samples/s", 76.7),
    "c4": ("pdearena shallow-water 2D Multi-ResNet (Unetbase-64_G), synthetic 2x3x96x192 fields, batch 16 per GPU, bf16",
           16, "samples/s", 61.6),
    "c5": ("wmh segmentation Multi-ResNet (Unetbase_G hidden 16, Haar encoder, 200 -> 13 odd extents), synthetic 2x200x200 slices, "
           "batch 32 per GPU", 32, "samples/s", 8.42),
}


def build(name: str, dev, rank: int = 0):
    """-> dict(model, loss_fn(*batch), host_batch() -> tuple of pinned tensors, opt kwargs for TrainStep)."""
    from unet_design_b200.pdearena.modules.twod_unetbase import Unetbase_G
    from unet_design_b200.wmh.model import Unetbase_G as WmhUnet
    gen = torch.Generator().manual_seed(rank)
    torch.manual_seed(1234)
    batch = CONFIGS[name][1]
    if name in ("c3", "c3u", "c4"):
        th, hw = (4, (128, 128)) if name != "c4" else (2, (96, 192))
        model = Unetbase_G(**_pde_cfg(th, name != "c3u")).to(dev)

        def host_batch():
            return (torch.randn(batch, th, 3, *hw, generator=gen).pin_memory(), torch.randn(batch, 1, 3, *hw, generator=gen).pin_memory())

        return dict(model=model, loss_fn=lambda x, y: custom_mse(model(x), y), host_batch=host_batch,
                    opt=dict(lr=2e-4, weight_decay=1e-5))
    if name == "c5":
        model = WmhUnet(hidden_channels=16, dwt_encoder=True).to(dev)

        def host_batch():
            return (torch.randn(batch, 2, 200, 200, generator=gen).pin_memory(),
                    (torch.rand(batch, 1, 200, 200, generator=gen) < 0.01).float().pin_memory())

        return dict(model=model, loss_fn=lambda x, m: dice_loss(m, model(x)), host_batch=host_batch, opt=dict(lr=2e-4))
    if name == "c1":
        from unet_design_b200.diff_mnist.unet import get_unet_wavelet
        model = get_unet_wavelet(32, 1, num_channels=32, dropout=0.0, num_res_blocks=2, dwt_encoder=True).to(dev)
        diff = _MnistDiffusion(dev)

        def host_batch():
            return (torch.randn(batch, 1, 32, 32, generator=gen).pin_memory(),)

        return dict(model=model, loss_fn=lambda x0: diff.loss(model, x0), host_batch=host_batch, opt=dict(lr=1e-3))
    raise KeyError(name)


def build_reference(name: str, dev):
    """The reference algorithm (oracle/ restatement, fp32) for the CPU baseline / GPU library comparator, or None."""
    from oracle import torch_ref_pde
    torch.manual_seed(1234)
    if name in ("c3", "c3u", "c4"):
        th = 4 if name != "c4" else 2
        model = torch_ref_pde.from_reference_cfg(_pde_cfg(th, name != "c3u")).to(dev)
        opt = torch.optim.AdamW(model.parameters(), lr=2e-4, weight_decay=1e-5)
        return model, opt, (lambda x, y: custom_mse(model(x), y))
    if name == "c5":
        model = torch_ref_pde.from_reference_cfg(dict(hidden_channels=16, dwt_encoder=True), wmh=True).to(dev)
        opt = torch.optim.Adam(model.parameters(), lr=2e-4)
        return model, opt, (lambda x, m: dice_loss(m, model(x)))
    return None            # c1: no fp32 restatement of UNet_wavelet in oracle/ (its CPU number is quoted from profiles/)
