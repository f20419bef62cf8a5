"""Generate tests/golden/*.pt by running the reference's OWN module classes.

Runs only in the build container (needs /root/reference, read-only).  The reference's
`pytorch_wavelets` dependency is absent from the image, so `sys.modules` maps it to
`oracle.pytorch_wavelets_restated` (parity of the Haar arithmetic itself is therefore
unpinned, see oracle/__init__.py); everything else -- ResBlock, UpSample, AttnBlock,
DTWBlock's scale + channel-tile logic, UNetWaveletEnc's level bookkeeping, the DDPM loss
-- is the reference's code, unmodified.

    python tools/make_golden.py            # rewrites tests/golden/
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import pytorch_wavelets_restated as pw  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_module(name: str, path: str):
    sys.modules["pytorch_wavelets"] = pw
    if "matplotlib" not in sys.modules:            # diff_mnist/mnist_diff/unet.py:6 imports pyplot
        mpl = types.ModuleType("matplotlib")
        mpl.pyplot = types.ModuleType("matplotlib.pyplot")
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = mpl.pyplot
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def grads_of(module, names):
    params = dict(module.named_parameters())
    return {n: params[n].grad.detach().clone() for n in names}


def cifar_goldens():
    m = load_reference_module("ref_cifar_model", f"{REF}/diff_cifar/model.py")
    d = load_reference_module("ref_cifar_diffusion", f"{REF}/diff_cifar/diffusion.py")

    # --- one ResBlock with a 1x1 shortcut, and one without (reference diff_cifar/model.py:122-169)
    for tag, cin, cout, attn in (("resblock_sc", 96, 64, False), ("resblock_id", 64, 64, False),
                                 ("resblock_attn", 64, 64, True)):
        torch.manual_seed(1234)
        blk = m.ResBlock(cin, cout, tdim=128, dropout=0.0, attn=attn).double().float()
        # conv2 starts at gain 1e-5: give it weight so its gradient path is exercised
        with torch.no_grad():
            blk.block2[-1].weight.mul_(1e5 * 0.5)
        torch.manual_seed(0)
        x = torch.randn(4, cin, 8, 8, requires_grad=True)
        temb = torch.randn(4, 128, requires_grad=True)
        y = blk(x, temb)
        gy = torch.randn_like(y)
        y.backward(gy)
        torch.save({"state": blk.state_dict(), "x": x.detach(), "temb": temb.detach(), "y": y.detach(),
                    "gy": gy, "gx": x.grad, "gtemb": temb.grad,
                    "gparams": {n: p.grad for n, p in blk.named_parameters()},
                    "cfg": dict(in_ch=cin, out_ch=cout, tdim=128, dropout=0.0, attn=attn)},
                   f"{OUT}/cifar_{tag}.pt")

    # --- UpSample (model.py:66-81)
    torch.manual_seed(1234)
    up = m.UpSample(32)
    torch.manual_seed(0)
    x = torch.randn(3, 32, 4, 4, requires_grad=True)
    y = up(x, None)
    gy = torch.randn_like(y)
    y.backward(gy)
    torch.save({"state": up.state_dict(), "x": x.detach(), "y": y.detach(), "gy": gy, "gx": x.grad,
                "gparams": {n: p.grad for n, p in up.named_parameters()}}, f"{OUT}/cifar_upsample.pt")

    # --- DTWBlock scale + tile logic (model.py:253-323)
    torch.manual_seed(0)
    cases = []
    for (shape, J, out_ch) in (((2, 3, 8, 8), 0, 32), ((2, 32, 8, 8), 1, 32), ((2, 32, 8, 8), 1, 80),
                               ((1, 5, 7, 9), 1, 12), ((1, 3, 16, 16), 2, 7), ((1, 2, 25, 13), 3, 2)):
        x = torch.randn(*shape)
        y = m.DTWBlock(J=J, out_channels=out_ch)(x)
        cases.append({"x": x, "J": J, "out_channels": out_ch, "y": y})
    torch.save(cases, f"{OUT}/cifar_dtwblock.pt")

    # --- whole models: Multi-ResNet (Haar encoder) and the residual U-Net arm (model.py:326-496)
    for tag, kw in (("multiresnet", dict(dwt_encoder=True, multi_res_loss=True)),
                    ("unet", dict(dwt_encoder=False, multi_res_loss=False))):
        torch.manual_seed(1234)
        cfg = dict(T=20, ch=32, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0, **kw)
        net = m.UNetWaveletEnc(**cfg)
        with torch.no_grad():  # lift the 1e-5-gain convs so every gradient is well above round-off
            for n, p in net.named_parameters():
                if p.dim() == 4 and p.abs().max() < 1e-3:
                    p.mul_(3e4)
        trainer = d.GaussianDiffusionTrainer(net, 1e-4, 0.02, 20, kw["multi_res_loss"], False, "cpu")
        torch.manual_seed(0)
        x0 = torch.randn(4, 3, 16, 16)
        # replay the trainer's own RNG draws so that oracle/torch_ref.loss_from can reproduce them
        torch.manual_seed(7)
        t = torch.randint(20, size=(4,))
        noise = torch.randn_like(x0)
        torch.manual_seed(7)
        loss, loss_list = trainer(x0, n_levels_used=-1)
        loss.backward()
        x_t = (d.extract(trainer.sqrt_alphas_bar, t, x0.shape) * x0
               + d.extract(trainer.sqrt_one_minus_alphas_bar, t, x0.shape) * noise)
        with torch.no_grad():
            out = net(x_t, t)
            out_1lvl = net(x_t[:, :, ::2, ::2].contiguous(), t, n_levels_used=1)
        torch.save({"cfg": cfg, "state": net.state_dict(), "x0": x0, "t": t, "noise": noise, "x_t": x_t,
                    "out": out, "out_1lvl": out_1lvl, "loss": loss.detach(),
                    "loss_list": [l.detach() for l in loss_list],
                    # every small gradient plus one conv weight per block kind (keeps the fixture ~1.5 MB)
                    "gparams": {n: p.grad for n, p in net.named_parameters() if p.grad is not None
                                and (p.numel() <= 4096 or n.endswith("block1.2.weight"))}},
                   f"{OUT}/cifar_{tag}.pt")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    cifar_goldens()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
