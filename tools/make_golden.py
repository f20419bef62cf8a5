"""Generate tests/golden/*.pt by running the reference's OWN module classes.

Runs only in the build container (needs /root/reference, read-only).  The reference's `pytorch_wavelets`
dependency is absent from the image, so `sys.modules` maps it to `oracle.pytorch_wavelets_restated` (parity of
the Haar arithmetic itself is therefore unpinned, see oracle/__init__.py); everything else -- ResBlock, UpSample,
AttnBlock, DTWBlock's scale + channel-tile logic, UNetWaveletEnc's level bookkeeping, the DDPM loss, pdearena's
conv blocks and Unetbase_G, wmh's odd-extent handling -- is the reference's code, unmodified.

Parameters are set by tests/det_init.py (a pure function of name and shape) so the fixtures carry inputs, outputs
and gradients only.

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
sys.path.insert(0, os.path.join(ROOT, "tests"))

from det_init import apply_det_init  # noqa: E402
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


def small_grads(module, also=()):
    return {n: p.grad.clone() for n, p in module.named_parameters()
            if p.grad is not None and (p.numel() <= 4096 or any(n.endswith(a) for a in also))}


def cifar_goldens():
    m = load_reference_module("ref_cifar_model", f"{REF}/diff_cifar/model.py")
    d = load_reference_module("ref_cifar_diffusion", f"{REF}/diff_cifar/diffusion.py")

    # --- ResBlock with a 1x1 shortcut, with an identity shortcut, with attention (diff_cifar/model.py:122-169)
    for tag, cin, cout, attn in (("resblock_sc", 96, 64, False), ("resblock_id", 64, 64, False),
                                 ("resblock_attn", 64, 64, True)):
        cfg = dict(in_ch=cin, out_ch=cout, tdim=128, dropout=0.0, attn=attn)
        blk = apply_det_init(m.ResBlock(**cfg))
        torch.manual_seed(0)
        x = torch.randn(4, cin, 8, 8, requires_grad=True)
        temb = torch.randn(4, 128, requires_grad=True)
        y = blk(x, temb)
        gy = torch.randn_like(y)
        y.backward(gy)
        torch.save({"cfg": cfg, "x": x.detach(), "temb": temb.detach(), "y": y.detach(), "gy": gy, "gx": x.grad,
                    "gtemb": temb.grad, "gparams": small_grads(blk, also=("block1.2.weight", "block2.3.weight"))},
                   f"{OUT}/cifar_{tag}.pt")

    # --- UpSample (model.py:66-81)
    up = apply_det_init(m.UpSample(32))
    torch.manual_seed(0)
    x = torch.randn(3, 32, 4, 4, requires_grad=True)
    y = up(x, None)
    gy = torch.randn_like(y)
    y.backward(gy)
    torch.save({"x": x.detach(), "y": y.detach(), "gy": gy, "gx": x.grad, "gparams": small_grads(up, also=("main.weight",))},
               f"{OUT}/cifar_upsample.pt")

    # --- DTWBlock scale + tile logic (model.py:253-323)
    torch.manual_seed(0)
    cases = []
    for (shape, J, out_ch) in (((2, 3, 8, 8), 0, 32), ((2, 32, 8, 8), 1, 32), ((2, 32, 8, 8), 1, 80),
                               ((1, 5, 7, 9), 1, 12), ((1, 3, 16, 16), 2, 7), ((1, 2, 25, 13), 3, 2)):
        x = torch.randn(*shape)
        cases.append({"x": x, "J": J, "out_channels": out_ch, "y": m.DTWBlock(J=J, out_channels=out_ch)(x)})
    torch.save(cases, f"{OUT}/cifar_dtwblock.pt")

    # --- whole models: Multi-ResNet (Haar encoder, multi-res loss) and the residual U-Net arm (model.py:326-496)
    for tag, kw in (("multiresnet", dict(dwt_encoder=True, multi_res_loss=True)),
                    ("unet", dict(dwt_encoder=False, multi_res_loss=False))):
        cfg = dict(T=20, ch=64, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0, **kw)
        net = apply_det_init(m.UNetWaveletEnc(**cfg))
        trainer = d.GaussianDiffusionTrainer(net, 1e-4, 0.02, 20, kw["multi_res_loss"], False, "cpu")
        torch.manual_seed(0)
        x0 = torch.randn(4, 3, 16, 16)
        torch.manual_seed(7)                      # replay the trainer's own draws so the oracle can reproduce them
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
        torch.save({"cfg": cfg, "x0": x0, "t": t, "noise": noise, "x_t": x_t, "out": out, "out_1lvl": out_1lvl,
                    "loss": loss.detach(), "loss_list": [l.detach() for l in loss_list],
                    "gparams": small_grads(net, also=("upblocks.0.0.block1.2.weight",))}, f"{OUT}/cifar_{tag}.pt")


def cifar_sampler_golden():
    """DDPM Algorithm 2 (diffusion.py:94-222) through the reference's own sampler on a small Multi-ResNet, with the
    loop's `torch.randn_like` draws recorded so that the oracle and the CUDA path can replay them."""
    m = load_reference_module("ref_cifar_model", f"{REF}/diff_cifar/model.py")
    d = load_reference_module("ref_cifar_diffusion", f"{REF}/diff_cifar/diffusion.py")
    cases = {}
    for var_type in ("fixedlarge", "fixedsmall"):
        cfg = dict(T=6, ch=64, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0, dwt_encoder=True, multi_res_loss=False)
        net = apply_det_init(m.UNetWaveletEnc(**cfg)).eval()
        sampler = d.GaussianDiffusionSampler(net, 1e-4, 0.02, 6, img_size=16, mean_type="epsilon", var_type=var_type)
        torch.manual_seed(3)
        x_T = torch.randn(2, 3, 16, 16)
        noises, real = [], torch.randn_like

        def recording(x):
            n = real(x)
            noises.append(n)
            return n

        torch.randn_like = recording
        try:
            with torch.no_grad():
                x_0 = sampler(x_T, n_levels_used=-1)
        finally:
            torch.randn_like = real
        cases[var_type] = {"cfg": cfg, "x_T": x_T, "noises": noises, "x_0": x_0}
    torch.save(cases, f"{OUT}/cifar_sampler.pt")


def pdearena_wmh_goldens():
    sys.modules["pytorch_wavelets"] = pw
    sys.path.insert(0, f"{REF}/pdearena")
    import pdearena.modules.twod_unet as ref_unet          # noqa: E402  (reference, read-only)
    import pdearena.modules.twod_unetbase as ref_base      # noqa: E402
    wmh = load_reference_module("ref_wmh_model", f"{REF}/wmh/model.py")

    # --- post-norm conv blocks (twod_unetbase.py:12-32, :148-161) and the pre-norm ResidualBlock (twod_unet.py:16-61)
    blocks = {}
    for tag, ctor in (("conv", lambda: ref_base.ConvBlock(32, 48)),
                      ("partial", lambda: ref_base.PartialResnetConvBlock(32, 48, activation="silu")),
                      ("full", lambda: ref_base.FullResnetConvBlock(32, 32)),
                      ("conv_nonorm", lambda: ref_base.ConvBlock(32, 32, norm=False)),
                      ("residual_sc", lambda: ref_unet.ResidualBlock(32, 64, norm=True, n_groups=8)),
                      ("residual_id", lambda: ref_unet.ResidualBlock(32, 32, norm=False))):
        blk = apply_det_init(ctor())
        torch.manual_seed(0)
        x = torch.randn(2, 32, 12, 10, requires_grad=True)
        y = blk(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        blocks[tag] = {"x": x.detach(), "y": y.detach(), "gy": gy, "gx": x.grad,
                       "gparams": {n: p.grad for n, p in blk.named_parameters()}}
    torch.save(blocks, f"{OUT}/pdearena_blocks.pt")

    # --- Unetbase_G: Multi-ResNet (Haar encoder, +1 extra ResNet layer), the residual U-Net arm, multi-res loss
    for tag, kw in (("multiresnet", dict(dwt_encoder=True, n_extra_resnet_layers=1)), ("unet", dict(dwt_encoder=False)),
                    ("multiresnet_mrl", dict(dwt_encoder=True, multi_res_loss=True))):
        cfg = dict(n_input_scalar_components=1, n_input_vector_components=1, n_output_scalar_components=1,
                   n_output_vector_components=1, time_history=2, time_future=1, hidden_channels=16, **kw)
        net = apply_det_init(ref_base.Unetbase_G(**cfg))
        torch.manual_seed(0)
        x = torch.randn(2, 2, 3, 32, 48)
        out = net(x)
        outs = out if isinstance(out, list) else [out]
        gys = [torch.randn_like(o) for o in outs]
        sum((o * g).sum() for o, g in zip(outs, gys)).backward()
        with torch.no_grad():
            out2 = net(x[..., ::4, ::4].contiguous(), n_levels_used=2) if kw.get("multi_res_loss") else None
        torch.save({"cfg": cfg, "x": x, "out": [o.detach() for o in outs], "gy": gys, "out_2lvl": out2,
                    "gparams": small_grads(net, also=("up.3.conv.conv1.weight",))},
                   f"{OUT}/pdearena_unetbase_g_{tag}.pt")

    # --- wmh Unetbase_G: odd extents 25 -> 13 with the decoder crop (Haar arm) / replicate pad (U-Net arm)
    for tag, kw in (("multiresnet", dict(dwt_encoder=True)), ("unet", dict(dwt_encoder=False))):
        cfg = dict(hidden_channels=16, **kw)
        net = apply_det_init(wmh.Unetbase_G(**cfg))
        torch.manual_seed(0)
        x = torch.randn(1, 2, 200, 200).to(torch.bfloat16).float()
        out = net(x)
        gy = torch.randn_like(out).to(torch.bfloat16).float()
        (out * gy).sum().backward()
        torch.save({"cfg": cfg, "x": x.to(torch.bfloat16), "out": out.detach().to(torch.bfloat16), "gy": gy.to(torch.bfloat16),
                    "gparams": small_grads(net)}, f"{OUT}/wmh_unetbase_g_{tag}.pt")


def mnist_goldens():
    """diff_mnist: guided-diffusion ResBlock with scale-shift norm (layers.py:250-338), AttentionBlock (:341-368),
    Upsample / Downsample (:195-247) and the UNet_wavelet container (mnist_diff/unet.py:75-556)."""
    sys.modules["pytorch_wavelets"] = pw
    load_reference_module("_mpl_stub_probe", f"{REF}/diff_cifar/model.py")     # installs the matplotlib stub
    sys.path.insert(0, f"{REF}/diff_mnist")
    import mnist_diff.unet as ref_unet                             # noqa: E402  (reference, read-only)
    import torch_ddpm.ddpm.models.unet.layers as ref_layers        # noqa: E402

    blocks = {}
    for tag, ctor, cin in (("res_scale_shift", lambda: ref_layers.ResBlock(128, 128, 0.0, out_channels=64, use_scale_shift_norm=True), 128),
                           ("res_plain_id", lambda: ref_layers.ResBlock(64, 128, 0.0, use_scale_shift_norm=False), 64),
                           ("attention", lambda: ref_layers.AttentionBlock(64, num_heads=4), 64),
                           ("upsample", lambda: ref_layers.Upsample(64, True), 64),
                           ("downsample_conv", lambda: ref_layers.Downsample(64, True), 64),
                           ("downsample_pool", lambda: ref_layers.Downsample(64, False), 64)):
        blk = apply_det_init(ctor())            # also replaces the zero-initialised convs, so every path carries signal
        torch.manual_seed(0)
        x = torch.randn(3, cin, 8, 8, requires_grad=True)
        emb = torch.randn(3, 128, requires_grad=True)
        y = blk(x, emb) if tag.startswith("res_") else blk(x)
        gy = torch.randn_like(y)
        y.backward(gy)
        blocks[tag] = {"x": x.detach(), "emb": emb.detach(), "y": y.detach(), "gy": gy, "gx": x.grad,
                       "gemb": emb.grad, "gparams": {n: p.grad for n, p in blk.named_parameters()}}
    torch.save(blocks, f"{OUT}/mnist_blocks.pt")

    for tag, kw in (("multiresnet", dict(dwt_encoder=True)), ("multiresnet_mrl", dict(dwt_encoder=True, multi_res_loss=True)),
                    ("unet", dict(dwt_encoder=False)), ("unet_pool", dict(dwt_encoder=False, avg_pool_down=True))):
        cfg = dict(image_size=32, image_channels=1, num_channels=32, dropout=0.0, num_res_blocks=1, **kw)
        net = apply_det_init(ref_unet.get_unet_wavelet(**cfg))
        torch.manual_seed(0)
        x = torch.randn(2, 1, 32, 32)
        t = torch.randint(30, (2, 1))
        out, _ = net(x, t)
        outs = out if isinstance(out, list) else [out]
        gys = [torch.randn_like(o) for o in outs]
        sum((o * g).sum() for o, g in zip(outs, gys)).backward()
        with torch.no_grad():
            out2, _ = net(x[..., ::4, ::4].contiguous(), t, n_levels_used=2)
        torch.save({"cfg": cfg, "x": x, "t": t, "out": [o.detach() for o in outs], "gy": gys,
                    "out_2lvl": out2 if isinstance(out2, list) else [out2],
                    "gparams": small_grads(net, also=("out_f_list.0.0.0.in_layers.2.weight",)),
                    "keys": {k: tuple(v.shape) for k, v in net.state_dict().items()}}, f"{OUT}/mnist_unet_wavelet_{tag}.pt")


def containers_r2_goldens():
    """Round-2 additions: the remaining container classes north_star names -- pdearena `Unetbase`
    (twod_unetbase.py:60-141), `twod_unet.Unet` ("Unetmod", twod_unet.py:389-548) and diff_mnist `UNetModel` through
    `get_unet` (torch_ddpm/ddpm/models/unet/unet.py:14-311, models/utils.py:5-53) -- plus a ReLU ConvBlock."""
    sys.modules["pytorch_wavelets"] = pw
    sys.path.insert(0, f"{REF}/pdearena")
    # twod_unet.py imports .fourier (torch.fft only) -- importable as is
    import pdearena.modules.twod_unet as ref_unet          # noqa: E402
    import pdearena.modules.twod_unetbase as ref_base      # noqa: E402
    common = dict(n_input_scalar_components=1, n_input_vector_components=1, n_output_scalar_components=1,
                  n_output_vector_components=1, time_history=2, time_future=1)
    for tag, ctor, cfg in (("unetbase", ref_base.Unetbase, dict(hidden_channels=16, activation="gelu", **common)),
                           ("unetbase_relu", ref_base.Unetbase, dict(hidden_channels=16, activation="relu", **common)),
                           ("unetmod", ref_unet.Unet, dict(hidden_channels=16, activation="gelu", norm=True, **common)),
                           ("unetmod_1x1_attn", ref_unet.Unet, dict(hidden_channels=16, activation="silu", norm=True,
                                                                    ch_mults=(1, 2), is_attn=(False, True), mid_attn=True,
                                                                    n_blocks=1, use1x1=True, **common))):
        net = apply_det_init(ctor(**cfg))
        torch.manual_seed(0)
        x = torch.randn(2, 2, 3, 32, 48) if "attn" not in tag else torch.randn(2, 2, 3, 16, 16)
        out = net(x)
        gy = torch.randn_like(out)
        (out * gy).sum().backward()
        torch.save({"cfg": cfg, "x": x, "out": [out.detach()], "gy": [gy], "out_2lvl": None,
                    "gparams": small_grads(net, also=("final.weight",)),
                    "keys": {k: tuple(v.shape) for k, v in net.state_dict().items()}}, f"{OUT}/pdearena_{tag}.pt")

    load_reference_module("_mpl_stub_probe", f"{REF}/diff_cifar/model.py")     # installs the matplotlib stub
    sys.path.insert(0, f"{REF}/diff_mnist")
    import torch_ddpm.ddpm.models.utils as ref_utils               # noqa: E402
    cfg = dict(image_size=32, image_channels=1, num_channels=32, dropout=0.0, num_res_blocks=1)
    net = apply_det_init(ref_utils.get_unet(**cfg))
    torch.manual_seed(0)
    x = torch.randn(2, 1, 32, 32)
    t = torch.randint(30, (2, 1))
    out = net(x, t)
    gy = torch.randn_like(out)
    (out * gy).sum().backward()
    with torch.no_grad():
        out2 = net(x[..., ::4, ::4].contiguous(), t, n_levels_used=2)
    torch.save({"cfg": cfg, "x": x, "t": t, "out": out.detach(), "gy": gy, "out_2lvl": out2,
                "gparams": small_grads(net, also=("output_blocks.0.0.in_layers.2.weight",)),
                "keys": {k: tuple(v.shape) for k, v in net.state_dict().items()}}, f"{OUT}/mnist_unetmodel.pt")


def state_dict_keys():
    """Key names and shapes of the reference's state_dicts (the checkpoint compatibility surface)."""
    m = load_reference_module("ref_cifar_model", f"{REF}/diff_cifar/model.py")
    sys.path.insert(0, f"{REF}/pdearena")
    import pdearena.modules.twod_unetbase as ref_base      # noqa: E402
    wmh = load_reference_module("ref_wmh_model", f"{REF}/wmh/model.py")
    out = {}
    for name, fixture, cls in (("cifar_multiresnet", "cifar_multiresnet.pt", m.UNetWaveletEnc),
                               ("cifar_unet", "cifar_unet.pt", m.UNetWaveletEnc),
                               ("pdearena_unetbase_g_multiresnet", "pdearena_unetbase_g_multiresnet.pt", ref_base.Unetbase_G),
                               ("pdearena_unetbase_g_unet", "pdearena_unetbase_g_unet.pt", ref_base.Unetbase_G),
                               ("wmh_unetbase_g_multiresnet", "wmh_unetbase_g_multiresnet.pt", wmh.Unetbase_G),
                               ("wmh_unetbase_g_unet", "wmh_unetbase_g_unet.pt", wmh.Unetbase_G)):
        cfg = torch.load(f"{OUT}/{fixture}", weights_only=False)["cfg"]
        out[name] = {"cfg": cfg, "keys": {k: tuple(v.shape) for k, v in cls(**cfg).state_dict().items()}}
    torch.save(out, f"{OUT}/state_dict_keys.pt")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "sampler":      # only the sampler fixture (the others are unchanged)
        cifar_sampler_golden()
        raise SystemExit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "r2":           # only the round-2 container fixtures
        containers_r2_goldens()
        raise SystemExit(0)
    cifar_goldens()
    cifar_sampler_golden()
    pdearena_wmh_goldens()
    mnist_goldens()
    containers_r2_goldens()
    state_dict_keys()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
