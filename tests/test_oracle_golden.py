"""The torch fp32 oracle (oracle/torch_ref.py) replays the golden vectors that
tools/make_golden.py produced from the reference's own classes (diff_cifar/model.py,
diff_cifar/diffusion.py).  This is what pins the conv-block oracle."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import haar_np, torch_ref

TOL = 2e-6  # same torch ops in the same order: round-off only


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("tag", ["resblock_sc", "resblock_id", "resblock_attn"])
def test_resblock_matches_reference(golden_dir, tag):
    g = _load(golden_dir, f"cifar_{tag}.pt")
    blk = torch_ref.ResBlock(**g["cfg"])
    blk.load_state_dict(g["state"])
    x = g["x"].clone().requires_grad_(True)
    temb = g["temb"].clone().requires_grad_(True)
    y = blk(x, temb)
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < TOL
    assert rel_err(x.grad, g["gx"]) < TOL
    assert rel_err(temb.grad, g["gtemb"]) < TOL
    for n, p in blk.named_parameters():
        assert rel_err(p.grad, g["gparams"][n]) < 1e-5, n


def test_upsample_matches_reference(golden_dir):
    g = _load(golden_dir, "cifar_upsample.pt")
    up = torch_ref.UpSample(32)
    up.load_state_dict(g["state"])
    x = g["x"].clone().requires_grad_(True)
    y = up(x, None)
    y.backward(g["gy"])
    assert rel_err(y, g["y"]) < TOL and rel_err(x.grad, g["gx"]) < TOL


def test_dtwblock_matches_reference(golden_dir):
    for case in _load(golden_dir, "cifar_dtwblock.pt"):
        y = torch_ref.DTWBlock(case["J"], case["out_channels"])(case["x"])
        assert y.shape == case["y"].shape
        assert rel_err(y, case["y"]) < TOL
        y_np = haar_np.dwtblock(case["x"].numpy(), case["J"], case["out_channels"])
        np.testing.assert_allclose(y_np, case["y"].numpy(), atol=2e-6)


@pytest.mark.parametrize("tag", ["multiresnet", "unet"])
def test_model_and_loss_match_reference(golden_dir, tag):
    g = _load(golden_dir, f"cifar_{tag}.pt")
    net = torch_ref.UNetWaveletEnc(**g["cfg"])
    missing = net.load_state_dict(g["state"], strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    trainer = torch_ref.GaussianDiffusionTrainer(net, 1e-4, 0.02, g["cfg"]["T"], g["cfg"]["multi_res_loss"])
    assert rel_err(trainer.q_sample(g["x0"], g["t"], g["noise"]), g["x_t"]) < TOL
    loss, loss_list = trainer.loss_from(g["x0"], g["t"], g["noise"])
    loss.backward()
    assert abs(float(loss.detach()) - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    for a, b in zip(loss_list, g["loss_list"]):
        assert abs(float(a) - float(b)) < 1e-5 * abs(float(b))
    params = dict(net.named_parameters())
    for n, gr in g["gparams"].items():
        assert rel_err(params[n].grad, gr) < 1e-4, n
    with torch.no_grad():
        out = net(g["x_t"], g["t"])
        out1 = net(g["x_t"][:, :, ::2, ::2].contiguous(), g["t"], n_levels_used=1)
    outs = out if isinstance(out, list) else [out]
    gouts = g["out"] if isinstance(g["out"], list) else [g["out"]]
    for a, b in zip(outs, gouts):
        assert rel_err(a, b) < TOL
    a1 = out1[-1] if isinstance(out1, list) else out1
    b1 = g["out_1lvl"][-1] if isinstance(g["out_1lvl"], list) else g["out_1lvl"]
    assert rel_err(a1, b1) < TOL
