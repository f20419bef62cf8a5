"""The torch fp32 oracle (oracle/torch_ref.py) replays the golden vectors that tools/make_golden.py produced from
the reference's own classes (diff_cifar/model.py, diff_cifar/diffusion.py).  This is what pins the conv-block
oracle; tolerance = round-off of the same torch ops in the same order."""
import numpy as np
import pytest
import torch

import golden_checks as gc
from oracle import haar_np, torch_ref

TOL = 2e-6


@pytest.mark.parametrize("tag", ["resblock_sc", "resblock_id", "resblock_attn"])
def test_resblock_matches_reference(tag):
    gc.check_cifar_resblock(torch_ref, tag, "cpu", TOL)


def test_upsample_matches_reference():
    gc.check_cifar_upsample(torch_ref, "cpu", TOL)


def test_dtwblock_matches_reference():
    gc.check_cifar_dtwblock(torch_ref, "cpu")
    for case in gc.load("cifar_dtwblock.pt"):
        y_np = haar_np.dwtblock(case["x"].numpy(), case["J"], case["out_channels"])
        np.testing.assert_allclose(y_np, case["y"].numpy(), atol=2e-6)


@pytest.mark.parametrize("tag", ["multiresnet", "unet"])
def test_model_and_loss_match_reference(tag):
    gc.check_cifar_model(torch_ref, torch_ref.GaussianDiffusionTrainer, tag, "cpu", 5e-6, 2e-4)
    g = gc.load(f"cifar_{tag}.pt")
    net = torch_ref.UNetWaveletEnc(**g["cfg"])
    trainer = torch_ref.GaussianDiffusionTrainer(net, 1e-4, 0.02, g["cfg"]["T"], g["cfg"]["multi_res_loss"])
    assert float((trainer.q_sample(g["x0"], g["t"], g["noise"]) - g["x_t"]).abs().max()) < 1e-6


def test_cifar_sampler_oracle_matches_reference():
    gc.check_cifar_sampler(torch_ref, lambda net, T, vt: torch_ref.GaussianDiffusionSampler(net, 1e-4, 0.02, T, "epsilon", vt),
                           "cpu", 1e-5)


@pytest.mark.parametrize("fixture,wmh", [("pdearena_unetbase_g_multiresnet.pt", False), ("pdearena_unetbase_g_unet.pt", False),
                                         ("pdearena_unetbase_g_multiresnet_mrl.pt", False),
                                         ("wmh_unetbase_g_multiresnet.pt", True), ("wmh_unetbase_g_unet.pt", True)])
def test_pde_family_oracle_matches_reference(fixture, wmh):
    """oracle/torch_ref_pde.py (pdearena / wmh `Unetbase_G`) against goldens recorded from the reference's own classes;
    the wmh fixtures store bf16-rounded tensors, hence their looser bound."""
    from oracle import torch_ref_pde
    tol = 1e-2 if wmh else 5e-6
    gc.check_unetbase_g(lambda **cfg: torch_ref_pde.from_reference_cfg(cfg, wmh=wmh), fixture, "cpu", tol, 50 * tol if not wmh else 3e-2)
