"""CPU checks of the drop-in modules' Python side (level bookkeeping, channel maps, odd-extent handling, autograd
glue, state_dict surface) against golden vectors of the reference's own classes, with the kernels replaced by the
test-only emulation of their contract (tests/_emulated_ops.py).  Tolerances are bf16-level: activations are bf16."""
import pytest
import torch

import golden_checks as gc

TOL = 2e-2


def test_state_dict_surface_matches_reference(emulated_ops):
    """Keys and shapes equal the reference's (recorded by tools/make_golden.py into tests/golden/state_dict_keys.pt)."""
    from unet_design_b200.diff_cifar.model import UNetWaveletEnc
    from unet_design_b200.pdearena.modules.twod_unetbase import Unetbase_G
    from unet_design_b200.wmh.model import Unetbase_G as WmhUnet
    ref = gc.load("state_dict_keys.pt")
    builders = {"cifar_multiresnet": lambda c: UNetWaveletEnc(**c), "cifar_unet": lambda c: UNetWaveletEnc(**c),
                "pdearena_unetbase_g_multiresnet": lambda c: Unetbase_G(**c), "pdearena_unetbase_g_unet": lambda c: Unetbase_G(**c),
                "wmh_unetbase_g_multiresnet": lambda c: WmhUnet(**c), "wmh_unetbase_g_unet": lambda c: WmhUnet(**c)}
    for name, build in builders.items():
        net = build(ref[name]["cfg"])
        ours = {k: tuple(v.shape) for k, v in net.state_dict().items()}
        assert ours == ref[name]["keys"], name
    # conv weights keep the [Cout, kh, kw, Cin] memory order through load_state_dict
    net = UNetWaveletEnc(**ref["cifar_multiresnet"]["cfg"])
    net.load_state_dict({k: torch.zeros(s) if "timembedding.0" not in k else net.state_dict()[k]
                         for k, s in ref["cifar_multiresnet"]["keys"].items()}, strict=True)
    assert net.upblocks[0][0].block1[2].weight.permute(0, 2, 3, 1).is_contiguous()


@pytest.mark.parametrize("tag", ["resblock_sc", "resblock_id", "resblock_attn"])
def test_cifar_resblock(emulated_ops, tag):
    from unet_design_b200.diff_cifar import model
    gc.check_cifar_resblock(model, tag, "cpu", TOL)


def test_cifar_upsample_and_dtwblock(emulated_ops):
    from unet_design_b200.diff_cifar import model
    gc.check_cifar_upsample(model, "cpu", TOL)
    gc.check_cifar_dtwblock(model, "cpu")


@pytest.mark.parametrize("tag", ["multiresnet", "unet"])
def test_cifar_model(emulated_ops, tag):
    from unet_design_b200.diff_cifar import model
    from unet_design_b200.diff_cifar.diffusion import GaussianDiffusionTrainer
    gc.check_cifar_model(model, GaussianDiffusionTrainer, tag, "cpu", 3 * TOL, 6e-2)


def test_pdearena_blocks(emulated_ops):
    from unet_design_b200.pdearena.modules import twod_unet, twod_unetbase
    gc.check_pdearena_blocks(twod_unetbase, twod_unet, "cpu", TOL)


@pytest.mark.parametrize("tag", ["multiresnet", "unet", "multiresnet_mrl"])
def test_pdearena_unetbase_g(emulated_ops, tag):
    from unet_design_b200.pdearena.modules.twod_unetbase import Unetbase_G
    gc.check_unetbase_g(Unetbase_G, f"pdearena_unetbase_g_{tag}.pt", "cpu", 3 * TOL, 8e-2)


@pytest.mark.parametrize("tag", ["multiresnet", "unet"])
def test_wmh_unetbase_g_odd_extents(emulated_ops, tag):
    from unet_design_b200.wmh.model import Unetbase_G
    gc.check_unetbase_g(Unetbase_G, f"wmh_unetbase_g_{tag}.pt", "cpu", 3 * TOL, 8e-2)


def test_mnist_blocks(emulated_ops):
    from unet_design_b200.diff_mnist import layers
    gc.check_mnist_blocks(layers, "cpu", TOL)


@pytest.mark.parametrize("tag", ["multiresnet", "multiresnet_mrl", "unet", "unet_pool"])
def test_mnist_unet_wavelet(emulated_ops, tag):
    from unet_design_b200.diff_mnist.unet import get_unet_wavelet
    gc.check_mnist_unet(get_unet_wavelet, tag, "cpu", 3 * TOL, 0.12)   # 4 levels deep through 1-channel bottlenecks


def test_cifar_sampler(emulated_ops):
    """Drop-in GaussianDiffusionSampler (DDPM Algorithm 2, diffusion.py:94-222) against the reference's sampler."""
    from unet_design_b200.diff_cifar import model
    from unet_design_b200.diff_cifar.diffusion import GaussianDiffusionSampler

    def make(net, T, vt):
        return GaussianDiffusionSampler(net, 1e-4, 0.02, T, img_size=16, mean_type="epsilon", var_type=vt)

    gc.check_cifar_sampler(model, make, "cpu", 3 * TOL)
    with pytest.raises(AssertionError):          # the reference's default mean_type is rejected by its own assert
        GaussianDiffusionSampler(None, 1e-4, 0.02, 4)
    # buffer surface of the reference class
    sd = make(torch.nn.Identity(), 6, "fixedlarge").state_dict()
    assert set(sd) == {"betas", "sqrt_recip_alphas_bar", "sqrt_recipm1_alphas_bar", "posterior_var",
                       "posterior_log_var_clipped", "posterior_mean_coef1", "posterior_mean_coef2"}
    assert all(v.dtype == torch.float64 for v in sd.values())


@pytest.mark.parametrize("tag", ["unetbase", "unetbase_relu", "unetmod", "unetmod_1x1_attn"])
def test_pdearena_unetbase_and_modern_unet(emulated_ops, tag):
    """The remaining pdearena containers north_star names: `Unetbase` (twod_unetbase.py:60-141) and `twod_unet.Unet`."""
    from unet_design_b200.pdearena.modules.twod_unet import Unet
    from unet_design_b200.pdearena.modules.twod_unetbase import Unetbase
    gc.check_unetbase_g(Unetbase if tag.startswith("unetbase") else Unet, f"pdearena_{tag}.pt", "cpu", 3 * TOL,
                        gc.CONTAINER_GRAD_TOL.get(tag, 8e-2))


def test_mnist_unetmodel_get_unet(emulated_ops):
    from unet_design_b200.diff_mnist.unet import get_unet
    gc.check_mnist_unetmodel(get_unet, "cpu", 3 * TOL, 0.12)


def test_mnist_precomputed_embedding_rows_are_one_shot(emulated_ops):
    """The batched `emb_layers` rows (diff_mnist.layers.precompute_emb_layers) are consumed by the forward that made them: nothing
    is left on the blocks afterwards, a second forward with other timesteps is not served stale rows, and a ResBlock called
    on its own still evaluates its own `emb_layers`."""
    import torch
    from unet_design_b200.diff_mnist.layers import ResBlock
    from unet_design_b200.diff_mnist.unet import get_unet_wavelet
    torch.manual_seed(0)
    model = get_unet_wavelet(32, 1, num_channels=32, num_res_blocks=1, dwt_encoder=True).eval()
    for p in model.parameters():                          # the zero-initialised output convs would hide the time embedding
        p.data.normal_(0.0, 0.05)
    x = torch.randn(2, 1, 32, 32)
    t1, t2 = torch.tensor([[3], [7]]), torch.tensor([[11], [2]])
    with torch.no_grad():
        a1 = model(x, t1)[0]
        assert not any("_emb_pre" in m.__dict__ for m in model.modules()), "precomputed rows left behind"
        a2 = model(x, t2)[0]
        a1_again = model(x, t1)[0]
    assert torch.equal(a1, a1_again)
    assert not torch.equal(a1, a2)
    blk = next(m for m in model.modules() if isinstance(m, ResBlock))
    blk.__dict__["_emb_pre"] = torch.zeros(1)             # a stale row must not survive a failed forward
    try:
        model(x[:, :, :3], t1)                              # bad extent: raises somewhere inside
    except Exception:
        pass
    model.__dict__.pop("_emb_blocks", None)
