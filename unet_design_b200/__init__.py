"""unet_design_b200 -- B200-native (sm_100a) hot path of FabianFalck/unet-design.

Haar wavelet encoder / decoder, up/down-sampling and the GroupNorm + SiLU/GELU + 3x3-conv residual
blocks as hand-written CUDA (TMA + tcgen05/TMEM implicit GEMM) behind the reference's own module API:

    from unet_design_b200.diff_cifar.model import UNetWaveletEnc      # diff_cifar/model.py:326
    from unet_design_b200.diff_cifar.diffusion import GaussianDiffusionTrainer

There is no CPU path: modules raise if the CUDA extension is missing or the tensors are not on a GPU.
"""
__version__ = "0.1.0"
