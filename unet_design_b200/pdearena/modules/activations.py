"""Activation names of pdearena/pdearena/modules/activations.py:3-9 mapped onto the fused GroupNorm+activation
kernel.  `gelu` (exact erf form) and `silu` are fused; the others are not used by the hot-path configs."""
ACTIVATION_REGISTRY = {"gelu": "gelu", "silu": "silu"}


def resolve(activation: str) -> str:
    act = ACTIVATION_REGISTRY.get(activation, None)
    if act is None:
        raise NotImplementedError(f"Activation {activation} not implemented by the B200 kernels (gelu, silu)")
    return act
