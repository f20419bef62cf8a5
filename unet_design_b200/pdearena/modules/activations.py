"""Activation names of pdearena/pdearena/modules/activations.py:3-9 mapped onto the fused GroupNorm+activation
kernel.  `gelu` (exact erf form), `silu` and `relu` are fused; `tanh` / `sigmoid` are not used by any config."""
ACTIVATION_REGISTRY = {"gelu": "gelu", "silu": "silu", "relu": "relu"}


def resolve(activation: str) -> str:
    act = ACTIVATION_REGISTRY.get(activation, None)
    if act is None:
        raise NotImplementedError(f"Activation {activation} not implemented by the B200 kernels (gelu, silu, relu)")
    return act
