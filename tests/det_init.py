"""Deterministic parameter initialisation shared by tools/make_golden.py (which runs the REFERENCE classes) and
the parity tests (which build the oracle / drop-in classes): every parameter is a pure function of its name and
shape, rounded to a bf16-representable value, so golden fixtures do not have to carry state_dicts."""
import math
import zlib

import torch


def det_tensor(name: str, shape, kind: str) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    t = torch.randn(tuple(shape), generator=g)
    if kind == "weight" and len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        t = t * (1.0 / math.sqrt(fan_in))
    elif kind == "norm_weight":
        t = 1.0 + 0.2 * t
    else:                                   # biases, norm biases
        t = 0.1 * t
    return t.to(torch.bfloat16).float()


def apply_det_init(module: torch.nn.Module) -> torch.nn.Module:
    norm_names = set()
    for mname, m in module.named_modules():
        if isinstance(m, torch.nn.GroupNorm):
            norm_names.add(mname)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if not p.requires_grad:
                continue
            owner = name.rsplit(".", 1)[0] if "." in name else ""
            leaf = name.rsplit(".", 1)[-1]
            if owner in norm_names:
                kind = "norm_weight" if leaf == "weight" else "bias"
            else:
                kind = "weight" if leaf == "weight" else "bias"
            p.copy_(det_tensor(name, p.shape, kind))
    return module
