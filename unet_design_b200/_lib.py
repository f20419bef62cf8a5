"""Loader of the in-tree sm_100a libraries.  There is NO fallback: if the CUDA extension is missing
or cannot be loaded, importing the ops raises, loudly."""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
CAPI_PATH = os.path.join(_HERE, "libunet_b200.so")
TORCH_LIB_PATH = os.path.join(_HERE, "libunet_b200_torch.so")

_capi = None
_ops = None


class ExtensionMissing(RuntimeError):
    pass


def capi() -> ctypes.CDLL:
    """The C-ABI shared library (include/unet_b200.h) as a ctypes handle."""
    global _capi
    if _capi is None:
        if not os.path.exists(CAPI_PATH):
            raise ExtensionMissing(
                f"{CAPI_PATH} is missing: build it with `python -m unet_design_b200.csrc.build` "
                "(or __graft_entry__.build()).  unet_design_b200 has no CPU / eager fallback.")
        _capi = ctypes.CDLL(CAPI_PATH, mode=ctypes.RTLD_GLOBAL)
        _capi.ub200_version.restype = ctypes.c_char_p
        _capi.ub200_status_string.restype = ctypes.c_char_p
    return _capi


def ops():
    """`torch.ops.unet_b200` (the PyTorch C++ extension over the C ABI)."""
    global _ops
    if _ops is None:
        capi()
        if not os.path.exists(TORCH_LIB_PATH):
            raise ExtensionMissing(
                f"{TORCH_LIB_PATH} is missing: build it with `python -m unet_design_b200.csrc.build`.")
        torch.ops.load_library(TORCH_LIB_PATH)
        _ops = torch.ops.unet_b200
    return _ops
