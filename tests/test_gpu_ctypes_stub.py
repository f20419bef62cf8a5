"""Runs the reference-side ctypes stub of INTEGRATION.md section 2 (a `DTWBlock` that binds `libunet_b200.so` with raw
device pointers, no torch extension, no package import) on the GPU and checks it against the numpy oracle, plus direct
ctypes calls of the Haar and conv entry points the same way a maintainer of the reference would make them."""
import ctypes
import math
import os
import re

import numpy as np
import pytest
import torch

from oracle import haar_np

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unet_design_b200", "libunet_b200.so")


def _stub_source():
    """The python block of INTEGRATION.md section 2 that defines `class DTWBlock`, verbatim."""
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    src = next(b for b in blocks if "class DTWBlock" in b and "ctypes.CDLL" in b)
    return src.replace('ctypes.CDLL("libunet_b200.so")', f'ctypes.CDLL({LIB!r})')


def test_integration_md_dtwblock_stub_matches_oracle():
    ns = {}
    exec(compile(_stub_source(), "INTEGRATION.md#2", "exec"), ns)
    torch.manual_seed(0)
    for shape, J, out_ch in (((2, 3, 32, 32), 0, 128), ((2, 128, 32, 32), 1, 128), ((1, 5, 7, 9), 1, 12), ((1, 2, 25, 13), 3, 2),
                             ((4, 3, 32, 32), 2, 3)):
        x = torch.randn(*shape)
        blk = ns["DTWBlock"](J, out_ch)
        y = blk(x.cuda().contiguous())
        torch.cuda.synchronize()
        np.testing.assert_array_equal(y.cpu().numpy(), haar_np.dwtblock(x.numpy(), J, out_ch))


def test_direct_ctypes_haar_roundtrip_and_status_codes():
    lib = ctypes.CDLL(LIB)
    lib.ub200_status_string.restype = ctypes.c_char_p
    i64, p = ctypes.c_int64, ctypes.c_void_p
    torch.manual_seed(1)
    x = torch.randn(3, 4, 25, 13, device="cuda")
    planes, h, w = 12, 25, 13
    h2, w2 = math.ceil(h / 2), math.ceil(w / 2)
    ll = torch.empty(3, 4, h2, w2, device="cuda")
    highs = torch.empty(3, 4, 3, h2, w2, device="cuda")
    stream = p(torch.cuda.current_stream().cuda_stream)
    assert lib.ub200_haar_dwt2d_fwd(p(x.data_ptr()), i64(planes), i64(h), i64(w), p(ll.data_ptr()), p(highs.data_ptr()), stream) == 0
    rec = torch.empty_like(x)
    assert lib.ub200_haar_idwt2d(p(ll.data_ptr()), p(highs.data_ptr()), i64(planes), i64(h2), i64(w2), i64(h), i64(w),
                                 p(rec.data_ptr()), stream) == 0
    torch.cuda.synchronize()
    r_ll, r_lh, r_hl, r_hh = haar_np.dwt2_level(x.cpu().numpy())
    np.testing.assert_array_equal(ll.cpu().numpy(), r_ll)
    np.testing.assert_array_equal(highs.cpu().numpy(), np.stack([r_lh, r_hl, r_hh], 2))
    assert float((rec - x).abs().max()) < 1e-5
    # error behaviour: bad arguments come back as negative status codes with a message, never an exception / abort
    rc = lib.ub200_haar_idwt2d(p(ll.data_ptr()), None, i64(planes), i64(h2), i64(w2), i64(h + 5), i64(w), p(rec.data_ptr()), stream)
    assert rc < 0 and lib.ub200_status_string(rc)
    # fused 3-level analysis through the multi entry point (host array of device pointers)
    x8 = torch.randn(2, 3, 64, 64, device="cuda")
    outs = [torch.empty(2, 3, 3, 64 >> j, 64 >> j, device="cuda") for j in (1, 2, 3)]
    ll3 = torch.empty(2, 3, 8, 8, device="cuda")
    arr = (ctypes.c_void_p * 3)(*[o.data_ptr() for o in outs])
    assert lib.ub200_haar_dwt2d_multi_fwd(p(x8.data_ptr()), i64(6), i64(64), i64(64), ctypes.c_int(3), p(ll3.data_ptr()), arr, stream) == 0
    torch.cuda.synchronize()
    r_l, r_h = haar_np.dwt2(x8.cpu().numpy(), 3)
    np.testing.assert_array_equal(ll3.cpu().numpy(), r_l)
    for o, r in zip(outs, r_h):
        np.testing.assert_array_equal(o.cpu().numpy(), r)
