"""DWT/iDWT size sweep of BASELINE config 5 / SURVEY.md §8(d): achieved GB/s (algorithmic bytes / CUDA-event time) for
the full 4-band DWT, iDWT, LL-only and DTWBlock kernels, fp32 NCHW.  `--one` runs a single 1 GiB DWT+iDWT+DTWBlock
(the ncu target)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200._lib import ops as raw  # noqa: E402

o = raw()
SHAPES = [(128, 3, 32, 32), (128, 128, 32, 32), (8, 64, 128, 128), (16, 64, 96, 192), (32, 16, 200, 200), (32, 128, 25, 25),
          (64, 64, 256, 256), (16, 64, 1024, 1024)]
if "--one" in sys.argv:
    SHAPES = [(64, 64, 256, 256)]


def timeit(fn, reps):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


rows = []
for shp in SHAPES:
    n, c, h, w = shp
    x = torch.randn(*shp, device="cuda")
    E = x.numel() * 4
    reps = 3 if "--one" in sys.argv else (10 if E > (1 << 28) else 50)
    ll, hi = o.haar_dwt2d_fwd(x, True)
    h2, w2 = ll.shape[-2:]
    cases = {
        "dwt_4band": (lambda: o.haar_dwt2d_fwd(x, True), E + 4 * ll.numel() * 4),
        "dwt_ll_only": (lambda: o.haar_dwt2d_fwd(x, False), E + ll.numel() * 4),
        "idwt_4band": (lambda: o.haar_idwt2d(ll, hi, h, w), 4 * ll.numel() * 4 + E),
        "dtwblock_J1_tile2": (lambda: o.dwtblock_fwd(x, 1, 2 * c), E + 2 * ll.numel() * 4),
        "dtwblock_J1_bwd": (lambda: o.dwtblock_bwd(ll, c, h, w, 1), ll.numel() * 4 + E),
    }
    for name, (fn, nbytes) in cases.items():
        ms = timeit(fn, reps)
        rows.append({"shape": shp, "case": name, "ms": ms, "algorithmic_bytes": nbytes, "GBps": nbytes / ms / 1e6,
                     "l2_resident": E < 100 * (1 << 20)})
        print(f"{str(shp):24s} {name:20s} {ms * 1e3:9.1f} us {nbytes / ms / 1e6:8.0f} GB/s" + ("  (fits L2)" if E < 100 * (1 << 20) else ""), flush=True)
    del x, ll, hi
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/haar_sweep.json", "w"), indent=0)
