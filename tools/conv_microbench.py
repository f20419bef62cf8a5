"""Micro-benchmark of the fprop kernel on one shape: back-to-back launches timed with CUDA events (L2-warm).
Env overrides (read once per process): UB200_FPROP_BN, UB200_FPROP_STAGES, UB200_FPROP_CLUSTER, UB200_FPROP_MSUB."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200 import ops  # noqa: E402
from unet_design_b200._lib import ops as raw  # noqa: E402

shapes = [(128, 4, 4, 256, 256, 3), (128, 4, 4, 512, 256, 3), (128, 4, 4, 256, 512, 3), (128, 4, 4, 256, 256, 1),
          (128, 8, 8, 256, 256, 3), (128, 8, 8, 512, 256, 3), (128, 8, 8, 256, 512, 1),
          (128, 16, 16, 256, 256, 3), (128, 16, 16, 256, 256, 1), (128, 16, 16, 256, 768, 1), (128, 16, 16, 768, 256, 1),
          (128, 32, 32, 256, 256, 3), (128, 32, 32, 128, 128, 3), (128, 32, 32, 256, 128, 3), (128, 32, 32, 128, 256, 1)]
if os.environ.get("SHAPES") == "small":
    shapes = shapes[:11]
elif os.environ.get("SHAPES"):
    shapes = [shapes[int(i)] for i in os.environ["SHAPES"].split(",")]
o = raw()
for (n, h, w, cin, cout, k) in shapes:
    a = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    wt = torch.randn(cout, cin, k, k, device="cuda").contiguous(memory_format=torch.channels_last)
    wp = ops.pack_weight(wt)
    out = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    for _ in range(5):
        o.conv_fprop(a, wp, k, cout, None, None, None, None, None, out, None, None)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    g = torch.cuda.CUDAGraph()                      # replayed graph: no host launch latency between the kernels
    with torch.cuda.graph(g):
        for _ in range(reps):
            o.conv_fprop(a, wp, k, cout, None, None, None, None, None, out, None, None)
    g.replay()
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / reps
    fl = 2.0 * n * h * w * cout * k * k * cin
    print(f"BN={os.environ.get('UB200_FPROP_BN','auto'):>4s} ST={os.environ.get('UB200_FPROP_STAGES','auto'):>4s} "
          f"shape {h:3d}x{w:<3d} {cin:4d}->{cout:<4d} k{k}: {us:8.1f} us  {fl / us / 1e6:8.1f} TFLOP/s  kblocks={k*k*cin//64}", flush=True)
