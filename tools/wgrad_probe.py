"""Correctness + timing probe of conv_wgrad on the config shapes (one process per env setting: the switches are read once).
Env: UB200_WGRAD_HALO (0/1), UB200_WGRAD_BO (descriptor base-offset mode 0/1/2), UB200_WGRAD_PAIR (0/1)."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200._lib import ops as raw  # noqa: E402

shapes = [(4, 8, 8, 256, 256, 3), (2, 16, 16, 256, 512, 1), (3, 8, 8, 384, 256, 1), (8, 32, 32, 128, 128, 3), (4, 16, 16, 256, 256, 3), (2, 32, 32, 384, 128, 3), (1, 64, 64, 64, 128, 3),
          (1, 128, 128, 64, 64, 3), (1, 96, 192, 64, 64, 3), (3, 16, 16, 64, 320, 3), (2, 48, 48, 128, 128, 3)]
big = [(128, 32, 32, 256, 256, 3), (128, 32, 32, 128, 128, 3), (128, 32, 32, 384, 128, 3), (128, 32, 32, 256, 128, 3),
       (128, 16, 16, 256, 256, 3), (128, 16, 16, 512, 256, 3), (128, 8, 8, 512, 256, 3), (128, 4, 4, 512, 256, 3),
       (128, 32, 32, 384, 128, 1), (128, 16, 16, 256, 768, 1), (128, 8, 8, 256, 256, 3), (128, 4, 4, 256, 256, 3),
       (128, 16, 16, 256, 256, 1), (128, 4, 4, 512, 256, 1), (128, 8, 8, 512, 256, 1)]
if os.environ.get("PROBE_BIG"):
    big = [big[int(i)] for i in os.environ["PROBE_BIG"].split(",")]
    shapes = shapes[:4]
o = raw()
tag = " ".join(f"{k[6:]}={os.environ[k]}" for k in sorted(os.environ) if k.startswith("UB200_WGRAD"))
torch.backends.cudnn.allow_tf32 = False
worst = 0.0
for (n, h, w, cin, cout, k) in shapes:
    g = torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16)
    a = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
    dw = torch.zeros(cout, k, k, cin, device="cuda")
    o.conv_wgrad(g, a, k, dw)
    ar = a.float().permute(0, 3, 1, 2)
    ref = torch.nn.grad.conv2d_weight(ar, (cout, cin, k, k), g.float().permute(0, 3, 1, 2), padding=k // 2)
    err = float((dw.permute(0, 3, 1, 2) - ref).norm() / ref.norm())
    worst = max(worst, err)
    print(f"[{tag}] check {n}x{h}x{w} {cin}->{cout} k{k}: rel err {err:.2e} {'OK' if err < 1e-2 else 'FAIL'}", flush=True)
print(f"[{tag}] worst rel err {worst:.2e}", flush=True)
if (worst < 1e-2 or os.environ.get("UB200_WGRAD_DEBUG", "0") != "0") and os.environ.get("PROBE_TIME", "1") != "0":
    for (n, h, w, cin, cout, k) in big:
        # PROBE_COLD=1: rotate over operand sets that together exceed the 126 MB L2, so every launch reads DRAM-cold operands
        nset = 1
        if os.environ.get("PROBE_COLD", "0") == "1":
            nset = max(2, int(400e6 // (2.0 * n * h * w * (cin + cout))) + 1)
        gs = [torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16) for _ in range(nset)]
        as_ = [torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16) for _ in range(nset)]
        g, a = gs[0], as_[0]
        dw = torch.zeros(cout, k, k, cin, device="cuda")
        for _ in range(3):
            o.conv_wgrad(g, a, k, dw)
        reps = 20
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            for r in range(reps):
                o.conv_wgrad(gs[r % nset], as_[r % nset], k, dw)
        gr.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        fl = 2.0 * n * h * w * cout * k * k * cin
        print(f"[{tag}] time {h:3d}x{w:<3d} {cin:4d}->{cout:<4d} k{k}: {us:8.1f} us {fl / us / 1e6:8.1f} TFLOP/s", flush=True)
