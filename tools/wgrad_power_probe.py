"""Is the wgrad microbench rate a power / clock limit?  Replays the 256->256 @ 32x32 wgrad back to back for ~3 s per setting while
nvidia-smi samples SM clock, power and throttle reasons; prints the rate of the last second and the samples."""
import os
import subprocess
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200._lib import ops as raw  # noqa: E402

o = raw()
n, h, w, cin, cout, k = 128, 32, 32, 256, 256, 3
g = torch.randn(n, h, w, cout, device="cuda").to(torch.bfloat16)
a = torch.randn(n, h, w, cin, device="cuda").to(torch.bfloat16)
dw = torch.zeros(cout, k, k, cin, device="cuda")
for _ in range(3):
    o.conv_wgrad(g, a, k, dw)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr):
    for _ in range(50):
        o.conv_wgrad(g, a, k, dw)
torch.cuda.synchronize()
smi = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader",
                        "-lms", "200"], stdout=subprocess.PIPE, text=True)
t_end = time.time() + 3.0
rates = []
while time.time() < t_end:
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    us = 1e3 * e0.elapsed_time(e1) / 500
    rates.append(2.0 * n * h * w * cout * k * k * cin / us / 1e6)
smi.terminate()
out = smi.stdout.read().strip().splitlines()
tag = " ".join(f"{k_[6:]}={os.environ[k_]}" for k_ in sorted(os.environ) if k_.startswith("UB200_WGRAD"))
print(f"[{tag}] TFLOP/s first {rates[0]:.0f} .. last {rates[-1]:.0f} (min {min(rates):.0f}, max {max(rates):.0f}, {len(rates)} windows)")
print("   smi:", " | ".join(out[::2][:10]))
