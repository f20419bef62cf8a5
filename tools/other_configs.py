"""Forward+backward timing of the other BASELINE configs through the drop-in modules (configs 1, 3, 4, 5 of BASELINE.json;
config 2 is bench.py).  Synthetic inputs of SURVEY.md §8(d) shapes; loss = MSE / Dice stand-in; torch.optim.Adam."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200.diff_mnist.unet import get_unet_wavelet  # noqa: E402
from unet_design_b200.pdearena.modules.twod_unetbase import Unetbase_G  # noqa: E402
from unet_design_b200.wmh.model import Unetbase_G as WmhUnet  # noqa: E402

dev = "cuda"


def run(name, model, make_batch, loss_fn, steps=8, warm=3):
    model = model.to(dev).train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    batch = make_batch()
    for i in range(warm + steps):
        if i == warm:
            torch.cuda.synchronize(); t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(model, batch)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    n = batch[0].shape[0]
    params = sum(p.numel() for p in model.parameters()) / 1e6
    print(f"{name:58s} {params:6.2f} M params  {dt * 1e3:8.2f} ms/step  {n / dt:9.1f} samples/s  loss {float(loss):.4f}", flush=True)


torch.manual_seed(0)
# C1: diff_mnist Multi-ResNet, 1x32x32, batch 64
run("C1 diff_mnist get_unet_wavelet(32,1,32,dwt_encoder=True) b64", get_unet_wavelet(32, 1, 32, dwt_encoder=True),
    lambda: (torch.randn(64, 1, 32, 32, device=dev), torch.randint(30, (64, 1), device=dev)),
    lambda m, b: ((m(b[0], b[1])[0] - b[0]) ** 2).mean())
# C3: pdearena Navier-Stokes, 8 x 4 x 3 x 128 x 128
for kw, tag in ((dict(dwt_encoder=True), "Multi-ResNet"), (dict(dwt_encoder=False), "U-Net"),
                (dict(dwt_encoder=True, n_extra_resnet_layers=3), "Multi-ResNet +3")):
    run(f"C3 pdearena NS Unetbase_G(hidden=64) {tag} b8", Unetbase_G(1, 1, 1, 1, 4, 1, 64, **kw),
        lambda: (torch.randn(8, 4, 3, 128, 128, device=dev), torch.randn(8, 1, 3, 128, 128, device=dev)),
        lambda m, b: ((m(b[0]) - b[1]) ** 2).mean())
# C4: pdearena shallow water, 16 x 2 x 3 x 96 x 192
run("C4 pdearena SW Unetbase_G(hidden=64) Multi-ResNet b16", Unetbase_G(1, 1, 1, 1, 2, 1, 64, dwt_encoder=True),
    lambda: (torch.randn(16, 2, 3, 96, 192, device=dev), torch.randn(16, 1, 3, 96, 192, device=dev)),
    lambda m, b: ((m(b[0]) - b[1]) ** 2).mean())
# C5: wmh segmentation, 32 x 2 x 200 x 200 (odd extents 25 -> 13)


def dice(m, b):
    p = m(b[0])
    inter = (p * b[1]).sum()
    return 1 - (2 * inter + 1) / (p.sum() + b[1].sum() + 1)


for kw, tag in ((dict(dwt_encoder=True), "Multi-ResNet"), (dict(dwt_encoder=False), "U-Net")):
    run(f"C5 wmh Unetbase_G(hidden=16) {tag} b32", WmhUnet(16, **kw),
        lambda: (torch.randn(32, 2, 200, 200, device=dev), (torch.rand(32, 1, 200, 200, device=dev) < 0.01).float()), dice)
