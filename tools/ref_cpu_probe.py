"""Times train steps of the UNMODIFIED reference classes (imported from /root/reference, read-only) on the CPU cores of the
BUILD container.  /root/reference does not exist on the GPU box, so these numbers cannot be re-measured beside the GPU
runs; bench.py quotes them (kind "reference", with this provenance) only where oracle/ has no restatement to time there
(config 1).  `pytorch_wavelets` resolves to oracle/pytorch_wavelets_restated (it is absent from the image).

    python tools/ref_cpu_probe.py        # writes profiles/r02_reference_cpu_probe.json
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_golden as mg  # noqa: E402  (reuses its loader of the reference modules)

REF = mg.REF


def timed(fn, steps=3, warmup=1):
    ts = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        fn()
        if i >= warmup:
            ts.append(time.perf_counter() - t0)
    return ts


def main():
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    out = {}
    # ---- config 1: diff_mnist Multi-ResNet, batch 64 (the reference's own CPU-runnable case)
    mg.load_reference_module("_mpl_stub_probe", f"{REF}/diff_cifar/model.py")
    sys.path.insert(0, f"{REF}/diff_mnist")
    import mnist_diff.unet as ref_unet
    torch.manual_seed(1234)
    net = ref_unet.get_unet_wavelet(32, 1, num_channels=32, dropout=0.0, num_res_blocks=2, dwt_encoder=True)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    n = 30
    betas = torch.linspace(0.1 / n, 20.0 / n, n)
    acp = torch.cumprod(1 - betas, 0)
    a, b = acp.sqrt(), (1 - acp).sqrt()
    x0 = torch.randn(64, 1, 32, 32)

    def step1():
        t = torch.randint(n, (64,))
        noise = torch.randn_like(x0)
        x_t = a[t].view(-1, 1, 1, 1) * x0 + b[t].view(-1, 1, 1, 1) * noise
        opt.zero_grad()
        o, _ = net(x_t, t.unsqueeze(-1))
        loss = torch.mean(torch.mean(torch.square(o - noise).reshape(64, -1), dim=-1))
        loss.backward()
        opt.step()

    ts = timed(step1, steps=5, warmup=2)
    out["c1"] = {"value": 64 * len(ts) / sum(ts), "unit": "images/s", "cores": cores,
                 "sample": f"{len(ts)} steps of batch 64, the reference's get_unet_wavelet(32, 1, 32, dwt_encoder=True) + Adam 1e-3, fp32 PyTorch CPU"}
    # ---- config 2 (for context next to the oracle port timed on the GPU box)
    m = mg.load_reference_module("ref_cifar_model", f"{REF}/diff_cifar/model.py")
    d = mg.load_reference_module("ref_cifar_diffusion", f"{REF}/diff_cifar/diffusion.py")
    torch.manual_seed(1234)
    net2 = m.UNetWaveletEnc(T=1000, ch=128, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1, dwt_encoder=True)
    tr = d.GaussianDiffusionTrainer(net2, 1e-4, 0.02, 1000, False, False, "cpu")
    opt2 = torch.optim.Adam(net2.parameters(), lr=2e-4)
    x2 = torch.rand(16, 3, 32, 32) * 2 - 1

    def step2():
        opt2.zero_grad()
        loss, _ = tr(x2, n_levels_used=-1)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net2.parameters(), 1.0)
        opt2.step()

    ts = timed(step2, steps=2, warmup=1)
    out["c2"] = {"value": 16 * len(ts) / sum(ts), "unit": "images/s", "cores": cores,
                 "sample": f"{len(ts)} steps of batch 16, the reference's UNetWaveletEnc (config 2) + Adam + clip, fp32 PyTorch CPU"}
    path = os.path.join(ROOT, "profiles", "r02_reference_cpu_probe.json")
    json.dump(out, open(path, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
