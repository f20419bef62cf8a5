"""DDPM sampler (Algorithm 2) step latency on config 2: CUDA-graph loop (time index on the device) vs eager launches."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200.diff_cifar.diffusion import GaussianDiffusionSampler  # noqa: E402
from unet_design_b200.diff_cifar.model import UNetWaveletEnc  # noqa: E402


def main():
    torch.manual_seed(0)
    T = int(os.environ.get("T", "100"))
    batch = int(os.environ.get("BATCH", "128"))
    net = UNetWaveletEnc(T=1000, ch=128, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1, dwt_encoder=True).cuda().eval()
    out = {"T_steps_timed": T, "batch": batch}
    for mode in ("graph", "eager"):
        s = GaussianDiffusionSampler(net, 1e-4, 0.02, T, img_size=32, mean_type="epsilon", var_type="fixedlarge").cuda().eval()
        s.use_cuda_graph = mode == "graph"
        x_T = torch.randn(batch, 3, 32, 32, device="cuda")
        s(x_T, -1)                                   # warm-up (and graph capture)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x0 = s(x_T, -1)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        assert torch.isfinite(x0).all()
        out[mode] = {"ms_per_reverse_step": 1e3 * dt / T, "images_per_s_at_1000_steps": batch / (dt / T * 1000)}
    print(json.dumps(out))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/sampler_bench.json", "w"), indent=1)


if __name__ == "__main__":
    main()
