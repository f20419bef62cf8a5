"""1k-step loss curves, reference algorithm vs the B200 path (north_star: "loss curves matched over 1k steps").

Both arms run the training loop of diff_cifar/main.py:397-429 (DDPM Algorithm-1 loss, clip_grad_norm_ 1.0, Adam 2e-4,
LambdaLR warm-up) on the BASELINE config-2 architecture from the SAME initial weights, the SAME synthetic data stream and
the SAME (t, noise) draws:

  ref   oracle/torch_ref.py on the GPU in fp32 (TF32 off) + torch.optim.Adam + LambdaLR + clip_grad_norm_
  b200  unet_design_b200 (bf16 tcgen05 kernels, Philox dropout, fused clip + Adam + EMA tail, one CUDA graph per step)

dropout = 0 makes the comparison deterministic up to bf16 rounding (value check: curves must track); dropout = 0.1 uses
different RNG streams (torch's vs Philox), so that run is a statistical check of the smoothed curves.

    python tools/loss_curve.py --steps 1000 --batch 128 --dropout 0.0 --out profiles/r02_loss_curve_dropout0.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CFG = dict(T=1000, ch=128, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dwt_encoder=True)


def smooth(xs, w):
    out, acc = [], 0.0
    for i, v in enumerate(xs):
        acc += v
        if i >= w:
            acc -= xs[i - w]
        out.append(acc / min(i + 1, w))
    return out


def run_curves(steps=1000, batch=128, dropout=0.0, seed=0, cfg=None, lr=2e-4, warmup=5000, use_graph=True, img=32):
    from oracle import torch_ref
    from unet_design_b200 import ops
    from unet_design_b200.diff_cifar.diffusion import GaussianDiffusionTrainer
    from unet_design_b200.diff_cifar.model import UNetWaveletEnc
    from unet_design_b200.train import TrainStep

    cfg = dict(cfg or CFG, dropout=dropout)
    dev = torch.device("cuda")
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(1234)
    ref = torch_ref.UNetWaveletEnc(**cfg).to(dev)                     # the reference's own initialisers
    net = UNetWaveletEnc(**cfg).to(dev)
    net.load_state_dict(ref.state_dict())
    ref.train(); net.train()
    T = cfg["T"]
    ref_trainer = torch_ref.GaussianDiffusionTrainer(ref, 1e-4, 0.02, T).to(dev)
    trainer = GaussianDiffusionTrainer(net, 1e-4, 0.02, T).to(dev)
    opt = torch.optim.Adam([p for p in ref.parameters() if p.requires_grad], lr=lr)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(s, warmup) / warmup)
    ops.seed_dropout(seed + 77)
    step = TrainStep(net, lambda x0, t, noise: trainer.loss_from(x0, t, noise)[0], lr=lr, warmup=warmup, grad_clip=1.0,
                     ema_decay=0.9999, use_cuda_graph=use_graph)
    gen = torch.Generator(device=dev).manual_seed(seed)
    losses = {"ref": [], "b200": []}
    ours = []
    t0 = time.time()
    for i in range(steps):
        x0 = torch.rand(batch, 3, img, img, device=dev, generator=gen) * 2 - 1
        t = torch.randint(T, (batch,), device=dev, generator=gen)
        noise = torch.randn(batch, 3, img, img, device=dev, generator=gen)
        opt.zero_grad(set_to_none=True)
        lref, _ = ref_trainer.loss_from(x0, t, noise)
        lref.backward()
        torch.nn.utils.clip_grad_norm_(ref.parameters(), 1.0)
        opt.step(); sched.step()
        losses["ref"].append(lref.detach())
        ours.append(step(x0, t, noise).clone())
    torch.cuda.synchronize()
    losses["ref"] = [float(v) for v in losses["ref"]]
    losses["b200"] = [float(v) for v in ours]
    return {"config": cfg, "steps": steps, "batch": batch, "dropout": dropout, "lr": lr, "warmup": warmup,
            "wall_s": time.time() - t0, "losses": losses}


def compare(res, window=50, marks=(100, 500, 1000)):
    a, b = smooth(res["losses"]["ref"], window), smooth(res["losses"]["b200"], window)
    rows = []
    for m in marks:
        if m <= len(a):
            rows.append({"step": m, "ref": a[m - 1], "b200": b[m - 1], "rel_diff": abs(a[m - 1] - b[m - 1]) / abs(a[m - 1])})
    first = min(20, len(a))
    early = max(abs(x - y) / abs(x) for x, y in zip(res["losses"]["ref"][:first], res["losses"]["b200"][:first]))
    return {"window": window, "marks": rows, "max_rel_diff_first_20_raw_steps": early,
            "final_ref": a[-1], "final_b200": b[-1]}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--batch", type=int, default=128)
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    res = run_curves(args.steps, args.batch, args.dropout)
    res["summary"] = compare(res)
    print(json.dumps(res["summary"], indent=1))
    if args.out:
        os.makedirs(os.path.dirname(os.path.join(ROOT, args.out)) or ".", exist_ok=True)
        with open(os.path.join(ROOT, args.out), "w") as f:
            json.dump(res, f)
