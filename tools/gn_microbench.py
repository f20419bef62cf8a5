"""GroupNorm+activation kernels alone: multi-pass (UB200_GN_FUSED=0) vs cluster-fused, per layer shape of config 2."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200 import _lib  # noqa: E402

SHAPES = [(128, 32, 32, 128), (128, 32, 32, 256), (128, 32, 32, 384), (128, 16, 16, 256), (128, 16, 16, 512), (128, 8, 8, 256),
          (128, 8, 8, 512), (128, 4, 4, 256), (128, 4, 4, 512)]


def timeit(fn, iters=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.zero_()
        torch.cuda._sleep(100000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters * 1e3


def main():
    ops = _lib.ops()
    out = {"fused": os.environ.get("UB200_GN_FUSED", "1"), "stream": os.environ.get("UB200_GN_STREAM", "policy"), "rows": []}
    for (n, h, w, c) in SHAPES:
        x = torch.randn(n, h, w, c, device="cuda").to(torch.bfloat16)
        gy = torch.randn_like(x)
        y, gx = torch.empty_like(x), torch.empty_like(x)
        gamma, beta = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
        stats = torch.empty(n, 32, 2, device="cuda")
        fwd = lambda: ops.gn_act_fwd(x, 32, stats, 1e-5, gamma, beta, None, None, 1, 0.0, 0, 0, None, None, y, True)
        bwd = lambda: ops.gn_act_bwd(gy, x, 32, stats, 1e-5, gamma, beta, None, None, 1, 0.0, 0, 0, None, gx, False, dg, db, None, None, None)
        tf, tb = timeit(fwd), timeit(bwd)
        mb = x.numel() * 2 / 1e6
        row = {"shape": [n, h, w, c], "fwd_us": round(tf, 1), "bwd_us": round(tb, 1), "fwd_GBs_2pass": round(2 * mb / tf * 1e3, 0),
               "bwd_GBs_3pass": round(3 * mb / tb * 1e3, 0)}
        print(row, flush=True)
        out["rows"].append(row)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open(f"gpurun_out/gn_microbench_fused{out['fused']}_stream{out['stream']}.json", "w"), indent=1)


if __name__ == "__main__":
    main()
