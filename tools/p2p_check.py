"""Peer-memory gradient all-reduce (csrc/p2p.cu) against NCCL, under torchrun on the GPUs of one box:
results equal to NCCL's within fp32 summation-order noise, BITWISE identical on every rank, replayable from a CUDA graph,
and the device time of one whole-arena reduction.  Exit code 0 = all checks passed (rank 0 prints one JSON line)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unet_design_b200.train import PeerAllReduce  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    n = 26_560_400 if len(sys.argv) < 2 else int(sys.argv[1])        # ~ the config-2 gradient arena (106 MB)
    n = n // 4 * 4
    peer = PeerAllReduce(dev)
    g = peer.allocate(n)
    out = {"world": world, "floats": n}
    ok = True
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for trial, (lo, hi) in enumerate([(0, n), (4, 4 + 1024), (n // 3 // 4 * 4, n // 2 // 4 * 4), (n - 8, n)]):
        g.copy_(torch.randn(n, device=dev, generator=gen))
        ref = g.clone()
        dist.all_reduce(ref[lo:hi], op=dist.ReduceOp.SUM)
        torch.cuda.synchronize()
        dist.barrier()
        peer(lo, hi)
        torch.cuda.synchronize()
        dist.barrier()
        err = float((g - ref).abs().max())
        untouched = torch.equal(g[:lo], ref[:lo]) and torch.equal(g[hi:], ref[hi:])
        gathered = [torch.empty(hi - lo, device=dev) for _ in range(world)]
        dist.all_gather(gathered, g[lo:hi].contiguous())
        same = all(torch.equal(gathered[0], t) for t in gathered[1:])
        out[f"trial{trial}"] = {"range": [lo, hi], "max_abs_err_vs_nccl": err, "bitwise_equal_across_ranks": same, "outside_untouched": untouched}
        ok = ok and err < 1e-4 and same and untouched
    # CUDA graph: 7 buckets on a side stream, replayed; the inputs are re-randomised before every replay
    src = torch.randn(n, device=dev, generator=gen)
    bounds = [i * (n // 7) // 4 * 4 for i in range(7)] + [n]
    side = torch.cuda.Stream(device=dev)
    graph = torch.cuda.CUDAGraph()
    g.copy_(src)
    torch.cuda.synchronize()
    with torch.cuda.graph(graph, capture_error_mode="thread_local"):
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for lo, hi in zip(bounds[:-1], bounds[1:]):
                peer(lo, hi)
        torch.cuda.current_stream().wait_stream(side)
    want = src.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    worst = 0.0
    for _ in range(5):
        g.copy_(src)
        torch.cuda.synchronize()
        dist.barrier()
        graph.replay()
        torch.cuda.synchronize()
        worst = max(worst, float((g - want).abs().max()))
    out["graph_replay_max_abs_err"] = worst
    ok = ok and worst < 1e-4
    # device time of a whole-arena reduction, ours and NCCL's (max over ranks)
    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)
    out["p2p_ms_whole_arena"] = round(timed(lambda: peer(0, n)), 4)
    out["nccl_ms_whole_arena"] = round(timed(lambda: dist.all_reduce(g, op=dist.ReduceOp.SUM)), 4)
    out["p2p_ms_16MB_bucket"] = round(timed(lambda: peer(0, 4 << 20)), 4)
    out["nccl_ms_16MB_bucket"] = round(timed(lambda: dist.all_reduce(g[:4 << 20], op=dist.ReduceOp.SUM)), 4)
    okt = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    out["ok"] = bool(float(okt) == 1.0)
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
