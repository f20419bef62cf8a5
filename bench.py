#!/usr/bin/env python
"""bench.py -- Multi-ResNet (Haar encoder) DDPM train throughput on B200, BASELINE.json's headline metric.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU algorithm (oracle port)

Workload (N = 1 and every N: weak scaling): BASELINE configs[1] = diff_cifar Multi-ResNet DDPM U-Net,
`UNetWaveletEnc(T=1000, ch=128, ch_mult=[1,2,2,2], attn=[1], num_res_blocks=2, dropout=0.1, dwt_encoder=True)`,
synthetic 3x32x32, batch 128 per GPU, bf16 compute / fp32 master weights, Adam 2e-4 + warm-up + clip 1.0 +
EMA 0.9999 (diff_cifar/hyperparams.py:36-55).  A "step" = noise draw, q-sample, forward, MSE, backward,
(all-reduce,) clip + Adam + EMA.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

CFG = dict(T=1000, ch=128, ch_mult=[1, 2, 2, 2], attn=[1], num_res_blocks=2, dropout=0.1, dwt_encoder=True)
BATCH_PER_GPU = 128
IMG = (3, 32, 32)
METRIC = "multi_resnet_ddpm_train_images_per_sec"
WORKLOAD = "diff_cifar Multi-ResNet (Haar encoder) DDPM train step, synthetic 3x32x32, batch 128 per GPU, bf16"


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's algorithm (oracle/torch_ref.py, pinned against the reference's own classes by
# tests/golden) on the host cores.  Used for `cpu_baseline` (bounded sample) and for --impl reference.
# ------------------------------------------------------------------------------------------------
def cpu_train_steps(batch: int, steps: int, warmup: int):
    from oracle import torch_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    net = torch_ref.UNetWaveletEnc(**CFG)
    trainer = torch_ref.GaussianDiffusionTrainer(net, 1e-4, 0.02, CFG["T"])
    opt = torch.optim.Adam([p for p in net.parameters() if p.requires_grad], lr=2e-4)
    ema = [p.detach().clone() for p in net.parameters() if p.requires_grad]
    x0 = torch.rand(batch, *IMG) * 2 - 1
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss, _ = trainer(x0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()
        with torch.no_grad():
            for e, p in zip(ema, (q for q in net.parameters() if q.requires_grad)):
                e.mul_(0.9999).add_(p, alpha=1 - 0.9999)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, cores


def reference_steps(name: str, batch: int, steps: int, warmup: int, dev="cpu", autocast=False, channels_last=False):
    """Timed train steps of the reference algorithm (oracle/ restatement) for a non-headline config: CPU baseline, or
    the GPU library (cuDNN) comparator when `dev` is a CUDA device.  Returns (seconds per step list, cores)."""
    import bench_configs as bc
    built = bc.build_reference(name, dev)
    if built is None:
        return None, os.cpu_count() or 1
    model, opt, loss_fn = built
    cores = os.cpu_count() or 1
    if str(dev) == "cpu":
        torch.set_num_threads(cores)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    torch.manual_seed(0)
    shapes = {"c3": ((batch, 4, 3, 128, 128), (batch, 1, 3, 128, 128)), "c3u": ((batch, 4, 3, 128, 128), (batch, 1, 3, 128, 128)),
              "c4": ((batch, 2, 3, 96, 192), (batch, 1, 3, 96, 192)), "c5": ((batch, 2, 200, 200), (batch, 1, 200, 200))}[name]
    x = torch.randn(*shapes[0], device=dev)
    y = torch.randn(*shapes[1], device=dev) if name != "c5" else (torch.rand(*shapes[1], device=dev) < 0.01).float()
    times = []
    for i in range(warmup + steps):
        if str(dev) != "cpu":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            loss = loss_fn(x, y)
        loss.backward()
        opt.step()
        if str(dev) != "cpu":
            torch.cuda.synchronize()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return times, cores


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == "c2":
        batch, unit, workload, metric = BATCH_PER_GPU, "images/s", WORKLOAD, METRIC
        times, cores = cpu_train_steps(batch, args.steps, args.warmup)
        kind, src = "port", "oracle/torch_ref.py"
    else:
        import bench_configs as bc
        workload, batch, unit, _ = bc.CONFIGS[args.config]
        metric = f"{args.config}_train_{unit.replace('/', '_per_')}"
        times, cores = reference_steps(args.config, batch, args.steps, args.warmup)
        kind, src = "port", "oracle/torch_ref_pde.py"
        if times is None:
            print(json.dumps({"impl": "reference", "unavailable": f"no fp32 restatement of config {args.config} in oracle/ "
                              "(its CPU number is quoted from profiles/r02_reference_cpu_probe.json)"}), flush=True)
            return
    total = sum(times)
    value = batch * len(times) / total
    sample = f"{len(times)} steps of the full per-GPU batch {batch}, fp32, PyTorch CPU ({src}), {cores} threads"
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload, "global_batch": batch, "note": "reference algorithm on host cores via the pinned oracle "
                   "restatement (the reference is pure Python and /root/reference does not travel to the GPU box); one process, "
                   "whatever --gpus says"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


def gpu_library_baseline(dev, steps=6, warmup=3):
    """The secondary comparator SURVEY.md 8(d) names: the reference algorithm (oracle/torch_ref.py) on the SAME GPU through
    PyTorch's library path -- cuDNN / cuBLAS / flash-attention, channels_last + bf16 autocast, and plain fp32 (TF32 on, the
    PyTorch default for cuDNN) -- same step definition (loss, backward, clip, Adam, EMA), eager launches."""
    from oracle import torch_ref
    out = {}
    for tag, autocast, cl in (("bf16_autocast_channels_last", True, True), ("fp32_tf32_convs", False, False)):
        torch.manual_seed(1234)
        net = torch_ref.UNetWaveletEnc(**CFG).to(dev)
        if cl:
            net = net.to(memory_format=torch.channels_last)
        trainer = torch_ref.GaussianDiffusionTrainer(net, 1e-4, 0.02, CFG["T"]).to(dev)
        params = [p for p in net.parameters() if p.requires_grad]
        opt = torch.optim.Adam(params, lr=2e-4, fused=True)
        ema = [p.detach().clone() for p in params]
        x0 = torch.rand(BATCH_PER_GPU, *IMG, device=dev) * 2 - 1
        ev = []
        for i in range(warmup + steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            opt.zero_grad(set_to_none=True)
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                loss, _ = trainer(x0)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 1.0)
            opt.step()
            torch._foreach_mul_(ema, 0.9999)
            torch._foreach_add_(ema, [p.detach() for p in params], alpha=1e-4)
            e1.record()
            if i >= warmup:
                ev.append((e0, e1))
        torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
        out[tag] = {"ms_per_step": ms, "images_per_s": BATCH_PER_GPU / (ms * 1e-3)}
        del net, trainer, opt, ema, params
        torch.cuda.empty_cache()
    out["what"] = ("oracle/torch_ref.py (the reference's model, pinned by goldens) on this GPU via cuDNN/cuBLAS/SDPA, eager, "
                   f"batch {BATCH_PER_GPU}, {steps} timed steps; not the product path")
    return out


# ------------------------------------------------------------------------------------------------
# per-kernel timing proxy (CUDA events around every conv launch of one eager step)
# ------------------------------------------------------------------------------------------------
class KernelTimer:
    def __init__(self, real):
        self._real, self.records = real, []

    def __getattr__(self, name):
        fn = getattr(self._real, name)
        if name not in ("conv_fprop", "conv_wgrad"):
            return fn

        def timed(*a):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(200000)     # keep the GPU busy ~100 us so the launch below is queued before e0 fires:
            e0.record()                   # the interval is then kernel execution, not host launch latency
            out = fn(*a)
            e1.record()
            if name == "conv_fprop":
                act, _, k, cout, a2 = a[0], a[1], a[2], a[3], a[4]
                n, h, w, cin = act.shape
                flops = 2.0 * n * h * w * cout * (k * k * cin + (a2.shape[3] if a2 is not None else 0))
            else:
                g, act, k = a[0], a[1], a[2]
                n, h, w, cin = act.shape
                flops = 2.0 * n * h * w * g.shape[3] * k * k * cin
            self.records.append((name, flops, e0, e1, (n, h, w, cin, (cout if name == "conv_fprop" else g.shape[3]), k,
                                                       (a2.shape[3] if name == "conv_fprop" and a2 is not None else 0))))
            return out
        return timed

    _overhead_ms = None

    def overhead_ms(self):
        """Event-pair overhead: the same bracket around a one-element fill (a ~1.5 us kernel) measures the fixed cost
        of the two event records plus the launch gap; everything above 1.5 us of it is subtracted from each record."""
        if KernelTimer._overhead_ms is None:
            t = torch.zeros(1, device="cuda")
            vals = []
            for _ in range(25):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(200000)
                e0.record(); t.zero_(); e1.record()
                torch.cuda.synchronize()
                vals.append(e0.elapsed_time(e1))
            vals.sort()
            KernelTimer._overhead_ms = max(0.0, vals[len(vals) // 2] - 1.5e-3)
        return KernelTimer._overhead_ms

    def _ms(self, e0, e1):
        return max(e0.elapsed_time(e1) - self.overhead_ms(), 1e-4)

    def table(self):
        torch.cuda.synchronize()
        rows = []
        for name, flops, e0, e1, shape in self.records:
            ms = self._ms(e0, e1)
            rows.append({"kernel": name, "n_h_w_cin_cout_k_cin2": shape, "us": 1e3 * ms, "tflops": flops / (ms * 1e-3) / 1e12})
        return rows

    def summary(self, raw=False):
        torch.cuda.synchronize()
        out = {}
        for name, flops, e0, e1, _ in self.records:
            d = out.setdefault(name, {"launches": 0, "flops": 0.0, "ms": 0.0})
            d["launches"] += 1; d["flops"] += flops; d["ms"] += (e0.elapsed_time(e1) if raw else self._ms(e0, e1))
        return out


def haar_roofline(peaks):
    """Haar DWT GB/s at a beyond-L2 size (64x64x256x256 fp32 = 1 GiB in, 1 GiB out per launch)."""
    from unet_design_b200._lib import ops as raw
    o = raw()
    x = torch.randn(64, 64, 256, 256, device="cuda")
    res = {}

    def timeit(fn, bytes_per_launch, reps=10):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        return {"ms": ms, "GBps": bytes_per_launch / ms / 1e6, "bytes": bytes_per_launch}

    E = x.numel() * 4
    res["dwt_4band_1GiB"] = timeit(lambda: o.haar_dwt2d_fwd(x, True), 2 * E)
    ll, hi = o.haar_dwt2d_fwd(x, True)
    res["idwt_4band_1GiB"] = timeit(lambda: o.haar_idwt2d(ll, hi, 256, 256), 2 * E)
    res["dwtblock_J1_LL_tile1"] = timeit(lambda: o.dwtblock_fwd(x, 1, 64), E * 1.25)
    del ll, hi
    x2 = torch.randn(128, 128, 32, 32, device="cuda")          # config-2 encoder shape (67 MB: lives in L2)
    res["dwtblock_J1_cfg2_128x128x32x32_L2resident"] = timeit(lambda: o.dwtblock_fwd(x2, 1, 128), x2.numel() * 4 * 1.25, reps=50)
    best = res["dwt_4band_1GiB"]
    return {"bound": "hbm", "kernel": "haar_dwt_vec4<true>", "achieved": best["GBps"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": best["GBps"] / peaks["hbm_gbs"], "frac_of_8TBps_nominal": best["GBps"] / 8000.0,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, `ncu --set full` capture of this kernel on this shape
            # (profiles/r01_ncu_full_haar_raw.csv): 1.0739 GB + 1.0263 GB vs 2.1475 GB algorithmic -> no re-reads
            "traffic": 2.1002e9, "traffic_source": "profiles/r01_ncu_full_haar_raw.csv",
            "peak_source": peaks["source"], "cases": res}


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu_arm(args):
    import torch.distributed as dist
    from unet_design_b200 import _lib, ops
    from unet_design_b200.diff_cifar.model import UNetWaveletEnc
    from unet_design_b200.train import DDPMTrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    torch.manual_seed(1234)
    net = UNetWaveletEnc(**CFG).to(dev)
    net.train()
    ops.seed_dropout(1234 + rank)
    step = DDPMTrainStep(net, T=CFG["T"], lr=2e-4, warmup=5000, grad_clip=1.0, ema_decay=0.9999,
                         use_cuda_graph=not args.no_graph, overlap_allreduce=not args.no_overlap)
    gen = torch.Generator().manual_seed(rank)
    host_batches = [(torch.rand(BATCH_PER_GPU, *IMG, generator=gen) * 2 - 1).pin_memory() for _ in range(4)]
    dev_batches = [b.to(dev) for b in host_batches]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # kernel launches of one step, counted on an eager pass (graph replays do not go through Python)
    l0 = ops.launches()
    step._body(dev_batches[0])
    launches_per_step = ops.launches() - l0
    torch.cuda.synchronize()

    # ---- device-resident inputs: W warm-up + K timed steps
    for i in range(max(args.warmup, 3)):
        step(dev_batches[i % 4])
    if args.profile_step:            # ncu --profile-from-start off: exactly one (eager) step between start/stop
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step._body(dev_batches[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(dev_batches[i % 4])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(loss)

    # ---- end to end through the public API: pinned host batch in, loss float out, every step
    for i in range(3):
        step.step_from_host(host_batches[i % 4])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step.step_from_host(host_batches[i % 4])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()

    phases = None
    if world > 1 and not args.no_graph and not step._graph_has_tail:
        ph = step.timed_phases(dev_batches[0])
        tp = torch.tensor(ph, device=dev, dtype=torch.float64)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        phases = {"fwd_bwd_graph_ms": float(tp[0]), "grad_allreduce_ms": float(tp[1]), "clip_adam_ema_ms": float(tp[2])}
    elif world > 1 and not args.no_graph and step._p2p is not None:
        # the collective lives inside the step graph (bucketed, on a forked stream): what can be timed separately is the
        # whole-arena peer-memory reduction run ALONE; the exposed part is ms_per_step(N) - ms_per_step(1)
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(6)]
        for e0, e1 in ev:
            barrier()
            e0.record()
            step._allreduce_range(0, step.arena.g.numel())
            e1.record()
        torch.cuda.synchronize()
        tp = torch.tensor([sorted(e0.elapsed_time(e1) for e0, e1 in ev[1:])[2]], device=dev, dtype=torch.float64)
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        phases = {"whole_step_graph_ms": None, "grad_allreduce_alone_ms": float(tp[0]),
                  "note": "bucketed peer-memory all-reduce + optimiser tail are inside the step graph; alone = one launch over the "
                          "whole 106 MB arena, median of 5, max over ranks"}
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])
        if phases is not None and "whole_step_graph_ms" in phases:
            phases["whole_step_graph_ms"] = ms / args.steps

    # ---- roofline of the dominant kernel, measured live: one eager step with CUDA events around every conv launch.
    # Every rank runs it (the step contains the gradient all-reduce); rank 0 reports.
    timer = KernelTimer(_lib.ops())
    _lib._ops = timer
    ops.enable_side_wgrad(False)      # the events below see only the current stream: time the wgrad launches on it
    step._body(dev_batches[0])
    ksum = timer.summary()
    ksum_raw = timer.summary(raw=True)
    if args.dump_kernels and rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(timer.table(), open(os.path.join(ROOT, "gpurun_out", "conv_launch_table.json"), "w"), indent=0)
    _lib._ops = timer._real
    barrier()

    line = None
    if rank == 0:
        imgs = BATCH_PER_GPU * world * args.steps
        fp = ksum.get("conv_fprop", {"flops": 0.0, "ms": 1.0, "launches": 0})
        wg = ksum.get("conv_wgrad", {"flops": 0.0, "ms": 1.0, "launches": 0})
        peak_tf = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        raw_fp, raw_wg = ksum_raw.get("conv_fprop", fp), ksum_raw.get("conv_wgrad", wg)
        ach = raw_fp["flops"] / (raw_fp["ms"] * 1e-3) / 1e12                 # no event-overhead subtraction
        ach_adj = fp["flops"] / (fp["ms"] * 1e-3) / 1e12
        roofline = {"bound": "tensor", "kernel": "conv_fprop_kernel (fprop + dgrad launches)", "achieved": ach, "peak": peak_tf,
                    "unit": "TFLOP/s", "frac": ach / peak_tf, "frac_of_burst_peak": ach / peaks["bf16_tflops"],
                    "achieved_event_overhead_adjusted": ach_adj,
                    # DRAM bytes of the largest launch (256->256 @ 32x32, batch 128, 106 us under ncu) in the per-launch metrics
                    # pass profiles/r02_conv_traffic_all_launches.csv: 68.3 MB read + 21.8 MB written (67 + 67 MB
                    # algorithmic; part of the output is still in L2 when the kernel ends), tensor pipe 83 % active
                    "traffic": 90.1e6, "traffic_source": "profiles/r02_conv_traffic_all_launches.csv (largest launch)",
                    "event_overhead_us_subtracted": 1e3 * timer.overhead_ms(),
                    "launches_per_step": fp["launches"],
                    "flops_per_step": fp["flops"], "ms_per_step": fp["ms"],
                    "peak_source": peaks["source"] + ", sustained figure (kernel timed inside a long step)",
                    "wgrad": {"achieved": raw_wg["flops"] / (raw_wg["ms"] * 1e-3) / 1e12, "launches_per_step": wg["launches"],
                              "ms_per_step": raw_wg["ms"], "frac": raw_wg["flops"] / (raw_wg["ms"] * 1e-3) / 1e12 / peak_tf,
                              "frac_of_burst_peak": raw_wg["flops"] / (raw_wg["ms"] * 1e-3) / 1e12 / peaks["bf16_tflops"],
                              "achieved_event_overhead_adjusted": wg["flops"] / (wg["ms"] * 1e-3) / 1e12},
                    "conv_share_of_step": (fp["ms"] + wg["ms"]) / (ms / args.steps)}
        haar = haar_roofline(peaks) if not args.skip_haar else None
        cpu = None
        if not args.skip_cpu:
            times, cores = cpu_train_steps(BATCH_PER_GPU, 4, 1)
            cpu = {"value": BATCH_PER_GPU * len(times) / sum(times), "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": f"{len(times)} steps of the full per-GPU batch {BATCH_PER_GPU}, same model and optimiser, fp32 PyTorch "
                             "CPU (oracle/torch_ref.py)"}
        lib = gpu_library_baseline(dev) if not args.skip_lib else None
        line = {
            "metric": METRIC, "value": imgs / (ms * 1e-3), "unit": "images/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": BATCH_PER_GPU * world, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph, "allreduce": ("none" if world == 1 else (((("bucketed peer-memory (NVLink P2P) all-reduce kernels" if step._p2p is not None else "bucketed NCCL all-reduces") + " captured inside the step graph, overlapped with backward") if step._graph_has_tail else ("one " + ("peer-memory" if step._p2p is not None else "NCCL") + " all-reduce of the gradient arena after the forward+backward graph")) if not args.no_graph else ("tail" if args.no_overlap else "bucketed, overlapped with backward"))), "l2_policy": "4 rotating input batches; activations+weights per step "
                       "(~3 GB) exceed the 126 MB L2", "loss": loss_val},
            "clocks": clocks,
            "e2e": {"value": imgs / e2e_s, "unit": "images/s", "h2d_bytes_per_step": BATCH_PER_GPU * 3 * 32 * 32 * 4,
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "dp_phases": phases,
            "dp_bucket_trace": ({"gradients": len(step.arena.params), "launched_after": [n for _, n in step.bucket_trace],
                                 "bucket_mb": [round((hi - lo) * 4 / 2**20, 1) for lo, hi, _ in step._buckets]}
                                if world > 1 and step._buckets else None),
            "roofline": roofline, "haar_roofline": haar, "cpu_baseline": cpu, "gpu_library_baseline": lib,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def run_gpu_arm_generic(args):
    """BASELINE configs 1, 3, 4, 5 through the model-agnostic `TrainStep` (one CUDA graph per step, same timing rules)."""
    import torch.distributed as dist
    import bench_configs as bc
    from unet_design_b200 import _lib, ops
    from unet_design_b200.train import TrainStep

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = measured_peaks()
    workload, batch, unit, gflop = bc.CONFIGS[args.config]
    built = bc.build(args.config, dev, rank)
    model = built["model"]
    model.train()
    step = TrainStep(model, built["loss_fn"], use_cuda_graph=not args.no_graph, overlap_allreduce=not args.no_overlap,
                     **built["opt"])
    host = [built["host_batch"]() for _ in range(4)]
    devb = [tuple(t.to(dev) for t in hb) for hb in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    l0 = ops.launches()
    step._body(*devb[0])
    launches_per_step = ops.launches() - l0
    for i in range(max(args.warmup, 3)):
        step(*devb[i % 4])
    if args.profile_step:            # ncu --profile-from-start off: exactly one (eager) step between start/stop
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        step._body(*devb[0])
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(*devb[i % 4])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    loss_val = float(loss)

    def from_host(hb):
        out = step(*[t.to(dev, non_blocking=True) for t in hb])
        return float(out.cpu())

    for i in range(3):
        from_host(host[i % 4])
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        from_host(host[i % 4])
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    barrier()
    if world > 1:
        t = torch.tensor([ms, e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_s = float(t[0]), float(t[1])

    timer = KernelTimer(_lib.ops())
    _lib._ops = timer
    ops.enable_side_wgrad(False)
    step._body(*devb[0])
    ksum = timer.summary(raw=True)
    _lib._ops = timer._real
    barrier()

    line = None
    if rank == 0:
        n = batch * world * args.steps
        peak_tf = peaks["bf16_tflops_sustained"] or peaks["bf16_tflops"]
        fl = sum(d["flops"] for d in ksum.values())
        tms = sum(d["ms"] for d in ksum.values())
        ach = fl / (tms * 1e-3) / 1e12
        value = n / (ms * 1e-3)
        roofline = {"bound": "tensor", "kernel": "conv_fprop_kernel + conv_wgrad_kernel (all conv launches of one step)",
                    "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                    "frac_of_burst_peak": ach / peaks["bf16_tflops"], "traffic": None,
                    "launches_per_step": sum(d["launches"] for d in ksum.values()), "flops_per_step": fl, "ms_per_step": tms,
                    "conv_share_of_step": tms / (ms / args.steps),
                    "whole_step_tflops": gflop * 1e9 * value / world / 1e12,
                    "whole_step_frac": gflop * 1e9 * value / world / 1e12 / peak_tf,
                    "peak_source": peaks["source"] + ", sustained figure; CUDA events around every conv launch of one eager step, raw"}
        cpu = lib = None
        if not args.skip_cpu:
            small = {"c3": 2, "c3u": 2, "c4": 2, "c5": 4}.get(args.config)
            if small:
                times, cores = reference_steps(args.config, small, 2, 1)
                cpu = {"value": small * len(times) / sum(times), "unit": unit, "cores": cores, "kind": "port",
                       "sample": f"{len(times)} steps of batch {small} (of {batch}), fp32 PyTorch CPU (oracle/torch_ref_pde.py)"}
            else:
                probe = os.path.join(ROOT, "profiles", "r02_reference_cpu_probe.json")
                ref = json.load(open(probe)).get(args.config) if os.path.exists(probe) else None
                if ref:
                    cpu = {"value": ref["value"], "unit": unit, "cores": ref["cores"], "kind": "reference",
                           "sample": ref["sample"] + " -- measured in the build container by tools/ref_cpu_probe.py (the reference "
                                     "cannot travel to the GPU box), NOT on this host"}
        if not args.skip_lib and args.config != "c1":
            lib = {}
            for tag, ac, cl in (("bf16_autocast_channels_last", True, True), ("fp32_tf32_convs", False, False)):
                times, _ = reference_steps(args.config, batch, 4, 2, dev=dev, autocast=ac, channels_last=cl)
                lib[tag] = {"ms_per_step": 1e3 * sum(times) / len(times), "samples_per_s": batch * len(times) / sum(times)}
            lib["what"] = "oracle/torch_ref_pde.py on this GPU via cuDNN (eager), full batch; not the product path"
        line = {
            "metric": f"{args.config}_train_{unit.replace('/', '_per_')}", "value": value, "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "global_batch": batch * world, "parallelism": f"dp{world}",
                       "cuda_graph": not args.no_graph, "l2_policy": "4 rotating input batches", "loss": loss_val},
            "clocks": clocks,
            "e2e": {"value": n / e2e_s, "unit": unit, "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host[0]),
                    "d2h_bytes_per_step": 4},
            "gpu_launches": launches_per_step * args.steps,
            "roofline": roofline, "cpu_baseline": cpu, "gpu_library_baseline": lib,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of one CUDA graph per step")
    ap.add_argument("--no-overlap", action="store_true", help="one gradient all-reduce after backward instead of bucketed overlap")
    ap.add_argument("--dump-kernels", action="store_true", help="write per-launch conv timings to gpurun_out/conv_launch_table.json")
    ap.add_argument("--profile-step", action="store_true", help="bracket one eager step with cudaProfilerStart/Stop (for ncu)")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-haar", action="store_true")
    ap.add_argument("--skip-lib", action="store_true", help="skip the cuDNN / library comparator leg")
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c3u", "c4", "c5"],
                    help="BASELINE.json config (c2 = the headline diff_cifar Multi-ResNet; c3u = the residual U-Net arm of c3)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.config != "c2":
        run_gpu_arm_generic(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
