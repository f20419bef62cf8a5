"""Training step of the Multi-ResNet DDPM (the hot loop of diff_cifar/main.py:397-429) on one B200 or,
data-parallel, on the GPUs of one box.

    step = DDPMTrainStep(model, T=1000, lr=2e-4, warmup=5000, grad_clip=1.0, ema_decay=0.9999)
    loss = step(x0)                       # x0 already on the device
    loss = step.step_from_host(x0_pinned) # the end-to-end call: H2D copy, step, loss read-back

What is B200-first about it (reference file:line in parentheses):

* parameters, gradients, Adam moments and the EMA copy live in four flat fp32 arenas; conv weights keep
  their [Cout, kh, kw, Cin] order inside the arena, so the tensor-core kernels read them in place;
* `clip_grad_norm_` + `Adam.step` + `LambdaLR` warm-up + `ema()` (main.py:425-429, :57-77, :90-91) are TWO
  kernels over the arena (sum of squares, then clip+Adam+EMA), with the step counter, learning-rate
  warm-up and dropout counter on the device: no host synchronisation anywhere in the step;
* the whole step (noise draw, q-sample, forward, loss, backward, optimiser) is captured in ONE CUDA graph and
  replayed, because the network is several hundred small-to-medium kernels and launch-bound otherwise;
* data parallel (`nn.DataParallel` at main.py:235-238 in the reference): one process per GPU, identical
  replicas, sum all-reduce of the gradient arena, the 1/world mean folded into the optimiser's grad_scale.
  Two modes: (a) CUDA graph of forward+backward, then ONE all-reduce of the 106 MB arena and the 2-kernel
  optimiser tail (default; ~0.4 ms of NVLink time per step); (b) eager (`use_cuda_graph=False`): arena buckets
  all-reduced by NCCL from post-accumulate hooks on a side stream as backward produces them, overlapping the
  rest of backward.
"""
from __future__ import annotations

import os

from typing import List, Optional

import torch
import torch.distributed as dist
from torch import nn

from . import ops
from .diff_cifar.diffusion import GaussianDiffusionTrainer


class _Tf32Matmul:
    """fp32 nn.Linear layers of the time-embedding path (tiny GEMMs, otherwise run as SIMT sgemm) on TF32 tensor
    cores for the duration of the step; the caller's global setting is restored afterwards."""

    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


class FlatArena:
    """Trainable parameters of `model` re-homed into one flat fp32 buffer (and a matching gradient buffer).
    Conv weights (4-D) keep channels_last order: the arena slice is viewed as [Cout,kh,kw,Cin] and permuted."""

    def __init__(self, model: nn.Module):
        params = [p for p in model.parameters() if p.requires_grad]
        assert params, "model has no trainable parameters"
        dev = params[0].device
        # modules may ask for some parameters to be neighbours (AttnBlock: q|k|v weights, q|k|v biases), so that their
        # concatenation is a view of the arena; members of a group are emitted together at the first member's position
        groups, member_of = [], {}
        for m in model.modules():
            if hasattr(m, "fused_param_groups"):
                gw, gb = m.fused_param_groups()
                if all(p.requires_grad and p.numel() % 4 == 0 for p in list(gw) + list(gb)):
                    groups.append((list(gw), list(gb)))
                    for grp in (gw, gb):
                        for p in grp:
                            member_of[id(p)] = grp
        ordered, seen = [], set()
        for p in params:
            if id(p) in seen:
                continue
            for q in member_of.get(id(p), [p]):
                ordered.append(q)
                seen.add(id(q))
        params = ordered
        self.fused_groups = groups
        self.params = params
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # keep every slice 16-byte aligned
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        total = self.offsets[-1]
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        self.g = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(params, self.offsets):
            view = self._view(self.p, p, off)
            view.copy_(p.data)
            p.data = view
            p.grad = self._view(self.g, p, off)

    @staticmethod
    def _view(flat, p, off):
        n = p.numel()
        if p.dim() == 4:
            co, ci, kh, kw = p.shape
            return flat[off:off + n].view(co, kh, kw, ci).permute(0, 3, 1, 2)
        return flat[off:off + n].view(p.shape)

    def rebind_grads(self):
        """Point every .grad back at the arena (after anything that may have replaced it)."""
        for p, off in zip(self.params, self.offsets):
            p.grad = self._view(self.g, p, off)


class PackedWeights:
    """bf16 GEMM operands of every conv weight, refreshed once per optimiser step.

    * `shadow`: bf16 copy of the whole parameter arena, written by the Adam kernel.  A conv weight's slice is its
      fprop / wgrad operand ([Cout, kh, kw, Cin], same order as the fp32 master).
    * `dgrad`: the transposed + 180-degree-rotated operands ([Cin_pad, kh, kw, Cout]) of all convs, produced by ONE
      kernel launch from a device-resident table.
    Registered with `ops.register_packed_weight`, so forward / backward calls pick them up instead of re-packing.
    Call `refresh_from_master()` after loading a state_dict into a model that already has a training step.
    """

    def __init__(self, arena: FlatArena):
        self.arena = arena
        dev = arena.p.device
        self.shadow = torch.empty(arena.p.numel(), dtype=torch.bfloat16, device=dev)
        rows, doff = [], 0
        self.entries = []
        for p, off in zip(arena.params, arena.offsets):
            if p.dim() == 4 and p.shape[0] % 16 == 0 and p.shape[1] % 16 == 0 and p.shape[2] == p.shape[3]:
                cout, cin, k, _ = p.shape
                n_fwd, n_bwd = cout * k * k * cin, (cin + 15) // 16 * 16 * k * k * cout
                rows.append([off, doff, cout, cin, k])
                self.entries.append((p, off, n_fwd, doff, n_bwd))
                doff += (n_bwd + 7) // 8 * 8                    # keep every operand 16-byte aligned
        # fused 1x1 groups (q|k|v): one more dgrad operand for the concatenated weight, everything else is a view
        off_of = {id(p): off for p, off in zip(arena.params, arena.offsets)}
        fused = []
        for gw, gb in getattr(arena, "fused_groups", []):
            cout, cin = sum(w.shape[0] for w in gw), gw[0].shape[1]
            if gw[0].dim() == 4 and gw[0].shape[2] == 1 and cin % 16 == 0 and cout % 16 == 0:
                off, boff = off_of[id(gw[0])], off_of[id(gb[0])]
                rows.append([off, doff, cout, cin, 1])
                fused.append((gw[0], off, boff, cout, cin, doff))
                doff += (cin * cout + 7) // 8 * 8
        self.dgrad = torch.empty(max(doff, 8), dtype=torch.bfloat16, device=dev)
        self.table = torch.tensor(rows, dtype=torch.int64, device=dev) if rows else None
        for p, off, n_fwd, d0, n_bwd in self.entries:
            ops.register_packed_weight(p, self.shadow[off:off + n_fwd], self.dgrad[d0:d0 + n_bwd])
        for first, off, boff, cout, cin, d0 in fused:
            n = cout * cin
            ops.register_fused_conv(first, self.shadow[off:off + n], self.dgrad[d0:d0 + n],
                                    arena.g[off:off + n].view(cout, 1, 1, cin).permute(0, 3, 1, 2),
                                    arena.p[boff:boff + cout], arena.g[boff:boff + cout])
        self.refresh_from_master()

    def refresh_dgrad(self):
        if self.table is not None:
            ops.pack_dgrad_weights_batched_(self.shadow, self.dgrad, self.table)

    def refresh_from_master(self):
        self.shadow.copy_(self.arena.p)
        self.refresh_dgrad()


class DDPMTrainStep:
    def __init__(self, model: nn.Module, T: int = 1000, beta_1: float = 1e-4, beta_T: float = 0.02, lr: float = 2e-4,
                 warmup: int = 5000, grad_clip: float = 1.0, ema_decay: float = 0.9999, multi_res_loss: bool = False,
                 betas=(0.9, 0.999), eps: float = 1e-8, use_cuda_graph: bool = True, process_group=None,
                 bucket_mb: float = 16.0, overlap_allreduce: bool = True):
        self.model = model
        self.device = next(model.parameters()).device
        self.trainer = GaussianDiffusionTrainer(model, beta_1, beta_T, T, multi_res_loss, False, self.device).to(self.device)
        self.lr, self.warmup, self.grad_clip, self.ema_decay = lr, warmup, grad_clip, ema_decay
        self.betas, self.eps = betas, eps
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or
                                                              (dist.is_available() and dist.is_initialized())) else 1
        if self.world > 1:      # identical replicas: rank 0's initial weights everywhere (one broadcast)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        self.arena = FlatArena(model)
        n = self.arena.p.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.ema = self.arena.p.clone()
        self.packed = PackedWeights(self.arena)
        for p, off in zip(self.arena.params, self.arena.offsets):
            ops.register_grad_sink(p, p.grad, None)
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        if self.device.type == "cuda":
            # second stream for the kernels that only feed the optimiser (ops._Side).  Not with the eager bucketed
            # all-reduce: its hooks launch a bucket's collective as soon as autograd has passed the parameters, which
            # would race with a weight gradient still queued on the side stream.
            eager_overlap = self.world > 1 and overlap_allreduce and not use_cuda_graph
            ops.enable_side_wgrad(os.environ.get("UB200_SIDE_WGRAD", "1") != "0" and not eager_overlap,
                                  int(os.environ.get("UB200_SIDE_WGRAD_PIXELS", str(1 << 30))))
            ops._Side.chansum = os.environ.get("UB200_SIDE_CHANSUM", "1") != "0" and not eager_overlap
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)       # 1-based after the first bump
        self.steps_done = 0
        self.use_graph = use_cuda_graph and self.device.type == "cuda"
        self._graph = None
        self._static_x0 = None
        self._static_loss = None
        self._buckets: List[tuple] = []
        self._comm_stream = None
        self._hooks_live = False
        self.overlap = overlap_allreduce
        if self.world > 1 and self.overlap:
            self._build_buckets(int(bucket_mb * (1 << 20) / 4))

    # ------------------------------------------------------------------ data parallel
    def _build_buckets(self, bucket_elems: int):
        """Arena ranges in REVERSE parameter order (the order backward fills them); each range is all-reduced
        as soon as its last gradient has been accumulated."""
        offs, params = self.arena.offsets, self.arena.params
        hi = offs[-1]
        members: List[int] = []
        for i in range(len(params) - 1, -1, -1):
            members.append(i)
            if hi - offs[i] >= bucket_elems or i == 0:
                self._buckets.append((offs[i], hi, tuple(members)))
                hi = offs[i]
                members = []
        self._bucket_of = {}
        for b, (_, _, mem) in enumerate(self._buckets):
            for i in mem:
                self._bucket_of[i] = b
        self._pending = [0] * len(self._buckets)
        if self.device.type == "cuda":
            self._comm_stream = torch.cuda.Stream(device=self.device)
        # note: post-accumulate hooks also fire for parameters whose Function returned None because its kernel
        # accumulated straight into the arena (ops gradient sinks), and they fire after that kernel was enqueued
        for i, p in enumerate(params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i: int):
        def hook(_param):
            if not self._hooks_live:
                return
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0:
                self._launch_allreduce(b)
        return hook

    def _launch_allreduce(self, b: int):
        lo, hi, _ = self._buckets[b]
        view = self.arena.g[lo:hi]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self._comm_stream):
                dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg)
        else:
            dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.pg)

    def _arm_buckets(self):
        for b, (_, _, mem) in enumerate(self._buckets):
            self._pending[b] = len(mem)

    def _finish_allreduce(self):
        """After backward: reduce the buckets that hold parameters this step did not touch (e.g. the coarse
        tails without the multi-resolution loss; their hooks never fire), then join the side stream.  The set of
        untouched parameters is the same on every rank, so the collective order stays consistent."""
        for b in range(len(self._buckets)):
            if self._pending[b] > 0:
                self._pending[b] = 0
                self._launch_allreduce(b)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    # ------------------------------------------------------------------ one step
    def _fwd_bwd(self, x0: torch.Tensor, overlap: bool) -> torch.Tensor:
        self.arena.g.zero_()
        self.step_dev.add_(1)
        self._hooks_live = overlap
        if overlap:
            self._arm_buckets()
        with _Tf32Matmul():
            loss, _ = self.trainer(x0)
            loss.backward()
        ops.join_side_stream()          # small-layer weight gradients ran on a second stream (ops._Side)
        return loss.detach()

    def _reduce_and_update(self, overlapped: bool):
        if self.world > 1:
            if overlapped:
                self._finish_allreduce()
            else:                      # one all-reduce of the whole arena after backward
                dist.all_reduce(self.arena.g, op=dist.ReduceOp.SUM, group=self.pg)
        self._update()

    def _update(self):
        self.sumsq.zero_()
        ops.sumsq_(self.arena.g, self.sumsq)
        # gradients hold the SUM over ranks: the mean (what DataParallel / DDP produce) is a grad_scale of 1/world
        ops.adam_ema_step_(self.arena.p, self.arena.g, self.m, self.v, self.ema, self.sumsq, self.grad_clip,
                           1.0 / self.world, self.lr, self.betas[0], self.betas[1], self.eps, self.ema_decay, 1,
                           self.warmup, self.step_dev, self.packed.shadow)
        self.packed.refresh_dgrad()                 # one launch: dgrad operands of every conv from the bf16 shadow
        ops.advance_dropout_state(self.device)

    def _body(self, x0: torch.Tensor) -> torch.Tensor:
        """Eager step; with several ranks the bucketed all-reduce overlaps backward."""
        overlap = self.world > 1 and self.overlap
        loss = self._fwd_bwd(x0, overlap)
        self._reduce_and_update(overlap)
        return loss

    def __call__(self, x0: torch.Tensor) -> torch.Tensor:
        self.steps_done += 1
        if not self.use_graph:
            return self._body(x0)
        if self._graph is None:
            self._capture(x0)
        self._static_x0.copy_(x0, non_blocking=True)
        self._graph.replay()
        if self.world > 1:             # the collective and the optimiser tail stay outside the graph (3 launches)
            self._reduce_and_update(False)
        return self._static_loss

    def timed_phases(self, x0: torch.Tensor, reps: int = 5):
        """Device time (ms, averaged) of the three phases of a data-parallel step: forward+backward graph, gradient
        all-reduce, clip+Adam+EMA tail.  Diagnostic only (bench.py reports it next to the step time)."""
        assert self.use_graph and self.world > 1 and self._graph is not None
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
        for r in range(reps):
            self._static_x0.copy_(x0, non_blocking=True)
            ev[r][0].record()
            self._graph.replay()
            ev[r][1].record()
            dist.all_reduce(self.arena.g, op=dist.ReduceOp.SUM, group=self.pg)
            ev[r][2].record()
            self._update()
            ev[r][3].record()
        torch.cuda.synchronize(self.device)
        return [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / reps for i in range(3)]

    def _capture(self, x0: torch.Tensor):
        """One CUDA graph for the whole step on a single GPU; forward + backward only when data-parallel (NCCL work
        captured into a graph dead-locked on this stack, so the gradient all-reduce is issued right after the replay)."""
        self._static_x0 = x0.clone()
        whole = self.world == 1
        # The warm-up runs real steps (allocator, tensor maps, cuBLAS handles): snapshot every piece of training state
        # they touch and put it back, so that the first replay is step 1 of the run exactly as in eager mode.
        state = [self.arena.p, self.m, self.v, self.ema, self.step_dev, self.packed.shadow, self.packed.dgrad,
                 ops.dropout_device_counter(self.device)]
        saved = [t.clone() for t in state]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):           # warm-up on a side stream
            for _ in range(3):
                self._fwd_bwd(self._static_x0, False)
                self._reduce_and_update(False)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for t, keep in zip(state, saved):
            t.copy_(keep)
        del saved
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._static_loss = self._fwd_bwd(self._static_x0, False)
            if whole:
                self._reduce_and_update(False)

    def step_from_host(self, x0_pinned: torch.Tensor) -> float:
        """The end-to-end call a user makes: batch in pinned host memory in, loss (a Python float) out."""
        x0 = x0_pinned.to(self.device, non_blocking=True)
        loss = self(x0)
        return float(loss.cpu())

    # ------------------------------------------------------------------ checkpoint surface
    def ema_state_dict(self):
        """`state_dict` of the EMA model (the reference keeps a deep copy, main.py:207-219)."""
        sd = {k: v.clone() for k, v in self.model.state_dict().items()}
        names = {id(p): n for n, p in self.model.named_parameters()}
        for p, off in zip(self.arena.params, self.arena.offsets):
            sd[names[id(p)]] = FlatArena._view(self.ema, p, off).clone()
        return sd
