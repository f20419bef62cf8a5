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

`TrainStep(model, loss_fn)` is the model-agnostic form (any of the reference's families); `DDPMTrainStep` binds it to the
diff_cifar loss.  `state_dict()` / `load_state_dict()` cover what diff_cifar/main.py:443-453 checkpoints.
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

    def __init__(self, model: nn.Module, grad_factory=None):
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
        # parameters whose gradients only appear at the very end of backward (a model says which: e.g. the batched time-embedding
        # path of diff_cifar, whose ONE backward launch serves every ResBlock) go to the front of the arena = into the LAST
        # data-parallel bucket; left in place they would hold every bucket back until backward is over
        late = set()
        for m in model.modules():
            if hasattr(m, "late_grad_params"):
                late.update(id(p) for p in m.late_grad_params())
        if late:
            ordered = [p for p in ordered if id(p) in late] + [p for p in ordered if id(p) not in late]
        params = ordered
        self.fused_groups = groups
        self.params = params
        sizes = [(p.numel() + 3) // 4 * 4 for p in params]          # keep every slice 16-byte aligned
        self.offsets = [0]
        for s in sizes:
            self.offsets.append(self.offsets[-1] + s)
        total = self.offsets[-1]
        self.p = torch.zeros(total, dtype=torch.float32, device=dev)
        # grad_factory(total) -> zeroed fp32 [total] buffer on `dev`: data parallel puts the gradients in peer-mapped memory
        self.g = grad_factory(total) if grad_factory is not None else torch.zeros(total, dtype=torch.float32, device=dev)
        assert self.g.numel() == total and self.g.dtype == torch.float32 and self.g.device == dev
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


class PeerAllReduce:
    """Sum all-reduce of ranges of a gradient arena over NVLink peer memory (csrc/p2p.cu): each rank's arena lives in a
    symmetric block that every other rank of the box has mapped through CUDA IPC, and one kernel launch per range does
    the two-shot reduce in place.  Being a plain launch it is captured into the step's CUDA graph on the comm stream,
    which is what NCCL could not be on this stack.  torch.distributed only carries the 64-byte handles."""

    def __init__(self, device: torch.device, group=None):
        self.device, self.group = device, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        assert 2 <= self.world <= 8, "peer all-reduce: 2..8 ranks of one box"
        self.arena = None
        self.bases: List[int] = []

    def allocate(self, floats: int) -> torch.Tensor:
        o = ops._ops()
        self.arena, handle, base = o.p2p_alloc(int(floats), self.device.index)
        blobs = [None] * self.world
        dist.all_gather_object(blobs, (os.uname().nodename, bytes(handle.numpy().tobytes())), group=self.group)
        if any(b[0] != blobs[self.rank][0] for b in blobs):
            raise RuntimeError("peer all-reduce: ranks on different hosts")
        for r, (_, h) in enumerate(blobs):
            self.bases.append(int(base) if r == self.rank else
                              int(o.p2p_open(torch.frombuffer(bytearray(h), dtype=torch.uint8), self.device.index)))
        dist.barrier(group=self.group)          # every rank has mapped every block before the first launch
        return self.arena

    def __call__(self, lo: int, hi: int, ctas: int = 0) -> None:
        ops._ops().p2p_allreduce(self.arena, self.bases, self.rank, int(lo), int(hi - lo), int(ctas))
        ops._count(1)


class TrainStep:
    """Model-agnostic fast training step: flat arenas, packed bf16 weight shadow, gradient sinks, the fused
    clip + Adam(W) + EMA tail and (optionally) the whole step as one CUDA graph, data-parallel over `process_group`.

        step = TrainStep(model, loss_fn, lr=2e-4, weight_decay=1e-5)      # pdearena: AdamW
        loss = step(x, y)                                                # tensors already on the device

    `loss_fn(*tensors, **static)` runs the model and returns a scalar loss (or a tuple whose first element is the loss).
    Tensor arguments are copied into static buffers when the step is a CUDA graph; keyword arguments must be hashable
    Python values (e.g. `n_levels_used`): graphs are captured per (tensor shapes, keyword values).
    Serves every model family of the reference: diff_cifar through `DDPMTrainStep` below, diff_mnist / pdearena /
    wmh through a loss closure (bench.py `--config`)."""

    def __init__(self, model: nn.Module, loss_fn, lr: float = 2e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, warmup: int = 0, grad_clip: float = 0.0, ema_decay: Optional[float] = None,
                 use_cuda_graph: bool = True, process_group=None, bucket_mb: float = 16.0,
                 overlap_allreduce: bool = True):
        self.model = model
        self.loss_fn = loss_fn
        self.device = next(model.parameters()).device
        self.lr, self.warmup, self.grad_clip, self.ema_decay = lr, warmup, grad_clip, ema_decay
        self.betas, self.eps, self.weight_decay = betas, eps, weight_decay
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (process_group is not None or
                                                              (dist.is_available() and dist.is_initialized())) else 1
        if self.world > 1:      # identical replicas: rank 0's initial weights everywhere (one broadcast)
            for t in list(model.parameters()) + list(model.buffers()):
                dist.broadcast(t.data, src=0, group=process_group)
        # data parallel on one box: gradients live in peer-mapped memory and are reduced by our own kernel (UB200_DP_P2P=0:
        # NCCL).  Any failure to set the peer mapping up (no IPC in this container, different hosts) falls back to NCCL.
        self._p2p = None
        grad_factory = None
        if self.world > 1 and self.device.type == "cuda" and self.world <= 8 and os.environ.get("UB200_DP_P2P", "1") != "0":
            ok = torch.zeros(1, device=self.device)
            try:
                peer = PeerAllReduce(self.device, process_group)
                n_floats = sum((p.numel() + 3) // 4 * 4 for p in model.parameters() if p.requires_grad)
                peer.allocate(n_floats)
                ok.fill_(1)
            except Exception as e:          # noqa: BLE001 -- reported, then the NCCL path runs
                peer = None
                print(f"[unet_design_b200] peer-memory all-reduce unavailable ({type(e).__name__}: {e}); using NCCL", flush=True)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=process_group)       # all ranks or none
            if float(ok) == 1.0:
                self._p2p = peer
                grad_factory = lambda total: peer.arena                     # noqa: E731
        self.arena = FlatArena(model, grad_factory)
        n = self.arena.p.numel()
        self.m = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.v = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.ema = self.arena.p.clone() if ema_decay is not None else None
        self.packed = PackedWeights(self.arena)
        # gradient sinks: kernels that accumulate (wgrad, GroupNorm dgamma / dbeta, bias sums) write straight into the arena;
        # `on_ready` tells the data-parallel bucket logic that the kernel producing this parameter's gradient is enqueued
        # (autograd's post-accumulate hooks do not fire for them: their Function returns None for the parameter)
        for i, (p, off) in enumerate(zip(self.arena.params, self.arena.offsets)):
            ops.register_grad_sink(p, p.grad, self._make_hook(i))
        self.sumsq = torch.zeros(1, dtype=torch.float32, device=self.device)
        if self.device.type == "cuda":
            # second stream for the kernels that only feed the optimiser (ops._Side); the bucketed all-reduce waits for it
            ops.enable_side_wgrad(os.environ.get("UB200_SIDE_WGRAD", "1") != "0",
                                  int(os.environ.get("UB200_SIDE_WGRAD_PIXELS", str(1 << 30))))
            ops._Side.chansum = os.environ.get("UB200_SIDE_CHANSUM", "1") != "0"
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=self.device)       # 1-based after the first bump
        self.steps_done = 0
        self.use_graph = use_cuda_graph and self.device.type == "cuda"
        self._graphs = {}              # (tensor shapes / dtypes, static kwargs) -> (graph, static inputs, static loss)
        self._graph_has_tail = False   # the captured graph includes the collective + optimiser tail
        self._dgrad_stale = False      # the optimiser has moved the weights since the dgrad operands were packed
        self._pack_stream = None
        self._buckets: List[tuple] = []
        self._comm_stream = None
        self._hooks_live = False
        self._probe, self._silent, self.bucket_trace, self._launched = False, frozenset(), [], []
        self._early_ctas = int(os.environ.get("UB200_P2P_EARLY_CTAS", "0"))
        self.overlap = overlap_allreduce and os.environ.get("UB200_DP_OVERLAP", "1") != "0"
        if self.world > 1 and self.overlap:
            # (peer-memory buckets pay two cross-GPU barriers each: 32 MB measured better than 16 MB at 2 GPUs)
            bucket_mb = float(os.environ.get("UB200_DP_BUCKET_MB", 32.0 if self._p2p is not None and bucket_mb == 16.0 else bucket_mb))
            self._build_buckets(int(bucket_mb * (1 << 20) / 4))
        # `model.load_state_dict(...)` after this point writes the fp32 masters in place (they are arena views); the
        # bf16 operands forward / backward actually read must follow, and before any training the EMA is by definition
        # a copy of the (now loaded) weights -- diff_cifar/main.py:217-221 deep-copies, then restores both models.
        model.register_load_state_dict_post_hook(lambda _m, _incompatible: self._after_model_load())

    def _after_model_load(self):
        self.packed.refresh_from_master()
        self._dgrad_stale = False
        if self.ema is not None and self.steps_done == 0:
            self.ema.copy_(self.arena.p)

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
        self._launched = [False] * len(self._buckets)
        self._seen = [False] * len(params)
        self._silent = frozenset()     # parameters the current step is known not to touch (found by a probe pass, see _capture)
        self._probe = False
        self.bucket_trace: List[tuple] = []   # (bucket, gradients reported so far) in launch order: diagnostic
        if self.device.type == "cuda":
            self._comm_stream = torch.cuda.Stream(device=self.device)
        # parameters whose gradient arrives through autograd's AccumulateGrad (no sink kernel) report through this hook;
        # the sink kernels report through `on_ready` (registered in __init__): whichever fires first counts, once per step
        for i, p in enumerate(params):
            p.register_post_accumulate_grad_hook(self._make_hook(i))

    def _make_hook(self, i: int):
        def hook(_param=None):
            if not self._hooks_live or self._seen[i]:
                return
            self._seen[i] = True
            if self._probe:
                return
            if i in self._silent:
                raise RuntimeError("data-parallel buckets: a parameter the probe pass saw no gradient for has produced one")
            b = self._bucket_of[i]
            self._pending[b] -= 1
            if self._pending[b] == 0 and not self._launched[b]:
                self._launch_allreduce(b)
        return hook

    def _allreduce_range(self, lo: int, hi: int, early: bool = False):
        if self._p2p is not None:
            # a bucket launched from a gradient hook runs NEXT TO backward kernels and has slack: throttled to a few CTAs
            self._p2p(lo, hi, self._early_ctas if early else 0)
        else:
            dist.all_reduce(self.arena.g[lo:hi], op=dist.ReduceOp.SUM, group=self.pg)

    def _launch_allreduce(self, b: int, early: bool = True):
        lo, hi, _ = self._buckets[b]
        self._launch_range(lo, hi, [b], early)

    def _launch_range(self, lo: int, hi: int, members, early: bool):
        if len(self._launched) != len(self._buckets):
            self._launched = [False] * len(self._buckets)
        for b in members:
            self._launched[b] = True
            self.bucket_trace.append((b, sum(getattr(self, "_seen", ()))))
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream(self.device))
            if ops._Side.stream is not None:      # weight-gradient kernels into this bucket may be queued on the second stream
                self._comm_stream.wait_stream(ops._Side.stream)
            with torch.cuda.stream(self._comm_stream):
                self._allreduce_range(lo, hi, early)
        else:
            self._allreduce_range(lo, hi, early)

    def _arm_buckets(self):
        for b, (_, _, mem) in enumerate(self._buckets):
            self._pending[b] = sum(1 for i in mem if i not in self._silent)
        self._launched = [False] * len(self._buckets)
        self._seen = [False] * len(self.arena.params)
        self.bucket_trace = []

    def _finish_allreduce(self):
        """After backward: reduce the buckets that hold parameters this step did not touch (e.g. the coarse
        tails without the multi-resolution loss; their hooks never fire), then join the side stream.  The set of
        untouched parameters is the same on every rank, so the collective order stays consistent."""
        # what is left when backward ends (the buckets that hold the last gradients, parameters this step did not touch) goes
        # out as ONE launch per contiguous arena range: each launch costs two cross-GPU barriers on the critical path
        left = sorted((self._buckets[b][0], self._buckets[b][1], b) for b in range(len(self._buckets)) if not self._launched[b])
        runs: List[list] = []
        for lo, hi, b in left:
            self._pending[b] = 0
            if runs and runs[-1][1] == lo:
                runs[-1][1] = hi
                runs[-1][2].append(b)
            else:
                runs.append([lo, hi, [b]])
        for lo, hi, members in runs:
            self._launch_range(lo, hi, members, early=False)
        if self._comm_stream is not None:
            torch.cuda.current_stream(self.device).wait_stream(self._comm_stream)

    def _rank_sync(self):
        """Host-level rendezvous of the ranks (one-off places only: warm-up, end of capture).  The peer-memory all-reduce
        spins on flags in device memory with a bounded wait, so ranks must not reach it tens of seconds apart."""
        if self.world > 1 and self._p2p is not None:
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.pg)

    # ------------------------------------------------------------------ one step
    def _loss(self, inputs, static) -> torch.Tensor:
        out = self.loss_fn(*inputs, **static)
        return out[0] if isinstance(out, (tuple, list)) else out

    def _fwd_bwd(self, inputs, static, overlap: bool) -> torch.Tensor:
        self.arena.g.zero_()
        self.step_dev.add_(1)
        # (a bucket's collective waits for the second stream too, see _launch_allreduce: weight-gradient kernels into it may
        # still be queued there)
        self._hooks_live = overlap
        if overlap:
            self._arm_buckets()
        # dgrad operands (transposed + rotated weights) of every conv, re-packed from the bf16 shadow the optimiser wrote at
        # the end of the previous step.  Only backward reads them, so the 0.15 ms launch runs on a second stream under
        # the forward pass instead of on the critical path behind the optimiser.
        pack_stream = None
        if self._dgrad_stale:
            if self.device.type == "cuda":
                if self._pack_stream is None:
                    self._pack_stream = torch.cuda.Stream(device=self.device)
                pack_stream = self._pack_stream
                pack_stream.wait_stream(torch.cuda.current_stream(self.device))
                with torch.cuda.stream(pack_stream):
                    self.packed.refresh_dgrad()
            else:
                self.packed.refresh_dgrad()
            self._dgrad_stale = False
        with _Tf32Matmul():
            loss = self._loss(inputs, static)
            if pack_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(pack_stream)
            loss.backward()
        self._hooks_live = False
        ops.join_side_stream()          # weight gradients ran on a second stream (ops._Side)
        self._overlapped = overlap
        return loss.detach()

    def _reduce_and_update(self, overlapped: bool):
        if self.world > 1:
            if overlapped:
                self._finish_allreduce()
            else:                      # one all-reduce of the whole arena after backward
                self._allreduce_range(0, self.arena.g.numel())
        self._update()

    def _update(self):
        if self.grad_clip > 0:
            self.sumsq.zero_()
            ops.sumsq_(self.arena.g, self.sumsq)
        # gradients hold the SUM over ranks: the mean (what DataParallel / DDP produce) is a grad_scale of 1/world
        ops.adam_ema_step_(self.arena.p, self.arena.g, self.m, self.v, self.ema,
                           self.sumsq if self.grad_clip > 0 else None, self.grad_clip,
                           1.0 / self.world, self.lr, self.betas[0], self.betas[1], self.eps,
                           self.ema_decay if self.ema_decay is not None else 0.0, 1,
                           self.warmup, self.step_dev, self.packed.shadow, self.weight_decay)
        self._dgrad_stale = True                    # re-packed under the next forward (see _fwd_bwd)
        ops.advance_dropout_state(self.device)

    def _body(self, *inputs, **static) -> torch.Tensor:
        """Eager step; with several ranks the bucketed all-reduce overlaps backward (when no second stream is in use)."""
        self._silent = frozenset()      # no probe pass in eager mode: untouched parameters are flushed after backward
        loss = self._fwd_bwd(inputs, static, self.world > 1 and self.overlap)
        self._reduce_and_update(self._overlapped)
        return loss

    def __call__(self, *inputs, **static) -> torch.Tensor:
        self.steps_done += 1
        if not self.use_graph:
            return self._body(*inputs, **static)
        key = (tuple((tuple(t.shape), t.dtype) for t in inputs), tuple(sorted(static.items())))
        entry = self._graphs.get(key)
        if entry is None:
            entry = self._capture(key, inputs, static)
        graph, static_in, static_loss = entry
        for dst, src in zip(static_in, inputs):
            dst.copy_(src, non_blocking=True)
        graph.replay()
        if self.world > 1 and not self._graph_has_tail:   # the collective and the optimiser tail follow the replay (3 launches)
            self._reduce_and_update(False)
        return static_loss

    def timed_phases(self, *inputs, reps: int = 5, **static):
        """Device time (ms, averaged) of the three phases of a data-parallel step: forward+backward graph, gradient
        all-reduce, clip+Adam+EMA tail.  Diagnostic only (bench.py reports it next to the step time)."""
        assert self.use_graph and self.world > 1 and self._graphs
        graph, static_in, _ = next(iter(self._graphs.values()))
        ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(reps)]
        for r in range(reps):
            for dst, src in zip(static_in, inputs):
                dst.copy_(src, non_blocking=True)
            ev[r][0].record()
            graph.replay()
            ev[r][1].record()
            self._allreduce_range(0, self.arena.g.numel())
            ev[r][2].record()
            self._update()
            ev[r][3].record()
        torch.cuda.synchronize(self.device)
        return [sum(e[i].elapsed_time(e[i + 1]) for e in ev) / reps for i in range(3)]

    def _training_state(self):
        state = [self.arena.p, self.m, self.v, self.step_dev, self.packed.shadow, self.packed.dgrad,
                 ops.dropout_device_counter(self.device)]
        return state + ([self.ema] if self.ema is not None else [])

    def _capture(self, key, inputs, static):
        """One CUDA graph for the whole step on a single GPU; forward + backward only when data-parallel (NCCL work
        captured into a graph dead-locked on this stack, so the gradient all-reduce is issued right after the replay)."""
        static_in = [t.clone() for t in inputs]
        # data parallel: by default the graph holds forward + backward and the collective + optimiser follow the replay;
        # UB200_DP_GRAPH_NCCL=1 captures the bucketed NCCL all-reduces (on the comm stream, forked from and joined to the
        # capture stream, so they overlap the rest of backward) and the optimiser tail into the same graph.  The capture
        # then runs in thread-local error mode: NCCL's watchdog thread polls CUDA events, which a global-mode capture forbids.
        # With the peer-memory all-reduce (self._p2p) the collectives are ordinary kernel launches: always captured.
        nccl_in_graph = self.world > 1 and (os.environ.get("UB200_DP_GRAPH_NCCL", "0") == "1" or
                                            (self._p2p is not None and os.environ.get("UB200_DP_GRAPH_P2P", "1") != "0"))
        whole = self.world == 1 or nccl_in_graph
        # The warm-up runs real steps (allocator, tensor maps, cuBLAS handles): snapshot every piece of training state
        # they touch and put it back, so that the first replay is step 1 of the run exactly as in eager mode.
        state = self._training_state()
        saved = [t.clone() for t in state]
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):           # warm-up on a side stream
            for it in range(3):
                # the first warm-up step doubles as a probe: which parameters report a gradient at all with these static
                # arguments (coarse tails without the multi-resolution loss never do).  The silent ones must not hold their
                # bucket back until the end of backward.
                self._probe = nccl_in_graph and self.overlap and it == 0
                self._fwd_bwd(static_in, static, self._probe)
                if self._probe:
                    self._silent = frozenset(i for i, seen in enumerate(self._seen) if not seen)
                    self._probe = False
                self._rank_sync()       # first-call initialisation skews the ranks by seconds; the peer kernel's waits are bounded
                self._reduce_and_update(False)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        for t, keep in zip(state, saved):
            t.copy_(keep)
        del saved
        torch.cuda.synchronize(self.device)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, capture_error_mode="thread_local" if self.world > 1 else "global"):
            static_loss = self._fwd_bwd(static_in, static, nccl_in_graph and self.overlap)
            if whole:
                self._reduce_and_update(self._overlapped)
        self._graph_has_tail = whole
        self._rank_sync()               # capture time differs per rank: line the ranks up before the first replay
        entry = (graph, static_in, static_loss)
        self._graphs[key] = entry
        return entry

    # ------------------------------------------------------------------ checkpoint surface
    def _named(self):
        names = {id(p): n for n, p in self.model.named_parameters()}
        return [(names[id(p)], p, off) for p, off in zip(self.arena.params, self.arena.offsets)]

    def ema_state_dict(self):
        """`state_dict` of the EMA model (the reference keeps a deep copy, main.py:207-219)."""
        assert self.ema is not None, "this step keeps no EMA (ema_decay=None)"
        sd = {k: v.clone() for k, v in self.model.state_dict().items()}
        for name, p, off in self._named():
            sd[name] = FlatArena._view(self.ema, p, off).clone()
        return sd

    def load_ema_state_dict(self, sd):
        """Restore the EMA weights from a model-shaped state_dict (`ema_model.load_state_dict` at main.py:221)."""
        assert self.ema is not None
        for name, p, off in self._named():
            FlatArena._view(self.ema, p, off).copy_(sd[name])

    def state_dict(self):
        """Everything a resumed run needs beyond the model weights: Adam moments per parameter name (the layout of
        `torch.optim.Adam.state_dict()['state']`: exp_avg / exp_avg_sq), the step counter (bias correction, LambdaLR
        warm-up), the EMA weights and the dropout RNG counters.  diff_cifar/main.py:443-453 saves net_model,
        ema_model, sched, optim and step."""
        opt = {name: {"exp_avg": FlatArena._view(self.m, p, off).clone(), "exp_avg_sq": FlatArena._view(self.v, p, off).clone()}
               for name, p, off in self._named()}
        return {"model": {k: v.clone() for k, v in self.model.state_dict().items()},
                "ema": self.ema_state_dict() if self.ema is not None else None,
                "optim": opt, "step": int(self.step_dev), "steps_done": self.steps_done,
                "dropout": {"seed": ops._DropoutState.seed, "host_offset": ops._DropoutState.host_offset,
                            "dev_offset": int(ops.dropout_device_counter(self.device))}}

    def load_state_dict(self, sd):
        self.steps_done = int(sd["steps_done"])          # before the model load: a resumed EMA is not re-initialised
        self.model.load_state_dict(sd["model"])          # post-hook refreshes the bf16 operands
        for name, p, off in self._named():
            FlatArena._view(self.m, p, off).copy_(sd["optim"][name]["exp_avg"])
            FlatArena._view(self.v, p, off).copy_(sd["optim"][name]["exp_avg_sq"])
        if self.ema is not None and sd.get("ema") is not None:
            self.load_ema_state_dict(sd["ema"])
        self.step_dev.fill_(int(sd["step"]))
        d = sd["dropout"]
        ops._DropoutState.seed, ops._DropoutState.host_offset = int(d["seed"]), int(d["host_offset"])
        ops.dropout_device_counter(self.device).fill_(int(d["dev_offset"]))
        self.packed.refresh_from_master()
        self._dgrad_stale = False


class DDPMTrainStep(TrainStep):
    """The diff_cifar training step (main.py:397-429): DDPM Algorithm 1 loss + clip 1.0 + Adam + LambdaLR warm-up + EMA.
    `step(x0, n_levels_used=..., n_downsample=...)` serves the staged (sequential-resolution) loop of main.py:397-423:
    one CUDA graph per (batch shape, level count)."""

    def __init__(self, model: nn.Module, T: int = 1000, beta_1: float = 1e-4, beta_T: float = 0.02, lr: float = 2e-4,
                 warmup: int = 5000, grad_clip: float = 1.0, ema_decay: float = 0.9999, multi_res_loss: bool = False,
                 betas=(0.9, 0.999), eps: float = 1e-8, use_cuda_graph: bool = True, process_group=None,
                 bucket_mb: float = 16.0, overlap_allreduce: bool = True, sequ_train_algo: bool = False):
        dev = next(model.parameters()).device
        self.trainer = GaussianDiffusionTrainer(model, beta_1, beta_T, T, multi_res_loss, sequ_train_algo, dev).to(dev)
        super().__init__(model, self._ddpm_loss, lr=lr, betas=betas, eps=eps, warmup=warmup, grad_clip=grad_clip,
                         ema_decay=ema_decay, use_cuda_graph=use_cuda_graph, process_group=process_group,
                         bucket_mb=bucket_mb, overlap_allreduce=overlap_allreduce)

    def _ddpm_loss(self, x0, n_levels_used: int = -1, n_downsample: int = 0):
        return self.trainer(x0, n_levels_used, n_downsample)

    def step_from_host(self, x0_pinned: torch.Tensor, **static) -> float:
        """The end-to-end call a user makes: batch in pinned host memory in, loss (a Python float) out."""
        x0 = x0_pinned.to(self.device, non_blocking=True)
        loss = self(x0, **static)
        return float(loss.cpu())
