"""CPU checks of the training-step host logic (flat arena, fused optimiser tail contract, data-parallel
buckets) with the kernels replaced by the test-only emulation, against the torch fp32 oracle model trained
with torch.optim.Adam + clip_grad_norm_ + EMA as diff_cifar/main.py:424-429 does."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import rel_err

CFG = dict(T=20, ch=32, ch_mult=[1, 1], attn=[1], num_res_blocks=1, dropout=0.0, dwt_encoder=True)


def _oracle_steps(state, x0s, seeds, lr=1e-3, warmup=2):
    from oracle import torch_ref
    net = torch_ref.UNetWaveletEnc(**CFG)
    net.load_state_dict(state)
    trainer = torch_ref.GaussianDiffusionTrainer(net, 1e-4, 0.02, CFG["T"])
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.Adam(params, lr=lr)
    sched = torch.optim.lr_scheduler.LambdaLR(opt, lambda s: min(s, warmup) / warmup)
    ema = [p.detach().clone() for p in params]
    losses = []
    for x0, seed in zip(x0s, seeds):
        opt.zero_grad()
        torch.manual_seed(seed)
        loss, _ = trainer(x0)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step(); sched.step()
        for e, p in zip(ema, params):
            e.mul_(0.99).add_(p.detach(), alpha=0.01)
        losses.append(float(loss.detach()))
    return net, ema, losses


def _make(state):
    from unet_design_b200.diff_cifar.model import UNetWaveletEnc
    from unet_design_b200.train import DDPMTrainStep
    net = UNetWaveletEnc(**CFG)
    net.load_state_dict(state)
    return net, DDPMTrainStep(net, T=CFG["T"], lr=1e-3, warmup=2, grad_clip=1.0, ema_decay=0.99, use_cuda_graph=False)


def _init_state():
    from oracle import torch_ref
    torch.manual_seed(1234)
    net = torch_ref.UNetWaveletEnc(**CFG)
    with torch.no_grad():
        for _, p in net.named_parameters():
            if p.dim() == 4 and p.abs().max() < 1e-3:
                p.mul_(3e4)
    return net.state_dict()


def test_flat_arena_views_and_three_steps_match_oracle(emulated_ops):
    state = _init_state()
    torch.manual_seed(0)
    x0s = [torch.randn(4, 3, 16, 16) for _ in range(3)]
    seeds = [11, 12, 13]
    ref, ref_ema, ref_losses = _oracle_steps(state, x0s, seeds)
    net, step = _make(state)
    # every trainable parameter and its gradient are views of the arenas; conv weights stay channels_last
    for p in step.arena.params:
        assert p.data.untyped_storage().data_ptr() == step.arena.p.untyped_storage().data_ptr()
        assert p.grad.untyped_storage().data_ptr() == step.arena.g.untyped_storage().data_ptr()
        if p.dim() == 4:
            assert p.permute(0, 2, 3, 1).is_contiguous()
    losses = []
    for x0, seed in zip(x0s, seeds):
        torch.manual_seed(seed)
        losses.append(float(step(x0)))
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 3e-2 * abs(b)
    rp = dict(ref.named_parameters())

    def close(a, b):
        # Adam turns gradients that are zero in exact arithmetic (time-embedding path through a GroupNorm with one
        # channel per group, softmax-invariant key bias) into lr-sized steps of round-off sign: allow 1.5 lr RMS.
        rms = float((a.detach() - b.detach()).pow(2).mean().sqrt())
        return rel_err(a, b) < 2e-3 or rms < 1.5e-3

    bad = [n for n, p in net.named_parameters() if p.requires_grad and not close(p, rp[n])]
    assert not bad, bad
    ema_sd = step.ema_state_dict()
    names = [n for n, p in ref.named_parameters() if p.requires_grad]
    assert all(close(ema_sd[n], e) for n, e in zip(names, ref_ema))
    assert int(step.step_dev) == 3 and set(ema_sd) == set(net.state_dict())


def _dp_worker(rank, world, port, state, x0, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import unet_design_b200._lib as lib
    from _emulated_ops import EmulatedOps
    lib._ops = EmulatedOps()
    torch.set_num_threads(1)
    from unet_design_b200.diff_cifar.model import UNetWaveletEnc
    from unet_design_b200.train import DDPMTrainStep
    net = UNetWaveletEnc(**CFG)
    if rank == 0:
        net.load_state_dict(state)          # rank 0's weights are broadcast by DDPMTrainStep
    step = DDPMTrainStep(net, T=CFG["T"], lr=1e-3, warmup=0, grad_clip=1.0, ema_decay=0.99, use_cuda_graph=False,
                         bucket_mb=0.05)
    assert len(step._buckets) > 3
    shard = x0.chunk(world)[rank]
    # same (t, noise) as the single-process run: draw for the full batch, keep this rank's rows
    torch.manual_seed(5)
    t = torch.randint(CFG["T"], size=(x0.shape[0],)).chunk(world)[rank]
    noise = torch.randn_like(x0).chunk(world)[rank]
    step.arena.g.zero_(); step.step_dev.add_(1); step._arm_buckets(); step._hooks_live = True
    loss, _ = step.trainer.loss_from(shard, t, noise)
    loss.backward()
    step._finish_allreduce()
    g = step.arena.g.clone() / world
    ret[rank] = (g, float(loss.detach()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_buckets_world2_gloo(emulated_ops):
    """world_size 2 over gloo: bucketed all-reduce from post-accumulate hooks reproduces the full-batch gradient."""
    state = _init_state()
    torch.manual_seed(1)
    x0 = torch.randn(4, 3, 16, 16)
    net, step = _make(state)
    torch.manual_seed(5)
    t = torch.randint(CFG["T"], size=(4,))
    noise = torch.randn_like(x0)
    step.arena.g.zero_()
    loss, _ = step.trainer.loss_from(x0, t, noise)
    loss.backward()
    g_full = step.arena.g.clone()
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + os.getpid() % 2000
    mp.spawn(_dp_worker, args=(2, port, state, x0, ret), nprocs=2, join=True)
    g0, l0 = ret[0]
    g1, l1 = ret[1]
    assert torch.equal(g0, g1)                               # both ranks hold the same reduced gradient
    assert abs(0.5 * (l0 + l1) - float(loss)) < 1e-2 * abs(float(loss))
    assert rel_err(g0, g_full) < 3e-2                        # bf16 activations: batch split changes rounding only


def test_arena_lays_out_fused_groups_contiguously(emulated_ops):
    """AttnBlock asks for q|k|v weights (and biases) to be neighbours in the parameter arena: their concatenation is then
    a view (no concat / split kernels).  The reorder must not change any value, and the EMA state_dict keeps the
    reference's per-parameter keys."""
    from unet_design_b200 import ops
    from unet_design_b200.diff_cifar.model import AttnBlock, UNetWaveletEnc
    from unet_design_b200.train import DDPMTrainStep

    torch.manual_seed(0)
    net = UNetWaveletEnc(**CFG)
    before = {k: v.clone() for k, v in net.state_dict().items()}
    step = DDPMTrainStep(net, T=CFG["T"], use_cuda_graph=False)
    after = net.state_dict()
    assert set(before) == set(after) and all(torch.equal(before[k], after[k]) for k in before)
    attn = [m for m in net.modules() if isinstance(m, AttnBlock)]
    assert attn
    off = {id(p): o for p, o in zip(step.arena.params, step.arena.offsets)}
    for a in attn:
        for grp in a.fused_param_groups():
            pos = [off[id(p)] for p in grp]
            assert pos == [pos[0] + i * grp[0].numel() for i in range(3)]          # back to back, in q, k, v order
            flat = step.arena.p[pos[0]:pos[0] + 3 * grp[0].numel()]
            want = torch.cat([p.detach().permute(0, 2, 3, 1).reshape(-1) if p.dim() == 4 else p.detach().reshape(-1) for p in grp])
            assert torch.equal(flat, want)
        assert ops.fused_conv_registered(a.proj_q.weight)
    ema = step.ema_state_dict()
    assert set(ema) == set(before) and all(torch.equal(ema[k], before[k]) for k in before)


def test_checkpoint_round_trip_resumes_identically(emulated_ops):
    """state_dict() / load_state_dict() of the training step (diff_cifar/main.py:443-453 saves net_model, ema_model,
    optim, sched, step): 2 steps + save, then A continues and B is rebuilt from the checkpoint; both take the same 3rd step."""
    state = _init_state()
    torch.manual_seed(0)
    x0s = [torch.randn(4, 3, 16, 16) for _ in range(3)]
    net, step = _make(state)
    for i, x0 in enumerate(x0s[:2]):
        torch.manual_seed(20 + i)
        step(x0)
    ckpt = step.state_dict()
    ckpt = {k: (v.copy() if isinstance(v, dict) else v) for k, v in ckpt.items()}
    torch.manual_seed(22)
    la = float(step(x0s[2]))
    net_b, step_b = _make(_init_state())                 # fresh model + step, then restore
    step_b.load_state_dict(ckpt)
    assert int(step_b.step_dev) == 2 and step_b.steps_done == 2
    torch.manual_seed(22)
    lb = float(step_b(x0s[2]))
    assert la == lb
    assert torch.equal(step.arena.p, step_b.arena.p) and torch.equal(step.m, step_b.m) and torch.equal(step.v, step_b.v)
    assert torch.equal(step.ema, step_b.ema) and torch.equal(step.packed.shadow, step_b.packed.shadow)


def test_model_load_after_step_construction_refreshes_operands(emulated_ops):
    """`net_model.load_state_dict(...)` AFTER the step was built (the reference restores after constructing everything,
    main.py:219-221): the bf16 operands the kernels read follow the new masters, and an untrained EMA follows too."""
    state = _init_state()
    net, step = _make(state)
    other = {k: (v + 0.25 if v.dtype.is_floating_point else v) for k, v in state.items()}
    net.load_state_dict(other)
    assert torch.equal(step.packed.shadow.float(), step.arena.p.to(torch.bfloat16).float())
    assert torch.equal(step.ema, step.arena.p)
    names = {id(p): n for n, p in net.named_parameters()}
    for p in step.arena.params:
        assert torch.equal(p.detach(), other[names[id(p)]])


def test_generic_train_step_adamw_matches_torch(emulated_ops):
    """TrainStep(model, loss_fn) with decoupled weight decay on a pdearena ResidualBlock against torch.optim.AdamW."""
    from unet_design_b200.pdearena.modules.twod_unet import ResidualBlock
    from unet_design_b200.train import TrainStep
    import copy
    torch.manual_seed(3)
    blk = ResidualBlock(16, 32, activation="gelu", norm=True, n_groups=1)
    ref = copy.deepcopy(blk)
    x, y = torch.randn(2, 16, 8, 8), torch.randn(2, 32, 8, 8)

    def loss_fn(xb, yb):
        return torch.nn.functional.mse_loss(blk(xb), yb)

    step = TrainStep(blk, loss_fn, lr=1e-2, weight_decay=0.1, use_cuda_graph=False)
    opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.1)
    for _ in range(3):
        step(x, y)
        opt.zero_grad()
        torch.nn.functional.mse_loss(ref(x), y).backward()
        opt.step()
    for (n, p), (_, q) in zip(blk.named_parameters(), ref.named_parameters()):
        assert rel_err(p, q) < 2e-2, n


def test_bucket_allreduce_waits_for_the_second_stream(emulated_ops, monkeypatch):
    """weight-gradient kernels run on ops._Side.stream, so a bucket's collective must wait for that
    stream as well as for the current one before it reads the gradient arena."""
    import contextlib
    from unet_design_b200 import ops
    from unet_design_b200 import train as tr
    state = _init_state()
    net, step = _make(state)

    class FakeStream:
        def __init__(self):
            self.waited = []

        def wait_stream(self, s):
            self.waited.append(s)

    comm, side, cur = FakeStream(), object(), object()
    step._buckets, step._comm_stream = [(0, 16, (0,))], comm
    monkeypatch.setattr(ops._Side, "stream", side)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: cur)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    reduced = []
    monkeypatch.setattr(tr.dist, "all_reduce", lambda t, op=None, group=None: reduced.append(t.numel()))
    step._launch_allreduce(0)
    assert comm.waited == [cur, side] and reduced == [16]
