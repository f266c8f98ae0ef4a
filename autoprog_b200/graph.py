"""Whole-step CUDA-graph capture of the training hot path.

The reference launches ~1300 kernels per step from Python and synchronises every step (main_prog.py:1035); at the
early AutoProg stages (128-160 px, depth 9-12) the GPU work per step is shorter than the host's launch time.  A
`GraphedTrainStep` captures zero-grad + forward + loss + backward + fused optimizer/EMA step ONCE per (resolution,
depth, batch) configuration and replays it.  Everything that changes between steps is read from DEVICE memory at
replay time: the mix-token box (host RNG -> pinned ring slot -> device int32[4], read by the flip and loss kernels),
the optimizer's lr / bias corrections (same staging), DropPath masks (CUDA RNG is graph-safe), inputs (static
buffers).  The tiny host -> device copies are issued eagerly ahead of each replay from a ring of pinned slots
(optim.PinnedRing), so the host may queue several steps ahead without overwriting a queued step's values.

`GraphCache` keeps one `GraphedTrainStep` per (input size, layer_num, batch) over ONE model / optimizer / flat state:
the super-net epochs of the AutoProg search draw a random (r, l) every step (main_prog.py:1824-1836, 1907-1910), which
a single captured configuration cannot serve.  `probe_throughput` is the forward+backward timing of
`validate_trainset(test_throughput=True)` (main_prog.py:1245-1298) that feeds the search objective.
"""
from __future__ import annotations

import random
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from .optim import PinnedRing
from .volo import rand_bbox


def _stage1_stride(net) -> int:
    """Input pixels per stage-1 token: stem stride x patch-projection stride (models/volo.py:355-373)."""
    pe = net.patch_embed
    s = pe.proj.stride[0]
    if getattr(pe, 'stem_conv', False):
        s *= pe.conv[0].stride[0]
    return int(s)


class _StateSnapshot:
    """Everything a training step mutates, so that the warm-up steps of a capture leave no trace: parameters, optimizer
    moments / step counter, EMA copies, BatchNorm buffers, bf16 shadows and the host / device RNG streams."""

    def __init__(self, net, optimizer):
        self.opt = optimizer
        fl = optimizer.flat
        self.flat = [(g.flat_p, g.flat_p.clone()) for g in fl.groups]
        self.flat += [(g.shadow, g.shadow.clone()) for g in fl.groups if g.shadow is not None]
        self.flat += [(t, t.clone()) for t in list(optimizer.exp_avg) + list(optimizer.exp_avg_sq)]
        self.flat += [(t, t.clone()) for ef in optimizer.ema_flats for t in ef]
        mods = [net] + list(optimizer.ema_models)
        self.flat += [(b, b.clone()) for m in mods for b in m.buffers()]
        self.step_count = optimizer.step_count
        self.np_state = np.random.get_state()
        self.cuda_rng = torch.cuda.get_rng_state()

    @torch.no_grad()
    def restore(self):
        for dst, src in self.flat:
            dst.copy_(src)
        self.opt.step_count = self.step_count
        np.random.set_state(self.np_state)
        torch.cuda.set_rng_state(self.cuda_rng)
        ops.invalidate_derived_caches()


class GraphedTrainStep:
    """One graph for the whole step -- also under `DistributedDataParallel` (autoprog_b200.ddp), where the bucketed
    NCCL all-reduces are captured inside the graph as a branch that overlaps the remaining backward (`_capture_ddp`).

    `restore_after_warmup=True` (default) makes construction free of side effects: the warm-up steps that precede the
    capture run on the real model / optimizer and are rolled back afterwards (one construction per
    (resolution, depth, batch) would otherwise advance the weights by `warmup` optimizer steps each time)."""

    def __init__(self, model, criterion, optimizer, example_input, example_target, bf16: bool = True, warmup: int = 3,
                 sample_config: Optional[dict] = None, pool=None, restore_after_warmup: bool = True,
                 ddp_mode: str = 'overlap'):
        self.model, self.criterion, self.optimizer, self.bf16 = model, criterion, optimizer, bf16
        self.net = model.module if hasattr(model, 'module') else model
        self.ddp = model if (hasattr(model, 'reduce_now') and getattr(model, 'world', 1) > 1) else None
        self.sample_config = dict(sample_config) if sample_config else None
        assert ddp_mode in ('overlap', 'split'), ddp_mode
        self.ddp_mode = ddp_mode
        dev = example_input.device
        self.x = example_input.clone()
        self.t = example_target.clone()
        self.loss = torch.zeros((), device=dev, dtype=torch.float32)      # outside the graph pool: survives other graphs
        self.box = PinnedRing((4,), torch.int32, dev)          # mix-token box: host RNG -> pinned slot -> device
        self.box_dev = self.box.dev
        self._activate()
        snap = _StateSnapshot(self.net, optimizer) if restore_after_warmup else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._host_prepare()
                self._stage()
                self._device_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        if snap is not None:
            snap.restore()
            torch.cuda.synchronize()
        # capture only records: no host state (RNG, optimizer step count) is advanced for it
        from . import kernels as K
        n0 = K.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        pool_arg = {} if pool is None else {'pool': pool}
        if self.ddp is None:
            with torch.cuda.graph(self.graph, **pool_arg):
                self.loss.copy_(self._device_step())
        else:
            self._capture_ddp(pool_arg)
        self.kernels_per_step = K.launch_count() - n0      # launches of this library captured in the graph(s)

    # ------------------------------------------------------------------------------------------------------------
    def _activate(self):
        """Point the model at this step's device-resident mix-token box and apply its elastic sub-net flags."""
        self.net._graph_box, self.net._graph_box_host = self.box_dev, self.box.current()
        if self.sample_config is not None:
            self.net.set_sample_config(self.sample_config)

    def _host_prepare(self):
        net = self.net
        if getattr(net, 'mix_token', False) and net.training:
            lam = np.random.beta(net.beta, net.beta)
            s = net.pooling_scale
            g = self.x.shape[-1] // _stage1_stride(net)        # stage-1 token grid
            box = rand_bbox((self.x.shape[0], g, g, 0), lam, scale=s)
            slot = self.box.next_slot()
            slot.copy_(torch.tensor([int(v) for v in box], dtype=torch.int32))
            self.net._graph_box_host = slot
        self.optimizer.prepare_step()

    def _stage(self):
        """Eager host -> device copies of the per-step scalars, queued ahead of the replay that reads them."""
        self.box.push()
        self.optimizer.stage_hyper()

    def _fwd_bwd(self):
        self.optimizer.zero_grad()
        with ops.autocast(enabled=self.bf16):
            out = self.model(self.x)
            loss = self.criterion(out, self.t)
        if self.ddp is not None:
            with self.ddp.no_sync():          # gradients are reduced outside the captured region
                loss.backward()
        else:
            loss.backward()
        # gradients that autograd allocated itself are copied into the flat buffer inside THIS captured segment: the
        # eager all-reduce between the graphs must already see them
        self.optimizer.flat.ensure_grad_views()
        return loss.detach()

    def _opt_step(self):
        self.optimizer.launch_step(stage=False)
        self.optimizer.update_ema_buffers()

    def _device_step(self):
        loss = self._fwd_bwd()
        if self.ddp is not None:
            self.ddp.reduce_now()
        self._opt_step()
        return loss

    # ------------------------------------------------------------------------------------------------------------
    def _capture_ddp(self, pool_arg):
        """Data-parallel capture.

        ddp_mode 'overlap' (default): ONE graph for the whole step with the bucketed NCCL all-reduces captured inside
        it.  The reducer's autograd hooks fire during the captured backward exactly as in eager mode: the moment a
        bucket's last gradient has been written, `all_reduce(async_op=True)` is recorded on NCCL's stream -- a parallel
        branch of the graph that joins again (work.wait()) in front of the optimizer kernels.  Bucket k therefore
        reduces while the backward kernels of the earlier layers run, which is what north_star asks for and what
        apex's `delay_allreduce=True` (main_prog.py:543) did not do.
        ddp_mode 'split': [zero-grad, forward, loss, backward] | eager all-reduce of all buckets | [optimizer + EMA]
        (the all-reduce is exposed); kept as the fallback for stacks whose NCCL cannot be captured."""
        if self.ddp_mode == 'overlap':
            with torch.cuda.graph(self.graph, **pool_arg):
                self.optimizer.zero_grad()
                with ops.autocast(enabled=self.bf16):
                    out = self.model(self.x)
                    loss = self.criterion(out, self.t)
                loss.backward()                      # hooks launch the bucket all-reduces; the engine callback joins them
                self.loss.copy_(loss.detach())
                self._opt_step()
            return
        with torch.cuda.graph(self.graph, **pool_arg):
            self.loss.copy_(self._fwd_bwd())
        self.graph_opt = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
            self._opt_step()

    def __call__(self, x=None, target=None):
        if self.net.__dict__.get('_graph_box') is not self.box_dev or self.sample_config is not None:
            self._activate()
        if x is not None:
            self.x.copy_(x, non_blocking=True)
        if target is not None:
            self.t.copy_(target, non_blocking=True)
        self._host_prepare()
        self._stage()
        self.graph.replay()
        if self.graph_opt is not None:
            self.ddp.reduce_now()
            self.graph_opt.replay()
        # the replayed optimizer kernel rewrote parameters and EMA copies through raw pointers (no `_version` bump):
        # derived copies cached by an eager forward between replays (EMA / model evaluation) are stale now
        ops.invalidate_derived_caches()
        return self.loss

    def close(self):
        if self.net.__dict__.get('_graph_box') is self.box_dev:
            self.net._graph_box = None
        ops.invalidate_derived_caches()


class GraphCache:
    """(input size, layer_num, batch) -> GraphedTrainStep over one model / optimizer.

    `step(x224, target_by_r, config)` resizes the batch to `config['input_size']` with the bilinear kernel
    (main_prog.py:1910), applies `set_sample_config(config)` (main_prog.py:1908) and replays -- capturing first if this
    configuration has not been seen.  All graphs share one memory pool: only one of them runs at a time and nothing in
    the pool is read after the replay returns (the loss is copied to a tensor outside it)."""

    def __init__(self, model, criterion, optimizer, bf16: bool = True, warmup: int = 2):
        self.model, self.criterion, self.optimizer, self.bf16, self.warmup = model, criterion, optimizer, bf16, warmup
        self.steps: Dict[Tuple[int, int, int], GraphedTrainStep] = {}
        self.pool = None
        self.captures = 0

    def get(self, x: torch.Tensor, target: torch.Tensor, config: dict) -> GraphedTrainStep:
        key = (int(x.shape[-1]), int(config['layer_num']), int(x.shape[0]))
        gs = self.steps.get(key)
        if gs is None:
            gs = GraphedTrainStep(self.model, self.criterion, self.optimizer, x, target, bf16=self.bf16, warmup=self.warmup,
                                  sample_config=config, pool=self.pool)
            if self.pool is None:
                self.pool = gs.graph.pool()
            self.steps[key] = gs
            self.captures += 1
        return gs

    def step(self, x: torch.Tensor, target: torch.Tensor, config: dict) -> torch.Tensor:
        from .progressive import resize_input
        r = int(config['input_size'])
        if x.shape[-1] != r:
            x = resize_input(x, r)
        return self.get(x, target, config)(x, target)

    def close(self):
        for gs in self.steps.values():
            gs.close()
        self.steps.clear()


def sample_configs(l_list: Sequence[int], r_list: Sequence[int], mode: str = 'random'):
    """main_prog.py:1824-1836 (python `random`, seeded with the epoch by the caller: every rank draws the same sub-net)."""
    if mode == 'random':
        config = {'min_layer_num': l_list[0], 'max_layer_num': l_list[-1], 'layer_num': random.choice(l_list),
                  'input_size': random.choice(r_list)}
    elif mode == 'smallest':
        config = {'min_layer_num': l_list[0], 'max_layer_num': l_list[-1], 'layer_num': l_list[0], 'input_size': r_list[0]}
    else:
        raise NotImplementedError(mode)
    config['token_label_size'] = config['input_size'] // 16
    return config, list(l_list).index(config['layer_num']), list(r_list).index(config['input_size'])


def probe_throughput(model, loss_fn, x: torch.Tensor, target: torch.Tensor, config: dict, steps: int = 50, warmup: int = 3,
                     bf16: bool = True) -> float:
    """Seconds per forward+backward of the sub-net `config` selects on a batch shaped like `x` -- the quantity
    `validate_trainset(test_throughput=True)` measures for the AutoProg objective (main_prog.py:1245-1298: train-mode
    forward, `loss_fn(output[0], target)`, backward, gradients discarded).  Timed on the device with CUDA events over a
    captured graph (no per-step host synchronisation); model / optimizer state is left untouched: gradients go to
    scratch `.grad`s that are dropped afterwards."""
    from .progressive import resize_input
    net = model.module if hasattr(model, 'module') else model
    net.set_sample_config(config)
    r = int(config['input_size'])
    if x.shape[-1] != r:
        x = resize_input(x, r)
    was_training = net.training
    net.train()
    saved = [(p, p.grad) for p in net.parameters()]
    np_state = np.random.get_state()

    def fwd_bwd():
        for p, _ in saved:
            p.grad = None
            p._apb_grad_claimed = False          # see ops.grad_dest: the flat view may be handed out again
        with ops.autocast(enabled=bf16):
            out = model(x)
            loss = loss_fn(out[0] if isinstance(out, tuple) else out, target)
        loss.backward()

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fwd_bwd()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fwd_bwd()
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(steps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    del g
    for p, gsaved in saved:
        p.grad = gsaved
    np.random.set_state(np_state)
    net.train(was_training)
    ops.invalidate_derived_caches()
    return e0.elapsed_time(e1) / steps / 1e3
