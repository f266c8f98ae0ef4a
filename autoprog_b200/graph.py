"""Whole-step CUDA-graph capture of the training hot path.

The reference launches ~1300 kernels per step from Python and synchronises every step (main_prog.py:1035); at the
early AutoProg stages (128-160 px, depth 9-12) the GPU work per step is shorter than the host's launch time.  A
`GraphedTrainStep` captures zero-grad + forward + loss + backward + fused optimizer/EMA step ONCE per (resolution,
depth, batch) configuration and replays it.  Everything that changes between steps is read from DEVICE memory at
replay time: the mix-token box (host RNG -> pinned ring slot -> device int32[4], read by the flip and loss kernels),
the optimizer's lr / bias corrections (same staging), DropPath masks (CUDA RNG is graph-safe), inputs (static
buffers).  The tiny host -> device copies are issued eagerly ahead of each replay from a ring of pinned slots
(optim.PinnedRing), so the host may queue several steps ahead without overwriting a queued step's values.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops
from .optim import PinnedRing
from .volo import rand_bbox


class GraphedTrainStep:
    """Single GPU: one graph for the whole step.  Under `DistributedDataParallel` (autoprog_b200.ddp) the step is two
    graphs -- [zero-grad, forward, loss, backward] and [optimizer + EMA] -- with the bucketed NCCL all-reduce of the
    flat gradient buffers issued eagerly in between (NCCL collectives are kept out of the captured region)."""

    def __init__(self, model, criterion, optimizer, example_input, example_target, bf16: bool = True, warmup: int = 3):
        self.model, self.criterion, self.optimizer, self.bf16 = model, criterion, optimizer, bf16
        self.net = model.module if hasattr(model, 'module') else model
        self.ddp = model if (hasattr(model, 'reduce_now') and getattr(model, 'world', 1) > 1) else None
        dev = example_input.device
        self.x = example_input.clone()
        self.t = example_target.clone()
        self.box = PinnedRing((4,), torch.int32, dev)          # mix-token box: host RNG -> pinned slot -> device
        self.box_dev = self.box.dev
        self.net._graph_box, self.net._graph_box_host = self.box_dev, self.box.current()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._host_prepare()
                self._stage()
                self._device_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # capture only records: no host state (RNG, optimizer step count) is advanced for it
        from . import kernels as K
        n0 = K.launch_count()
        self.graph = torch.cuda.CUDAGraph()
        self.graph_opt = None
        if self.ddp is None:
            with torch.cuda.graph(self.graph):
                self.loss = self._device_step()
        else:
            with torch.cuda.graph(self.graph):
                self.loss = self._fwd_bwd()
            self.graph_opt = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph_opt, pool=self.graph.pool()):
                self._opt_step()
        self.kernels_per_step = K.launch_count() - n0      # launches of this library captured in the graph(s)

    def _host_prepare(self):
        net = self.net
        if getattr(net, 'mix_token', False) and net.training:
            lam = np.random.beta(net.beta, net.beta)
            s = net.pooling_scale
            g = self.x.shape[-1] // 8                      # stage-1 token grid
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
        # gradients that autograd allocated itself (e.g. the cuDNN stem convolutions) are copied into the flat buffer
        # inside THIS captured segment: the eager all-reduce between the two graphs must already see them
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

    def __call__(self, x=None, target=None):
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
        return self.loss

    def close(self):
        self.net._graph_box = None
