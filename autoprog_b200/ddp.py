"""Data-parallel gradient reducer: bucketed all-reduce over NCCL, overlapped with backward.

Replaces apex `DistributedDataParallel(delay_allreduce=True)` / torch DDP (main_prog.py:538-549, 1413-1424).
One process per GPU; gradients live in the flat buffers of `FlatState`, buckets are contiguous slices of them, so a
bucket's all-reduce starts (async, on NCCL's stream) the moment its last gradient has been accumulated and there are
no flatten/unflatten copies.  Parameters that received no gradient in a step (elastic depth: identity layers,
models/volo.py:141, 231) keep zeros in their slice on every rank -- all ranks sample the same sub-net
(main_prog.py:1861), so the reduction stays consistent, which is the property apex's delay_allreduce provided.
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import List, Optional

import torch
import torch.distributed as dist
import torch.nn as nn

from .flat import FlatState


class _Bucket:
    __slots__ = ('flat', 'n_params', 'ready', 'work')

    def __init__(self, flat, n_params):
        self.flat, self.n_params, self.ready, self.work = flat, n_params, 0, None


class DistributedDataParallel(nn.Module):
    def __init__(self, module: nn.Module, flat: Optional[FlatState] = None, bucket_mb: float = 32.0, process_group=None,
                 weight_decay: float = 0.0, broadcast: bool = True):
        super().__init__()
        self.module = module
        self.flat = flat or FlatState(module, weight_decay)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._sync = True
        self._armed = False
        self.buckets: List[_Bucket] = []
        cap = int(bucket_mb * 1024 * 1024 / 4)
        self._bucket_of = {}
        for g in self.flat.groups:
            start, count = 0, 0
            for i, (p, o) in enumerate(zip(g.params, g.offsets)):
                end = g.offsets[i + 1] if i + 1 < len(g.params) else g.numel
                count += 1
                self._bucket_of[p] = len(self.buckets)
                if end - start >= cap or i + 1 == len(g.params):
                    self.buckets.append(_Bucket(g.flat_g[start:end], count))
                    start, count = end, 0
        if broadcast and self.world > 1:
            for g in self.flat.groups:
                dist.broadcast(g.flat_p, 0, group=self.pg)
                if g.shadow is not None:
                    g.shadow.copy_(g.flat_p)
            for b in module.buffers():
                dist.broadcast(b, 0, group=self.pg)
        for p in self._bucket_of:
            p.register_post_accumulate_grad_hook(self._hook)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    @contextmanager
    def no_sync(self):
        """Gradient accumulation (`update=False` steps of prog/scaler.py): skip the all-reduce."""
        old, self._sync = self._sync, False
        try:
            yield
        finally:
            self._sync = old

    # ---- autograd hooks -------------------------------------------------------------------
    def _hook(self, p):
        if not self._sync or self.world == 1:
            return
        if not self._armed:
            self._armed = True
            torch.autograd.Variable._execution_engine.queue_callback(self._finalize)
        # a gradient that autograd allocated itself (not written in place by a kernel, see ops.grad_dest) must be in the
        # flat buffer BEFORE its bucket's all-reduce is launched
        view = getattr(p, '_apb_grad_view', None)
        if view is not None and p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
            p.grad = view
        b = self.buckets[self._bucket_of[p]]
        b.ready += 1
        if b.ready == b.n_params:
            self._launch(b)

    def _launch(self, b: _Bucket):
        if dist.get_backend(self.pg) == 'nccl':
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.AVG, group=self.pg, async_op=True)
        else:
            b.work = dist.all_reduce(b.flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)

    def _finalize(self):
        self.flat.ensure_grad_views()
        for b in self.buckets:                      # buckets holding parameters without gradient this step
            if b.work is None:
                self._launch(b)
        avg_needed = dist.get_backend(self.pg) != 'nccl'
        for b in self.buckets:
            b.work.wait()
            if avg_needed:
                b.flat.div_(self.world)
            b.work, b.ready = None, 0
        self._armed = False

    def reduce_now(self):
        """Synchronous reduction of all buckets (for callers that run backward under `no_sync`)."""
        if self.world == 1:
            return
        self.flat.ensure_grad_views()               # gradients autograd allocated itself -> flat buffer, before reducing
        for b in self.buckets:
            self._launch(b)
        self._finalize()
