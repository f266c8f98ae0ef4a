"""Fused AdamW + multi-EMA optimizer over flat parameter storage.

Semantics = `torch.optim.AdamW` (what timm's `create_optimizer(opt='adamw')` builds, main_prog.py:484) with timm's
no-weight-decay filter, followed by `ModelEmaV2.update` for every attached EMA model (main_prog.py:1032-1033) --
executed as ONE kernel launch per parameter group that also refreshes the bf16 compute copies.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import kernels as K
from . import ops
from .flat import FlatState


class PinnedRing:
    """Host -> device staging of a few per-step scalars (lr / bias corrections, the mix-token box) that kernels read
    from DEVICE memory.  `slots` pinned host buffers feed one device buffer; a slot is rewritten only after the async
    copy that read it has executed, so the host may run several steps ahead of the GPU without corrupting a step that
    is still queued (a single pinned buffer would be overwritten by step i+1 before step i's copy has run)."""

    def __init__(self, shape, dtype, device, slots: int = 8):
        self.host = [torch.zeros(shape, dtype=dtype).pin_memory() for _ in range(slots)]
        self.events: List[Optional[torch.cuda.Event]] = [None] * slots
        self.dev = torch.zeros(shape, dtype=dtype, device=device)
        self.i = 0

    def next_slot(self) -> torch.Tensor:
        """The host buffer to fill for the coming step."""
        self.i = (self.i + 1) % len(self.host)
        if self.events[self.i] is not None:
            self.events[self.i].synchronize()
        return self.host[self.i]

    def current(self) -> torch.Tensor:
        return self.host[self.i]

    def push(self):
        """Enqueue current host slot -> device buffer on the current stream (never inside a graph capture)."""
        self.dev.copy_(self.host[self.i], non_blocking=True)
        if self.events[self.i] is None:
            self.events[self.i] = torch.cuda.Event()
        self.events[self.i].record()


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model: nn.Module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.05,
                 ema_models: Sequence[nn.Module] = (), ema_decays: Sequence[float] = (), flat: Optional[FlatState] = None):
        assert len(ema_models) == len(ema_decays)
        self.flat = flat or FlatState(model, weight_decay)
        groups = [dict(params=g.params, weight_decay=wd) for g, wd in zip(self.flat.groups, self.flat.weight_decays)]
        super().__init__(groups, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.exp_avg = [torch.zeros_like(g.flat_p) for g in self.flat.groups]
        self.exp_avg_sq = [torch.zeros_like(g.flat_p) for g in self.flat.groups]
        self.step_count = 0
        dev = self.flat.groups[0].flat_p.device
        self._hyper = PinnedRing((len(self.flat.groups), 3), torch.float32, dev)   # {lr, 1-b1^t, sqrt(1-b2^t)} per group
        self.ema_models = list(ema_models)
        self.ema_decays = [float(d) for d in ema_decays]
        self.ema_flats = [self.flat.flat_like(e) for e in self.ema_models]     # [ema][group] -> flat fp32
        for g in self.flat.groups:                                            # publish the bf16 copies the kernel keeps fresh
            if g.shadow is not None:
                for p, s in zip(g.params, g.views(g.shadow)):
                    if p.dim() == 2:
                        ops.register_shadow(p, s)

    def zero_grad(self, set_to_none: bool = False):
        self.flat.zero_grad()

    def prepare_step(self):
        """Host part of a step: bump the step counter and stage {lr, bias corrections} in pinned memory."""
        self.step_count += 1
        h = self._hyper.next_slot()
        for gi, group in enumerate(self.param_groups):
            b1, b2 = group['betas']
            h[gi, 0] = group['lr']
            h[gi, 1] = 1.0 - b1 ** self.step_count
            h[gi, 2] = math.sqrt(1.0 - b2 ** self.step_count)

    def stage_hyper(self):
        """Enqueue this step's hyper-parameters host -> device (eager; graph replays read the device copy)."""
        self._hyper.push()

    @torch.no_grad()
    def launch_step(self, stage: bool = True):
        """Device part of a step.  stage=False: CUDA-graph capturable -- the kernels read lr / bias corrections from
        device memory that `stage_hyper()` refreshes eagerly before every replay."""
        self.flat.ensure_grad_views()
        if stage:
            self.stage_hyper()
        for gi, (g, group) in enumerate(zip(self.flat.groups, self.param_groups)):
            b1, b2 = group['betas']
            K.adamw_ema(g.flat_p, g.flat_g, self.exp_avg[gi], self.exp_avg_sq[gi], self._hyper.dev[gi], b1, b2, group['eps'],
                        group['weight_decay'], [ef[gi] for ef in self.ema_flats], self.ema_decays, g.shadow)
        ops.invalidate_derived_caches()

    # ---- checkpointing: torch.optim.AdamW layout (per-parameter `step` / `exp_avg` / `exp_avg_sq`), so what the
    # reference's CheckpointSaver stores (`optimizer.state_dict()`, main_prog.py:600-606) and `resume_checkpoint`
    # reloads interchanges with a stock AdamW built over the same parameter order
    def state_dict(self):
        state, groups, idx = {}, [], 0
        for gi, (g, group) in enumerate(zip(self.flat.groups, self.param_groups)):
            ids = []
            for m, v in zip(g.views(self.exp_avg[gi]), g.views(self.exp_avg_sq[gi])):
                state[idx] = {'step': torch.tensor(float(self.step_count)), 'exp_avg': m.detach().clone(),
                              'exp_avg_sq': v.detach().clone()}
                ids.append(idx)
                idx += 1
            pg = {k: v for k, v in group.items() if k != 'params'}
            for k, v in (('amsgrad', False), ('maximize', False), ('foreach', None), ('capturable', False),
                         ('differentiable', False), ('fused', None), ('decoupled_weight_decay', True)):
                pg.setdefault(k, v)           # the keys torch.optim.AdamW's step() looks up in a loaded group
            pg['params'] = ids
            groups.append(pg)
        return {'state': state, 'param_groups': groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict['param_groups']
        if len(groups) != len(self.param_groups) or any(len(a['params']) != len(b['params'])
                                                         for a, b in zip(groups, self.param_groups)):
            raise ValueError('FusedAdamW.load_state_dict: parameter groups do not match this optimizer')
        st = state_dict['state']
        steps = set()
        for gi, (g, saved, group) in enumerate(zip(self.flat.groups, groups, self.param_groups)):
            group.update({k: v for k, v in saved.items() if k != 'params'})
            for pid, m, v in zip(saved['params'], g.views(self.exp_avg[gi]), g.views(self.exp_avg_sq[gi])):
                ent = st.get(pid, st.get(str(pid)))
                if ent is None:            # parameter that never received a gradient in the saved run
                    m.zero_(); v.zero_()
                    continue
                m.copy_(ent['exp_avg'].reshape(m.shape))
                v.copy_(ent['exp_avg_sq'].reshape(v.shape))
                steps.add(int(float(ent['step'])))
        if len(steps) > 1:
            raise ValueError(f'FusedAdamW.load_state_dict: one step counter for all parameters is required, got {sorted(steps)}')
        self.step_count = steps.pop() if steps else 0
        self.flat.refresh_shadows()        # a resume usually loads the model weights next to the optimizer state

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        self.prepare_step()
        self.launch_step()
        self.update_ema_buffers()
        return loss

    @torch.no_grad()
    def update_ema_buffers(self):
        """ModelEmaV2 also averages buffers (BatchNorm running stats); they are tiny, so plain torch ops -- batched
        with the multi-tensor `_foreach_*` forms: two launches per EMA instead of two per buffer."""
        src = dict(self.flat.model.named_buffers())
        for ema, d in zip(self.ema_models, self.ema_decays):
            fl_dst, fl_src = [], []
            for n, b in ema.named_buffers():
                if b.dtype.is_floating_point:
                    fl_dst.append(b)
                    fl_src.append(src[n].detach())
                else:
                    b.copy_(src[n])
            if fl_dst:
                torch._foreach_mul_(fl_dst, d)
                torch._foreach_add_(fl_dst, fl_src, alpha=1.0 - d)
