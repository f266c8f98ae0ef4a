"""Flat parameter / gradient storage shared by the fused optimizer and the data-parallel reducer.

All trainable parameters of a model are re-pointed into two contiguous fp32 buffers (weight-decay group, no-decay
group -- timm's `create_optimizer` rule: ndim <= 1, names ending in `.bias`, and `model.no_weight_decay()` get no
decay); their `.grad`s are views into matching flat gradient buffers.  One kernel launch then updates a whole group
(AdamW + every EMA copy + the bf16 compute copy), and the reducer all-reduces contiguous slices of the gradient
buffers without any flatten / unflatten copies.

Parameter order inside a group is REVERSE registration order, i.e. roughly the order in which backward produces
the gradients, so a bucket (a contiguous slice) completes early and its all-reduce overlaps the rest of backward.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Tuple

import torch
import torch.nn as nn


def split_decay(model: nn.Module, weight_decay: float) -> Tuple[List[Tuple[str, nn.Parameter]], List[Tuple[str, nn.Parameter]]]:
    skip = set(model.no_weight_decay()) if hasattr(model, 'no_weight_decay') else set()
    decay, no_decay = [], []
    for name, p in model.named_parameters():
        if not p.requires_grad:
            continue
        if weight_decay == 0 or p.ndim <= 1 or name.endswith('.bias') or name in skip:
            no_decay.append((name, p))
        else:
            decay.append((name, p))
    return decay, no_decay


class FlatGroup:
    """One contiguous fp32 parameter buffer + gradient buffer (+ optional bf16 shadow) for a list of parameters."""

    def __init__(self, named: List[Tuple[str, nn.Parameter]], want_shadow: bool = True):
        named = list(reversed(named))
        self.names = [n for n, _ in named]
        self.params = [p for _, p in named]
        dev = self.params[0].device if self.params else torch.device('cpu')
        self.offsets, off = [], 0
        for p in self.params:
            self.offsets.append(off)
            off += (p.numel() + 7) // 8 * 8           # keep every tensor 32-byte (fp32) / 16-byte (bf16) aligned
        self.numel = off
        self.flat_p = torch.zeros(off, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(off, device=dev, dtype=torch.float32)
        self.shadow = torch.zeros(off, device=dev, dtype=torch.bfloat16) if (want_shadow and dev.type == 'cuda') else None
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + p.numel()].view(p.shape)
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
                # destination for kernels that write a parameter gradient in place (ops.grad_dest): when `.grad` is
                # None at backward time the wgrad GEMM / reductions write straight into the flat buffer and autograd
                # adopts that view -- no per-parameter accumulate kernel
                p._apb_grad_view = p.grad
        if self.shadow is not None:
            self.shadow.copy_(self.flat_p)

    def views(self, flat: torch.Tensor):
        return [flat[o:o + p.numel()].view(p.shape) for p, o in zip(self.params, self.offsets)]

    def ensure_grad_views(self):
        """Re-attach `.grad` views (e.g. after zero_grad(set_to_none=True) let autograd allocate fresh tensors)."""
        for p, o in zip(self.params, self.offsets):
            view = self.flat_g[o:o + p.numel()].view(p.shape)
            if p.grad is None:
                p.grad = view
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
                p.grad = view

    def zero_grad(self):
        """Zero the flat buffer and detach the `.grad` views: backward re-attaches them (direct writes, see above)."""
        self.flat_g.zero_()
        for p in self.params:
            p.grad = None
            p._apb_grad_claimed = False        # ops.grad_dest hands the view out once per backward


class FlatState:
    def __init__(self, model: nn.Module, weight_decay: float, want_shadow: bool = True):
        decay, no_decay = split_decay(model, weight_decay)
        self.groups: List[FlatGroup] = []
        self.weight_decays: List[float] = []
        for named, wd in ((decay, weight_decay), (no_decay, 0.0)):
            if named:
                self.groups.append(FlatGroup(named, want_shadow))
                self.weight_decays.append(wd)
        self.model = model

    def flat_like(self, other: nn.Module) -> List[torch.Tensor]:
        """Flat fp32 buffers holding `other`'s parameters (an EMA deep copy) in this state's layout; `other`'s
        parameters are re-pointed into them so the EMA module stays a normal nn.Module."""
        named = dict(other.named_parameters())
        outs = []
        for g in self.groups:
            flat = torch.zeros_like(g.flat_p)
            with torch.no_grad():
                for n, p, o in zip(g.names, g.params, g.offsets):
                    q = named[n]
                    flat[o:o + p.numel()].copy_(q.detach().reshape(-1))
                    q.data = flat[o:o + p.numel()].view(q.shape)
            outs.append(flat)
        return outs

    def zero_grad(self):
        for g in self.groups:
            g.zero_grad()

    @torch.no_grad()
    def refresh_shadows(self):
        """Re-cast the bf16 compute copies from the fp32 parameters (after a load that wrote the parameters in place)."""
        for g in self.groups:
            if g.shadow is not None:
                g.shadow.copy_(g.flat_p)

    def ensure_grad_views(self):
        for g in self.groups:
            g.ensure_grad_views()
