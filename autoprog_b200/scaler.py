"""Loss scalers with the reference's uniform call signature (prog/scaler.py:17-74).

`__call__(loss, optimizer, clip_grad=None, clip_mode='norm', parameters=None, create_graph=False, update=True)`;
`update=False` only accumulates gradients (batch splits, main_prog.py:971).  bf16 needs no loss scaling, so the
B200 AMP path is `Bf16Scaler` (== NoScaler semantics under `autoprog_b200.autocast()`); `NativeScaler` keeps the
fp16 GradScaler flow for callers that insist on it.  Apex is not required.
"""
import torch


def unitwise_norm(x, norm_type=2.0):
    """timm.utils.agc.unitwise_norm: whole-tensor norm for scalars / vectors, per-unit (dim 0) norm otherwise."""
    if x.ndim <= 1:
        return x.norm(norm_type)
    return x.norm(norm_type, dim=tuple(range(1, x.ndim)), keepdim=True)


def dispatch_clip_grad(parameters, value, mode='norm', norm_type=2.0):
    """timm.utils.clip_grad.dispatch_clip_grad ('norm' | 'value' | 'agc')."""
    if mode == 'norm':
        torch.nn.utils.clip_grad_norm_(parameters, value, norm_type=norm_type)
    elif mode == 'value':
        torch.nn.utils.clip_grad_value_(parameters, value)
    elif mode == 'agc':
        # timm adaptive_clip_grad: UNIT-wise norms (per output row for ndim > 1 tensors), eps 1e-3 on the parameter norm
        for p in parameters:
            if p.grad is None:
                continue
            pd, g = p.detach(), p.grad.detach()
            max_norm = unitwise_norm(pd, norm_type).clamp_(min=1e-3).mul_(value)
            gn = unitwise_norm(g, norm_type)
            clipped = g * (max_norm / gn.clamp(min=1e-6))
            g.copy_(torch.where(gn < max_norm, g, clipped))
    else:
        raise AssertionError(f'Unknown clip mode ({mode}).')


class NoScaler:
    state_dict_key = 'no_scaler'

    def __call__(self, loss, optimizer, clip_grad=None, clip_mode='norm', parameters=None, create_graph=False,
                 update=True):
        loss.backward(create_graph=create_graph)
        if update:
            if clip_grad is not None:
                dispatch_clip_grad(parameters, clip_grad, mode=clip_mode)
            optimizer.step()

    def state_dict(self):
        return None

    def load_state_dict(self, state_dict):
        pass


class Bf16Scaler(NoScaler):
    """bf16 autocast has fp32's exponent range: backward + (clip) + step, no scale factor to maintain."""
    state_dict_key = 'bf16_scaler'


class NativeScaler:
    state_dict_key = 'amp_scaler'

    def __init__(self):
        self._scaler = torch.amp.GradScaler('cuda')

    def __call__(self, loss, optimizer, clip_grad=None, clip_mode='norm', parameters=None, create_graph=False,
                 update=True):
        self._scaler.scale(loss).backward(create_graph=create_graph)
        if update:
            if clip_grad is not None:
                assert parameters is not None
                self._scaler.unscale_(optimizer)
                dispatch_clip_grad(parameters, clip_grad, mode=clip_mode)
            self._scaler.step(optimizer)
            self._scaler.update()

    def state_dict(self):
        return self._scaler.state_dict()

    def load_state_dict(self, state_dict):
        self._scaler.load_state_dict(state_dict)


ApexScaler = Bf16Scaler   # the reference's preferred AMP entry (prog/scaler.py:17-34) maps onto the bf16 path
