import torch


def dispatch_clip_grad(parameters, value, mode='norm', norm_type=2.0):
    if mode == 'norm':
        torch.nn.utils.clip_grad_norm_(parameters, value, norm_type=norm_type)
    elif mode == 'value':
        torch.nn.utils.clip_grad_value_(parameters, value)
    else:
        raise ValueError(mode)
