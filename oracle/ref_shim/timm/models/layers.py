import collections.abc
import torch
import torch.nn as nn


def to_2tuple(x):
    if isinstance(x, collections.abc.Iterable):
        return tuple(x)
    return (x, x)


def trunc_normal_(tensor, mean=0., std=1., a=-2., b=2.):
    return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)


class DropPath(nn.Module):
    """timm 0.4.5 semantics: x.div(keep) * floor(keep + U[0,1)) per sample, train only."""

    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        if self.drop_prob == 0. or not self.training:
            return x
        keep = 1 - self.drop_prob
        shape = (x.shape[0],) + (1,) * (x.ndim - 1)
        mask = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
        mask.floor_()
        if getattr(self, 'record', None) is not None:   # golden generation: expose the Bernoulli draws
            self.record.append(mask.reshape(-1).clone())
        return x.div(keep) * mask
