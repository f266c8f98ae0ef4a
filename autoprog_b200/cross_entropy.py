"""Token-labeling losses -- drop-in for the reference's `loss/cross_entropy.py` (same class names / signatures).

`TokenLabelCrossEntropy` runs the fused single-pass CUDA kernel (forward + gradient, csrc/tlce.cu); the three small
variants are expressed through the same kernel.  CUDA only, no eager fallback.
"""
import torch
import torch.nn as nn

from . import ops


def _area(bb):
    bbx1, bby1, bbx2, bby2 = (int(v) for v in bb)
    return (bbx2 - bbx1) * (bby2 - bby1)


def _soft_ce(x, target):
    """mean_rows(sum_c -t log_softmax(x)) through the fused kernel: rows become 'tokens' of one-image batches."""
    rows, C = x.shape
    if target.shape[0] != rows:
        target = target.repeat(rows // target.shape[0], 1)
    dummy_cls = torch.zeros(rows, C, device=x.device, dtype=x.dtype)
    t3 = torch.zeros(rows, C, 3, device=x.device, dtype=torch.float32)
    t3[:, :, 2] = target.float()
    return ops.TokenLabelCEFn.apply(dummy_cls, x.reshape(rows, 1, C), t3, 0, 0.0, 1.0)


class SoftTargetCrossEntropy(nn.Module):
    """loss/cross_entropy.py:21-36."""

    def forward(self, x, target):
        return _soft_ce(x, target)


class TokenLabelSoftTargetCrossEntropy(nn.Module):
    """loss/cross_entropy.py:92-109: image-level soft target only."""

    def forward(self, x, target):
        if target.dim() == 3 and target.shape[-1] == 2:
            target = target[:, :, 1]
        return _soft_ce(x, target)


class TokenLabelCrossEntropy(nn.Module):
    """loss/cross_entropy.py:112-156: cls_weight * CE(cls) + dense_weight * CE(all tokens)."""

    def __init__(self, dense_weight=1.0, cls_weight=1.0, mixup_active=True, classes=1000):
        super().__init__()
        self.CE = SoftTargetCrossEntropy()
        self.dense_weight, self.cls_weight = dense_weight, cls_weight
        self.mixup_active, self.classes = mixup_active, classes
        assert dense_weight + cls_weight > 0

    def _targets(self, target):
        return target

    def forward(self, x, target):
        output, aux_output, bb = x
        target = self._targets(target)
        if target.dtype != torch.float32:
            target = target.float()
        return ops.TokenLabelCEFn.apply(output, aux_output, target, _area(bb), float(self.cls_weight),
                                        float(self.dense_weight), getattr(bb, 'dev', None))


class TokenLabelGTCrossEntropy(TokenLabelCrossEntropy):
    """loss/cross_entropy.py:39-89: the cls target is mixed with the ground truth (ratio 0.9, or 0.5 when they agree)."""

    def __init__(self, dense_weight=1.0, cls_weight=1.0, mixup_active=True, smoothing=0.1, classes=1000):
        super().__init__(dense_weight, cls_weight, mixup_active, classes)
        self.smoothing = smoothing

    def _targets(self, target):
        if target.dim() == 2:
            return target
        gt, t_cls = target[:, :, 0], target[:, :, 1]
        ratio = (0.9 - 0.4 * (gt.max(-1)[1] == t_cls.max(-1)[1])).unsqueeze(-1)
        mixed = t_cls * ratio + gt * (1 - ratio)
        return torch.cat([target[:, :, :1], mixed.unsqueeze(-1), target[:, :, 2:]], dim=2)
