"""CPU oracle of the token-label TARGET builder (`tlt.data.create_token_label_target`).  TEST INFRASTRUCTURE ONLY.

The reference calls it every step on the caller side of the loss (main_prog.py:983-1004, 1919-1932) but its source lives
in the un-vendored third-party package `tlt==0.1.0` (Dockerfile:6), NOT under /root/reference -> **parity unpinned**: this
file restates the published TokenLabeling recipe and anchors on the reference's call sites and on
the layout its loss consumes (loss/cross_entropy.py:146-148: target [B, C, 2 + N], slot 0 = ground truth, slot 1 =
class-level soft label, slots 2.. = dense token labels, class-major).

Recipe (per sample; label map item = [3, 5, Hm, Wm]: plane 0 top-5 scores, plane 1 top-5 class ids, plane 2 carries the
augmentation record at [2, 0, 0, 0:6] = crop box x1, y1, x2, y2 (normalised), horizontal-flip flag, ground-truth class):
  1. scatter the top-5 scores into a dense map M[C, Hm, Wm];
  2. RoIAlign (torchvision.ops.roi_align, spatial_scale 1, adaptive sampling, aligned=False) of the crop box
     (x * Wm - 0.5, y * Hm - 0.5) to label_size x label_size; mirror horizontally when the flip flag is set;
     the class-level label is the same RoIAlign to 1 x 1;
  3. softmax over the class dimension;
  4. smoothing: value * on + off with off = smoothing / C, on = 1 - smoothing + off; slot 0 = smoothed one-hot of the
     ground-truth class.
1-D integer targets give the smoothed one-hot [B, C].
"""
from __future__ import annotations

import torch
from torchvision.ops import roi_align


def dense_label_map(scores: torch.Tensor, ids: torch.Tensor, num_classes: int) -> torch.Tensor:
    """[B, 5, Hm, Wm] scores / class ids -> [B, C, Hm, Wm] (duplicates of a class at one pixel add up)."""
    B, K, H, W = scores.shape
    dense = torch.zeros(B, num_classes, H, W, dtype=scores.dtype)
    dense.scatter_add_(1, ids.long(), scores)
    return dense


def create_token_label_target(target: torch.Tensor, num_classes: int, smoothing: float = 0.1, label_size: int = 1,
                              apply_softmax: bool = True) -> torch.Tensor:
    off = smoothing / num_classes
    on = 1.0 - smoothing + off
    if target.dim() == 1:
        out = torch.full((target.shape[0], num_classes), off, dtype=torch.float64)
        out.scatter_(1, target.long().view(-1, 1), on)
        return out
    B, _, K, Hm, Wm = target.shape
    t = target.double()
    dense = dense_label_map(t[:, 0], t[:, 1], num_classes)
    rec = t[:, 2, 0, 0, :6]
    boxes = torch.stack([rec[:, 0] * Wm - 0.5, rec[:, 1] * Hm - 0.5, rec[:, 2] * Wm - 0.5, rec[:, 3] * Hm - 0.5], dim=1)
    rois = [boxes[b:b + 1] for b in range(B)]
    tok = roi_align(dense, rois, (label_size, label_size))            # [B, C, L, L]
    cls = roi_align(dense, rois, (1, 1))                              # [B, C, 1, 1]
    flip = rec[:, 4] > 0.5
    tok = torch.where(flip.view(B, 1, 1, 1), tok.flip(3), tok)
    if apply_softmax:
        tok, cls = torch.softmax(tok, dim=1), torch.softmax(cls, dim=1)
    gt = torch.full((B, num_classes), off, dtype=torch.float64)
    gt.scatter_(1, rec[:, 5].long().view(-1, 1), on)
    return torch.cat([gt.unsqueeze(2), cls.reshape(B, num_classes, 1) * on + off,
                      tok.reshape(B, num_classes, label_size * label_size) * on + off], dim=2)
