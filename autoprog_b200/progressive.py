"""Progressive-learning schedule (host integer / float math that decides every stage's shapes).

Drop-in for the reference's `prog/progressive.py`: same function names, arguments and return values.
"""
import numpy as np


def make_divisible(v, divisor=8, min_value=None, round_limit=.9):
    """Round v to a multiple of divisor, never shrinking by more than 10 % (prog/progressive.py:34-40)."""
    lower = min_value or divisor
    rounded = max(lower, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < round_limit * v else rounded


def _ramp(start, num_stages):
    return np.linspace(start, 1., num_stages)


def progressive_schedule(args, r_max=224, h_max=12, l_max=18):
    """Per-stage (start epoch, resolution, heads, layers, RandAugment policy, drop-path, random-erase prob, crop scale)
    lists (prog/progressive.py:4-31).  `args` is the trainer's argparse namespace."""
    n = args.num_stages
    epochs = [int(v) for v in np.linspace(0, args.epochs, n + 1) // 1][:-1]
    res = [make_divisible(v, 32) for v in _ramp(args.r_scale, n) * r_max]
    heads = [make_divisible(v, 2) for v in _ramp(args.h_scale, n) * h_max]
    depth = [make_divisible(v, 1) for v in _ramp(args.l_scale, n) * l_max]
    assert isinstance(args.aa, str) and args.aa.startswith('rand')
    magnitude_max = float(args.aa.split('-')[1].lstrip('m'))
    magnitudes = [round(max(0., v)) for v in _ramp(args.aa_scale, n) * magnitude_max]
    aug = ['rand-m{}-mstd0.5-inc1'.format(m) if m > 0 else '' for m in magnitudes]
    drop_path = [max(0., v) for v in _ramp(args.dp_scale, n) * args.drop_path]
    erase = [max(0., v) for v in _ramp(args.re_scale, n) * args.reprob]
    lo = _ramp(args.resize_scale[0], n) * args.scale[0]
    hi = _ramp(args.resize_scale[1], n) * args.scale[1]
    crop = [[max(0., a), max(0., b)] for a, b in zip(lo, hi)]
    return epochs, res, heads, depth, aug, drop_path, erase, crop


def resize_input(x, r: int, out_dtype=None):
    """The per-step resolution switch of the trainer (main_prog.py:973-974 and :1910):
    `F.interpolate(input, size=(r, r), mode='bilinear', align_corners=False)` as one kernel; identity when the batch
    already has resolution r.  x: [B, C, H, W] fp32 on the GPU."""
    from . import kernels as K
    if x.shape[-2] == r and x.shape[-1] == r and (out_dtype is None or out_dtype == x.dtype):
        return x
    return K.bilinear_resize(x.contiguous(), r, r, out_dtype)
