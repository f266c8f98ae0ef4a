"""CPU oracle for the AutoProg VOLO/DeiT training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a plain torch-on-CPU *restatement* (functional, state_dict driven, fp64 capable) of
the arithmetic in the reference's `models/volo.py`, `models/submodels.py`, `loss/cross_entropy.py`,
`prog/progressive.py` and `prog/helpers.py:254-262`.  It is the checker for the CUDA path:

  * only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
    legs may import it; the product package `autoprog_b200` never does (and has no CPU fallback);
  * PARITY PIN: the reference ships no tests or golden vectors, so the pin is the
    reference itself: `oracle/gen_golden.py` imports the reference modules from /root/reference
    (through `oracle/ref_shim`, a stand-in for un-vendored timm) inside the build container, runs
    them on seeded inputs and writes `tests/golden/*.pt`;  `tests/test_oracle_golden.py` checks this
    restatement against those fixtures.  The DeiT part (`vit_forward`) restates timm 0.4.5's
    VisionTransformer, which is NOT under /root/reference -> **parity unpinned** for DeiT.

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# --------------------------------------------------------------------------------------
# host integer math (prog/progressive.py:4-40, prog/helpers.py:254-262, models/volo.py:598-616)
# --------------------------------------------------------------------------------------


def make_divisible(v, divisor=8, min_value=None, round_limit=.9):
    """prog/progressive.py:34-40."""
    floor_ = min_value or divisor
    out = max(floor_, int(v + divisor / 2) // divisor * divisor)
    if out < round_limit * v:
        out += divisor
    return out


def progressive_schedule(epochs, num_stages, r_scale, h_scale, l_scale, aa, aa_scale, drop_path, dp_scale,
                         reprob, re_scale, scale, resize_scale, r_max=224, h_max=12, l_max=18):
    """prog/progressive.py:4-31 with the argparse fields spelled out as arguments."""
    lin = lambda lo: np.linspace(lo, 1., num_stages)
    e = [int(i) for i in np.linspace(0, epochs, num_stages + 1) // 1][:-1]
    r = [make_divisible(i, 32) for i in lin(r_scale) * r_max]
    h = [make_divisible(i, 2) for i in lin(h_scale) * h_max]
    l = [make_divisible(i, 1) for i in lin(l_scale) * l_max]
    m_max = float(aa.split('-')[1].lstrip('m'))
    m = [round(max(0., i)) for i in lin(aa_scale) * m_max]
    aa_l = ['rand-m{}-mstd0.5-inc1'.format(k) if k > 0 else '' for k in m]
    dp = [max(0., i) for i in lin(dp_scale) * drop_path]
    re = [max(0., i) for i in lin(re_scale) * reprob]
    rs = [[max(0., a), max(0., b)] for a, b in zip(lin(resize_scale[0]) * scale[0], lin(resize_scale[1]) * scale[1])]
    return e, r, h, l, aa_l, dp, re, rs


def new_idx(idx, prev_l, new_l):
    """prog/helpers.py:254-258: which old layer a layer of the grown model is cloned from."""
    q = new_l // prev_l * prev_l
    keep = prev_l - new_l % prev_l
    a = idx * prev_l // q
    if a < keep:
        return a
    return (idx + keep) * prev_l // (q + prev_l)


def get_new_layer_idx(prev_l, new_l):
    """prog/helpers.py:261-262."""
    return [i for i in range(new_l) if new_idx(i, prev_l, new_l) == new_idx(i - 1, prev_l, new_l)]


def stage_layers(l):
    """models/submodels.py:20-25 / models/volo.py:602-607: [outlooker, transformer, 0, 0] split of depth l."""
    if l > 2:
        l0 = make_divisible(l * 0.23, 2)
        return [l0, l - l0, 0, 0]
    return [1, 1, 0, 0]


def identity_layer_plan(layer_num, min_layer_num, max_layer_num):
    """models/volo.py:598-616: per real stage, the indices flagged `is_identity_layer`."""
    l0 = make_divisible(layer_num * 0.23, 2)
    l0n = make_divisible(min_layer_num * 0.23, 2)
    l0x = make_divisible(max_layer_num * 0.23, 2)
    cur = [l0, layer_num - l0, 0, 0]
    lo = [l0n, min_layer_num - l0n, 0, 0]
    hi = [l0x, max_layer_num - l0x, 0, 0]
    plan = []
    for i in range(4):
        fresh = get_new_layer_idx(prev_l=lo[i], new_l=hi[i])
        grown = cur[i] - lo[i]
        plan.append(fresh if grown == 0 else fresh[:-grown])
    return plan


# --------------------------------------------------------------------------------------
# elementary ops
# --------------------------------------------------------------------------------------


def linear(x: Tensor, w: Tensor, b: Optional[Tensor] = None) -> Tensor:
    y = x @ w.t()
    return y if b is None else y + b


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.LayerNorm over the last dim (models/volo.py:472 default eps=1e-5; DeiT 1e-6 models/deit.py:66)."""
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x: Tensor) -> Tensor:
    """exact-erf GELU (nn.GELU default, models/volo.py:118,151)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def mlp(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """models/volo.py:161-167 (dropout p=0)."""
    hdn = gelu(linear(x, sd[pre + 'fc1.weight'], sd[pre + 'fc1.bias']))
    return linear(hdn, sd[pre + 'fc2.weight'], sd[pre + 'fc2.bias'])


def avgpool2_ceil(x: Tensor) -> Tensor:
    """AvgPool2d(2,2,ceil_mode=True) on NHWC (models/volo.py:75,87): edge windows average in-bounds pixels only."""
    B, H, W, C = x.shape
    h, w = (H + 1) // 2, (W + 1) // 2
    xp = x.new_zeros(B, 2 * h, 2 * w, C)
    xp[:, :H, :W] = x
    s = xp[:, 0::2, 0::2] + xp[:, 1::2, 0::2] + xp[:, 0::2, 1::2] + xp[:, 1::2, 1::2]
    ones = x.new_zeros(1, 2 * h, 2 * w, 1)
    ones[:, :H, :W] = 1
    cnt = ones[:, 0::2, 0::2] + ones[:, 1::2, 0::2] + ones[:, 0::2, 1::2] + ones[:, 1::2, 1::2]
    return s / cnt


# --------------------------------------------------------------------------------------
# OutlookAttention core (models/volo.py:77-103), k=3, pad=1, stride=2
# --------------------------------------------------------------------------------------


def _windows(v: Tensor, h: int, w: int) -> Tensor:
    """unfold (models/volo.py:83-85): [B,H,W,C] -> [B,h,w,9,C]; Q = qi*3+qj at pixel (2i-1+qi, 2j-1+qj), zero pad."""
    B, H, W, C = v.shape
    vp = v.new_zeros(B, 2 * h + 1, 2 * w + 1, C)
    vp[:, 1:1 + H, 1:1 + W] = v
    cols = [vp[:, qi:qi + 2 * h:2, qj:qj + 2 * w:2] for qi in range(3) for qj in range(3)]
    return torch.stack(cols, dim=3)


def _fold(o: Tensor, H: int, W: int) -> Tensor:
    """F.fold (models/volo.py:97-98): [B,h,w,9,C] -> [B,H,W,C], overlapping windows summed."""
    B, h, w, _, C = o.shape
    yp = o.new_zeros(B, 2 * h + 1, 2 * w + 1, C)
    for ki in range(3):
        for kj in range(3):
            yp[:, ki:ki + 2 * h:2, kj:kj + 2 * w:2] += o[:, :, :, ki * 3 + kj]
    return yp[:, 1:1 + H, 1:1 + W]


def outlook_probs(logits: Tensor, heads: int, scale: float) -> Tensor:
    """models/volo.py:88-92: [B,h,w,heads*81] -> softmax over the last 9 of [B,h,w,heads,9,9]."""
    B, h, w, _ = logits.shape
    return torch.softmax(logits.reshape(B, h, w, heads, 9, 9) * scale, dim=-1)


def outlook_core(v: Tensor, logits: Tensor, heads: int, scale: float) -> Tensor:
    """unfold -> softmax(scale*logits) @ windows -> fold (models/volo.py:83-98).  v [B,H,W,C], logits [B,h,w,heads*81]."""
    B, H, W, C = v.shape
    h, w = logits.shape[1], logits.shape[2]
    d = C // heads
    A = outlook_probs(logits, heads, scale)                        # [B,h,w,heads,P,Q]
    win = _windows(v, h, w).reshape(B, h, w, 9, heads, d)          # [B,h,w,Q,heads,d]
    out = torch.einsum('bijnpq,bijqnd->bijpnd', A, win).reshape(B, h, w, 9, C)
    return _fold(out, H, W)


def outlook_core_bwd(v: Tensor, logits: Tensor, dy: Tensor, heads: int, scale: float) -> Tuple[Tensor, Tensor]:
    """Closed-form backward of `outlook_core`; returns (dv, dlogits)."""
    B, H, W, C = v.shape
    h, w = logits.shape[1], logits.shape[2]
    d = C // heads
    A = outlook_probs(logits, heads, scale)
    win = _windows(v, h, w).reshape(B, h, w, 9, heads, d)
    dyw = _windows(dy, h, w).reshape(B, h, w, 9, heads, d)         # transpose of fold == unfold
    dA = torch.einsum('bijpnd,bijqnd->bijnpq', dyw, win)
    dl = scale * A * (dA - (A * dA).sum(-1, keepdim=True))
    dwin = torch.einsum('bijnpq,bijpnd->bijqnd', A, dyw).reshape(B, h, w, 9, C)
    dv = _fold(dwin, H, W)                                         # transpose of unfold == fold
    return dv, dl.reshape(B, h, w, heads * 81)


def outlook_attention(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """OutlookAttention.forward, models/volo.py:77-103 (qkv_bias=False; attn/proj have bias)."""
    C = x.shape[-1]
    scale = (C // heads) ** -0.5
    v = linear(x, sd[pre + 'v.weight'], sd.get(pre + 'v.bias'))
    logits = linear(avgpool2_ceil(x), sd[pre + 'attn.weight'], sd[pre + 'attn.bias'])
    y = outlook_core(v, logits, heads, scale)
    return linear(y, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


# --------------------------------------------------------------------------------------
# MHSA / class attention (models/volo.py:185-201, 261-308)
# --------------------------------------------------------------------------------------


def mhsa_core(qkv: Tensor, heads: int, scale: float) -> Tensor:
    """softmax(q k^T * scale) v per head; qkv [B,N,3*C] laid out (3, heads, d) -> [B,N,C] (models/volo.py:188-197)."""
    B, N, C3 = qkv.shape
    C = C3 // 3
    d = C // heads
    t = qkv.reshape(B, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    p = torch.softmax((q @ k.transpose(-2, -1)) * scale, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, N, C)


def mhsa(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """Attention.forward, models/volo.py:185-201; x [B,H,W,C]."""
    B, H, W, C = x.shape
    qkv = linear(x.reshape(B, H * W, C), sd[pre + 'qkv.weight'], sd.get(pre + 'qkv.bias'))
    o = mhsa_core(qkv, heads, (C // heads) ** -0.5)
    return linear(o, sd[pre + 'proj.weight'], sd[pre + 'proj.bias']).reshape(B, H, W, C)


def class_attn_core(q: Tensor, kv: Tensor, heads: int, scale: float) -> Tensor:
    """cls query against all tokens: q [B,C], kv [B,N,2C] laid out (2, heads, d) -> [B,C] (models/volo.py:264-275)."""
    B, N, C2 = kv.shape
    C = C2 // 2
    d = C // heads
    t = kv.reshape(B, N, 2, heads, d).permute(2, 0, 3, 1, 4)
    k, v = t[0], t[1]
    qh = q.reshape(B, heads, 1, d) * scale
    p = torch.softmax(qh @ k.transpose(-2, -1), dim=-1)
    return (p @ v).reshape(B, C)


def class_block(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    """ClassBlock.forward + ClassAttention.forward, models/volo.py:261-277, 304-308; x [B,1+N,C]."""
    C = x.shape[-1]
    xn = layer_norm(x, sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'])
    kv = linear(xn, sd[pre + 'attn.kv.weight'], sd.get(pre + 'attn.kv.bias'))
    q = linear(xn[:, 0], sd[pre + 'attn.q.weight'], sd.get(pre + 'attn.q.bias'))
    a = class_attn_core(q, kv, heads, (C // heads) ** -0.5)
    cls = x[:, 0] + linear(a, sd[pre + 'attn.proj.weight'], sd[pre + 'attn.proj.bias'])
    cls = cls + mlp(layer_norm(cls, sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias']), sd, pre + 'mlp.')
    return torch.cat([cls[:, None], x[:, 1:]], dim=1)


# --------------------------------------------------------------------------------------
# position-embedding bicubic resize (models/volo.py:580-596)
# --------------------------------------------------------------------------------------


def _cubic_w(t: float, A: float = -0.75) -> List[float]:
    def near(x):   # |x| <= 1
        return ((A + 2) * x - (A + 3)) * x * x + 1
    def far(x):    # 1 < |x| < 2
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A
    return [far(t + 1), near(t), near(1 - t), far(2 - t)]


def bicubic_matrix(n_in: int, n_out: int) -> Tensor:
    """[n_out, n_in] fp64 matrix of upsample_bicubic2d(align_corners=False) driven by scale_factor=(n_out+0.1)/n_in."""
    sf = (n_out + 0.1) / n_in
    assert int(math.floor(n_in * sf)) == n_out
    M = torch.zeros(n_out, n_in, dtype=torch.float64)
    for o in range(n_out):
        src = (o + 0.5) / sf - 0.5
        i0 = math.floor(src)
        ws = _cubic_w(src - i0)
        for k in range(4):
            j = min(max(i0 - 1 + k, 0), n_in - 1)
            M[o, j] += ws[k]
    return M


def bilinear_matrix(n_in: int, n_out: int) -> Tensor:
    """[n_out, n_in] fp64 matrix of upsample_bilinear2d(align_corners=False, size=n_out): the input resolution switch
    `F.interpolate(input, size=(r, r), mode='bilinear', align_corners=False)` (main_prog.py:973-974, 1910)."""
    M = torch.zeros(n_out, n_in, dtype=torch.float64)
    scale = n_in / n_out
    for o in range(n_out):
        src = max((o + 0.5) * scale - 0.5, 0.0)
        i0 = min(int(math.floor(src)), n_in - 1)
        i1 = min(i0 + 1, n_in - 1)
        lam = src - i0
        M[o, i0] += 1.0 - lam
        M[o, i1] += lam
    return M


def resize_input(x: Tensor, r: int) -> Tensor:
    """x [B,C,H,W] -> [B,C,r,r], the trainer's per-step bilinear resolution switch (main_prog.py:973-974)."""
    My = bilinear_matrix(x.shape[2], r).to(x)
    Mx = bilinear_matrix(x.shape[3], r).to(x)
    return torch.einsum('yh,xw,bchw->bcyx', My, Mx, x)


def pos_embed_resize(pos: Tensor, h0: int, w0: int) -> Tensor:
    """VOLO.interpolate_pos_encoding, models/volo.py:580-596: pos [1,h,w,C] -> [1,h0,w0,C] (identity if equal)."""
    _, h, w, C = pos.shape
    if h == h0 and w == w0:
        return pos
    My = bicubic_matrix(h, h0).to(pos)
    Mx = bicubic_matrix(w, w0).to(pos)
    return torch.einsum('yh,xw,bhwc->byxc', My, Mx, pos)


# --------------------------------------------------------------------------------------
# token-label loss (loss/cross_entropy.py:30-36, 62-89, 101-109, 136-156)
# --------------------------------------------------------------------------------------


def soft_ce(x: Tensor, t: Tensor) -> Tensor:
    """SoftTargetCrossEntropy.forward, loss/cross_entropy.py:30-36."""
    if x.shape[0] != t.shape[0]:
        t = t.repeat(x.shape[0] // t.shape[0], 1)
    return (-t * F.log_softmax(x, dim=-1)).sum(-1).mean()


def token_label_ce(x_cls: Tensor, x_aux: Tensor, bbox: Sequence[int], target: Tensor, dense_weight: float = 1.0,
                   cls_weight: float = 1.0, gt_mix: bool = False) -> Tensor:
    """TokenLabelCrossEntropy.forward (loss/cross_entropy.py:136-156); gt_mix=True -> TokenLabelGTCrossEntropy (:62-89)."""
    B, N, C = x_aux.shape
    if target.dim() == 2:
        t_cls = target
        t_aux = target.repeat(1, N).reshape(B * N, C)
    else:
        t_cls = target[:, :, 1]
        if gt_mix:
            gt = target[:, :, 0]
            ratio = (0.9 - 0.4 * (gt.max(-1)[1] == t_cls.max(-1)[1])).unsqueeze(-1)
            t_cls = t_cls * ratio + gt * (1 - ratio)
        t_aux = target[:, :, 2:].transpose(1, 2).reshape(-1, C)
    x1, y1, x2, y2 = [int(b) for b in bbox]
    lam = 1 - ((x2 - x1) * (y2 - y1) / N)
    if lam < 1:
        t_cls = lam * t_cls + (1 - lam) * t_cls.flip(0)
    return cls_weight * soft_ce(x_cls, t_cls) + dense_weight * soft_ce(x_aux.reshape(-1, C), t_aux)


def token_label_ce_grads(x_cls, x_aux, bbox, target, dense_weight=1.0, cls_weight=1.0):
    """Closed-form gradients of `token_label_ce`: returns (loss, d x_cls, d x_aux)."""
    B, N, C = x_aux.shape
    if target.dim() == 2:
        t_cls = target
        t_aux = target[:, None, :].expand(B, N, C)
    else:
        t_cls = target[:, :, 1]
        t_aux = target[:, :, 2:].transpose(1, 2)
    x1, y1, x2, y2 = [int(b) for b in bbox]
    lam = 1 - ((x2 - x1) * (y2 - y1) / N)
    if lam < 1:
        t_cls = lam * t_cls + (1 - lam) * t_cls.flip(0)
    lc = F.log_softmax(x_cls, -1)
    la = F.log_softmax(x_aux, -1)
    loss = cls_weight * (-(t_cls * lc).sum(-1)).mean() + dense_weight * (-(t_aux * la).sum(-1)).mean()
    dc = (cls_weight / B) * (lc.exp() * t_cls.sum(-1, keepdim=True) - t_cls)
    da = (dense_weight / (B * N)) * (la.exp() * t_aux.sum(-1, keepdim=True) - t_aux)
    return loss, dc, da


# --------------------------------------------------------------------------------------
# mix-token (models/volo.py:319-339, 649-660, 684-691)
# --------------------------------------------------------------------------------------


def rand_bbox(grid_h: int, grid_w: int, lam: float, rng=np.random) -> Tuple[int, int, int, int]:
    """models/volo.py:319-339 on the pooled grid (size[1]//scale, size[2]//scale); draws cx then cy from `rng`."""
    Wg, Hg = grid_h, grid_w          # the reference names dim-1 'W' and dim-2 'H'
    cut = np.sqrt(1. - lam)
    cw, ch = int(Wg * cut), int(Hg * cut)
    cx = rng.randint(Wg)
    cy = rng.randint(Hg)
    return (int(np.clip(cx - cw // 2, 0, Wg)), int(np.clip(cy - ch // 2, 0, Hg)),
            int(np.clip(cx + cw // 2, 0, Wg)), int(np.clip(cy + ch // 2, 0, Hg)))


def flip_in_box(x: Tensor, box: Sequence[int]) -> Tensor:
    """models/volo.py:655-658 / 687-689: inside box (dim1 in [b0,b2), dim2 in [b1,b3)) take the batch-flipped sample."""
    b0, b1, b2, b3 = box
    out = x.clone()
    out[:, b0:b2, b1:b3] = x.flip(0)[:, b0:b2, b1:b3]
    return out


# --------------------------------------------------------------------------------------
# whole VOLO forward (models/volo.py:376-396, 618-694)
# --------------------------------------------------------------------------------------


class VoloArch:
    """Static description of a VOLO variant (models/volo.py:697-821, models/submodels.py:9-41)."""

    def __init__(self, layers, embed_dims, num_heads, mlp_ratio=3, stem_hidden=64, img_size=224, num_classes=1000,
                 post_layers=2):
        self.layers, self.embed_dims, self.num_heads = list(layers), list(embed_dims), list(num_heads)
        self.mlp_ratio, self.stem_hidden, self.img_size = mlp_ratio, stem_hidden, img_size
        self.num_classes, self.post_layers = num_classes, post_layers

    @staticmethod
    def named(name: str, **kw) -> 'VoloArch':
        table = {
            'volo_d1': ([4, 4, 8, 2], [192, 384, 384, 384], [6, 12, 12, 12], 3, 64),
            'volo_d2': ([6, 4, 10, 4], [256, 512, 512, 512], [8, 16, 16, 16], 3, 64),
            'volo_d3': ([8, 8, 16, 4], [256, 512, 512, 512], [8, 16, 16, 16], 3, 64),
            'volo_d4': ([8, 8, 16, 4], [384, 768, 768, 768], [12, 16, 16, 16], 3, 64),
            'volo_d5': ([12, 12, 20, 4], [384, 768, 768, 768], [12, 16, 16, 16], 4, 128),
        }
        if name in table:
            L, E, Hh, r, s = table[name]
            return VoloArch(L, E, Hh, r, s, **kw)
        parts = name.split('_')          # volo_h{h}_l{l}, models/submodels.py:16-28
        hh, l = int(parts[1].lstrip('h')), int(parts[2].lstrip('l'))
        return VoloArch(stage_layers(l), [hh * 16, hh * 32, hh * 32, hh * 32], [hh // 2, hh, hh, hh], 3, 64, **kw)


def batch_norm_train(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    """nn.BatchNorm2d in train mode on NCHW (batch statistics, biased variance); models/volo.py:358-366."""
    mu = x.mean((0, 2, 3), keepdim=True)
    var = ((x - mu) ** 2).mean((0, 2, 3), keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def batch_norm_eval(x, w, b, rm, rv, eps=1e-5):
    return (x - rm.view(1, -1, 1, 1)) / torch.sqrt(rv.view(1, -1, 1, 1) + eps) * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)


def patch_embed(x: Tensor, sd: Dict[str, Tensor], train: bool) -> Tensor:
    """PatchEmbed.forward models/volo.py:355-380: conv7x7 s2 + BN + ReLU, 2x(conv3x3 + BN + ReLU), conv4x4 s4; -> NHWC."""
    p = 'patch_embed.'
    for i, (k, s, pad) in zip((0, 3, 6), ((7, 2, 3), (3, 1, 1), (3, 1, 1))):
        x = F.conv2d(x, sd[f'{p}conv.{i}.weight'], None, stride=s, padding=pad)
        bn = f'{p}conv.{i + 1}.'
        if train:
            x = batch_norm_train(x, sd[bn + 'weight'], sd[bn + 'bias'])
        else:
            x = batch_norm_eval(x, sd[bn + 'weight'], sd[bn + 'bias'], sd[bn + 'running_mean'], sd[bn + 'running_var'])
        x = torch.relu(x)
    k = sd[p + 'proj.weight'].shape[-1]
    x = F.conv2d(x, sd[p + 'proj.weight'], sd[p + 'proj.bias'], stride=k)
    return x.permute(0, 2, 3, 1)


def downsample(x: Tensor, sd: Dict[str, Tensor], pre: str) -> Tensor:
    """Downsample.forward models/volo.py:392-396: conv2x2 s2 on NHWC."""
    y = F.conv2d(x.permute(0, 3, 1, 2), sd[pre + 'proj.weight'], sd[pre + 'proj.bias'], stride=2)
    return y.permute(0, 2, 3, 1)


def drop_path_rates(arch: VoloArch, dpr: float) -> List[List[float]]:
    """models/volo.py:408-409, 429-430: only transformer stages receive drop_path_rate (outlooker call omits it :495-500)."""
    tot = sum(arch.layers)
    rates = []
    for i, n in enumerate(arch.layers):
        if i == 0:
            rates.append([0.0] * n)
        else:
            rates.append([dpr * (j + sum(arch.layers[:i])) / (tot - 1) for j in range(n)])
    return rates


def volo_forward(sd: Dict[str, Tensor], x: Tensor, arch: VoloArch, train: bool = True,
                 bbox: Optional[Sequence[int]] = None, skip: Optional[List[List[int]]] = None,
                 keep_masks: Optional[Dict[str, Tensor]] = None, keep_prob: Optional[Dict[str, float]] = None):
    """VOLO.forward models/volo.py:644-694.

    bbox: mix-token box on the pooled grid (train only; None/(0,0,0,0) = no mixing).
    skip: identity-layer plan from `identity_layer_plan` (models/volo.py:609-616).
    keep_masks / keep_prob: per-block DropPath draws keyed by block prefix: keep_masks[pre] = [mask_attn, mask_mlp]
        (each [B] of 0/1; the block calls its DropPath twice per forward, models/volo.py:232-233) and the keep
        probability (timm 0.4.5 DropPath: x / keep * mask).
    Returns (x_cls, x_aux, bbox) in train mode, x_cls + 0.5*max_n(x_aux) in eval mode.
    """
    x = patch_embed(x, sd, train)
    bbox = tuple(int(b) for b in bbox) if (bbox is not None and train) else (0, 0, 0, 0)
    if train:
        x = flip_in_box(x, [2 * b for b in bbox])

    def branch(pre, which, val):
        if keep_masks is not None and keep_masks.get(pre):
            m = keep_masks[pre][which].to(val.dtype).view(-1, *([1] * (val.dim() - 1)))
            return val / keep_prob[pre] * m
        return val

    net_idx, real_stage = 0, 0
    for si, n in enumerate(arch.layers):
        if net_idx == 2:       # models/volo.py:627-629: pos-embed is added before network[2]
            x = x + pos_embed_resize(sd['pos_embed'], x.shape[1], x.shape[2])
        heads = arch.num_heads[si]
        for li in range(n):
            if skip is not None and li in skip[real_stage]:
                continue
            pre = f'network.{net_idx}.{li}.'
            xn = layer_norm(x, sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'])
            if si == 0:
                a = outlook_attention(xn, sd, pre + 'attn.', heads)
            else:
                a = mhsa(xn, sd, pre + 'attn.', heads)
            x = x + branch(pre, 0, a)
            x = x + branch(pre, 1, mlp(layer_norm(x, sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias']), sd, pre + 'mlp.'))
        net_idx += 1
        real_stage += 1
        if si == 0:
            x = downsample(x, sd, f'network.{net_idx}.')
            net_idx += 1
    B, Hh, Ww, C = x.shape
    x = x.reshape(B, Hh * Ww, C)
    x = torch.cat([sd['cls_token'].expand(B, -1, -1), x], dim=1)
    for i in range(arch.post_layers):
        x = class_block(x, sd, f'post_network.{i}.', arch.num_heads[-1])
    x = layer_norm(x, sd['norm.weight'], sd['norm.bias'])
    x_cls = linear(x[:, 0], sd['head.weight'], sd['head.bias'])
    x_aux = linear(x[:, 1:], sd['aux_head.weight'], sd['aux_head.bias'])
    if not train:
        return x_cls + 0.5 * x_aux.max(1)[0]
    x_aux = flip_in_box(x_aux.reshape(B, Hh, Ww, -1), bbox).reshape(B, Hh * Ww, -1)
    return x_cls, x_aux, bbox


# --------------------------------------------------------------------------------------
# DeiT (models/deit.py:62-179 -> timm 0.4.5 VisionTransformer, NOT in /root/reference: parity unpinned)
# --------------------------------------------------------------------------------------


def vit_mhsa(x: Tensor, sd: Dict[str, Tensor], pre: str, heads: int) -> Tensor:
    B, N, C = x.shape
    qkv = linear(x, sd[pre + 'qkv.weight'], sd.get(pre + 'qkv.bias'))
    o = mhsa_core(qkv, heads, (C // heads) ** -0.5)
    return linear(o, sd[pre + 'proj.weight'], sd[pre + 'proj.bias'])


def vit_forward(sd: Dict[str, Tensor], x: Tensor, depth: int, heads: int, patch: int = 16, eps: float = 1e-6,
                skip: Optional[Sequence[int]] = None, dense: bool = False):
    """timm-0.4.5 VisionTransformer.forward as used by models/deit.py:64-66 (restated from the public package).

    dense=True additionally returns aux_head logits on the patch tokens (token-labeling extension, no reference).
    """
    x = F.conv2d(x, sd['patch_embed.proj.weight'], sd['patch_embed.proj.bias'], stride=patch)
    B, C, Hh, Ww = x.shape
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([sd['cls_token'].expand(B, -1, -1), x], dim=1) + sd['pos_embed']
    for i in range(depth):
        if skip is not None and i in skip:
            continue
        pre = f'blocks.{i}.'
        x = x + vit_mhsa(layer_norm(x, sd[pre + 'norm1.weight'], sd[pre + 'norm1.bias'], eps), sd, pre + 'attn.', heads)
        x = x + mlp(layer_norm(x, sd[pre + 'norm2.weight'], sd[pre + 'norm2.bias'], eps), sd, pre + 'mlp.')
    x = layer_norm(x, sd['norm.weight'], sd['norm.bias'], eps)
    out = linear(x[:, 0], sd['head.weight'], sd['head.bias'])
    if dense:
        return out, linear(x[:, 1:], sd['aux_head.weight'], sd['aux_head.bias'])
    return out
