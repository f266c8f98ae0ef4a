"""VOLO on the autoprog_b200 kernels -- drop-in for the reference's `models/volo.py`.

Same module tree, parameter names and nn container types as the reference (so `state_dict`s, timm's
ModelEmaV2 deep copies, `prog/helpers.py`-style weight inheritance and DDP wrappers keep working), same
constructor / factory signatures, same `forward` contract:

    train: (x_cls [B,classes], x_aux [B,N,classes], (bbx1, bby1, bbx2, bby2))      models/volo.py:694
    eval : x_cls + 0.5 * max_n x_aux                                               models/volo.py:681-682

Only the arithmetic differs: every forward/backward runs through the sm_100a kernels of `_apb.so`
(see ops.py); there is no eager / CPU fallback -- CUDA tensors are required.
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from . import kernels as K
from . import ops
from .helpers import get_new_layer_idx
from .progressive import make_divisible
from .registry import register_model

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def _cfg(url='', **kwargs):
    cfg = dict(url=url, num_classes=1000, input_size=(3, 224, 224), pool_size=None, crop_pct=.96,
               interpolation='bicubic', mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD,
               first_conv='patch_embed.proj', classifier='head')
    cfg.update(kwargs)
    return cfg


default_cfgs = {'volo': _cfg(crop_pct=0.96), 'volo_large': _cfg(crop_pct=1.15)}


class DropPath(nn.Module):
    """Stochastic depth, timm 0.4.5 semantics: x / keep * floor(keep + U[0,1)) per sample (train only).

    Inside the fused blocks only `sample_scale` is used: it returns the per-sample factor that the
    residual-add + LayerNorm kernel multiplies the branch by."""

    def __init__(self, drop_prob: float = 0.):
        super().__init__()
        self.drop_prob = float(drop_prob)
        self.forced = None          # tests: list of masks consumed in call order
        self._pooled = None         # per-sample scale pre-drawn for this forward by DropPathPool (one RNG launch / step)

    def sample_scale(self, batch: int, device) -> Optional[torch.Tensor]:
        if self.drop_prob == 0. or not self.training:
            return None
        keep = 1.0 - self.drop_prob
        if self.forced:
            mask = self.forced.pop(0).to(device=device, dtype=torch.float32)
        elif self._pooled and self._pooled[-1].shape[0] == batch:
            return self._pooled.pop()
        else:
            mask = torch.floor(keep + torch.rand(batch, device=device, dtype=torch.float32))
        return (mask / keep).contiguous()

    def forward(self, x):
        rs = self.sample_scale(x.shape[0], x.device)
        if rs is None:
            return x
        return _ScaleRows.apply(x, rs)


class DropPathPool:
    """Draws the stochastic-depth factors of ALL DropPath modules of a model in one go at the start of a forward:
    one `rand` + three elementwise launches per step instead of four tiny launches per residual branch (36 branches in
    volo_d1).  Same distribution as the per-module draw (x / keep * floor(keep + U[0,1)), timm 0.4.5 DropPath); the
    order in which random numbers are consumed differs from the reference, which has no effect on the statistics."""

    DRAWS = 2

    def __init__(self, model: nn.Module):
        self.mods = [m for m in model.modules() if isinstance(m, DropPath) and m.drop_prob > 0.]
        self.keep = None

    def draw(self, batch: int, device):
        if not self.mods:
            return
        if self.keep is None or self.keep.device != device:
            self.keep = torch.tensor([1.0 - m.drop_prob for m in self.mods], device=device, dtype=torch.float32)[:, None]
        # a block calls its DropPath module once per residual branch: DRAWS independent factors per module and step
        u = torch.rand(self.DRAWS, len(self.mods), batch, device=device, dtype=torch.float32)
        scales = u.add_(self.keep).floor_().div_(self.keep)
        for i, m in enumerate(self.mods):
            m._pooled = [scales[d, i] for d in range(self.DRAWS)]


class _ScaleRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, rs):
        ctx.rs = rs
        return K.scale_cast(x.contiguous(), x.dtype, rs)

    @staticmethod
    def backward(ctx, dy):
        return K.scale_cast(dy.contiguous(), dy.dtype, ctx.rs), None


def _drop_scale(mod, batch, device):
    return mod.sample_scale(batch, device) if isinstance(mod, DropPath) else None


class Linear(nn.Linear):
    """nn.Linear whose forward is the tcgen05 (bf16) / CUDA-core (fp32) GEMM."""

    def forward(self, x):
        return ops.LinearFn.apply(x, self.weight, self.bias)


class LayerNorm(nn.LayerNorm):
    def forward(self, x):
        return ops.LayerNormFn.apply(x, self.weight, self.bias, self.eps)


class GELU(nn.GELU):
    def forward(self, x):
        return ops.GeluFn.apply(x)


class OutlookAttention(nn.Module):
    """models/volo.py:48-103.  v / attn / proj Linear containers; the unfold-softmax-matmul-fold core is one kernel."""

    def __init__(self, dim, num_heads, kernel_size=3, padding=1, stride=1, qkv_bias=False, qk_scale=None, attn_drop=0.,
                 proj_drop=0.):
        super().__init__()
        self.num_heads, self.kernel_size, self.padding, self.stride = num_heads, kernel_size, padding, stride
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.v = Linear(dim, dim, bias=qkv_bias)
        self.attn = Linear(dim, kernel_size ** 4 * num_heads)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)
        self.unfold = nn.Unfold(kernel_size=kernel_size, padding=padding, stride=stride)      # kept for tree parity
        self.pool = nn.AvgPool2d(kernel_size=stride, stride=stride, ceil_mode=True)

    def _check(self, dim):
        if (self.kernel_size, self.padding, self.stride) != (3, 1, 2) or dim // self.num_heads != 32:
            raise NotImplementedError('autoprog_b200 OutlookAttention kernels cover kernel 3 / padding 1 / stride 2 / '
                                      'head_dim 32 (every VOLO variant of the reference); got '
                                      f'k={self.kernel_size} p={self.padding} s={self.stride} hd={dim // self.num_heads}')
        if self.attn_drop.p != 0. or self.proj_drop.p != 0.:
            raise NotImplementedError('attn_drop / proj_drop > 0 are not used by any reference config')

    def forward(self, x):
        self._check(x.shape[-1])
        v = self.v(x)
        logits = self.attn(ops.AvgPool2Fn.apply(x))
        y = ops.OutlookCoreFn.apply(v, logits, self.num_heads, self.scale)
        return self.proj(y)


class Mlp(nn.Module):
    """models/volo.py:147-167."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        self.fc1 = Linear(in_features, hidden_features)
        self.act = GELU() if act_layer in (nn.GELU, GELU) else act_layer()
        self.fc2 = Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)

    def forward(self, x):
        return self.drop(self.fc2(self.drop(self.act(self.fc1(x)))))


class _ResidualBlock(nn.Module):
    """Shared stream plumbing of Outlooker / Transformer: forward_stream carries (x, pending branch, its scale)."""

    def set_sample_config(self, is_identity_layer=False):
        self.is_identity_layer = is_identity_layer

    def _fused_ok(self):
        return isinstance(self.mlp.act, GELU) and self.mlp.drop.p == 0.

    def forward(self, x):
        if getattr(self, 'is_identity_layer', False):
            return x
        x1, z, rs = self.forward_stream(x, None, None)
        return ops.ResidualAddFn.apply(x1, z, rs, x1.dtype)

    def forward_stream(self, x, r, rs):
        if getattr(self, 'is_identity_layer', False):
            return x, r, rs
        if not self._fused_ok():
            raise NotImplementedError('fused blocks need GELU activation and drop=0 (all reference configs)')
        B = x.shape[0]
        rs_attn = _drop_scale(self.drop_path, B, x.device)
        rs_mlp = _drop_scale(self.drop_path, B, x.device)
        x1, z = self._fused(x, r, rs, rs_attn)
        return x1, z, rs_mlp


class Outlooker(_ResidualBlock):
    """models/volo.py:106-144."""

    def __init__(self, dim, kernel_size, padding, stride=1, num_heads=1, mlp_ratio=3., attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm, qkv_bias=False, qk_scale=None):
        super().__init__()
        self.norm1 = _make_norm(norm_layer, dim)
        self.attn = OutlookAttention(dim, num_heads, kernel_size=kernel_size, padding=padding, stride=stride,
                                     qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = _make_norm(norm_layer, dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)

    def _fused(self, x, r, rs, rs_attn):
        a = self.attn
        a._check(x.shape[-1])
        if a.v.bias is not None:
            raise NotImplementedError('qkv_bias=True is not used by the reference VOLO configs')
        return ops.OutlookerFn.apply(x, r, rs, rs_attn, a.num_heads, self.norm1.eps, self.norm1.weight, self.norm1.bias,
                                     a.v.weight, a.attn.weight, a.attn.bias, a.proj.weight, a.proj.bias,
                                     self.norm2.weight, self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias,
                                     self.mlp.fc2.weight, self.mlp.fc2.bias)


class Attention(nn.Module):
    """models/volo.py:170-201."""

    def __init__(self, dim, num_heads=8, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.scale = qk_scale or (dim // num_heads) ** -0.5
        self.qkv = Linear(dim, dim * 3, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = Linear(dim, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        shp = x.shape
        B, Cc = shp[0], shp[-1]
        qkv = self.qkv(x).reshape(B, -1, 3 * Cc)
        o = ops.MhsaCoreFn.apply(qkv, self.num_heads, self.scale)
        return self.proj(o).reshape(shp)


class Transformer(_ResidualBlock):
    """models/volo.py:204-234."""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, attn_drop=0., drop_path=0.,
                 act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = _make_norm(norm_layer, dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = _make_norm(norm_layer, dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer)

    def _fused(self, x, r, rs, rs_attn):
        a = self.attn
        return ops.TransformerFn.apply(x, r, rs, rs_attn, a.num_heads, self.norm1.eps, self.norm1.weight, self.norm1.bias,
                                       a.qkv.weight, a.qkv.bias, a.proj.weight, a.proj.bias, self.norm2.weight,
                                       self.norm2.bias, self.mlp.fc1.weight, self.mlp.fc1.bias, self.mlp.fc2.weight,
                                       self.mlp.fc2.bias)


class ClassAttention(nn.Module):
    """models/volo.py:237-277: the cls token queries all tokens."""

    def __init__(self, dim, num_heads=8, head_dim=None, qkv_bias=False, qk_scale=None, attn_drop=0., proj_drop=0.):
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = head_dim if head_dim is not None else dim // num_heads
        self.scale = qk_scale or self.head_dim ** -0.5
        inner = self.head_dim * num_heads
        self.kv = Linear(dim, inner * 2, bias=qkv_bias)
        self.q = Linear(dim, inner, bias=qkv_bias)
        self.attn_drop = nn.Dropout(attn_drop)
        self.proj = Linear(inner, dim)
        self.proj_drop = nn.Dropout(proj_drop)

    def forward(self, x):
        B = x.shape[0]
        kv = self.kv(x)
        q = self.q(x[:, 0])
        o = ops.ClassAttnCoreFn.apply(q, kv, self.num_heads, self.scale)
        return self.proj(o).reshape(B, 1, -1)

    def forward_pair(self, n_cls, n_tok):
        """Same arithmetic with the (normalised) class token [B,1,C] and patch tokens [B,N,C] kept in two tensors: the k / v
        rows of both come from the same Linear, the attention kernel reads key 0 from one buffer and keys 1.. from the other."""
        B = n_tok.shape[0]
        kv_tok = self.kv(n_tok)
        kv_cls = self.kv(n_cls).reshape(B, -1)
        q = self.q(n_cls.reshape(B, -1))
        o = ops.ClassAttnCoreSplitFn.apply(q, kv_cls, kv_tok, self.num_heads, self.scale)
        return self.proj(o).reshape(B, 1, -1)


class ClassBlock(nn.Module):
    """models/volo.py:280-308: only the cls row is updated."""

    def __init__(self, dim, num_heads, head_dim=None, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm):
        super().__init__()
        self.norm1 = _make_norm(norm_layer, dim)
        self.attn = ClassAttention(dim, num_heads=num_heads, head_dim=head_dim, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                   attn_drop=attn_drop, proj_drop=drop)
        self.drop_path = DropPath(drop_path) if drop_path > 0. else nn.Identity()
        self.norm2 = _make_norm(norm_layer, dim)
        self.mlp = Mlp(in_features=dim, hidden_features=int(dim * mlp_ratio), act_layer=act_layer, drop=drop)

    def forward(self, x):
        # same arithmetic as the reference (cls = x[:, :1]; ...; cat([cls, x[:, 1:]])) with the slice / cat glue done by
        # two small autograd Functions: the generic slice / cat backward zero-fills and re-adds [B, 1+N, C] tensors
        # several times per block for what is a one-row update
        cls, xt = _SplitCls.apply(x)
        cls = cls + self.drop_path(self.attn(self.norm1(xt))).to(cls.dtype)
        cls = cls + self.drop_path(self.mlp(self.norm2(cls))).to(cls.dtype)
        return _JoinCls.apply(cls, xt, 1)

    def forward_pair(self, cls, tok):
        """The reference's forward on (cls [B,1,C], tok [B,N,C]) instead of cat([cls, tok]): only the class row changes in a
        ClassBlock, so the model carries it separately and never copies the N patch tokens (models/volo.py:300-308 slices
        and re-concatenates them in every block; LayerNorm is row-wise, so norm1 on the two parts is the same arithmetic)."""
        cls = cls + self.drop_path(self.attn.forward_pair(self.norm1(cls), self.norm1(tok))).to(cls.dtype)
        cls = cls + self.drop_path(self.mlp(self.norm2(cls))).to(cls.dtype)
        return cls


class _SplitCls(torch.autograd.Function):
    """x [B, 1+N, C] -> (copy of row 0, x itself).  Backward adds the cls-row gradient into row 0 of the gradient that
    arrives for the pass-through output (a fresh tensor produced by autograd's accumulation or by _JoinCls)."""

    @staticmethod
    def forward(ctx, x):
        ctx.rows = x.shape[1]
        return x[:, :1].clone(), x.view_as(x)

    @staticmethod
    def backward(ctx, dcls, dx):
        if dx is None:
            dx = torch.zeros((dcls.shape[0], ctx.rows, dcls.shape[2]), device=dcls.device, dtype=dcls.dtype)
        if dcls is not None:
            dx[:, :1] += dcls
        return dx


class _JoinCls(torch.autograd.Function):
    """out = [cls ; src[:, off:]] along dim 1 (off = 1: src still carries its old cls row, off = 0: src are tokens only)."""

    @staticmethod
    def forward(ctx, cls, src, off):
        ctx.off = off
        B, n, C = src.shape
        out = torch.empty((B, 1 + n - off, C), device=src.device, dtype=src.dtype)
        out[:, :1] = cls
        out[:, 1:] = src[:, off:]
        return out

    @staticmethod
    def backward(ctx, dout):
        dcls = dout[:, :1].clone()
        if ctx.off == 0:
            return dcls, dout[:, 1:], None
        dsrc = dout.clone()
        dsrc[:, :1].zero_()
        return dcls, dsrc, None


def _make_norm(norm_layer, dim):
    return LayerNorm(dim) if norm_layer is nn.LayerNorm else norm_layer(dim)


def get_block(block_type, **kargs):
    if block_type == 'ca':
        return ClassBlock(**kargs)
    raise ValueError(block_type)


def rand_bbox(size, lam, scale=1):
    """models/volo.py:319-339 (numpy host RNG: draws cx then cy)."""
    gw, gh = size[1] // scale, size[2] // scale
    cut = np.sqrt(1. - lam)
    cw, ch = int(gw * cut), int(gh * cut)
    cx = np.random.randint(gw)
    cy = np.random.randint(gh)
    return (np.clip(cx - cw // 2, 0, gw), np.clip(cy - ch // 2, 0, gh),
            np.clip(cx + cw // 2, 0, gw), np.clip(cy + ch // 2, 0, gh))


class PatchEmbed(nn.Module):
    """models/volo.py:342-380.  bf16 mode: the 7x7 RGB stem conv is im2col + our tcgen05 GEMM (ops.StemConvFn), the two
    3x3 64->64 convs run through cuDNN (library implicit GEMM), BatchNorm+ReLU is our fused kernel, and the patch
    projection conv (kernel == stride) is patchify + our GEMM.  fp32 parity mode: all three stem convs via cuDNN."""

    def __init__(self, img_size=224, stem_conv=False, stem_stride=1, patch_size=8, in_chans=3, hidden_dim=64,
                 embed_dim=384):
        super().__init__()
        assert patch_size in [4, 8, 16]
        self.stem_conv = stem_conv
        if stem_conv:
            layers = [nn.Conv2d(in_chans, hidden_dim, kernel_size=7, stride=stem_stride, padding=3, bias=False),
                      nn.BatchNorm2d(hidden_dim), nn.ReLU(inplace=True)]
            for _ in range(2):
                layers += [nn.Conv2d(hidden_dim, hidden_dim, kernel_size=3, stride=1, padding=1, bias=False),
                           nn.BatchNorm2d(hidden_dim), nn.ReLU(inplace=True)]
            self.conv = nn.Sequential(*layers)
        self.proj = nn.Conv2d(hidden_dim, embed_dim, kernel_size=patch_size // stem_stride,
                              stride=patch_size // stem_stride)
        self.num_patches = (img_size // patch_size) ** 2

    def _stem(self, x):
        """conv (cuDNN) -> fused BatchNorm+ReLU kernel, three times; NCHW logical / channels-last physical."""
        mods = list(self.conv)
        i = 0
        while i < len(mods):
            m = mods[i]
            fuse = (isinstance(m, nn.Conv2d) and i + 2 < len(mods) and isinstance(mods[i + 1], nn.BatchNorm2d)
                    and isinstance(mods[i + 2], nn.ReLU) and K.bn_supported(mods[i + 1].num_features)
                    and mods[i + 1].affine and mods[i + 1].track_running_stats and mods[i + 1].momentum is not None)
            if not fuse:
                x = m(x.contiguous(memory_format=torch.channels_last) if isinstance(m, nn.Conv2d) else x)
                i += 1
                continue
            bn = mods[i + 1]
            nhwc = None
            if (torch.is_autocast_enabled('cuda') and not x.requires_grad and m.bias is None and m.groups == 1
                    and m.dilation == (1, 1) and m.in_channels <= 4 and m.stride[0] == m.stride[1]
                    and m.padding[0] == m.padding[1] and isinstance(m.padding[0], int)
                    and K.tc_supported(8, m.out_channels, 8)):
                # few input channels (the 7x7 stem conv on RGB): im2col + tcgen05 GEMM instead of the library conv
                nhwc = ops.StemConvFn.apply(x, m.weight, m.stride[0], m.padding[0])
            elif (not torch.is_autocast_enabled('cuda') and x.dtype == torch.float32 and m.bias is None and m.groups == 1
                  and m.dilation == (1, 1) and isinstance(m.padding[0], int)):
                # fp32 parity mode: library convolution with the (ill-conditioned) weight gradient accumulated in fp64
                x = ops.ParityConvFn.apply(x.contiguous(memory_format=torch.channels_last), m.weight, m.stride, m.padding)
            else:
                x = m(x.contiguous(memory_format=torch.channels_last))
            use_batch = bn.training
            if use_batch:
                with torch.no_grad():
                    bn.num_batches_tracked += 1
            y = ops.BNReLUFn.apply(nhwc if nhwc is not None else x.permute(0, 2, 3, 1), bn.weight, bn.bias,
                                   bn.running_mean, bn.running_var, float(bn.momentum), float(bn.eps), use_batch)
            x = y.permute(0, 3, 1, 2)
            i += 3
        return x

    def forward_nhwc(self, x):
        if self.stem_conv:
            if torch.is_autocast_enabled('cuda'):
                x = self._stem(x)
            else:   # fp32 parity mode: keep cuDNN off TF32
                with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                    x = self._stem(x)
        x = x.permute(0, 2, 3, 1)   # free for channels_last
        return ops.PatchConvFn.apply(x, self.proj.weight, self.proj.bias, self.proj.kernel_size[0])

    def forward(self, x):
        return self.forward_nhwc(x).permute(0, 3, 1, 2)   # reference returns B, C, H, W


class Downsample(nn.Module):
    """models/volo.py:383-396: conv p x p stride p between the stages, NHWC in / NHWC out."""

    def __init__(self, in_embed_dim, out_embed_dim, patch_size):
        super().__init__()
        self.proj = nn.Conv2d(in_embed_dim, out_embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        return ops.PatchConvFn.apply(x, self.proj.weight, self.proj.bias, self.proj.kernel_size[0])


def _stage_rates(index, layers, drop_path_rate):
    return [drop_path_rate * (i + sum(layers[:index])) / (sum(layers) - 1) for i in range(layers[index])]


def outlooker_blocks(block_fn, index, dim, layers, num_heads=1, kernel_size=3, padding=1, stride=1, mlp_ratio=3.,
                     qkv_bias=False, qk_scale=None, attn_drop=0, drop_path_rate=0., **kwargs):
    """models/volo.py:399-417."""
    return nn.Sequential(*[block_fn(dim, kernel_size=kernel_size, padding=padding, stride=stride, num_heads=num_heads,
                                    mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop,
                                    drop_path=dpr) for dpr in _stage_rates(index, layers, drop_path_rate)])


def transformer_blocks(block_fn, index, dim, layers, num_heads, mlp_ratio=3., qkv_bias=False, qk_scale=None,
                       attn_drop=0, drop_path_rate=0., **kwargs):
    """models/volo.py:420-441."""
    return nn.Sequential(*[block_fn(dim, num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                                    attn_drop=attn_drop, drop_path=dpr)
                           for dpr in _stage_rates(index, layers, drop_path_rate)])


class VOLO(nn.Module):
    """models/volo.py:444-694 (constructor arguments, attributes and outputs identical)."""

    def __init__(self, layers, img_size=224, in_chans=3, num_classes=1000, patch_size=8, stem_hidden_dim=64,
                 embed_dims=None, num_heads=None, downsamples=None, outlook_attention=None, mlp_ratios=None,
                 qkv_bias=False, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=nn.LayerNorm, post_layers=None, return_mean=False, return_dense=True, mix_token=True,
                 pooling_scale=2, out_kernel=3, out_stride=2, out_padding=1):
        super().__init__()
        self.num_classes = num_classes
        self.patch_embed = PatchEmbed(stem_conv=True, stem_stride=2, patch_size=patch_size, in_chans=in_chans,
                                      hidden_dim=stem_hidden_dim, embed_dim=embed_dims[0])
        grid = img_size // patch_size // pooling_scale
        self.pos_embed = nn.Parameter(torch.zeros(1, grid, grid, embed_dims[-1]))
        self.pos_drop = nn.Dropout(p=drop_rate)

        network = []
        for i in range(len(layers)):
            if outlook_attention[i]:
                # NB: like the reference (:495-500) the outlooker stage is built WITHOUT drop_path_rate
                network.append(outlooker_blocks(Outlooker, i, embed_dims[i], layers, downsample=downsamples[i],
                                                num_heads=num_heads[i], kernel_size=out_kernel, stride=out_stride,
                                                padding=out_padding, mlp_ratio=mlp_ratios[i], qkv_bias=qkv_bias,
                                                qk_scale=qk_scale, attn_drop=attn_drop_rate, norm_layer=norm_layer))
            else:
                network.append(transformer_blocks(Transformer, i, embed_dims[i], layers, num_heads[i],
                                                  mlp_ratio=mlp_ratios[i], qkv_bias=qkv_bias, qk_scale=qk_scale,
                                                  drop_path_rate=drop_path_rate, attn_drop=attn_drop_rate,
                                                  norm_layer=norm_layer))
            if downsamples[i]:
                network.append(Downsample(embed_dims[i], embed_dims[i + 1], 2))
        self.network = nn.ModuleList(network)

        self.post_network = None
        if post_layers is not None:
            self.post_network = nn.ModuleList([
                get_block(post_layers[i], dim=embed_dims[-1], num_heads=num_heads[-1], mlp_ratio=mlp_ratios[-1],
                          qkv_bias=qkv_bias, qk_scale=qk_scale, attn_drop=attn_drop_rate, drop_path=0.,
                          norm_layer=norm_layer) for i in range(len(post_layers))])
            self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dims[-1]))
            nn.init.trunc_normal_(self.cls_token, std=.02)

        self.return_mean = return_mean
        self.return_dense = return_dense
        if return_dense:
            assert not return_mean, "cannot return both mean and dense"
        self.mix_token = mix_token
        self.pooling_scale = pooling_scale
        if mix_token:
            self.beta = 1.0
            assert return_dense, "return all tokens if mix_token is enabled"
        if return_dense:
            self.aux_head = Linear(embed_dims[-1], num_classes) if num_classes > 0 else nn.Identity()
        self.norm = _make_norm(norm_layer, embed_dims[-1])
        self.embed_dim = embed_dims[-1]
        self.head = Linear(embed_dims[-1], num_classes) if num_classes > 0 else nn.Identity()

        nn.init.trunc_normal_(self.pos_embed, std=.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    @torch.jit.ignore
    def no_weight_decay(self):
        return {'pos_embed', 'cls_token'}

    def get_classifier(self):
        return self.head

    def reset_classifier(self, num_classes):
        self.num_classes = num_classes
        self.head = Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def interpolate_pos_encoding(self, x):
        """models/volo.py:580-596: pos_embed resized (bicubic, scale_factor=(h0+0.1)/h) to x's grid; fp32 [1,h0,w0,C]."""
        _, h0, w0, Cc = x.shape
        h, w = self.pos_embed.shape[1], self.pos_embed.shape[2]
        if h == h0 and w == w0:
            return self.pos_embed
        zero = torch.zeros(1, h0, w0, Cc, device=self.pos_embed.device, dtype=torch.float32)
        return ops.PosEmbedAddFn.apply(zero, self.pos_embed)

    def set_sample_config(self, config: dict):
        """models/volo.py:598-616: flag the not-yet-grown layers of the super-net as identity."""
        split = lambda n: (lambda a: [a, n - a, 0, 0])(make_divisible(n * 0.23, 2))
        cur, lo, hi = split(config['layer_num']), split(config['min_layer_num']), split(config['max_layer_num'])
        skip = []
        for i in range(4):
            fresh = get_new_layer_idx(prev_l=lo[i], new_l=hi[i])
            grown = cur[i] - lo[i]
            skip.append(fresh if grown == 0 else fresh[:-grown])
        real = 0
        for stage in self.network:
            if isinstance(stage, (nn.Sequential, nn.ModuleList)):
                for li, blk in enumerate(stage):
                    blk.set_sample_config(is_identity_layer=li in skip[real])
                real += 1

    # ---- forward pieces (names as in the reference, :618-642) -------------------------------
    def forward_embeddings(self, x):
        return self.patch_embed.forward_nhwc(x)

    def _flush(self, x, r, rs, dtype=None):
        if r is None:
            return x if dtype is None or x.dtype == dtype else ops.CastFn.apply(x, dtype)
        return ops.ResidualAddFn.apply(x, r, rs, dtype or x.dtype)

    def forward_tokens(self, x):
        x = ops.CastFn.apply(x, torch.float32) if x.dtype != torch.float32 else x    # residual stream is fp32
        r = rs = None
        for idx, stage in enumerate(self.network):
            if idx == 2:
                x = ops.PosEmbedAddFn.apply(self._flush(x, r, rs), self.pos_embed)
                r = rs = None
                x = self.pos_drop(x)
            if isinstance(stage, nn.Sequential) and all(isinstance(b, _ResidualBlock) for b in stage):
                for blk in stage:
                    x, r, rs = blk.forward_stream(x, r, rs)
            else:
                x = stage(self._flush(x, r, rs))
                r = rs = None
        x = self._flush(x, r, rs)
        return x.reshape(x.shape[0], -1, x.shape[-1])

    def forward_cls(self, x):
        cls_tokens = self.cls_token.expand(x.shape[0], -1, -1).to(x.dtype)
        x = _JoinCls.apply(cls_tokens, x, 0)
        for block in self.post_network:
            x = block(x)
        return x

    def forward_cls_pair(self, tok):
        """forward_cls without ever building cat([cls, tokens]): returns the final class token [B,1,C]."""
        cls = self.cls_token.expand(tok.shape[0], -1, -1).to(tok.dtype).contiguous()
        for block in self.post_network:
            cls = block.forward_pair(cls, tok)
        return cls

    def forward(self, x):
        if not x.is_cuda:
            raise RuntimeError('autoprog_b200.VOLO runs on CUDA (sm_100a) only; there is no CPU fallback')
        if self.training:
            pool = self.__dict__.get('_dp_pool')
            if pool is None:
                pool = self.__dict__['_dp_pool'] = DropPathPool(self)
            pool.draw(x.shape[0], x.device)
        x = self.forward_embeddings(x)

        graph_box = getattr(self, '_graph_box', None)      # set by graph.GraphedTrainStep: device-resident bbox
        if self.mix_token and self.training and graph_box is not None:
            patch_h, patch_w = x.shape[1] // self.pooling_scale, x.shape[2] // self.pooling_scale
            bbx1, bby1, bbx2, bby2 = (int(v) for v in self._graph_box_host.tolist())
            x = ops.FlipInBoxDevFn.apply(x, graph_box, self.pooling_scale)
        elif self.mix_token and self.training:
            lam = np.random.beta(self.beta, self.beta)
            patch_h, patch_w = x.shape[1] // self.pooling_scale, x.shape[2] // self.pooling_scale
            bbx1, bby1, bbx2, bby2 = (int(v) for v in rand_bbox(x.size(), lam, scale=self.pooling_scale))
            s = self.pooling_scale
            x = ops.FlipInBoxFn.apply(x, (s * bbx1, s * bby1, s * bbx2, s * bby2))
        else:
            bbx1, bby1, bbx2, bby2 = 0, 0, 0, 0

        x = self.forward_tokens(x)
        if (self.post_network is not None and not self.return_mean
                and all(isinstance(b, ClassBlock) for b in self.post_network)):
            # class token and patch tokens stay two tensors through the class blocks, the final norm and the two heads
            cls = self.forward_cls_pair(x)
            x_cls = self.head(self.norm(cls).reshape(cls.shape[0], -1))
            if not self.return_dense:
                return x_cls
            x_aux = self.aux_head(self.norm(x))
        else:
            if self.post_network is not None:
                x = self.forward_cls(x)
            x = self.norm(x)
            if self.return_mean:
                return self.head(x.float().mean(1).to(x.dtype))
            x_cls = self.head(x[:, 0])
            if not self.return_dense:
                return x_cls
            x_aux = self.aux_head(x[:, 1:])
        if not self.training:
            return x_cls + 0.5 * x_aux.max(1)[0]
        if self.mix_token and self.training:
            B, _, ncls = x_aux.shape
            if graph_box is not None:
                x_aux = ops.FlipInBoxDevFn.apply(x_aux.reshape(B, patch_h, patch_w, ncls), graph_box, 1)
                x_aux = x_aux.reshape(B, patch_h * patch_w, ncls)
                return x_cls, x_aux, ops.DevBox((bbx1, bby1, bbx2, bby2), graph_box)
            x_aux = ops.FlipInBoxFn.apply(x_aux.reshape(B, patch_h, patch_w, ncls), (bbx1, bby1, bbx2, bby2))
            x_aux = x_aux.reshape(B, patch_h * patch_w, ncls)
        return x_cls, x_aux, (bbx1, bby1, bbx2, bby2)


_VARIANTS = {
    #          layers            embed_dims              heads             mlp  stem  cfg
    'volo_d1': ([4, 4, 8, 2], [192, 384, 384, 384], [6, 12, 12, 12], 3, 64, 'volo'),          # models/volo.py:697-727
    'volo_d2': ([6, 4, 10, 4], [256, 512, 512, 512], [8, 16, 16, 16], 3, 64, 'volo'),         # :730-750
    'volo_d3': ([8, 8, 16, 4], [256, 512, 512, 512], [8, 16, 16, 16], 3, 64, 'volo'),         # :753-773
    'volo_d4': ([8, 8, 16, 4], [384, 768, 768, 768], [12, 16, 16, 16], 3, 64, 'volo_large'),  # :776-796
    'volo_d5': ([12, 12, 20, 4], [384, 768, 768, 768], [12, 16, 16, 16], 4, 128, 'volo_large'),  # :799-821
}


def build_volo(layers, embed_dims, num_heads, mlp_ratio, stem_hidden_dim=64, cfg='volo', **kwargs):
    kwargs.pop('pretrained', None)
    if stem_hidden_dim != 64:
        kwargs.setdefault('stem_hidden_dim', stem_hidden_dim)
    model = VOLO(list(layers), embed_dims=list(embed_dims), num_heads=list(num_heads), mlp_ratios=[mlp_ratio] * 4,
                 downsamples=[True, False, False, False], outlook_attention=[True, False, False, False],
                 post_layers=['ca', 'ca'], **kwargs)
    model.default_cfg = default_cfgs[cfg]
    return model


def _factory(name):
    spec = _VARIANTS[name]

    def fn(pretrained=False, **kwargs):
        return build_volo(*spec[:5], cfg=spec[5], **kwargs)
    fn.__name__ = name
    fn.__doc__ = f'{name}: layers {spec[0]}, dims {spec[1]}, heads {spec[2]} (reference models/volo.py factories).'
    return register_model(fn)


volo_d1, volo_d2, volo_d3, volo_d4, volo_d5 = (_factory(n) for n in _VARIANTS)
