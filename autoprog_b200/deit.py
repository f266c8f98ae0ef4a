"""DeiT on the autoprog_b200 kernels -- drop-in for the reference's `models/deit.py` factories.

The reference builds these from timm 0.4.5's `VisionTransformer` (models/deit.py:62-179), which is NOT vendored under
/root/reference, so the arithmetic is restated from the public package (patch16 conv -> cat cls -> + pos_embed ->
depth x [x += Attn(LN(x)); x += Mlp(LN(x))] -> LN -> head(x[:, 0]); LayerNorm eps 1e-6, qkv_bias=True, mlp_ratio 4) and
its parity is UNPINNED (oracle/volo_cpu.py:vit_forward carries the same note).  Parameter names follow timm
(`patch_embed.proj`, `cls_token`, `pos_embed`, `blocks.{i}.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}`, `norm`, `head`)
so the public DeiT checkpoints load.  Blocks are the same fused `Transformer` blocks VOLO's stage 2 uses.

Extensions that the reference only sketches (prog/helpers.py:753 'TODO: deit'): `set_sample_config` (elastic depth by
analogy with VOLO: the newest layers of the grown model are identity) and `return_dense=True` (token-labeling aux head).
"""
from __future__ import annotations

from functools import partial

import torch
import torch.nn as nn

from . import ops
from .helpers import get_new_layer_idx
from .registry import register_model
from .volo import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD, LayerNorm, Linear, Transformer

__all__ = ['deit_tiny_patch16_224', 'deit_small_patch16_224', 'deit_base_patch16_224', 'deit_tiny_distilled_patch16_224',
           'deit_small_distilled_patch16_224', 'deit_base_distilled_patch16_224', 'deit_base_patch16_384',
           'deit_base_distilled_patch16_384']


def _cfg(url='', **kwargs):
    cfg = dict(url=url, num_classes=1000, input_size=(3, 224, 224), pool_size=None, crop_pct=.9, interpolation='bicubic',
               mean=IMAGENET_DEFAULT_MEAN, std=IMAGENET_DEFAULT_STD, first_conv='patch_embed.proj', classifier='head')
    cfg.update(kwargs)
    return cfg


class PatchEmbed(nn.Module):
    """timm PatchEmbed: Conv2d(kernel = stride = patch) -> tokens; here patchify + tcgen05 GEMM."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768):
        super().__init__()
        self.img_size, self.patch_size = (img_size, img_size), (patch_size, patch_size)
        self.num_patches = (img_size // patch_size) ** 2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, x):
        x = x.permute(0, 2, 3, 1)                       # NCHW image -> NHWC view (made contiguous by the Function)
        y = ops.PatchConvFn.apply(x, self.proj.weight, self.proj.bias, self.patch_size[0])
        return y.reshape(y.shape[0], -1, y.shape[-1])   # [B, N, D]


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12, num_heads=12,
                 mlp_ratio=4., qkv_bias=True, qk_scale=None, drop_rate=0., attn_drop_rate=0., drop_path_rate=0.,
                 norm_layer=None, return_dense=False, n_prefix=1):
        super().__init__()
        norm_layer = norm_layer or partial(LayerNorm, eps=1e-6)
        self.num_classes, self.num_features, self.embed_dim = num_classes, embed_dim, embed_dim
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + n_prefix, embed_dim))
        self.pos_drop = nn.Dropout(p=drop_rate)
        dpr = [x.item() for x in torch.linspace(0, drop_path_rate, depth)]
        self.blocks = nn.ModuleList([
            Transformer(embed_dim, num_heads, mlp_ratio=mlp_ratio, qkv_bias=qkv_bias, qk_scale=qk_scale,
                        attn_drop=attn_drop_rate, drop_path=dpr[i], norm_layer=norm_layer) for i in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        self.return_dense = return_dense
        if return_dense:
            self.aux_head = Linear(embed_dim, num_classes) if num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.pos_embed, std=.02)
        nn.init.trunc_normal_(self.cls_token, std=.02)
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

    def reset_classifier(self, num_classes, global_pool=''):
        self.num_classes = num_classes
        self.head = Linear(self.embed_dim, num_classes) if num_classes > 0 else nn.Identity()

    def set_sample_config(self, config: dict):
        """Elastic depth (extension): layers that are new when growing min_layer_num -> max_layer_num and not yet
        reached by layer_num are identity, mirroring VOLO.set_sample_config (models/volo.py:598-616)."""
        fresh = get_new_layer_idx(prev_l=config['min_layer_num'], new_l=config['max_layer_num'])
        grown = config['layer_num'] - config['min_layer_num']
        skip = fresh if grown == 0 else fresh[:-grown]
        for i, blk in enumerate(self.blocks):
            blk.set_sample_config(is_identity_layer=i in skip)

    def _prefix(self, B):
        return [self.cls_token.expand(B, -1, -1)]

    def forward_features(self, x):
        if not x.is_cuda:
            raise RuntimeError('autoprog_b200 models run on CUDA (sm_100a) only; there is no CPU fallback')
        t = self.patch_embed(x)
        x = torch.cat([p.to(torch.float32) for p in self._prefix(t.shape[0])] + [t.to(torch.float32)], dim=1) + self.pos_embed
        x = self.pos_drop(x)
        r = rs = None
        for blk in self.blocks:
            x, r, rs = blk.forward_stream(x, r, rs)
        if r is not None:
            x = ops.ResidualAddFn.apply(x, r, rs, x.dtype)
        return self.norm(x)

    def forward(self, x):
        x = self.forward_features(x)
        out = self.head(x[:, 0])
        if self.return_dense:
            return out, self.aux_head(x[:, 1:])
        return out


class DistilledVisionTransformer(VisionTransformer):
    """models/deit.py:20-59: extra distillation token + head; eval returns the average of both heads."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, n_prefix=2, **kwargs)
        self.dist_token = nn.Parameter(torch.zeros(1, 1, self.embed_dim))
        self.head_dist = Linear(self.embed_dim, self.num_classes) if self.num_classes > 0 else nn.Identity()
        nn.init.trunc_normal_(self.dist_token, std=.02)
        self.head_dist.apply(self._init_weights)

    def _prefix(self, B):
        return [self.cls_token.expand(B, -1, -1), self.dist_token.expand(B, -1, -1)]

    def forward(self, x):
        x = self.forward_features(x)
        a, b = self.head(x[:, 0]), self.head_dist(x[:, 1])
        return (a, b) if self.training else (a + b) / 2


_SPECS = {  # name: (embed_dim, depth, heads, distilled, img_size)   -- models/deit.py:62-179
    'deit_tiny_patch16_224': (192, 12, 3, False, 224), 'deit_small_patch16_224': (384, 12, 6, False, 224),
    'deit_base_patch16_224': (768, 12, 12, False, 224), 'deit_tiny_distilled_patch16_224': (192, 12, 3, True, 224),
    'deit_small_distilled_patch16_224': (384, 12, 6, True, 224), 'deit_base_distilled_patch16_224': (768, 12, 12, True, 224),
    'deit_base_patch16_384': (768, 12, 12, False, 384), 'deit_base_distilled_patch16_384': (768, 12, 12, True, 384),
}


def _factory(name):
    dim, depth, heads, distilled, img = _SPECS[name]

    def fn(pretrained=False, **kwargs):
        if pretrained:
            raise RuntimeError('no network access: load a checkpoint with load_state_dict instead')
        kwargs.setdefault('img_size', img)
        cls = DistilledVisionTransformer if distilled else VisionTransformer
        model = cls(patch_size=16, embed_dim=dim, depth=depth, num_heads=heads, mlp_ratio=4, qkv_bias=True,
                    norm_layer=partial(LayerNorm, eps=1e-6), **kwargs)
        model.default_cfg = _cfg()
        return model
    fn.__name__ = name
    return register_model(fn)


for _n in _SPECS:
    globals()[_n] = _factory(_n)
