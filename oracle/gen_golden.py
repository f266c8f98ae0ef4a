"""Generate tests/golden/*.pt|json by running the REFERENCE's own PyTorch code on CPU (fp64).

Run inside the build container only (needs /root/reference; it does not exist on the GPU box):

    python oracle/gen_golden.py

The reference has no tests or golden vectors of its own, so these fixtures --
outputs of the unmodified reference modules on seeded inputs -- are what pins the oracle
(`oracle/volo_cpu.py`) and, through it, the CUDA path.  Weights/inputs are stored in fp32 and
up-cast to fp64 before the reference runs, so the stored values are exact inputs.
"""
import json
import os
import sys
from types import SimpleNamespace

import numpy as np

np.int = int  # models/volo.py:327-328 uses the alias removed in numpy>=1.24
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, 'ref_shim'), '/root/reference']

import torch  # noqa: E402
import torch.nn as nn  # noqa: E402

import models.volo as RV  # noqa: E402  (reference)
import loss.cross_entropy as RL  # noqa: E402  (reference)
import prog.progressive as RP  # noqa: E402  (reference)
import prog.helpers as RH  # noqa: E402  (reference)
from timm.models.layers import DropPath  # noqa: E402  (shim)

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)


def f32(t):
    return t.detach().to(torch.float32).clone()


def requantize_(module):
    """Round parameters to fp32-representable values so fixtures can store them in fp32 exactly."""
    with torch.no_grad():
        for p in list(module.parameters()) + list(module.buffers()):
            if p.dtype.is_floating_point:
                p.copy_(p.float().double())


def gen_outlook():
    cases = {}
    for seed, (name, (B, H, W, heads)) in enumerate({'even_8x8': (2, 8, 8, 2), 'odd_9x7': (2, 9, 7, 2),
                                                     'odd_5x6_h3': (1, 5, 6, 3)}.items()):
        torch.manual_seed(100 + seed)
        dim = 32 * heads
        m = RV.OutlookAttention(dim, heads, kernel_size=3, padding=1, stride=2).double()
        with torch.no_grad():
            for p in m.parameters():
                p.copy_(torch.randn_like(p) * 0.3)
        requantize_(m)
        x = torch.randn(B, H, W, dim).float().double().requires_grad_(True)
        dy = torch.randn(B, H, W, dim).float().double()
        y = m(x)
        y.backward(dy)
        cases[name] = dict(heads=heads, sd={k: f32(v) for k, v in m.state_dict().items()}, x=f32(x), dy=f32(dy),
                           y=y.detach().clone(), dx=f32(x.grad),
                           grads={k: f32(p.grad) for k, p in m.named_parameters()})
    torch.save(cases, os.path.join(OUT, 'outlook_attention.pt'))


def small_volo(img_size, layers=(2, 3, 0, 0), h=2, dpr=0.0, num_classes=12):
    torch.manual_seed(7)
    m = RV.VOLO(list(layers), img_size=img_size, num_classes=num_classes, stem_hidden_dim=8,
                embed_dims=[h * 16, h * 32, h * 32, h * 32], num_heads=[h // 2, h, h, h], mlp_ratios=[3, 3, 3, 3],
                downsamples=[True, False, False, False], outlook_attention=[True, False, False, False],
                post_layers=['ca', 'ca'], drop_path_rate=dpr).double()
    with torch.no_grad():   # spread the weights so every path carries signal
        for n, p in m.named_parameters():
            if p.dim() > 1:
                p.mul_(6.0)
            elif 'bias' in n:
                p.copy_(torch.randn_like(p) * 0.1)
            else:
                p.copy_(1 + 0.2 * torch.randn_like(p))
    requantize_(m)
    return m


def run_volo_case(m, r, B, train, seed, dense_weight=0.5, sample_cfg=None, C=12):
    torch.manual_seed(seed)
    np.random.seed(seed)
    x = torch.randn(B, 3, r, r).float().double().requires_grad_(True)
    m.train(train)
    if sample_cfg is not None:
        m.set_sample_config(sample_cfg)
    recs = {}
    for n, mod in m.named_modules():
        if isinstance(mod, DropPath):
            mod.record = []
            recs[n] = mod
    m.zero_grad()
    out = m(x)
    case = dict(r=r, train=train, x=f32(x), sample_cfg=sample_cfg)
    if not train:
        case['out'] = out.detach().clone()
        return case
    x_cls, x_aux, bbox = out
    N = x_aux.shape[1]
    target = torch.softmax(torch.randn(B, C, 2 + N) * 2, dim=1).float().double()
    crit = RL.TokenLabelCrossEntropy(dense_weight=dense_weight, cls_weight=1.0, classes=C)
    loss = crit(out, target)
    loss.backward()
    case.update(bbox=[int(b) for b in bbox], x_cls=x_cls.detach().clone(), x_aux=x_aux.detach().clone(),
                target=f32(target), loss=loss.detach().clone(), dx=f32(x.grad), dense_weight=dense_weight,
                grads={k: (f32(p.grad) if p.grad is not None else None) for k, p in m.named_parameters()},
                drop_masks={n.replace('.drop_path', '.'): [t.clone() for t in mod.record] for n, mod in recs.items()},
                drop_keep={n.replace('.drop_path', '.'): 1 - mod.drop_prob for n, mod in recs.items()})
    return case


def gen_volo():
    m = small_volo(img_size=64)
    fx = dict(arch=dict(layers=[2, 3, 0, 0], embed_dims=[32, 64, 64, 64], num_heads=[1, 2, 2, 2], stem_hidden=8,
                        img_size=64, num_classes=12),
              sd={k: f32(v) for k, v in m.state_dict().items()}, cases={})
    # eval first: train-mode forwards update the BatchNorm running stats, and `sd` is the pre-run snapshot
    with torch.no_grad():   # non-trivial running stats for the eval path
        for n, b in m.named_buffers():
            if n.endswith('running_mean'):
                b.copy_((torch.randn_like(b) * 0.1).float().double())
            elif n.endswith('running_var'):
                b.copy_((1 + 0.3 * torch.rand_like(b)).float().double())
    fx['sd'] = {k: f32(v) for k, v in m.state_dict().items()}
    fx['cases']['eval_r80'] = run_volo_case(m, 80, 2, False, 14)
    fx['cases']['train_r64'] = run_volo_case(m, 64, 4, True, 11)
    fx['cases']['train_r96_bicubic'] = run_volo_case(m, 96, 3, True, 12)
    fx['cases']['train_r104_oddgrid'] = run_volo_case(m, 104, 2, True, 13)
    torch.save(fx, os.path.join(OUT, 'volo_small.pt'))

    # elastic depth + DropPath: super-net of depth 6 ([2,4]) run as a depth-4 sub-net (min 3, max 6)
    m = small_volo(img_size=64, layers=(2, 4, 0, 0), dpr=0.3)
    fx = dict(arch=dict(layers=[2, 4, 0, 0], embed_dims=[32, 64, 64, 64], num_heads=[1, 2, 2, 2], stem_hidden=8,
                        img_size=64, num_classes=12, drop_path_rate=0.3),
              sd={k: f32(v) for k, v in m.state_dict().items()}, cases={})
    cfg = {'layer_num': 5, 'min_layer_num': 4, 'max_layer_num': 6}
    fx['cases']['train_r64_elastic_dp'] = run_volo_case(m, 64, 4, True, 21, sample_cfg=cfg)
    fx['identity_flags'] = {n: bool(getattr(mod, 'is_identity_layer', False)) for n, mod in m.named_modules()
                            if hasattr(mod, 'set_sample_config') and n}
    torch.save(fx, os.path.join(OUT, 'volo_small_elastic.pt'))


def gen_losses():
    torch.manual_seed(3)
    B, N, C = 4, 9, 10
    x_cls = (torch.randn(B, C) * 2).float().double()
    x_aux = (torch.randn(B, N, C) * 2).float().double()
    t3 = torch.softmax(torch.randn(B, C, 2 + N), 1).float().double() * 1.3   # need not sum to one
    t2 = torch.softmax(torch.randn(B, C), 1).float().double()
    fx = dict(x_cls=x_cls, x_aux=x_aux, t3=t3, t2=t2, cases={})
    for bbox in [(0, 0, 0, 0), (0, 1, 2, 3), (0, 0, 3, 3)]:
        for tname, t in (('t3', t3), ('t2', t2)):
            for wd, wc in ((0.5, 1.0), (1.0, 0.0)):
                xc = x_cls.clone().requires_grad_(True)
                xa = x_aux.clone().requires_grad_(True)
                l = RL.TokenLabelCrossEntropy(dense_weight=wd, cls_weight=wc, classes=C)((xc, xa, bbox), t)
                l.backward()
                fx['cases'][f'tlce|{bbox}|{tname}|{wd}|{wc}'] = dict(loss=l.detach(), dcls=xc.grad, daux=xa.grad)
        l = RL.TokenLabelGTCrossEntropy(dense_weight=0.5, cls_weight=1.0, classes=C)((x_cls, x_aux, bbox), t3)
        fx['cases'][f'gt|{bbox}'] = dict(loss=l)
    fx['cases']['soft'] = dict(loss=RL.SoftTargetCrossEntropy()(x_cls, t2))
    fx['cases']['soft_rep'] = dict(loss=RL.SoftTargetCrossEntropy()(x_aux.reshape(-1, C)[:8], t2))
    fx['cases']['tlsoft'] = dict(loss=RL.TokenLabelSoftTargetCrossEntropy()(x_cls, t3[:, :, :2]))
    torch.save(fx, os.path.join(OUT, 'losses.pt'))


def gen_tables():
    args = SimpleNamespace(num_stages=4, r_scale=0.5, h_scale=1., l_scale=0.5, aa_scale=0.5, dp_scale=0., re_scale=0.,
                           resize_scale=[1., 1.], aa='rand-m9-mstd0.5-inc1', drop_path=0.1, reprob=0.25,
                           scale=[0.08, 1.0], epochs=100)
    e, r, h, l, aa, dp, re, rs = RP.progressive_schedule(args, r_max=224, h_max=12, l_max=18)
    tab = {'train_autoprog_sh': dict(e=e, r=r, h=h, l=l, aa=aa, dp=dp, re=re, resize=rs)}
    args2 = SimpleNamespace(**{**vars(args), 'num_stages': 3, 'r_scale': 0.4, 'l_scale': 0.34, 'dp_scale': -0.5,
                               're_scale': -0.5, 'epochs': 300, 'aa_scale': 0.})
    e, r, h, l, aa, dp, re, rs = RP.progressive_schedule(args2, r_max=384, h_max=16, l_max=24)
    tab['alt'] = dict(e=e, r=r, h=h, l=l, aa=aa, dp=dp, re=re, resize=rs)
    tab['make_divisible'] = [[v, d, RP.make_divisible(v, d)] for v in (1, 2.07, 3.5, 4.14, 9, 15, 17.9, 100, 112, 149.3, 224)
                             for d in (1, 2, 8, 32)]
    tab['new_idx'] = {f'{p}->{n}': [RH.new_idx(i, p, n) for i in range(n)] for p, n in
                      ((2, 4), (7, 14), (7, 8), (7, 11), (4, 4), (9, 18), (3, 5), (2, 3), (4, 7))}
    tab['new_layer_idx'] = {f'{p}->{n}': RH.get_new_layer_idx(p, n) for p, n in
                            ((2, 4), (7, 14), (7, 8), (7, 11), (4, 4), (9, 18), (3, 5), (2, 3), (4, 7))}
    # identity-layer flags of the real super-net config (models/volo.py:598-616) via a tiny-width clone of the tree
    flags = {}
    for cur in (9, 12, 15, 18):
        m = RV.VOLO([4, 14, 0, 0], img_size=32, num_classes=2, stem_hidden_dim=4, embed_dims=[32, 32, 32, 32],
                    num_heads=[1, 1, 1, 1], mlp_ratios=[1, 1, 1, 1], downsamples=[True, False, False, False],
                    outlook_attention=[True, False, False, False], post_layers=['ca', 'ca'])
        m.set_sample_config({'layer_num': cur, 'min_layer_num': 9, 'max_layer_num': 18})
        flags[str(cur)] = [[i for i, b in enumerate(m.network[s]) if b.is_identity_layer] for s in (0, 2)]
    tab['identity_flags_9_18'] = flags
    # module tree / state_dict contract
    import models.submodels as RS
    ref = RV.volo_d1(img_size=224)
    tab['state_dict_volo_d1'] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    tab['modules_volo_d1'] = {n: type(m).__name__ for n, m in ref.named_modules()}
    ref = RS.model_variant(variant='volo_h12_l18', img_size=224, drop_path_rate=0.1)
    tab['state_dict_volo_h12_l18'] = {k: list(v.shape) for k, v in ref.state_dict().items()}
    tab['drop_path_volo_h12_l18'] = {n: m.drop_prob for n, m in ref.named_modules() if type(m).__name__ == 'DropPath'}
    json.dump(tab, open(os.path.join(OUT, 'tables.json'), 'w'), indent=1)


def gen_posembed():
    torch.manual_seed(5)
    m = RV.VOLO([1, 1, 0, 0], img_size=112, num_classes=2, stem_hidden_dim=4, embed_dims=[32, 8, 8, 8],
                num_heads=[1, 1, 1, 1], mlp_ratios=[1, 1, 1, 1], downsamples=[True, False, False, False],
                outlook_attention=[True, False, False, False], post_layers=None, return_dense=False, mix_token=False,
                return_mean=True).double()
    with torch.no_grad():
        m.pos_embed.copy_(torch.randn_like(m.pos_embed).float().double())
    fx = dict(pos=m.pos_embed.detach().clone(), out={})
    for g in ((4, 4), (5, 5), (6, 6), (8, 8), (10, 10), (12, 12), (14, 14), (24, 24), (9, 11)):
        probe = torch.zeros(1, g[0], g[1], 8, dtype=torch.float64)
        fx['out'][g] = m.interpolate_pos_encoding(probe).detach().clone()
    torch.save(fx, os.path.join(OUT, 'pos_embed.pt'))


if __name__ == '__main__':
    gen_outlook()
    gen_volo()
    gen_losses()
    gen_tables()
    gen_posembed()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
