"""CPU-side checks: the C-ABI library loads and exports every symbol of include/autoprog_b200.h (no compute calls),
and the host logic (schedule, layer maps, registry, module tree) matches fixtures produced by the reference."""
import ctypes
import json
import os
import re
from types import SimpleNamespace

import pytest
import torch

import autoprog_b200 as A
from autoprog_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, 'tests', 'golden')
TAB = json.load(open(os.path.join(G, 'tables.json')))


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'autoprog_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(apb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), 'run `python -m autoprog_b200.build` (or __graft_entry__.build()) first'
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 35
    for s in syms:
        assert hasattr(lib, s), f'{s} declared in the header but not exported'
    assert set(_lib.SIGNATURES) == set(syms), set(_lib.SIGNATURES) ^ set(syms)
    assert _lib.lib().apb_abi_version() == 1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'autoprog_b200')
    for f in os.listdir(pkg):
        if f.endswith('.py'):
            txt = open(os.path.join(pkg, f)).read()
            assert 'import oracle' not in txt and 'from oracle' not in txt, f


def test_cpu_tensors_fail_loudly():
    m = A.create_model('model_variant', variant='volo_h2_l4', img_size=64, num_classes=10)
    with pytest.raises(RuntimeError, match='CUDA'):
        m(torch.randn(1, 3, 64, 64))
    with pytest.raises(RuntimeError):
        A.TokenLabelCrossEntropy()((torch.randn(2, 4), torch.randn(2, 3, 4), (0, 0, 0, 0)), torch.rand(2, 4, 5))


def test_progressive_schedule_tables():
    args = SimpleNamespace(num_stages=4, r_scale=0.5, h_scale=1., l_scale=0.5, aa_scale=0.5, dp_scale=0., re_scale=0.,
                           resize_scale=[1., 1.], aa='rand-m9-mstd0.5-inc1', drop_path=0.1, reprob=0.25,
                           scale=[0.08, 1.0], epochs=100)
    t = TAB['train_autoprog_sh']
    e, r, h, l, aa, dp, re_, rs = A.progressive_schedule(args, r_max=224, h_max=12, l_max=18)
    assert (e, r, h, l, aa, rs) == (t['e'], t['r'], t['h'], t['l'], t['aa'], t['resize'])
    assert dp == pytest.approx(t['dp']) and re_ == pytest.approx(t['re'])
    args2 = SimpleNamespace(**{**vars(args), 'num_stages': 3, 'r_scale': 0.4, 'l_scale': 0.34, 'dp_scale': -0.5,
                               're_scale': -0.5, 'epochs': 300, 'aa_scale': 0.})
    t = TAB['alt']
    e, r, h, l, aa, dp, re_, rs = A.progressive_schedule(args2, r_max=384, h_max=16, l_max=24)
    assert (e, r, h, l, aa) == (t['e'], t['r'], t['h'], t['l'], t['aa'])
    assert dp == pytest.approx(t['dp']) and re_ == pytest.approx(t['re'])
    for v, d, want in TAB['make_divisible']:
        assert A.make_divisible(v, d) == want


def test_layer_index_maps():
    for key, want in TAB['new_idx'].items():
        p, n = map(int, key.split('->'))
        assert [A.new_idx(i, p, n) for i in range(n)] == want
        assert A.get_new_layer_idx(p, n) == TAB['new_layer_idx'][key]


def test_set_sample_config_identity_flags():
    for cur, want in TAB['identity_flags_9_18'].items():
        m = A.VOLO([4, 14, 0, 0], img_size=32, num_classes=2, stem_hidden_dim=4, embed_dims=[32, 32, 32, 32],
                   num_heads=[1, 1, 1, 1], mlp_ratios=[1, 1, 1, 1], downsamples=[True, False, False, False],
                   outlook_attention=[True, False, False, False], post_layers=['ca', 'ca'])
        m.set_sample_config({'layer_num': int(cur), 'min_layer_num': 9, 'max_layer_num': 18})
        got = [[i for i, b in enumerate(m.network[s]) if b.is_identity_layer] for s in (0, 2)]
        assert got == want


def test_module_tree_matches_reference():
    m = A.create_model('volo_d1', img_size=224)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == TAB['state_dict_volo_d1']
    assert list(m.state_dict().keys()) == list(TAB['state_dict_volo_d1'].keys())
    import torch.nn as nn
    base = {'Linear': nn.Linear, 'LayerNorm': nn.LayerNorm, 'Conv2d': nn.Conv2d, 'BatchNorm2d': nn.BatchNorm2d,
            'Sequential': nn.Sequential, 'ModuleList': nn.ModuleList}
    mods = dict(m.named_modules())
    for name, tname in TAB['modules_volo_d1'].items():
        assert name in mods, name
        if tname in base:   # prog/helpers.py dispatches on these container types (isinstance)
            assert isinstance(mods[name], base[tname]), (name, tname)
        elif tname not in ('Dropout', 'Identity', 'GELU', 'ReLU', 'Unfold', 'AvgPool2d'):
            assert type(mods[name]).__name__ == tname, (name, tname)
    assert len(m.network) == 5 and len(m.network[0]) == 4
    assert m.no_weight_decay() == {'pos_embed', 'cls_token'}
    assert m.num_classes == 1000 and m.mix_token and m.pooling_scale == 2 and m.beta == 1.0
    assert m.default_cfg['crop_pct'] == 0.96 and m.get_classifier() is m.head


def test_model_variant_any_size_and_drop_path_rates():
    m = A.create_model('model_variant', variant='volo_h12_l18', img_size=224, drop_path_rate=0.1, drop_rate=None)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == TAB['state_dict_volo_h12_l18']
    got = {n: mod.drop_prob for n, mod in m.named_modules() if type(mod).__name__ == 'DropPath'}
    want = TAB['drop_path_volo_h12_l18']
    assert set(got) == set(want)
    for k in want:
        assert got[k] == pytest.approx(want[k])
    n = {l: sum(p.numel() for p in A.create_model('model_variant', variant=f'volo_h12_l{l}').parameters())
         for l in (9, 18)}
    assert n[18] == 26632040 and n[9] < n[18]
    import copy
    copy.deepcopy(m)   # ModelEmaV2 relies on deepcopy


def test_registry_kwarg_filtering():
    m = A.create_model('volo_d1', pretrained=False, num_classes=10, drop_rate=None, drop_connect_rate=0.2,
                       drop_path_rate=None, bn_tf=False, bn_momentum=None, bn_eps=None, img_size=64)
    assert m.head.out_features == 10 and m.pos_embed.shape == (1, 4, 4, 384)
    assert any(getattr(mod, 'drop_prob', 0) > 0 for mod in m.modules())
    with pytest.raises(RuntimeError):
        A.create_model('nope')


def test_class_block_glue_functions_match_slice_and_cat():
    """ClassBlock's slice / cat glue (volo._SplitCls / _JoinCls, pure torch) == x[:, :1] ... torch.cat([cls, x[:, 1:]])
    of the reference (models/volo.py:304-308, 634-638), values and gradients."""
    import torch
    from autoprog_b200.volo import _JoinCls, _SplitCls
    torch.manual_seed(0)
    x = torch.randn(3, 5, 4, dtype=torch.double, requires_grad=True)
    w = torch.randn(5, 4, dtype=torch.double)
    g = torch.randn(3, 5, 4, dtype=torch.double)

    def f_ref(x):
        cls = x[:, :1]
        cls = cls + (x * w).sum(1, keepdim=True).tanh()
        return torch.cat([cls * 2, x[:, 1:]], 1)

    def f_new(x):
        cls, xt = _SplitCls.apply(x)
        cls = cls + (xt * w).sum(1, keepdim=True).tanh()
        return _JoinCls.apply(cls * 2, xt, 1)

    y1 = f_ref(x); y1.backward(g); g1 = x.grad.clone(); x.grad = None
    y2 = f_new(x); y2.backward(g)
    assert torch.equal(y1, y2) and torch.allclose(g1, x.grad, atol=1e-14)
    c = torch.randn(1, 1, 4, dtype=torch.double, requires_grad=True)
    t = torch.randn(3, 4, 4, dtype=torch.double, requires_grad=True)
    y1 = torch.cat((c.expand(3, -1, -1), t), 1); y1.backward(g)
    a1, b1 = c.grad.clone(), t.grad.clone(); c.grad = None; t.grad = None
    y2 = _JoinCls.apply(c.expand(3, -1, -1), t, 0); y2.backward(g)
    assert torch.equal(y1, y2) and torch.allclose(a1, c.grad, atol=1e-14) and torch.allclose(b1, t.grad, atol=1e-14)


def test_drop_path_pool_draws_once_per_step():
    """volo.DropPathPool: one RNG call per forward provides two factors per DropPath module (one per residual branch),
    each 0 or 1/keep as in timm 0.4.5 DropPath (x / keep * floor(keep + U)); modules fall back to their own draw after."""
    import torch
    import torch.nn as nn
    from autoprog_b200.volo import DropPath, DropPathPool
    torch.manual_seed(0)
    net = nn.Sequential(DropPath(0.1), nn.Identity(), DropPath(0.0), DropPath(0.5)).train()
    pool = DropPathPool(net)
    assert len(pool.mods) == 2                              # drop_prob 0 modules are not pooled
    pool.draw(64, torch.device('cpu'))
    for m in pool.mods:
        keep = 1.0 - m.drop_prob
        a = m.sample_scale(64, torch.device('cpu'))
        b = m.sample_scale(64, torch.device('cpu'))
        assert not m._pooled                               # both pooled draws consumed
        for t in (a, b):
            assert t.shape == (64,) and bool(((t == 0) | ((t - 1.0 / keep).abs() < 1e-6)).all())
        assert not torch.equal(a, b) or m.drop_prob == 0.0
        c = m.sample_scale(64, torch.device('cpu'))         # third call in the same step: own draw, same law
        assert bool(((c == 0) | ((c - 1.0 / keep).abs() < 1e-6)).all())
    assert net[2].sample_scale(64, torch.device('cpu')) is None
    net.eval()
    pool.draw(64, torch.device('cpu'))
    assert pool.mods[0].sample_scale(64, torch.device('cpu')) is None    # eval: identity


def test_scaler_call_signature_and_update_gating():
    """prog/scaler.py:17-74 contract: __call__(loss, optimizer, clip_grad, clip_mode, parameters, create_graph, update);
    update=False only accumulates gradients (batch splits, main_prog.py:971), update=True clips then steps."""
    import torch
    import torch.nn as nn
    from autoprog_b200.scaler import ApexScaler, Bf16Scaler, NoScaler, dispatch_clip_grad
    assert ApexScaler is Bf16Scaler and issubclass(Bf16Scaler, NoScaler) and Bf16Scaler.state_dict_key == 'bf16_scaler'
    torch.manual_seed(0)
    lin = nn.Linear(4, 3)
    opt = torch.optim.SGD(lin.parameters(), lr=0.1)
    x, y = torch.randn(5, 4), torch.randn(5, 3)
    w0 = lin.weight.detach().clone()
    sc = Bf16Scaler()
    sc(((lin(x) - y) ** 2).mean(), opt, update=False)
    g1 = lin.weight.grad.clone()
    assert torch.equal(lin.weight, w0)                                  # no step yet
    sc(((lin(x) - y) ** 2).mean(), opt, clip_grad=1e-3, clip_mode='norm', parameters=lin.parameters(), update=True)
    assert not torch.equal(lin.weight, w0)
    total = torch.sqrt(sum((p.grad ** 2).sum() for p in lin.parameters()))
    assert float(total) <= 1e-3 * 1.001                                 # accumulated (2 x g1) gradient was clipped
    assert torch.allclose(lin.weight.grad / lin.weight.grad.norm(), g1 / g1.norm(), atol=1e-6)
    assert sc.state_dict() is None
    for mode in ('value', 'agc'):
        lin.zero_grad()
        ((lin(x) - y) ** 2).mean().backward()
        dispatch_clip_grad(list(lin.parameters()), 1e-4, mode=mode)
        if mode == 'value':
            assert float(lin.weight.grad.abs().max()) <= 1e-4 + 1e-12
    try:
        dispatch_clip_grad(list(lin.parameters()), 1.0, mode='nope')
        raise RuntimeError('expected AssertionError')
    except AssertionError:
        pass


def test_flat_state_grouping_alignment_and_views():
    """flat.FlatState: timm's no-weight-decay rule (ndim <= 1, '.bias', model.no_weight_decay()), 8-element aligned
    slices in reverse registration order, parameters / gradients re-pointed into the flat buffers, EMA copies laid
    out identically, and ensure_grad_views() adopting gradients autograd allocated itself."""
    import copy
    import torch
    import torch.nn as nn
    from autoprog_b200.flat import FlatState, split_decay

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.pos_embed = nn.Parameter(torch.randn(1, 3, 5))
            self.fc = nn.Linear(5, 7)
            self.norm = nn.LayerNorm(7)

        def no_weight_decay(self):
            return {'pos_embed'}

        def forward(self, x):
            return self.norm(self.fc(x + self.pos_embed))

    torch.manual_seed(0)
    net = Net()
    decay, no_decay = split_decay(net, 0.05)
    assert [n for n, _ in decay] == ['fc.weight']
    assert sorted(n for n, _ in no_decay) == ['fc.bias', 'norm.bias', 'norm.weight', 'pos_embed']
    ref = {n: p.detach().clone() for n, p in net.named_parameters()}
    ema = copy.deepcopy(net)
    flat = FlatState(net, weight_decay=0.05, want_shadow=False)
    assert flat.weight_decays == [0.05, 0.0]
    for g in flat.groups:
        assert all(o % 8 == 0 for o in g.offsets) and g.numel % 8 == 0
        for p, o in zip(g.params, g.offsets):
            assert p.data_ptr() == g.flat_p[o:].data_ptr()                   # parameter lives in the flat buffer
    assert flat.groups[1].names == ['norm.bias', 'norm.weight', 'fc.bias', 'pos_embed']   # reverse registration order
    for n, p in net.named_parameters():
        assert torch.equal(p, ref[n])
    ema_flats = flat.flat_like(ema)
    for g, ef in zip(flat.groups, ema_flats):
        assert ef.shape == g.flat_p.shape and torch.equal(ef, g.flat_p)      # same layout, same (copied) values
    flat.zero_grad()
    assert all(p.grad is None for p in net.parameters())
    net(torch.randn(2, 3, 5)).sum().backward()                               # CPU: autograd allocates the gradients
    auto = {n: p.grad.clone() for n, p in net.named_parameters()}
    flat.ensure_grad_views()
    for g in flat.groups:
        for n, p, o in zip(g.names, g.params, g.offsets):
            assert p.grad.data_ptr() == g.flat_g[o:].data_ptr() and torch.equal(p.grad, auto[n])


def test_grad_dest_hands_the_flat_view_out_once_per_backward():
    """ops.grad_dest: a kernel may write a parameter's gradient straight into its flat-buffer view only ONCE per backward.
    `.grad` stays None while autograd collects the contributions of a parameter that is used twice in one forward (e.g.
    ClassAttention.kv on the class token and on the patch tokens), so a second hand-out would let the second kernel
    overwrite the first one's output; zero_grad re-arms the view, an existing `.grad` (accumulation steps) blocks it."""
    import torch
    import torch.nn as nn
    from autoprog_b200 import ops
    from autoprog_b200.flat import FlatState

    net = nn.Linear(4, 4)
    flat = FlatState(net, weight_decay=0.05, want_shadow=False)
    flat.zero_grad()
    w = net.weight
    first = ops.grad_dest(w)
    assert first is not None and first.data_ptr() == w._apb_grad_view.data_ptr()
    assert ops.grad_dest(w) is None                       # second use in the same backward: autograd accumulates instead
    flat.zero_grad()
    assert ops.grad_dest(w) is not None                   # re-armed by the next step's zero_grad
    flat.zero_grad()
    w.grad = torch.zeros_like(w)
    assert ops.grad_dest(w) is None                       # gradient accumulation over several backward passes
    assert ops.grad_dest(None) is None
    assert ops.grad_dest(nn.Parameter(torch.zeros(2))) is None     # not managed by a FlatState
