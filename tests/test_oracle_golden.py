"""Pins oracle/volo_cpu.py against fixtures produced by the reference's own code (oracle/gen_golden.py)."""
import json
import os

import pytest
import torch

from oracle import volo_cpu as O

G = os.path.join(os.path.dirname(__file__), 'golden')


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).norm() / (b.detach().double().norm() + 1e-300))


def test_schedule_tables():
    tab = json.load(open(os.path.join(G, 'tables.json')))
    t = tab['train_autoprog_sh']
    e, r, h, l, aa, dp, re, rs = O.progressive_schedule(100, 4, 0.5, 1., 0.5, 'rand-m9-mstd0.5-inc1', 0.5, 0.1, 0., 0.25,
                                                        0., [0.08, 1.0], [1., 1.], 224, 12, 18)
    assert (e, r, h, l, aa) == (t['e'], t['r'], t['h'], t['l'], t['aa'])
    assert dp == pytest.approx(t['dp']) and re == pytest.approx(t['re']) and rs == t['resize']
    assert r == [128, 160, 192, 224] and l == [9, 12, 15, 18]
    t = tab['alt']
    e, r, h, l, aa, dp, re, rs = O.progressive_schedule(300, 3, 0.4, 1., 0.34, 'rand-m9-mstd0.5-inc1', 0., 0.1, -0.5, 0.25,
                                                        -0.5, [0.08, 1.0], [1., 1.], 384, 16, 24)
    assert (e, r, h, l, aa) == (t['e'], t['r'], t['h'], t['l'], t['aa'])
    assert dp == pytest.approx(t['dp']) and re == pytest.approx(t['re'])
    for v, d, want in tab['make_divisible']:
        assert O.make_divisible(v, d) == want
    for key, want in tab['new_idx'].items():
        p, n = map(int, key.split('->'))
        assert [O.new_idx(i, p, n) for i in range(n)] == want
        assert O.get_new_layer_idx(p, n) == tab['new_layer_idx'][key]
    for cur, want in tab['identity_flags_9_18'].items():
        plan = O.identity_layer_plan(int(cur), 9, 18)
        assert [plan[0], plan[1]] == want


def test_outlook_attention_module():
    fx = torch.load(os.path.join(G, 'outlook_attention.pt'))
    for name, c in fx.items():
        sd = {'a.' + k: v.double().requires_grad_(True) for k, v in c['sd'].items()}
        x = c['x'].double().requires_grad_(True)
        y = O.outlook_attention(x, sd, 'a.', c['heads'])
        assert rel(y, c['y']) < 1e-12, name
        y.backward(c['dy'].double())
        assert rel(x.grad, c['dx']) < 1e-6, name
        for k, g in c['grads'].items():
            assert rel(sd['a.' + k].grad, g) < 1e-6, (name, k)


def test_outlook_core_closed_form_backward():
    torch.manual_seed(0)
    for (B, H, W, heads) in [(2, 7, 6, 2), (1, 8, 8, 1), (2, 5, 9, 3)]:
        h, w = (H + 1) // 2, (W + 1) // 2
        v = torch.randn(B, H, W, heads * 32, dtype=torch.float64, requires_grad=True)
        lg = torch.randn(B, h, w, heads * 81, dtype=torch.float64, requires_grad=True)
        dy = torch.randn(B, H, W, heads * 32, dtype=torch.float64)
        O.outlook_core(v, lg, heads, 0.17).backward(dy)
        dv, dl = O.outlook_core_bwd(v.detach(), lg.detach(), dy, heads, 0.17)
        assert rel(dv, v.grad) < 1e-13 and rel(dl, lg.grad) < 1e-13


def test_pos_embed_bicubic():
    fx = torch.load(os.path.join(G, 'pos_embed.pt'))
    for g, want in fx['out'].items():
        got = O.pos_embed_resize(fx['pos'], g[0], g[1])
        assert got.shape == want.shape
        assert rel(got, want) < 1e-12, g


def _volo_case(fx, c, arch):
    sd = {k: v.double().requires_grad_(v.dtype.is_floating_point) for k, v in fx['sd'].items()}
    x = c['x'].double().requires_grad_(True)
    skip = None
    if c.get('sample_cfg'):
        s = c['sample_cfg']
        skip = O.identity_layer_plan(s['layer_num'], s['min_layer_num'], s['max_layer_num'])
    if not c['train']:
        out = O.volo_forward(sd, x, arch, train=False, skip=skip)
        assert rel(out, c['out']) < 1e-11
        return
    out = O.volo_forward(sd, x, arch, train=True, bbox=c['bbox'], skip=skip, keep_masks=c.get('drop_masks'),
                         keep_prob=c.get('drop_keep'))
    assert rel(out[0], c['x_cls']) < 1e-11 and rel(out[1], c['x_aux']) < 1e-11
    loss = O.token_label_ce(out[0], out[1], out[2], c['target'].double(), dense_weight=c['dense_weight'])
    assert abs(float(loss) - float(c['loss'])) < 1e-11 * abs(float(c['loss']))
    loss.backward()
    assert rel(x.grad, c['dx']) < 2e-6
    for k, g in c['grads'].items():
        if g is None:
            assert sd[k].grad is None or float(sd[k].grad.abs().max()) == 0.0, k
        else:
            assert rel(sd[k].grad, g) < 2e-6, k


@pytest.mark.parametrize('case', ['train_r64', 'train_r96_bicubic', 'train_r104_oddgrid', 'eval_r80'])
def test_volo_small(case):
    fx = torch.load(os.path.join(G, 'volo_small.pt'))
    a = fx['arch']
    arch = O.VoloArch(a['layers'], a['embed_dims'], a['num_heads'], 3, a['stem_hidden'], a['img_size'], a['num_classes'])
    _volo_case(fx, fx['cases'][case], arch)


def test_volo_small_elastic_droppath():
    fx = torch.load(os.path.join(G, 'volo_small_elastic.pt'))
    a = fx['arch']
    arch = O.VoloArch(a['layers'], a['embed_dims'], a['num_heads'], 3, a['stem_hidden'], a['img_size'], a['num_classes'])
    rates = O.drop_path_rates(arch, a['drop_path_rate'])
    c = fx['cases']['train_r64_elastic_dp']
    for j, rate in enumerate(rates[1]):
        assert abs((1 - rate) - c['drop_keep'][f'network.2.{j}.']) < 1e-12
    _volo_case(fx, c, arch)
    plan = O.identity_layer_plan(5, 4, 6)
    for name, flag in fx['identity_flags'].items():
        _, net, li = name.split('.')
        assert flag == (int(li) in plan[0 if net == '0' else 1])


def test_losses():
    fx = torch.load(os.path.join(G, 'losses.pt'))
    for key, c in fx['cases'].items():
        parts = key.split('|')
        if parts[0] == 'tlce':
            bbox, t = eval(parts[1]), fx[parts[2]]
            wd, wc = float(parts[3]), float(parts[4])
            loss, dc, da = O.token_label_ce_grads(fx['x_cls'], fx['x_aux'], bbox, t, wd, wc)
            assert abs(float(loss - c['loss'])) < 1e-12
            assert rel(dc, c['dcls']) < 1e-12 or float(c['dcls'].abs().max()) == 0
            assert rel(da, c['daux']) < 1e-12
            assert abs(float(O.token_label_ce(fx['x_cls'], fx['x_aux'], bbox, t, wd, wc) - c['loss'])) < 1e-12
        elif parts[0] == 'gt':
            got = O.token_label_ce(fx['x_cls'], fx['x_aux'], eval(parts[1]), fx['t3'], 0.5, 1.0, gt_mix=True)
            assert abs(float(got - c['loss'])) < 1e-12
    C = fx['x_cls'].shape[-1]
    assert abs(float(O.soft_ce(fx['x_cls'], fx['t2']) - fx['cases']['soft']['loss'])) < 1e-12
    assert abs(float(O.soft_ce(fx['x_aux'].reshape(-1, C)[:8], fx['t2']) - fx['cases']['soft_rep']['loss'])) < 1e-12
    assert abs(float(O.soft_ce(fx['x_cls'], fx['t3'][:, :, 1]) - fx['cases']['tlsoft']['loss'])) < 1e-12


def test_oracle_resize_input_matches_the_reference_call():
    """main_prog.py:973-974 / :1910 call F.interpolate(input, size=(r, r), mode='bilinear', align_corners=False);
    the oracle's separable-matrix restatement must equal that exact call (fp64)."""
    import torch.nn.functional as F
    torch.manual_seed(0)
    for (H, W, r) in [(224, 224, 128), (224, 224, 160), (224, 224, 192), (224, 224, 224), (37, 53, 20), (20, 24, 57)]:
        x = torch.randn(2, 3, H, W, dtype=torch.float64)
        ref = F.interpolate(x, size=(r, r), mode='bilinear', align_corners=False)
        got = O.resize_input(x, r)
        assert float((got - ref).abs().max()) < 1e-12, (H, W, r)
