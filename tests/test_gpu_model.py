"""Whole-model parity on the GPU against fixtures produced by the reference's own code (tests/golden/volo_small*.pt):
logits, loss and every parameter gradient, fp32 (1e-5) and bf16 (2e-2)."""
import os

import numpy as np
import pytest
import torch

import autoprog_b200 as A
from gpu_util import G, need_gpu, rel

pytestmark = pytest.mark.gpu


def build(fx, dev):
    a = fx['arch']
    m = A.VOLO(a['layers'], img_size=a['img_size'], num_classes=a['num_classes'], stem_hidden_dim=a['stem_hidden'],
               embed_dims=a['embed_dims'], num_heads=a['num_heads'], mlp_ratios=[3, 3, 3, 3],
               downsamples=[True, False, False, False], outlook_attention=[True, False, False, False],
               post_layers=['ca', 'ca'], drop_path_rate=a.get('drop_path_rate', 0.0))
    m.load_state_dict(fx['sd'])
    return m.to(dev)


def run_case(m, c, dev, bf16, seed):
    np.random.seed(seed)
    m.train(c['train'])
    m.zero_grad(set_to_none=True)
    if c.get('sample_cfg'):
        m.set_sample_config(c['sample_cfg'])
    for name, masks in (c.get('drop_masks') or {}).items():
        blk = m.get_submodule(name.rstrip('.'))
        if masks:
            blk.drop_path.forced = [t.float() for t in masks]
    x = c['x'].to(dev).requires_grad_(True)
    with A.autocast(enabled=bf16):
        out = m(x)
        if not c['train']:
            return out, None, x
        crit = A.TokenLabelCrossEntropy(dense_weight=c['dense_weight'], cls_weight=1.0, classes=12)
        loss = crit(out, c['target'].to(dev))
    loss.backward()
    return out, loss, x


SEEDS = {'train_r64': 11, 'train_r96_bicubic': 12, 'train_r104_oddgrid': 13, 'eval_r80': 14, 'train_r64_elastic_dp': 21}


@pytest.mark.parametrize('case', ['train_r64', 'train_r104_oddgrid'])
def test_volo_small_golden_bf16_training_path(case):
    """bf16 with an input that needs no gradient -- what training does: the 7x7 stem conv runs as im2col + tcgen05 GEMM."""
    dev = need_gpu()
    fx = torch.load(os.path.join(G, 'volo_small.pt'))
    m = build(fx, dev)
    c = fx['cases'][case]
    np.random.seed(SEEDS[case])
    m.train(True)
    m.zero_grad(set_to_none=True)
    with A.autocast(enabled=True):
        out = m(c['x'].to(dev))
        crit = A.TokenLabelCrossEntropy(dense_weight=c['dense_weight'], cls_weight=1.0, classes=12)
        loss = crit(out, c['target'].to(dev))
    loss.backward()
    assert abs(float(loss) - float(c['loss'])) < 2e-2 * abs(float(c['loss']))
    g = dict(m.named_parameters())['patch_embed.conv.0.weight'].grad
    # per-tensor bound of the bf16 golden test (8 x 2e-2): the first layer's gradient carries the whole net's bf16 noise
    assert g is not None and rel(g, c['grads']['patch_embed.conv.0.weight']) < 0.16, rel(g, c['grads']['patch_embed.conv.0.weight'])
    both = [(p.grad.double().cpu(), c['grads'][n].double()) for n, p in m.named_parameters()
            if p.grad is not None and c['grads'].get(n) is not None]
    num = sum(float((a - b).pow(2).sum()) for a, b in both)
    den = sum(float(b.pow(2).sum()) for a, b in both)
    assert (num / den) ** 0.5 < 2e-2, (num / den) ** 0.5


@pytest.mark.parametrize('bf16', [False, True])
@pytest.mark.parametrize('case', ['train_r64', 'train_r96_bicubic', 'train_r104_oddgrid', 'eval_r80'])
def test_volo_small_golden(case, bf16):
    dev = need_gpu()
    fx = torch.load(os.path.join(G, 'volo_small.pt'))
    m = build(fx, dev)
    c = fx['cases'][case]
    out, loss, x = run_case(m, c, dev, bf16, SEEDS[case])
    t = 2e-2 if bf16 else 1e-5
    if not c['train']:
        assert rel(out, c['out']) < t, rel(out, c['out'])
        return
    assert list(out[2]) == c['bbox']
    assert rel(out[0], c['x_cls']) < t and rel(out[1], c['x_aux']) < t, (rel(out[0], c['x_cls']), rel(out[1], c['x_aux']))
    assert abs(float(loss) - float(c['loss'])) < t * abs(float(c['loss']))
    if not bf16:   # the image gradient is not a training quantity; in bf16 it crosses the cuDNN stem and is not pinned
        assert rel(x.grad, c['dx']) < t, rel(x.grad, c['dx'])
    params = dict(m.named_parameters())
    # per-tensor diagnostic on tensors big enough for a norm-wise error to be meaningful (tiny BatchNorm/bias vectors are
    # sums with heavy cancellation; they are covered by the whole-gradient bound below)
    worst = max(((rel(params[k].grad, g), k) for k, g in c['grads'].items() if g is not None and (not bf16 or g.numel() >= 256)),
                key=lambda z: z[0])
    # bf16: per-tensor bar 2e-2 on all but the tiniest tensors; whole-gradient vector within 2e-2
    flat = torch.cat([params[k].grad.flatten().double().cpu() for k, g in c['grads'].items() if g is not None])
    ref = torch.cat([g.flatten().double() for g in c['grads'].values() if g is not None])
    assert float((flat - ref).norm() / ref.norm()) < t, float((flat - ref).norm() / ref.norm())
    assert worst[0] < (8 * t if bf16 else 2 * t), worst   # bf16: cuDNN stem tensors are the noisiest (toy weights x6)


@pytest.mark.parametrize('bf16', [False, True])
def test_volo_elastic_depth_and_droppath(bf16):
    dev = need_gpu()
    fx = torch.load(os.path.join(G, 'volo_small_elastic.pt'))
    m = build(fx, dev)
    c = fx['cases']['train_r64_elastic_dp']
    out, loss, x = run_case(m, c, dev, bf16, SEEDS['train_r64_elastic_dp'])
    t = 2e-2 if bf16 else 1e-5
    for name, flag in fx['identity_flags'].items():
        assert bool(getattr(m.get_submodule(name), 'is_identity_layer', False)) == flag
    assert list(out[2]) == c['bbox']
    assert rel(out[0], c['x_cls']) < t and rel(out[1], c['x_aux']) < t
    assert abs(float(loss) - float(c['loss'])) < t * abs(float(c['loss']))
    params = dict(m.named_parameters())
    for k, g in c['grads'].items():
        if g is None:   # identity layers receive no gradient (main_prog.py:543 relies on this)
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0, k
    flat = torch.cat([params[k].grad.flatten().double().cpu() for k, g in c['grads'].items() if g is not None])
    ref = torch.cat([g.flatten().double() for g in c['grads'].values() if g is not None])
    assert float((flat - ref).norm() / ref.norm()) < t


def test_volo_d1_full_size_step_runs_and_is_deterministic():
    """BASELINE config: volo_d1 @224 bf16 fwd+loss+bwd; equal seeds give bit-identical loss and gradients."""
    dev = need_gpu()
    torch.manual_seed(0)
    m = A.create_model('volo_d1', img_size=224).to(dev)
    x = torch.randn(8, 3, 224, 224, device=dev)
    tgt = torch.softmax(torch.randn(8, 1000, 198, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    res = []
    for _ in range(2):
        np.random.seed(0)
        m.zero_grad(set_to_none=True)
        with A.autocast():
            out = m(x)
            loss = crit(out, tgt)
        loss.backward()
        res.append((loss.item(), m.head.weight.grad.clone(), m.network[0][0].attn.attn.weight.grad.clone()))
    assert out[0].shape == (8, 1000) and out[1].shape == (8, 196, 1000)
    assert np.isfinite(res[0][0]) and abs(res[0][0] - 10.36) < 0.5   # ~ (1 + 0.5) * ln(1000) at init
    assert res[0][0] == res[1][0] and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])


def test_opt_in_launch_modes_are_bit_identical():
    """The opt-in / switchable launch modes change scheduling only: programmatic dependent launch (apb_set_pdl), the wgrad
    side stream (ops.SIDE_WGRAD) and the four-stage GEMM kernels must reproduce the default path's loss and every gradient
    bit for bit (volo_d1 @ 224, bf16)."""
    from autoprog_b200 import kernels as K, ops
    dev = need_gpu()
    torch.manual_seed(0)
    m = A.create_model('volo_d1', img_size=224).to(dev)
    x = torch.randn(4, 3, 224, 224, device=dev)
    tgt = torch.softmax(torch.randn(4, 1000, 198, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)

    def run():
        np.random.seed(0)
        m.zero_grad(set_to_none=True)
        with A.autocast():
            loss = crit(m(x), tgt)
        loss.backward()
        torch.cuda.synchronize()
        return loss.item(), {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}

    base = run()
    side = ops.SIDE_WGRAD
    try:
        for name, enter, leave in [('pdl', lambda: K.set_pdl(True), lambda: K.set_pdl(False)),
                                   ('no side stream', lambda: setattr(ops, 'SIDE_WGRAD', False), lambda: setattr(ops, 'SIDE_WGRAD', side)),
                                   ('4-stage gemm', lambda: K.debug_gemm_switches(0, False), lambda: K.debug_gemm_switches(0, True))]:
            enter()
            try:
                loss, grads = run()
            finally:
                leave()
            assert loss == base[0], name
            assert grads.keys() == base[1].keys(), name
            bad = [k for k in grads if not torch.equal(grads[k], base[1][k])]
            assert not bad, (name, bad[:5])
    finally:
        K.set_pdl(False); K.debug_gemm_switches(0, True); ops.SIDE_WGRAD = side


@pytest.mark.parametrize('bf16', [False, True])
def test_deit_against_oracle(bf16):
    """DeiT (timm ViT restated; parity unpinned upstream): kernels vs oracle.vit_forward on a small config, incl. elastic depth."""
    from oracle import volo_cpu as O
    from autoprog_b200.deit import VisionTransformer
    from functools import partial
    from autoprog_b200.volo import LayerNorm
    dev = need_gpu()
    torch.manual_seed(3)
    m = VisionTransformer(img_size=64, patch_size=16, num_classes=10, embed_dim=64, depth=4, num_heads=2, mlp_ratio=4,
                          qkv_bias=True, norm_layer=partial(LayerNorm, eps=1e-6), return_dense=True).to(dev)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if p.dim() > 1:
                p.mul_(8.0)
            elif 'bias' in n:
                p.normal_(0, 0.1)
    m.set_sample_config({'layer_num': 3, 'min_layer_num': 2, 'max_layer_num': 4})
    skip = [i for i, b in enumerate(m.blocks) if getattr(b, 'is_identity_layer', False)]
    assert len(skip) == 1
    x = torch.randn(3, 3, 64, 64, device=dev)
    with A.autocast(enabled=bf16):
        out, aux = m(x)
    (out.float().square().sum() + aux.float().square().mean()).backward()
    sd = {k: v.detach().double().cpu().requires_grad_(True) for k, v in m.state_dict().items()}
    ro, ra = O.vit_forward(sd, x.double().cpu(), depth=4, heads=2, patch=16, eps=1e-6, skip=skip, dense=True)
    (ro.square().sum() + ra.square().mean()).backward()
    t = 2e-2 if bf16 else 1e-5
    assert rel(out, ro) < t and rel(aux, ra) < t, (rel(out, ro), rel(aux, ra))
    params = dict(m.named_parameters())
    flat = torch.cat([params[k].grad.flatten().double().cpu() for k in params if sd[k].grad is not None and params[k].grad is not None])
    ref = torch.cat([sd[k].grad.flatten() for k in params if sd[k].grad is not None and params[k].grad is not None])
    assert float((flat - ref).norm() / ref.norm()) < t
    for k in params:   # identity layer: no gradient
        if k.startswith(f'blocks.{skip[0]}.'):
            assert params[k].grad is None or float(params[k].grad.abs().max()) == 0.0


def test_flat_direct_gradient_writes_match_autograd_accumulation():
    """FusedAdamW's flat storage: gradients written in place by the kernels == gradients accumulated by autograd,
    and a second backward without zero_grad accumulates (batch-splits path, prog/scaler.py update=False)."""
    from autoprog_b200.optim import FusedAdamW
    dev = need_gpu()
    torch.manual_seed(0)
    m = A.create_model('model_variant', variant='volo_h2_l4', img_size=64, num_classes=16).to(dev)
    x = torch.randn(4, 3, 64, 64, device=dev)
    tgt = torch.softmax(torch.randn(4, 16, 18, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)

    def run():
        np.random.seed(3)
        with A.autocast():
            loss = crit(m(x), tgt)
        loss.backward()
    m.zero_grad(set_to_none=True)
    run()
    ref = {n: p.grad.detach().clone() for n, p in m.named_parameters()}
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    opt.zero_grad()
    assert all(p.grad is None for p in m.parameters())
    run()
    flat_ptrs = {id(p): v.data_ptr() for g in opt.flat.groups for p, v in zip(g.params, g.views(g.flat_g))}
    n_direct = 0
    for n, p in m.named_parameters():
        assert torch.equal(p.grad, ref[n]), n
        n_direct += int(p.grad.data_ptr() == flat_ptrs[id(p)])
    assert n_direct >= 0.6 * len(ref), n_direct          # almost every gradient landed in the flat buffer directly
    opt.flat.ensure_grad_views()
    for n, p in m.named_parameters():
        assert p.grad.data_ptr() == flat_ptrs[id(p)] and torch.equal(p.grad, ref[n]), n
    run()                                                # no zero_grad: accumulate
    for n, p in m.named_parameters():
        assert rel(p.grad, 2 * ref[n]) < 1e-6, n


def test_cuda_graph_step_matches_eager():
    """GraphedTrainStep (whole step captured once, mix-token box / lr read from memory at replay) == eager steps."""
    import copy
    from autoprog_b200.optim import FusedAdamW
    from autoprog_b200.graph import GraphedTrainStep
    dev = need_gpu()
    torch.manual_seed(0)
    base = A.create_model('model_variant', variant='volo_h2_l4', img_size=64, num_classes=16).to(dev)
    x = torch.randn(8, 3, 64, 64, device=dev)
    tgt = torch.softmax(torch.randn(8, 16, 18, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    losses = {}
    boxes = {}
    for mode in ('eager', 'graph'):
        m = copy.deepcopy(base)
        ema = copy.deepcopy(m).eval()
        opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05, ema_models=[ema], ema_decays=[0.99])
        np.random.seed(5)
        out = []
        if mode == 'eager':
            for _ in range(2):
                opt.zero_grad()
                with A.autocast():
                    o = m(x)
                    loss = crit(o, tgt)
                loss.backward()
                opt.step()
                out.append(float(loss))
        else:
            # construction runs 3 warm-up steps and rolls them back (weights, moments, EMA, RNG): replay i == eager step i
            step = GraphedTrainStep(m, crit, opt, x, tgt, bf16=True, warmup=3)
            out = [float(step()), float(step())]
            step.close()
        losses[mode] = out
        boxes[mode] = {n: p.detach().clone() for n, p in m.named_parameters()}
        boxes[mode]['__ema'] = next(ema.parameters()).detach().clone()
    # same data, same boxes, same optimizer state
    for i in (0, 1):
        assert abs(losses['eager'][i] - losses['graph'][i]) < 2e-3 * abs(losses['eager'][i]), (i, losses)
    worst = max(rel(boxes['graph'][n], boxes['eager'][n]) for n in boxes['eager'])
    assert worst < 5e-3, worst
