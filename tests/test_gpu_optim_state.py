"""FusedAdamW checkpoint interchange with torch.optim.AdamW, and bf16 shadow freshness after in-place parameter loads
(optimizer state was dropped from checkpoints; shadows went stale after load_state_dict / MoGrow loaders)."""
import copy

import numpy as np
import pytest
import torch

import autoprog_b200 as A
from autoprog_b200.flat import split_decay
from autoprog_b200.optim import FusedAdamW
from gpu_util import need_gpu, rel

pytestmark = pytest.mark.gpu


def _model(dev, seed=0):
    torch.manual_seed(seed)
    return A.create_model('model_variant', variant='volo_h2_l4', img_size=64, num_classes=16).to(dev)


def _step(m, opt, x, tgt, crit, bf16=True):
    np.random.seed(3)
    opt.zero_grad()
    with A.autocast(enabled=bf16):
        loss = crit(m(x), tgt)
    loss.backward()
    opt.step()
    return float(loss)


def _torch_adamw(m, lr, wd):
    decay, no_decay = split_decay(m, wd)
    # same parameter order as the flat groups (reverse registration order) so that state ids line up
    return torch.optim.AdamW([dict(params=[p for _, p in reversed(decay)], weight_decay=wd),
                              dict(params=[p for _, p in reversed(no_decay)], weight_decay=0.0)], lr=lr)


def test_state_dict_roundtrip_and_torch_interchange():
    dev = need_gpu()
    x = torch.randn(4, 3, 64, 64, device=dev)
    tgt = torch.softmax(torch.randn(4, 16, 18, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    m1 = _model(dev)
    o1 = FusedAdamW(m1, lr=1e-3, weight_decay=0.05)
    for _ in range(3):
        _step(m1, o1, x, tgt, crit, bf16=False)
    sd_opt, sd_model = copy.deepcopy(o1.state_dict()), copy.deepcopy(m1.state_dict())
    assert len(sd_opt['state']) == sum(len(g['params']) for g in sd_opt['param_groups']) > 0
    assert all(float(s['step']) == 3.0 for s in sd_opt['state'].values())
    # (a) resume into a fresh FusedAdamW: optimizer first, then weights (the reference order, main_prog.py:1359-1388)
    m2 = _model(dev, seed=1)
    o2 = FusedAdamW(m2, lr=1e-3, weight_decay=0.05)
    o2.load_state_dict(sd_opt)
    m2.load_state_dict(sd_model)
    assert o2.step_count == 3
    l1 = _step(m1, o1, x, tgt, crit, bf16=False)
    l2 = _step(m2, o2, x, tgt, crit, bf16=False)
    assert l1 == l2
    for (n, p), q in zip(m1.named_parameters(), m2.parameters()):
        # (fp32 mode runs the stem convolutions through cuDNN, whose weight-gradient kernels are not bitwise reproducible)
        assert torch.equal(p, q) or (n.startswith('patch_embed.conv') and rel(q, p) < 1e-6), n
    # (b) the same checkpoint drives a stock torch.optim.AdamW over the same parameter order
    m3 = _model(dev, seed=2)
    m3.load_state_dict(sd_model)
    o3 = _torch_adamw(m3, 1e-3, 0.05)
    o3.load_state_dict(sd_opt)
    np.random.seed(3)
    o3.zero_grad(set_to_none=True)
    crit(m3(x), tgt).backward()
    o3.step()
    worst = max(rel(q, p) for p, q in zip(m1.parameters(), m3.parameters()))
    assert worst < 1e-5, worst
    # (c) and back: a torch AdamW checkpoint loads into FusedAdamW
    m4 = _model(dev, seed=4)
    o4 = FusedAdamW(m4, lr=1e-3, weight_decay=0.05)
    o4.load_state_dict(o3.state_dict())
    assert o4.step_count == 4
    assert rel(o4.exp_avg[0], o1.exp_avg[0]) < 1e-5 and rel(o4.exp_avg_sq[0], o1.exp_avg_sq[0]) < 1e-5


def test_bf16_shadow_follows_in_place_loads():
    """load_state_dict AFTER the optimizer was built: the next bf16 forward must see the loaded weights."""
    dev = need_gpu()
    x = torch.randn(2, 3, 64, 64, device=dev)
    m = _model(dev, seed=0)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)          # publishes the bf16 shadows
    m.eval()
    with torch.no_grad(), A.autocast():
        before = m(x).float()
    donor = _model(dev, seed=9)
    m.load_state_dict(donor.state_dict())                    # in-place copy_ into the flat views
    donor.eval()
    with torch.no_grad(), A.autocast():
        after, want = m(x).float(), donor(x).float()
    assert rel(before, want) > 1e-2                          # the two models really differ
    assert torch.equal(after, want)
    # MoGrow-style single-parameter overwrite (prog/helpers.py: p.copy_(q))
    with torch.no_grad():
        m.head.weight.copy_(donor.head.weight * 0.5)
        donor.head.weight.mul_(0.5)
    with torch.no_grad(), A.autocast():
        assert torch.equal(m(x).float(), donor(x).float())
    del opt


def test_graph_replays_do_not_freeze_eval_weights():
    """Evaluations between graph replays must see the weights the replayed optimizer kernel wrote (cached
    conv / padded / cast copies kept matching because raw-pointer updates do not bump `_version`)."""
    from autoprog_b200.graph import GraphedTrainStep
    dev = need_gpu()
    m = _model(dev)
    ema = copy.deepcopy(m).eval()
    for p in ema.parameters():
        p.requires_grad_(False)
    opt = FusedAdamW(m, lr=5e-3, weight_decay=0.05, ema_models=[ema], ema_decays=[0.5])
    x = torch.randn(8, 3, 64, 64, device=dev)
    tgt = torch.softmax(torch.randn(8, 16, 18, device=dev), 1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    p0 = {n: p.detach().clone() for n, p in m.named_parameters()}
    step_count0 = opt.step_count
    gs = GraphedTrainStep(m, crit, opt, x, tgt, bf16=True, warmup=3)
    # construction is side-effect free: warm-up steps are rolled back
    assert opt.step_count == step_count0
    assert all(torch.equal(p, p0[n]) for n, p in m.named_parameters())

    def ev(net):
        net.eval()
        with torch.no_grad(), A.autocast():
            out = net(x).float()
        return out

    outs = []
    for _ in range(3):
        gs()
        m.train()
        outs.append((ev(ema), ev(copy.deepcopy(ema))))     # deep copy: fresh objects, nothing cached
        m.train()
    for cached, fresh in outs:
        assert torch.equal(cached, fresh)
    assert rel(outs[0][0], outs[2][0]) > 1e-4               # and the EMA really moved between evaluations
    gs.close()
