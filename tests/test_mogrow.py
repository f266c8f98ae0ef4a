"""MoGrow / elastic-depth API contract: the reference's own `prog/helpers.py` loaders must run
unmodified on autoprog_b200 modules and give the same result as on the reference modules; our own depth-growth
helper must agree with them.  Needs /root/reference (present in the build container only)."""
import copy
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not mounted')


def _ref():
    import numpy as np
    np.int = int
    for p in (os.path.join(ROOT, 'oracle', 'ref_shim'), REF):
        if p not in sys.path:
            sys.path.insert(0, p)
    import models.volo as RV
    import prog.helpers as RH
    return RV, RH


def _mine(l, **kw):
    import autoprog_b200 as A
    return A.create_model('model_variant', variant=f'volo_h2_l{l}', img_size=64, num_classes=10, **kw)


def _theirs(RV, l):
    from autoprog_b200.submodels import variant_layers
    return RV.VOLO(variant_layers(l), img_size=64, num_classes=10, embed_dims=[32, 64, 64, 64], num_heads=[1, 2, 2, 2],
                   mlp_ratios=[3] * 4, downsamples=[True, False, False, False],
                   outlook_attention=[True, False, False, False], post_layers=['ca', 'ca'])


def test_reference_mogrow_runs_on_our_modules_and_matches():
    RV, RH = _ref()
    torch.manual_seed(0)
    old_mine = _mine(5)
    old_ref = _theirs(RV, 5)
    old_ref.load_state_dict(old_mine.state_dict())                      # identical module tree => strict load works
    emas_mine, emas_ref = [], []
    for k in range(4):
        torch.manual_seed(10 + k)
        e = copy.deepcopy(old_mine)
        with torch.no_grad():
            for p in e.parameters():
                p.add_(torch.randn_like(p) * 0.01)
        r = copy.deepcopy(old_ref)
        r.load_state_dict(e.state_dict())
        emas_mine.append(SimpleNamespace(module=e))
        emas_ref.append(SimpleNamespace(module=r))
    torch.manual_seed(1)
    new_mine = _mine(9)
    new_ref = _theirs(RV, 9)
    new_ref.load_state_dict(new_mine.state_dict())
    assert len(new_mine.network[2]) > len(old_mine.network[2])
    # the reference loader, unmodified, on our classes and on its own classes
    RH.load_slice_clone_ema(new_mine, emas_mine[3], emas_mine)
    RH.load_slice_clone_ema(new_ref, emas_ref[3], emas_ref)
    sd_m, sd_r = new_mine.state_dict(), new_ref.state_dict()
    assert list(sd_m) == list(sd_r)
    for k in sd_m:
        assert torch.equal(sd_m[k], sd_r[k]), k
    # our own helper agrees with the reference loader
    from autoprog_b200.helpers import load_slice_clone_ema
    torch.manual_seed(1)
    again = _mine(9)
    load_slice_clone_ema(again, emas_mine[3], emas_mine)
    for k, v in again.state_dict().items():
        assert torch.equal(v, sd_m[k]), k
    # spot check the mapping itself: layer i <- EMA-3 layer new_idx(i)
    src = emas_mine[3].module.state_dict()
    for i in range(len(new_mine.network[2])):
        j = RH.new_idx(i, len(old_mine.network[2]), len(new_mine.network[2]))
        assert torch.equal(sd_m[f'network.2.{i}.attn.qkv.weight'], src[f'network.2.{j}.attn.qkv.weight'])


def test_reference_load_super_and_slice_run_on_our_modules():
    RV, RH = _ref()
    torch.manual_seed(2)
    big_mine, big_ref = _mine(9), _theirs(RV, 9)
    big_ref.load_state_dict(big_mine.state_dict())
    small_mine, small_ref = _mine(5), _theirs(RV, 5)
    small_ref.load_state_dict(small_mine.state_dict())
    RH.load_super(small_mine, big_mine, base_layer=5, model_name='volo')
    RH.load_super(small_ref, big_ref, base_layer=5, model_name='volo')
    for k, v in small_mine.state_dict().items():
        assert torch.equal(v, small_ref.state_dict()[k]), k
    m2, r2 = _mine(9), _theirs(RV, 9)
    r2.load_state_dict(m2.state_dict())
    RH.load_slice_clone(m2, small_mine)
    RH.load_slice_clone(r2, small_ref)
    for k, v in m2.state_dict().items():
        assert torch.equal(v, r2.state_dict()[k]), k


def test_width_growth_is_rejected():
    import autoprog_b200 as A
    from autoprog_b200.helpers import load_slice_clone_ema
    a = A.create_model('model_variant', variant='volo_h2_l5', img_size=64, num_classes=10)
    b = A.create_model('model_variant', variant='volo_h4_l5', img_size=64, num_classes=10)
    with pytest.raises(NotImplementedError):
        load_slice_clone_ema(b, a)
