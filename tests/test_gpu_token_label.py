"""Token-label target builder (csrc/token_label.cu) against the oracle restatement of tlt's recipe (torchvision roi_align on
the dense scattered map, oracle/token_label_cpu.py) -- parity unpinned upstream (tlt is not in the reference tree)."""
import pytest
import torch

import autoprog_b200 as A
from gpu_util import need_gpu, rel
from oracle import token_label_cpu as OT

pytestmark = pytest.mark.gpu


def _maps(B, C, Hm, Wm, seed, boxes):
    g = torch.Generator().manual_seed(seed)
    t = torch.zeros(B, 3, 5, Hm, Wm)
    t[:, 0] = torch.randn(B, 5, Hm, Wm, generator=g) * 4
    t[:, 1] = torch.randint(0, C, (B, 5, Hm, Wm), generator=g).float()
    for b, rec in enumerate(boxes):
        t[b, 2, 0, 0, :6] = torch.tensor(rec, dtype=torch.float32)
    return t


@pytest.mark.parametrize('C,Hm,Wm,L', [(40, 18, 18, 14), (1000, 18, 18, 14), (1000, 18, 18, 7), (33, 9, 13, 5), (1000, 14, 14, 1),
                                       (100, 18, 18, 24)])
@pytest.mark.parametrize('softmax', [True, False])
def test_token_label_target_vs_oracle(C, Hm, Wm, L, softmax):
    dev = need_gpu()
    boxes = [[0.1, 0.2, 0.8, 0.9, 0, 3], [0.0, 0.0, 1.0, 1.0, 1, 7], [0.3, 0.1, 0.55, 0.45, 0, C - 1],
             [0.62, 0.05, 0.66, 0.97, 1, 0],          # narrower than one map pixel: roi width clamps to 1
             [-0.1, -0.2, 1.3, 1.1, 0, 5]]            # box leaving the map: samples outside contribute zero
    t = _maps(len(boxes), C, Hm, Wm, C + L, boxes)
    t[0, 1, :, 4, 4] = 2.0                           # one class twice in a pixel's top-5: scores add up
    ref = OT.create_token_label_target(t, C, 0.1, L, apply_softmax=softmax)
    out = A.create_token_label_target(t.to(dev), C, 0.1, L, apply_softmax=softmax)
    assert out.shape == (len(boxes), C, 2 + L * L) and out.dtype == torch.float32
    assert torch.equal(out[:, :, 0].cpu().double().argmax(1), ref[:, :, 0].argmax(1))
    assert rel(out, ref) < 1e-5, rel(out, ref)
    assert float((out.cpu().double() - ref).abs().max()) < 5e-6 * max(1.0, float(ref.abs().max()))


def test_token_label_target_feeds_the_loss_and_onehot():
    dev = need_gpu()
    B, C, L = 4, 1000, 14
    t = _maps(B, C, 18, 18, 1, [[0.05 * b, 0.1, 0.6 + 0.1 * b, 0.9, b & 1, 10 * b] for b in range(B)])
    tgt = A.create_token_label_target(t.to(dev), C, 0.1, L)
    x_cls = torch.randn(B, C, device=dev)
    x_aux = torch.randn(B, L * L, C, device=dev)
    loss = A.TokenLabelCrossEntropy(dense_weight=0.5)((x_cls, x_aux, (1, 2, 5, 6)), tgt)
    from oracle import volo_cpu as O
    ref = O.token_label_ce(x_cls.double().cpu(), x_aux.double().cpu(), (1, 2, 5, 6), OT.create_token_label_target(t, C, 0.1, L), 0.5)
    assert abs(float(loss) - float(ref)) < 1e-5 * abs(float(ref))
    labels = torch.tensor([3, 999, 0, 17], device=dev)
    oh = A.create_token_label_target(labels, C, 0.1)
    assert rel(oh, OT.create_token_label_target(labels.cpu(), C, 0.1)) < 1e-6
