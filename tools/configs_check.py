"""Smoke-run the BASELINE configs (stage sub-nets, volo_d2@384, deit_small) for shape coverage; prints ms/step."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autoprog_b200 as A
from autoprog_b200.optim import FusedAdamW
dev = torch.device('cuda:0')
def run(name, res, B, kw=None, steps=5, img_size=224, deit=False):
    torch.manual_seed(0); np.random.seed(0)
    m = A.create_model(name, img_size=img_size, **(kw or {})).to(dev)
    opt = FusedAdamW(m, lr=1e-3, weight_decay=0.05)
    crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
    x = torch.randn(B, 3, res, res, device=dev); g = res // 16
    t = torch.softmax(torch.randn(B, 1000, 2 + g * g, device=dev), 1)
    def step():
        opt.zero_grad()
        with A.autocast():
            out = m(x)
            loss = crit(out, t) if not deit else out.float().logsumexp(-1).mean()
        loss.backward(); opt.step(); return loss
    for _ in range(3): l = step()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(steps): l = step()
    torch.cuda.synchronize(); ms = (time.perf_counter() - t0) / steps * 1e3
    print(f'{name:28s} kw={kw} res={res} B={B}: {ms:8.2f} ms/step  {B/ms*1e3:9.1f} img/s  loss={float(l):.4f}', flush=True)
    del m, opt; torch.cuda.empty_cache()
for l, r in ((9, 128), (12, 160), (15, 192), (18, 224)):
    run('model_variant', r, 128, dict(variant=f'volo_h12_l{l}', drop_path_rate=0.1))
run('volo_d2', 384, 32, img_size=384)
run('deit_small_patch16_224', 224, 128, deit=True)
