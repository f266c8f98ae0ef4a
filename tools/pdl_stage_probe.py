"""PDL on / off for the whole training step at the AutoProg stages (two captures per stage, replayed alternately)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autoprog_b200 as A
from autoprog_b200 import kernels as K
from autoprog_b200.optim import FusedAdamW
from autoprog_b200.graph import GraphedTrainStep
dev = torch.device('cuda:0'); B = 128
for l, r in ((9, 128), (12, 160), (15, 192), (18, 224)):
    g = r // 16
    xin = torch.randn(B, 3, r, r, device=dev); tg = torch.softmax(torch.randn(B, 1000, 2 + g * g, device=dev), 1)
    steps = {}
    for on in (False, True):
        K.set_pdl(on)
        torch.manual_seed(0)
        m = A.create_model('model_variant', variant=f'volo_h12_l{l}', img_size=224, drop_path_rate=0.1).to(dev)
        opt = FusedAdamW(m, lr=1e-4, weight_decay=0.05)
        steps[on] = GraphedTrainStep(m, A.TokenLabelCrossEntropy(dense_weight=0.5), opt, xin, tg, bf16=True, warmup=3)
    res = {False: [], True: []}
    for rep in range(3):
        for on in (False, True):
            for _ in range(2):
                steps[on]()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                steps[on]()
            e1.record(); torch.cuda.synchronize()
            res[on].append(e0.elapsed_time(e1) / 10)
    print(f'l{l} r{r}: pdl off {min(res[False]):.3f} ms  on {min(res[True]):.3f} ms  ({[round(v, 3) for v in res[False]]} / {[round(v, 3) for v in res[True]]})', flush=True)
    for s in steps.values():
        s.close()
    del steps, m, opt
    torch.cuda.empty_cache()
K.set_pdl(False)
