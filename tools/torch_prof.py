"""Which torch (non-library) kernels remain in a step and which aten ops launch them (torch.profiler, one eager step)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autoprog_b200 as A
from autoprog_b200.optim import FusedAdamW
from torch.profiler import profile, ProfilerActivity
dev = torch.device('cuda:0'); torch.manual_seed(0); np.random.seed(0)
model = A.create_model('volo_d1', img_size=224, drop_path_rate=0.1).to(dev)
decays = [0.998, 0.9986, 0.999, 0.9996]
emas = [copy.deepcopy(model).eval() for _ in decays]
opt = FusedAdamW(model, lr=2e-4, weight_decay=0.05, ema_models=emas, ema_decays=decays)
crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
x = torch.randn(128, 3, 224, 224, device=dev); t = torch.softmax(torch.randn(128, 1000, 198, device=dev), 1)
def step():
    opt.zero_grad()
    with A.autocast():
        loss = crit(model(x), t)
    loss.backward(); opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=False, record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith('aten::') and e.device_time_total > 0]
rows.sort(key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    print(f'{e.key:38s} n={e.count:4d} self_cuda={e.self_device_time_total:9.1f} us  shapes={str(e.input_shapes)[:110]}')
