"""One training step of the bench workload inside a cudaProfiler range (for `ncu --profile-from-start off`)."""
import os, sys, copy
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autoprog_b200 as A
from autoprog_b200.optim import FusedAdamW

B = int(os.environ.get('B', 128)); res = int(os.environ.get('RES', 224)); warm = int(os.environ.get('WARM', 3))
dev = torch.device('cuda:0'); torch.manual_seed(0); np.random.seed(0)
model = A.create_model(os.environ.get('MODEL', 'volo_d1'), img_size=224, drop_path_rate=0.1).to(dev)
decays = [0.998, 0.9986, 0.999, 0.9996]
emas = [copy.deepcopy(model).eval() for _ in decays]
opt = FusedAdamW(model, lr=2e-4, weight_decay=0.05, ema_models=emas, ema_decays=decays)
crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
g = res // 16
x = torch.randn(B, 3, res, res, device=dev); t = torch.softmax(torch.randn(B, 1000, 2 + g * g, device=dev), 1)
def step():
    opt.zero_grad()
    with A.autocast():
        loss = crit(model(x), t)
    loss.backward(); opt.step()
for _ in range(warm): step()
torch.cuda.synchronize(); torch.cuda.profiler.start(); step(); torch.cuda.synchronize(); torch.cuda.profiler.stop()
