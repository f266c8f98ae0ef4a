"""Soak: N graph-replayed training steps of volo_d1 (B=64, 224 px, bf16, fused AdamW + 4 EMA) on a FIXED synthetic batch.
The loss must fall monotonically-ish from ln(1000)-level to well below it (the model memorises the batch), stay finite, and the
EMA weights must track the model.  Prints the loss every 20 steps and ms / step."""
import copy, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import autoprog_b200 as A
from autoprog_b200.optim import FusedAdamW
from autoprog_b200.graph import GraphedTrainStep

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
dev = torch.device('cuda:0'); torch.manual_seed(0); np.random.seed(0)
B = 64
model = A.create_model('volo_d1', img_size=224, drop_path_rate=0.1).to(dev)
decays = [0.998, 0.9986, 0.999, 0.9996]
emas = [copy.deepcopy(model).eval() for _ in decays]
opt = FusedAdamW(model, lr=5e-4, weight_decay=0.05, ema_models=emas, ema_decays=decays)
crit = A.TokenLabelCrossEntropy(dense_weight=0.5)
# every image = its own smooth pattern (a random 7 x 7 map, upsampled) + noise, so the batch can be told apart
x = torch.nn.functional.interpolate(torch.randn(B, 3, 7, 7, device=dev) * 2, size=(224, 224), mode='bilinear') + 0.3 * torch.randn(B, 3, 224, 224, device=dev)
labels = torch.randint(0, 1000, (B,), device=dev)
t = torch.full((B, 1000, 198), 0.1 / 1000, device=dev)
t.scatter_(1, labels.view(B, 1, 1).expand(B, 1, 198), 0.9)           # every slot: smoothed one-hot of the image's label
step = GraphedTrainStep(model, crit, opt, x, t, bf16=True, warmup=2)
losses = []
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(steps):
    loss = step(x, t)
    if i % 20 == 0 or i == steps - 1:
        losses.append(float(loss))
        print(f'step {i:4d}  loss {losses[-1]:.4f}', flush=True)
torch.cuda.synchronize()
print(f'{(time.perf_counter() - t0) / steps * 1e3:.2f} ms / step (B={B}, incl. the loss read every 20 steps)')
assert all(np.isfinite(losses)), losses
assert losses[-1] < 0.75 * losses[0], (losses[0], losses[-1])
w = dict(model.named_parameters())['network.0.0.attn.v.weight']
e = dict(emas[0].named_parameters())['network.0.0.attn.v.weight']
print('ema[0] rel distance to model:', float(((w - e).norm() / w.norm()).detach()))
assert torch.isfinite(e).all() and float(((w - e).norm() / w.norm()).detach()) < 1.0
print('soak ok')
