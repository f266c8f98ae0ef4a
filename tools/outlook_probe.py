import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); torch.manual_seed(0); bf = torch.bfloat16
B = 128
v = torch.randn(B, 28, 28, 192, device=dev).to(bf); lg = torch.randn(B, 14, 14, 488, device=dev).to(bf); dy = torch.randn_like(v)
for _ in range(3):
    K.outlook_fwd(v, lg, 6, 32 ** -0.5)
    K.outlook_bwd(v, lg, dy, 6, 32 ** -0.5)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    vs = [torch.randn_like(v) for _ in range(3)]
    def t(fn, n=20):
        for i in range(3): fn(i % 3)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i % 3)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    tf = t(lambda i: K.outlook_fwd(vs[i], lg, 6, 32 ** -0.5))
    tb = t(lambda i: K.outlook_bwd(vs[i], lg, dy, 6, 32 ** -0.5))
    es = 2
    fby = (2 * v.numel() + B * 196 * 486) * es; bby = (3 * v.numel() + 2 * B * 196 * 486) * es
    print(f'outlook B={B} 28x28x192: fwd {tf:.1f} us ({fby / tf / 1e3:.0f} GB/s)  bwd {tb:.1f} us ({bby / tb / 1e3:.0f} GB/s)  fwd+bwd {(fby + bby) / (tf + tb) / 1e3:.0f} GB/s')
