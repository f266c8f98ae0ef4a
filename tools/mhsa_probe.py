"""MHSA core timing at the volo_d1 stage-2 shape (B=128, N=196, 12 heads, D=32) and the DeiT-small shape (D=64)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); torch.manual_seed(0); bf = torch.bfloat16
for (B, N, H, D) in [(128, 196, 12, 32), (128, 197, 6, 64)]:
    qkvs = [torch.randn(B, N, 3 * H * D, device=dev).to(bf) for _ in range(4)]
    do = torch.randn(B, N, H * D, device=dev).to(bf)
    out, lse = K.mhsa_fwd(qkvs[0], H, D ** -0.5)
    def t(fn, n=20):
        for i in range(3): fn(i % 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i % 4)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    tf = t(lambda i: K.mhsa_fwd(qkvs[i], H, D ** -0.5))
    tb = t(lambda i: K.mhsa_bwd(qkvs[i], out, do, lse, H, D ** -0.5))
    print(f'B={B} N={N} heads={H} D={D}: fwd {tf:.1f} us  bwd {tb:.1f} us', flush=True)
