import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0')
M, N, Kd = 25088, 1152, 384
a = torch.randn(M, Kd, device=dev).bfloat16(); w = torch.randn(N, Kd, device=dev).bfloat16(); bias = torch.randn(N, device=dev)
for epi in (0, 1):
    for _ in range(3):
        K.gemm(a, w, M, N, Kd, bias=bias, epilogue=epi)
torch.cuda.synchronize()
if len(sys.argv) > 1:
    import time
    for epi in (0, 1):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): K.gemm(a, w, M, N, Kd, bias=bias, epilogue=epi)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print(f'epi={epi}: {ms*1e3:.1f} us  {2*M*N*Kd/ms/1e9:.0f} TFLOP/s')
