"""Run one GEMM shape a few times (for ncu captures): python tools/gemm_one.py M N K ta tb [f32]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
M, N, Kd, ta, tb = (int(v) for v in sys.argv[1:6])
f32 = len(sys.argv) > 6 and sys.argv[6] == 'f32'
dev = torch.device('cuda:0'); bf = torch.bfloat16
As = [torch.randn((Kd, M) if ta else (M, Kd), device=dev).to(bf) for _ in range(3)]
Bs = [torch.randn((Kd, N) if tb else (N, Kd), device=dev).to(bf) for _ in range(3)]
out = torch.empty(M, N, device=dev, dtype=torch.float32 if f32 else bf)
rs = torch.empty(M, device=dev) if f32 else None
for i in range(3):
    K.gemm(As[i], Bs[i], M, N, Kd, trans_a=bool(ta), trans_b=bool(tb), out=out, out_dtype=out.dtype, rowsum_out=rs)
torch.cuda.synchronize()
