import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); bf = torch.bfloat16
M, N, Kd = 25088, 1152, 384
a = torch.randn(M, Kd, device=dev).to(bf); w = torch.randn(N, Kd, device=dev).to(bf); bias = torch.randn(N, device=dev)
dy = torch.randn(M, Kd, device=dev).to(bf); w2 = torch.randn(Kd, N, device=dev).to(bf)
for _ in range(3):
    out, aux = K.gemm(a, w, M, N, Kd, bias=bias, epilogue=K.EPI_GELU)
    K.gemm(dy, w2, M, N, Kd, trans_b=True, epilogue=K.EPI_DGELU, aux=aux)
torch.cuda.synchronize()
