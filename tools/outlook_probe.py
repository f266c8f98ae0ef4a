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
