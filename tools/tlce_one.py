"""A few TokenLabelCrossEntropy launches at the bench shape (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); bf = torch.bfloat16; B = 128
xa = torch.randn(B, 196, 1000, device=dev).to(bf); xc = torch.randn(B, 1000, device=dev).to(bf)
tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
for _ in range(3):
    K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5)
torch.cuda.synchronize()
