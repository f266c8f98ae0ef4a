"""One launch each of the HBM-bound headline kernels at BASELINE sizes (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); torch.manual_seed(0); bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
B = 128
v = torch.randn(B, 28, 28, 192, device=dev).to(bf); lg = torch.randn(B, 14, 14, 488, device=dev).to(bf); dy = torch.randn_like(v); y = torch.empty_like(v)
xa = torch.randn(B, 196, 1000, device=dev).to(bf); xc = torch.randn(B, 1000, device=dev).to(bf); tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
for _ in range(2):
    lib().apb_outlook_fwd_fma(v.data_ptr(), lg.data_ptr(), y.data_ptr(), B, 28, 28, 6, 32 ** -0.5, 488, st())
    lib().apb_outlook_fwd_mma(v.data_ptr(), lg.data_ptr(), y.data_ptr(), B, 28, 28, 6, 32 ** -0.5, 488, st())
    K.outlook_bwd(v, lg, dy, 6, 32 ** -0.5)
    K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5)
torch.cuda.synchronize()
