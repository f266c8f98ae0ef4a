"""A few launches of the OutlookAttention kernels at one grid (for ncu captures): python tools/outlook_one.py [HW] [B]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); torch.manual_seed(0); bf = torch.bfloat16
HW = int(sys.argv[1]) if len(sys.argv) > 1 else 28
B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
h = HW // 2
st = torch.cuda.current_stream().cuda_stream
vs = [torch.randn(B, HW, HW, 192, device=dev).to(bf) for _ in range(3)]
lg = torch.randn(B, h, h, 488, device=dev).to(bf); dy = torch.randn_like(vs[0]); y = torch.empty_like(vs[0])
s = 32 ** -0.5
for i in range(3):
    check(lib().apb_outlook_fwd_fma(vs[i].data_ptr(), lg.data_ptr(), y.data_ptr(), B, HW, HW, 6, s, 488, st), 'fma')
    check(lib().apb_outlook_fwd_mma(vs[i].data_ptr(), lg.data_ptr(), y.data_ptr(), B, HW, HW, 6, s, 488, st), 'mma')
    K.outlook_bwd(vs[i], lg, dy, 6, s)
torch.cuda.synchronize()
