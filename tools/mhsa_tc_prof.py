"""One launch each of the tcgen05 MHSA kernels at the volo_d1 stage-2 shape (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); bf = torch.bfloat16
B, N, H = 128, int(os.environ.get('N', 196)), 12
st = lambda: torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B, N, 3 * H * 32, device=dev).to(bf); do = torch.randn(B, N, H * 32, device=dev).to(bf)
out = torch.empty(B, N, H * 32, device=dev, dtype=bf); lse = torch.empty(B, H, N, device=dev); dq = torch.empty_like(qkv); ws = torch.empty(B * H * N, device=dev)
for _ in range(2):
    check(lib().apb_mhsa_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, 32 ** -0.5, st()), 'f')
    check(lib().apb_mhsa_bwd_tc(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), ws.data_ptr(), B, N, H, 32, 32 ** -0.5, st()), 'b')
torch.cuda.synchronize()
