"""Run the non-GEMM hot kernels once each at BASELINE sizes (for ncu --set full captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); torch.manual_seed(0)
B = 128; bf = torch.bfloat16
rows, C = B * 196, 384
x = torch.randn(rows, C, device=dev); r = torch.randn(rows, C, device=dev).to(bf); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
dy = torch.randn(rows, C, device=dev).to(bf); dres = torch.randn(rows, C, device=dev)
qkv = torch.randn(B, 196, 3 * 384, device=dev).to(bf); do = torch.randn(B, 196, 384, device=dev).to(bf)
v = torch.randn(B, 28, 28, 192, device=dev).to(bf); lg = torch.randn(B, 14, 14, 488, device=dev).to(bf); dyo = torch.randn_like(v)
xa = torch.randn(B, 196, 1000, device=dev).to(bf); xc = torch.randn(B, 1000, device=dev).to(bf); tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
big = torch.randn(B * 784, 576, device=dev).to(bf)
for it in range(2):
    xs, y, mean, rstd = K.ln_fwd(x, g, b, 1e-5, bf, r=r)
    K.ln_bwd(dy, xs, mean, rstd, g, dres=dres, want_dr=True)
    out, lse = K.mhsa_fwd(qkv, 12, 32 ** -0.5)
    K.mhsa_bwd(qkv, out, do, lse, 12, 32 ** -0.5)
    K.outlook_fwd(v, lg, 6, 32 ** -0.5)
    K.outlook_bwd(v, lg, dyo, 6, 32 ** -0.5)
    K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5)
    K.colsum(big, 576)
torch.cuda.synchronize()
