"""OutlookAttention core timing at the AutoProg stage shapes: fma-gather vs mma.sync forward, backward, TLCE."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); torch.manual_seed(0); bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
def t(fn, n=20):
    for i in range(3): fn(i % 3)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i % 3)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B = 128
for HW in (28, 24, 20, 16):
    h = HW // 2
    vs = [torch.randn(B, HW, HW, 192, device=dev).to(bf) for _ in range(3)]
    lg = torch.randn(B, h, h, 488, device=dev).to(bf); dy = torch.randn_like(vs[0]); y = torch.empty_like(vs[0])
    s = 32 ** -0.5
    tf = t(lambda i: lib().apb_outlook_fwd_fma(vs[i].data_ptr(), lg.data_ptr(), y.data_ptr(), B, HW, HW, 6, s, 488, st()))
    tm = t(lambda i: lib().apb_outlook_fwd_mma(vs[i].data_ptr(), lg.data_ptr(), y.data_ptr(), B, HW, HW, 6, s, 488, st()))
    dv = torch.empty_like(vs[0]); dl = torch.empty_like(lg)
    tb = t(lambda i: lib().apb_outlook_bwd_fma(vs[i].data_ptr(), lg.data_ptr(), dy.data_ptr(), dv.data_ptr(), dl.data_ptr(), B, HW, HW, 6, s, 488, st()))
    tbm = t(lambda i: lib().apb_outlook_bwd_mma(vs[i].data_ptr(), lg.data_ptr(), dy.data_ptr(), dv.data_ptr(), dl.data_ptr(), B, HW, HW, 6, s, 488, st()))
    fby = (2 * vs[0].numel() + B * h * h * 486) * 2; bby = (3 * vs[0].numel() + 2 * B * h * h * 486) * 2
    print(f'outlook B={B} {HW}x{HW}x192: fwd fma {tf:.1f} us ({fby / tf / 1e3:.0f} GB/s)  fwd mma {tm:.1f} us  bwd fma {tb:.1f} us ({bby / tb / 1e3:.0f} GB/s)  bwd mma {tbm:.1f} us', flush=True)
xa = torch.randn(B, 196, 1000, device=dev).to(bf); xc = torch.randn(B, 1000, device=dev).to(bf); tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
tt = t(lambda i: K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5))
print(f'tlce B=128 N=196 C=1000 bf16: {tt:.1f} us ({xa.numel() * 8 / tt / 1e3:.0f} GB/s)')
