"""Direct check + timing of the tcgen05 MHSA kernels (apb_mhsa_fwd_tc / apb_mhsa_bwd_tc) against torch fp32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200._lib import lib, check
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); bf = torch.bfloat16
st = lambda: torch.cuda.current_stream().cuda_stream
which = sys.argv[1] if len(sys.argv) > 1 else 'all'

def ref(qkv, do, H, scale):
    B, N, _ = qkv.shape
    x = qkv.float().requires_grad_(True)
    q, k, v = x.reshape(B, N, 3, H, 32).permute(2, 0, 3, 1, 4)
    a = torch.softmax(q @ k.transpose(-1, -2) * scale, -1)
    o = (a @ v).permute(0, 2, 1, 3).reshape(B, N, H * 32)
    o.backward(do.float())
    lse = torch.logsumexp(q @ k.transpose(-1, -2) * scale, -1)
    return o.detach(), lse.detach(), x.grad

def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())

for (B, N, H) in [(2, 196, 2), (1, 49, 3), (3, 197, 2), (2, 144, 12), (1, 224, 1), (2, 64, 3), (1, 1, 1), (1, 129, 1), (5, 100, 4), (37, 196, 12)]:
    torch.manual_seed(N + H)
    qkv = torch.randn(B, N, 3 * H * 32, device=dev).to(bf)
    do = torch.randn(B, N, H * 32, device=dev).to(bf)
    scale = 32 ** -0.5
    o_ref, lse_ref, dqkv_ref = ref(qkv, do, H, scale)
    out = torch.full((B, N, H * 32), float('nan'), device=dev, dtype=bf)
    lse = torch.full((B, H, N), float('nan'), device=dev)
    msg = f'B={B} N={N} H={H}:'
    if which in ('all', 'fwd'):
        check(lib().apb_mhsa_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, scale, st()), 'fwd_tc')
        torch.cuda.synchronize()
        msg += f' fwd out {rel(out, o_ref):.2e} lse {rel(lse, lse_ref):.2e}'
    if which in ('all', 'bwd'):
        o_in = o_ref.to(bf)
        dqkv = torch.full_like(qkv, float('nan'))
        wsb = torch.empty(B * H * N, device=dev)
        check(lib().apb_mhsa_bwd_tc(qkv.data_ptr(), o_in.data_ptr(), do.data_ptr(), lse_ref.contiguous().data_ptr(), dqkv.data_ptr(), wsb.data_ptr(), B, N, H, 32, scale, st()), 'bwd_tc')
        torch.cuda.synchronize()
        d = dqkv.reshape(B, N, 3, H * 32); r = dqkv_ref.reshape(B, N, 3, H * 32)
        msg += f' bwd dq {rel(d[:, :, 0], r[:, :, 0]):.2e} dk {rel(d[:, :, 1], r[:, :, 1]):.2e} dv {rel(d[:, :, 2], r[:, :, 2]):.2e}'
    print(msg, flush=True)

# timing at the volo_d1 stage-2 shape and the earlier stages
for (B, N, H) in [(128, 196, 12), (128, 144, 12), (128, 100, 12), (128, 64, 12)]:
    qkvs = [torch.randn(B, N, 3 * H * 32, device=dev).to(bf) for _ in range(4)]
    do = torch.randn(B, N, H * 32, device=dev).to(bf)
    out = torch.empty(B, N, H * 32, device=dev, dtype=bf); lse = torch.empty(B, H, N, device=dev); dq = torch.empty_like(qkvs[0])
    ws = torch.empty(B * H * N, device=dev)
    scale = 32 ** -0.5
    def t(fn, n=20):
        for i in range(3): fn(i % 4)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i % 4)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    res = {}
    if which in ('all', 'fwd'):
        res['fwd_tc'] = t(lambda i: lib().apb_mhsa_fwd_tc(qkvs[i].data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, scale, st()))
    res['fwd_mma'] = t(lambda i: lib().apb_mhsa_fwd_mma(qkvs[i].data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, scale, st())) if hasattr(lib(), 'apb_mhsa_fwd_mma') else None
    if which in ('all', 'bwd'):
        res['bwd_tc'] = t(lambda i: lib().apb_mhsa_bwd_tc(qkvs[i].data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), ws.data_ptr(), B, N, H, 32, scale, st()))
    res['bwd_mma'] = t(lambda i: lib().apb_mhsa_bwd_mma(qkvs[i].data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), ws.data_ptr(), B, N, H, 32, scale, st())) if hasattr(lib(), 'apb_mhsa_bwd_mma') else None
    print(f'B={B} N={N} heads={H}: ' + '  '.join(f'{k} {v:.1f} us' for k, v in res.items() if v is not None), flush=True)
