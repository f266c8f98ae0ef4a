"""Per-warp event timeline of CTA 0 of the tcgen05 MHSA kernels (apb_debug_mhsa_trace)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); bf = torch.bfloat16
B, N, H = 128, 196, 12
st = lambda: torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B, N, 3 * H * 32, device=dev).to(bf); do = torch.randn(B, N, H * 32, device=dev).to(bf)
out = torch.empty(B, N, H * 32, device=dev, dtype=bf); lse = torch.empty(B, H, N, device=dev); dq = torch.empty_like(qkv); ws = torch.empty(B * H * N, device=dev)
which = sys.argv[1] if len(sys.argv) > 1 else 'bwd'
def run():
    if which == 'fwd':
        check(lib().apb_mhsa_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, 32 ** -0.5, st()), 'f')
    else:
        check(lib().apb_mhsa_bwd_tc(qkv.data_ptr(), out.data_ptr(), do.data_ptr(), lse.data_ptr(), dq.data_ptr(), ws.data_ptr(), B, N, H, 32, 32 ** -0.5, st()), 'b')
check(lib().apb_mhsa_fwd_tc(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, N, H, 32, 32 ** -0.5, st()), 'f')
run(); torch.cuda.synchronize()
buf = torch.zeros(18, 1024, dtype=torch.int64, device=dev)
lib().apb_debug_mhsa_trace(buf.data_ptr())
run(); torch.cuda.synchronize()
lib().apb_debug_mhsa_trace(None)
b = buf.cpu()
t0 = min(int(v) & 0xFFFFFFFFFFFF for v in b.flatten().tolist() if v != 0)
names_f = {1: 'wait_S', 2: 'got_S', 3: 'p1done', 4: 'xchg', 5: 'arriveP', 6: 'got_O', 20: 'S0', 21: 'S1', 30: 'PV0', 31: 'PV1', 40: 'PV0end', 41: 'PV1end'}
names_b = {1: 'wait_sdp', 2: 'got_sdp', 3: 'wait_mmadone', 4: 'got_mmadone', 5: 'arrive_pds', 6: 'wait_dkv', 7: 'got_dkv', 8: 'dkv_stored', 9: 'wait_dq', 10: 'got_dq',
           20: 'SDP0', 21: 'SDP1', 30: 'PROD0', 31: 'PROD1', 40: 'PROD0end', 41: 'PROD1end'}
names = names_f if which == 'fwd' else names_b
for w in (0, 4, 8, 12, 17):
    ev = [(int(v) >> 48, (int(v) & 0xFFFFFFFFFFFF) - t0) for v in b[w].tolist() if v != 0]
    print(f'--- warp {w}: {len(ev)} events')
    prev = 0
    for i, (e, t) in enumerate(ev[:70]):
        print(f'   {t:8d} (+{t - prev:6d}) {names.get(e, e)}')
        prev = t
