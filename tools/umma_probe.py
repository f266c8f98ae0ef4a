"""Descriptor-convention probe (csrc/umma_probe.cu): prints the max error of one tcgen05.mma per layout mode."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0'); torch.manual_seed(0)
A = torch.randn(128, 64, device=dev).to(torch.bfloat16)
B = torch.randn(64, 64, device=dev).to(torch.bfloat16)
names = {0: 'B MN-major SW128, N=32 at byte offset inside the row', 1: 'A,B K-major SW64 (64-byte rows)',
         2: 'B MN-major SW64', 3: 'A MN-major SW128 (M=128) + B MN-major SW128'}
for mode in (3, 0, 1, 2):
    for off in (0, 32):
        D = torch.full((128, 32), float('nan'), device=dev)
        check(lib().apb_debug_umma_probe(A.data_ptr(), B.data_ptr(), D.data_ptr(), mode, off, torch.cuda.current_stream().cuda_stream), 'probe')
        torch.cuda.synchronize()
        ref = A.float() @ B.float()[:, off:off + 32]
        err = float((D - ref).abs().max())
        print(f'mode {mode} ({names[mode]}) off={off}: max err {err:.4f}  {"OK" if err < 0.05 else "MISMATCH"}', flush=True)
