"""Cycles per tcgen05.mma by operand layout (csrc/umma_timing.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200._lib import lib, check
dev = torch.device('cuda:0')
names = {0: 'SS K-major SW64', 1: 'TS + B MN SW64 (N=32)', 2: 'SS A MN SW64 + B MN SW64', 3: 'SS A K SW64 + B MN SW64', 4: 'SS K-major SW128', 5: 'SS A MN SW128 + B MN SW128'}
for kind, N in [(0, 208), (0, 96), (0, 32), (1, 32), (2, 32), (3, 32), (4, 208), (4, 32), (5, 32), (5, 192)]:
    for reps in (8, 64):
        out = torch.zeros(4, dtype=torch.int64, device=dev)
        check(lib().apb_debug_umma_timing(out.data_ptr(), kind, N, reps, torch.cuda.current_stream().cuda_stream), 't')
        torch.cuda.synchronize()
        o = out.tolist()
        print(f'kind {kind} ({names[kind]}) N={N} reps={reps}: issue {o[2]} cyc ({o[2]/reps:.1f}/mma), issue+complete {o[3]} cyc ({o[3]/reps:.1f}/mma)', flush=True)
