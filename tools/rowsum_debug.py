import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); torch.manual_seed(0)
for (M, N, Kd) in [(384, 384, 25088), (256, 128, 1024)]:
    a = torch.randn(Kd, M, device=dev).bfloat16(); b = torch.randn(Kd, N, device=dev).bfloat16()
    lib = K.lib(); st = torch.cuda.current_stream().cuda_stream
    split = int(lib.apb_gemm_tc_suggest_split(M, N, Kd)); slots = int(lib.apb_gemm_tc_rowsum_slots(N, split))
    parts = torch.zeros(split, M, N, device=dev); rparts = torch.full((slots, M), 7.0, device=dev)
    rc = lib.apb_gemm_tc_rowsum(a.data_ptr(), b.data_ptr(), parts.data_ptr(), None, None, M, N, Kd, 1, 1, 0, K.BF16 if hasattr(K,'BF16') else 1, K.F32, split, rparts.data_ptr(), st)
    torch.cuda.synchronize()
    print('rc', rc, 'split', split, 'slots', slots, 'nan per slot', torch.isnan(rparts).sum(1).tolist()[:12], 'untouched', (rparts == 7.0).sum(1).tolist()[:12])
    ref = a.float().sum(0)
    print('sum err', float((rparts.sum(0) - ref).abs().max()), 'ref max', float(ref.abs().max()))
    print(rparts[:4, :6])
