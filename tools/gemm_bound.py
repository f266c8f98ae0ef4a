"""What bounds the tcgen05 GEMM main loop: time one shape with loads / MMAs / stores switched off (apb_debug_gemm_switches).
    python tools/gemm_bound.py [4|5]      # four- or five-stage 128 x 192 kernels (default 5)"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == 'child':
    import torch
    from autoprog_b200 import kernels as K
    K.debug_gemm_switches(int(sys.argv[2]), sys.argv[3] == '5')
    dev = torch.device('cuda:0'); bf = torch.bfloat16
    for (M, N, Kd, tb) in [(25088, 384, 1152, 0), (25088, 1152, 384, 0), (25088, 384, 384, 0), (100352, 192, 576, 0), (25088, 1152, 384, 1), (25088, 384, 1152, 1)]:
        a = torch.randn(M, Kd, device=dev).to(bf); w = torch.randn((Kd, N) if tb else (N, Kd), device=dev).to(bf)
        out = torch.empty(M, N, device=dev, dtype=bf)
        f = lambda: K.gemm(a, w, M, N, Kd, trans_b=bool(tb), out=out)
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3): f()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(30): f()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 30 * 1e3
        print(f'  {M}x{N}x{Kd}: {us:7.1f} us  {2 * M * N * Kd / us / 1e6:7.0f} TFLOP/s-equivalent', flush=True)
else:
    for dbg, what in [(0, 'normal'), (1, 'no TMA loads'), (2, 'no MMAs'), (3, 'no loads, no MMAs'), (4, 'no stores'), (6, 'no MMAs, no stores'), (7, 'barriers + epilogue math only')]:
        stages = sys.argv[1] if len(sys.argv) > 1 else '5'
        print(f'switches={dbg} ({what}), one-CTA kernel, {stages} stages', flush=True)
        subprocess.run([sys.executable, __file__, 'child', str(dbg), stages], env=dict(os.environ, APB_GEMM_PAIR='0'))
