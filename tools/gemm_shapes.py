"""Time every distinct GEMM shape of one volo_d1 training step (B=128, 224 px) in isolation: the tcgen05 kernel
(autoprog_b200.kernels.gemm, split-K included) next to cuBLAS (torch.matmul) on the same operands.

    python tools/gemm_shapes.py            # prints one line per shape

Buffers rotate over 4 copies so operands are not L2-resident between iterations.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K

dev = torch.device('cuda:0')
bf = torch.bfloat16
# (M, N, K, trans_a, trans_b, epilogue, out_f32, launches per step)
SHAPES = [
    (1152, 384, 25088, 1, 1, 0, 1, 28), (25088, 1152, 384, 0, 0, 1, 0, 14), (25088, 384, 1152, 0, 1, 0, 0, 28),
    (384, 1152, 25088, 1, 1, 0, 1, 14), (25088, 1152, 384, 0, 1, 2, 0, 14), (25088, 384, 1152, 0, 0, 0, 0, 14),
    (25088, 1152, 384, 0, 0, 0, 0, 14), (100352, 576, 192, 0, 0, 1, 0, 4), (384, 384, 25088, 1, 1, 0, 1, 14),
    (100352, 576, 192, 0, 1, 2, 0, 4), (192, 192, 100352, 1, 1, 0, 1, 8), (25088, 384, 384, 0, 1, 0, 0, 14),
    (25088, 384, 384, 0, 0, 0, 0, 14), (192, 576, 100352, 1, 1, 0, 1, 4), (576, 192, 100352, 1, 1, 0, 1, 4),
    (100352, 192, 576, 0, 0, 0, 0, 4), (100352, 192, 192, 0, 0, 0, 0, 8), (100352, 192, 192, 0, 1, 0, 0, 8),
    (100352, 192, 576, 0, 1, 0, 0, 4), (488, 192, 25088, 1, 1, 0, 1, 4), (25088, 192, 488, 0, 1, 0, 0, 4),
    (25088, 488, 192, 0, 0, 0, 0, 4),
]
NBUF, ITERS = 4, 12


def timed(fn):
    """Device time per call: ITERS calls captured in one CUDA graph and replayed (an eager loop is HOST-bound below ~15 us
    per call here: ctypes + four cuTensorMapEncode calls per launch)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for i in range(3):
            fn(i % NBUF)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(ITERS):
            fn(i % NBUF)
    g.replay()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / ITERS * 1e3)
    return best   # us


def main():
    tot_ours = tot_blas = 0.0
    for M, N, Kd, ta, tb, epi, of32, cnt in SHAPES + [(25088, 1000, 384, 0, 0, 0, 0, 1), (100352, 1024, 192, 0, 0, 0, 0, 1)]:
        As = [torch.randn((Kd, M) if ta else (M, Kd), device=dev).to(bf) for _ in range(NBUF)]
        Bs = [torch.randn((Kd, N) if tb else (N, Kd), device=dev).to(bf) for _ in range(NBUF)]
        aux = torch.randn(M, N, device=dev).to(bf) if epi == 2 else None
        bias = torch.randn(N, device=dev) if epi == 1 else None
        odt = torch.float32 if of32 else bf
        out = torch.empty(M, N, device=dev, dtype=odt)

        def ours(i):
            K.gemm(As[i], Bs[i], M, N, Kd, trans_a=bool(ta), trans_b=bool(tb), bias=bias, epilogue=epi, aux=aux,
                   out=out if epi != 1 else None, out_dtype=odt)

        def blas(i):
            a = As[i].t() if ta else As[i]
            b = Bs[i] if tb else Bs[i].t()
            torch.matmul(a, b)          # bf16 out, no epilogue: a lower bound for what a library call costs here

        t_p = None
        if not ta and not tb and epi == 0 and not of32:
            def pair(i):
                K.check(K.lib().apb_gemm_tc_pair(As[i].data_ptr(), Bs[i].data_ptr(), out.data_ptr(), None, M, N, Kd, K.BF16,
                                                 torch.cuda.current_stream().cuda_stream), 'pair')
            t_p = timed(pair)
        t_o, t_b = timed(ours), timed(blas)
        fl = 2.0 * M * N * Kd
        tot_ours += t_o * cnt
        tot_blas += t_b * cnt
        print(f'M={M:6d} N={N:5d} K={Kd:6d} ta={ta} tb={tb} epi={epi} f32={of32} x{cnt:2d}: ours {t_o:7.1f} us {fl / t_o / 1e6:6.0f} TF/s'
              f' | cuBLAS(plain bf16) {t_b:7.1f} us {fl / t_b / 1e6:6.0f} TF/s | split={K.lib().apb_gemm_tc_suggest_split(M, N, Kd) if of32 else 1}' + (f' | pair(cta_group::2) {t_p:7.1f} us {fl / t_p / 1e6:6.0f} TF/s' if t_p else ''),
              flush=True)
    print(f'per-step total: ours {tot_ours / 1e3:.3f} ms, cuBLAS plain {tot_blas / 1e3:.3f} ms')


if __name__ == '__main__':
    main()
