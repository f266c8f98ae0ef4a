"""Programmatic dependent launch: time per launch of back-to-back kernels inside a CUDA graph, attribute on vs off."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K
dev = torch.device('cuda:0'); bf = torch.bfloat16


def graph_time(fn, n=100, reps=5):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n):
            fn()
    g.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def both(name, fn):
    r = {}
    for on in (False, True):
        K.set_pdl(on)
        r[on] = graph_time(fn)
    print(f'{name}: {r[False]:.2f} us / launch plain, {r[True]:.2f} us with PDL ({r[False] - r[True]:+.2f})', flush=True)


for (M, N, Kd) in [(25088, 384, 384), (25088, 1152, 384), (25088, 384, 1152), (100352, 192, 192), (1568, 384, 384)]:
    a = torch.randn(M, Kd, device=dev).to(bf); w = torch.randn(N, Kd, device=dev).to(bf); out = torch.empty(M, N, device=dev, dtype=bf)
    both(f'gemm NT {M}x{N}x{Kd}', lambda: K.gemm(a, w, M, N, Kd, out=out))
M, N, Kd = 384, 1152, 25088
a = torch.randn(Kd, M, device=dev).to(bf); w = torch.randn(Kd, N, device=dev).to(bf); out = torch.empty(M, N, device=dev)
both('wgrad TN 384x1152x25088 (split-K + reduce: 2 launches)', lambda: K.gemm(a, w, M, N, Kd, trans_a=True, trans_b=True, out=out, out_dtype=torch.float32))
x = torch.randn(25088, 384, device=dev); r = torch.randn(25088, 384, device=dev).to(bf)
g_ = torch.ones(384, device=dev); b_ = torch.zeros(384, device=dev)
both('layernorm fwd 25088x384 (fp32 stream + bf16 branch)', lambda: K.ln_fwd(x, g_, b_, 1e-5, bf, r=r))

# heterogeneous chain as in an MLP block: LN -> fc1 -> fc2 -> LN
M = 25088
x = torch.randn(M, 384, device=dev); r = torch.randn(M, 384, device=dev).to(bf)
w1 = torch.randn(1152, 384, device=dev).to(bf); w2 = torch.randn(384, 1152, device=dev).to(bf)
h = torch.empty(M, 1152, device=dev, dtype=bf); o = torch.empty(M, 384, device=dev, dtype=bf)
def mlp():
    _, y, _, _ = K.ln_fwd(x, g_, b_, 1e-5, bf, r=r)
    K.gemm(y, w1, M, 1152, 384, out=h)
    K.gemm(h, w2, M, 384, 1152, out=o)
both('chain LN -> fc1 -> fc2 (3 launches, per launch)', mlp)

# the whole training step, two captures in one process (attribute off / on), replayed alternately
import copy
import autoprog_b200 as A
from autoprog_b200.optim import FusedAdamW
from autoprog_b200.graph import GraphedTrainStep
steps = {}
B = 128
xin = torch.randn(B, 3, 224, 224, device=dev); tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
for on in (False, True):
    K.set_pdl(on)
    torch.manual_seed(0)
    m = A.create_model('volo_d1', img_size=224, drop_path_rate=0.1).to(dev)
    opt = FusedAdamW(m, lr=1e-4, weight_decay=0.05)
    steps[on] = GraphedTrainStep(m, A.TokenLabelCrossEntropy(dense_weight=0.5), opt, xin, tg, bf16=True, warmup=3)
for rep in range(3):
    for on in (False, True):
        for _ in range(2):
            steps[on]()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            steps[on]()
        e1.record(); torch.cuda.synchronize()
        print(f'train step (graph, B=128, no EMA) pdl={int(on)}: {e0.elapsed_time(e1) / 10:.3f} ms', flush=True)
