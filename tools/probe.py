"""Quick device timings of the main kernels at BASELINE sizes (CUDA events, L2 flushed between iterations)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from autoprog_b200 import kernels as K

dev = torch.device('cuda:0')
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
for dtype, nm, es in ((torch.bfloat16, 'bf16', 2), (torch.float32, 'f32', 4)):
    B, H, W, heads = 128, 28, 28, 6
    v = torch.randn(B, H, W, heads * 32, device=dev).to(dtype)
    lg = torch.randn(B, 14, 14, heads * 81, device=dev).to(dtype)
    dy = torch.randn_like(v)
    s = 32 ** -0.5
    for simt in (True, False):
        tf = timeit(lambda: K.outlook_fwd(v, lg, heads, s, simt=simt))
        tb = timeit(lambda: K.outlook_bwd(v, lg, dy, heads, s, simt=simt))
        elems_f = 2 * v.numel() + lg.numel()
        elems_b = 3 * v.numel() + 2 * lg.numel()
        res[f'outlook_{nm}_{"simt" if simt else "main"}'] = dict(fwd_ms=tf, bwd_ms=tb, fwd_GBs=elems_f * es / tf / 1e6,
                                                                  bwd_GBs=elems_b * es / tb / 1e6,
                                                                  fwdbwd_GBs=(elems_f + elems_b) * es / (tf + tb) / 1e6)
    xa = torch.randn(B, 196, 1000, device=dev).to(dtype)
    xc = torch.randn(B, 1000, device=dev).to(dtype)
    tg = torch.softmax(torch.randn(B, 1000, 198, device=dev), 1)
    t = timeit(lambda: K.tlce_fwd_bwd(xc, xa, tg, 4, 1.0, 0.5))
    res[f'tlce_{nm}'] = dict(ms=t, GBs=B * 196 * 1000 * (2 * es + 4) / t / 1e6)
    x = torch.randn(B * 784, 192, device=dev)
    r = torch.randn(B * 784, 192, device=dev).to(dtype)
    g = torch.ones(192, device=dev); b = torch.zeros(192, device=dev)
    t = timeit(lambda: K.ln_fwd(x, g, b, 1e-5, dtype, r=r))
    res[f'ln_fwd_{nm}'] = dict(ms=t, GBs=x.numel() * (8 + 2 * es) / t / 1e6)
    M, N, Kd = 25088, 1152, 384
    a = torch.randn(M, Kd, device=dev).to(dtype); w = torch.randn(N, Kd, device=dev).to(dtype)
    for force in (True, False):
        K._FORCE_SIMT = force
        try:
            t = timeit(lambda: K.gemm(a, w, M, N, Kd))
            res[f'gemm_{nm}_{"simt" if force else "auto"}'] = dict(ms=t, TFs=2 * M * N * Kd / t / 1e9)
        except Exception as e:
            res[f'gemm_{nm}_{"simt" if force else "auto"}'] = str(e)[:100]
    K._FORCE_SIMT = False
    qkv = torch.randn(B, 196, 3 * 384, device=dev).to(dtype)
    t = timeit(lambda: K.mhsa_fwd(qkv, 12, s))
    res[f'mhsa_fwd_{nm}'] = dict(ms=t)
print(json.dumps(res, indent=1))
os.makedirs('gpurun_out', exist_ok=True)
json.dump(res, open('gpurun_out/probe.json', 'w'), indent=1)
