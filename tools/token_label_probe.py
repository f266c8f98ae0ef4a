"""Token-label target builder: csrc/token_label.cu against the same recipe in stock torch / torchvision ops on this GPU
(scatter_add -> roi_align -> softmax -> smoothing; float32).  B=128, C=1000, 18x18 label maps, 14x14 tokens."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import autoprog_b200 as A

dev = torch.device('cuda:0')


def torch_recipe(t, C, smoothing, L):
    from torchvision.ops import roi_align
    B, _, K, Hm, Wm = t.shape
    off = smoothing / C
    on = 1.0 - smoothing + off
    dense = torch.zeros(B, C, Hm, Wm, device=t.device).scatter_add_(1, t[:, 1].long(), t[:, 0])
    rec = t[:, 2, 0, 0, :6]
    idx = torch.arange(B, device=t.device, dtype=t.dtype).view(B, 1)
    boxes = torch.cat([idx, torch.stack([rec[:, 0] * Wm - 0.5, rec[:, 1] * Hm - 0.5, rec[:, 2] * Wm - 0.5, rec[:, 3] * Hm - 0.5], 1)], 1)
    tok = roi_align(dense, boxes, (L, L))
    cls = roi_align(dense, boxes, (1, 1))
    tok = torch.where((rec[:, 4] > 0.5).view(B, 1, 1, 1), tok.flip(3), tok)
    tok, cls = torch.softmax(tok, 1), torch.softmax(cls, 1)
    gt = torch.full((B, C), off, device=t.device).scatter_(1, rec[:, 5].long().view(-1, 1), on)
    return torch.cat([gt.unsqueeze(2), cls.reshape(B, C, 1) * on + off, tok.reshape(B, C, L * L) * on + off], 2)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for B, C, Hm, L in [(128, 1000, 18, 14), (128, 1000, 18, 7), (128, 1000, 18, 24)]:
    g = torch.Generator().manual_seed(0)
    t = torch.zeros(B, 3, 5, Hm, Hm)
    t[:, 0] = torch.randn(B, 5, Hm, Hm, generator=g) * 4
    t[:, 1] = torch.randint(0, C, (B, 5, Hm, Hm), generator=g).float()
    for b in range(B):
        x1, y1 = 0.4 * torch.rand(2, generator=g)
        t[b, 2, 0, 0, :6] = torch.tensor([x1, y1, x1 + 0.3 + 0.3 * torch.rand(1, generator=g).item(), y1 + 0.3 + 0.3 * torch.rand(1, generator=g).item(), b & 1, b])
    t = t.to(dev)
    ours = A.create_token_label_target(t, C, 0.1, L)
    try:
        ref = torch_recipe(t, C, 0.1, L)
        err = float((ours - ref).abs().max())
        t_ref = timeit(lambda: torch_recipe(t, C, 0.1, L))
    except Exception as e:      # torchvision CUDA ops missing
        err, t_ref = float('nan'), float('nan')
        print('torch recipe unavailable:', e)
    t_ours = timeit(lambda: A.create_token_label_target(t, C, 0.1, L))
    out_mb = ours.numel() * 4 / 1e6
    print(f'B={B} C={C} map {Hm}x{Hm} -> {L}x{L}: ours {t_ours:.1f} us ({out_mb / t_ours * 1e6 / 1e6:.2f} TB/s of {out_mb:.0f} MB output)'
          f' | torch ops {t_ref:.1f} us | max abs diff {err:.2e}', flush=True)
