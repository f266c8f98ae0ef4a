"""Thin tensor-level wrappers over the C ABI (no autograd here; see ops.py).

Every function takes/returns contiguous CUDA tensors, allocates outputs with torch's caching allocator,
launches on torch's current stream and raises `KernelError` on failure.  No CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import os

import torch

from ._lib import check, lib

F32, BF16 = 0, 1
_CODES = {torch.float32: F32, torch.bfloat16: BF16}


def dt(t: torch.Tensor) -> int:
    try:
        return _CODES[t.dtype]
    except KeyError:
        raise TypeError(f'autoprog_b200 kernels support float32 / bfloat16 tensors, got {t.dtype}') from None


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError('autoprog_b200 kernels need CUDA tensors (there is no CPU fallback)')
    if not t.is_contiguous():
        raise RuntimeError(f'autoprog_b200 kernels need contiguous tensors, got strides {t.stride()} for shape {tuple(t.shape)}')
    return t.data_ptr()


def _st() -> int:
    return torch.cuda.current_stream().cuda_stream


# ------------------------------------------------------------------ outlook attention core
def outlook_fwd(v: torch.Tensor, logits: torch.Tensor, heads: int, scale: float, simt: bool = False) -> torch.Tensor:
    """logits [B,h,w,lpitch] with heads*81 <= lpitch < heads*81+8 (tail = padding)."""
    B, H, W, Cc = v.shape
    assert Cc == heads * 32, 'OutlookAttention kernels are built for head_dim 32'
    lpitch = logits.shape[-1]
    assert logits.shape[:3] == (B, (H + 1) // 2, (W + 1) // 2) and logits.dtype == v.dtype
    y = torch.empty_like(v)
    fn = lib().apb_outlook_fwd_simt if simt else lib().apb_outlook_fwd
    check(fn(_p(v), _p(logits), _p(y), B, H, W, heads, scale, lpitch, dt(v), _st()), 'outlook_fwd')
    return y


def outlook_bwd(v, logits, dy, heads: int, scale: float, simt: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    B, H, W, _ = v.shape
    dv = torch.empty_like(v)
    dl = torch.empty_like(logits)
    fn = lib().apb_outlook_bwd_simt if simt else lib().apb_outlook_bwd
    check(fn(_p(v), _p(logits), _p(dy), _p(dv), _p(dl), B, H, W, heads, scale, logits.shape[-1], dt(v), _st()), 'outlook_bwd')
    return dv, dl


# ------------------------------------------------------------------ token-label CE
_TLCE_TICKETS = {}     # (device index, stream) -> zero int32 ticket of the single-launch loss kernel (re-armed by the kernel)


def _tlce_ticket(device: torch.device) -> torch.Tensor:
    key = (device.index, _st())
    t = _TLCE_TICKETS.get(key)
    if t is None:
        t = _TLCE_TICKETS[key] = torch.zeros(1, device=device, dtype=torch.int32)
    return t


def tlce_fwd_bwd(x_cls, x_aux, target, box_area: int, w_cls: float, w_dense: float, box_dev=None, single_launch: bool = True):
    B, N, Cc = x_aux.shape
    assert x_cls.shape == (B, Cc) and x_cls.dtype == x_aux.dtype
    assert target.dtype == torch.float32
    is3d = target.dim() == 3
    assert tuple(target.shape) == ((B, Cc, 2 + N) if is3d else (B, Cc)), f'target shape {tuple(target.shape)}'
    loss = torch.empty((), device=x_aux.device, dtype=torch.float32)
    d_cls = torch.empty_like(x_cls)
    d_aux = torch.empty_like(x_aux)
    ws = torch.empty(int(lib().apb_tlce_workspace_floats(B, N)), device=x_aux.device, dtype=torch.float32)
    ticket = _tlce_ticket(x_aux.device) if single_launch else None
    check(lib().apb_tlce_fwd_bwd(_p(x_cls), _p(x_aux), _p(target), int(is3d), B, N, Cc, int(box_area), _p(box_dev), w_cls, w_dense,
                                 _p(loss), _p(d_cls), _p(d_aux), _p(ws), _p(ticket), dt(x_aux), _st()), 'tlce_fwd_bwd')
    return loss, d_cls, d_aux


def scale_lazy(a: torch.Tensor, b: torch.Tensor, g: torch.Tensor, applied: torch.Tensor) -> None:
    """a, b *= g / applied in place (no memory traffic when the factor is 1); applied <- g.  See apb_scale_lazy."""
    gs = g.to(torch.float32).reshape(1).contiguous()
    check(lib().apb_scale_lazy(_p(a), a.numel(), _p(b), b.numel(), _p(gs), _p(applied), dt(a), _st()), 'scale_lazy')


def scale_by_scalar(x: torch.Tensor, scalar: torch.Tensor) -> torch.Tensor:
    out = torch.empty_like(x)
    s = scalar.to(torch.float32).reshape(1).contiguous()
    check(lib().apb_scale_by_scalar(_p(x), _p(out), x.numel(), _p(s), dt(x), _st()), 'scale_by_scalar')
    return out


# ------------------------------------------------------------------ layer norm (+ residual)
def ln_fwd(x, gamma, beta, eps: float, out_dtype, r=None, rs=None, rows_per_sample: int = 1, want_sum: bool = False,
           want_y: bool = True):
    """xs = x + rs[b]*r ; y = LN(xs).  Returns (xs or None, y or None, mean, rstd)."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    xs = torch.empty_like(x) if (want_sum or r is not None) else None
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype) if want_y else None
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    if r is not None:
        assert r.shape == x.shape and r.dtype == out_dtype
    check(lib().apb_ln_fwd(_p(x), _p(r), _p(rs), rows_per_sample, _p(gamma), _p(beta), _p(xs), _p(y), _p(mean), _p(rstd),
                           rows, Cc, eps, dt(x), _CODES[out_dtype], _st()), 'ln_fwd')
    return xs, y, mean, rstd


def ln_bwd(dy, xs, mean, rstd, gamma, dres=None, want_dr: bool = False, rs=None, rows_per_sample: int = 1, dg_out=None,
           db_out=None):
    """Returns (dxs, dr, dgamma, dbeta); dxs = dres + LN'(dy) (stream dtype of xs); dr = rs[b]*dxs in dy's dtype.
    dg_out / db_out: optional fp32 destinations (views of the flat gradient buffer) written in place."""
    Cc = xs.shape[-1]
    rows = xs.numel() // Cc
    dxs = torch.empty_like(xs)
    dg = dg_out if dg_out is not None else torch.empty(Cc, device=xs.device, dtype=torch.float32)
    db = db_out if db_out is not None else torch.empty(Cc, device=xs.device, dtype=torch.float32)
    ws = torch.empty(int(lib().apb_ln_bwd_workspace_floats(Cc)), device=xs.device, dtype=torch.float32)
    if dres is not None:
        assert dres.dtype == xs.dtype and dres.shape == xs.shape
    dr = torch.empty(xs.shape, device=xs.device, dtype=dy.dtype) if want_dr else None
    check(lib().apb_ln_bwd(_p(dy), _p(xs), _p(mean), _p(rstd), _p(gamma), _p(dres), _p(dxs), _p(dr), _p(rs), rows_per_sample,
                           _p(dg), _p(db), 0, _p(ws), rows, Cc, dt(xs), dt(dy), _st()), 'ln_bwd')
    if dg_out is not None:
        dg = dg.detach()           # fresh aliases so autograd can adopt them as `.grad` without cloning
    if db_out is not None:
        db = db.detach()
    return dxs, dr, dg, db


def colsum(a: torch.Tensor, Cc: Optional[int] = None, out=None) -> torch.Tensor:
    """fp32 column sums of a viewed as [rows, C] (bias gradients, batch reductions)."""
    Cc = Cc or a.shape[-1]
    rows = a.numel() // Cc
    if out is None:
        out = torch.empty(Cc, device=a.device, dtype=torch.float32)
    ws = torch.empty(max(1, int(lib().apb_colsum_workspace_floats(rows, Cc))), device=a.device, dtype=torch.float32)
    check(lib().apb_colsum(_p(a), rows, Cc, _p(out), 0, _p(ws), dt(a), _st()), 'colsum')
    return out


# ------------------------------------------------------------------ GEMM
EPI_NONE, EPI_GELU, EPI_DGELU, EPI_ACC = 0, 1, 2, 3
_FORCE_SIMT = False   # tests flip this to run the bf16 model on the CUDA-core GEMM


def gemm(a, b, M: int, N: int, K: int, trans_a: bool = False, trans_b: bool = False, bias=None, epilogue: int = EPI_NONE,
         aux=None, out=None, out_dtype=None, rowsum_out=None):
    """C[M,N] = epi(A(m,k) B(n,k) + bias).  bf16 inputs -> tcgen05 kernel, fp32 inputs -> CUDA-core fp32 kernel.

    rowsum_out (fp32 [M], tcgen05 path only -- check `gemm_uses_tc` first): also receives sum_k A(m,k), i.e. the bias
    gradient when this is the wgrad GEMM of a Linear (A = dY^T)."""
    assert a.dtype == b.dtype, (a.dtype, b.dtype)
    out_dtype = out_dtype or a.dtype
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    if epilogue == EPI_GELU and aux is None:
        aux = torch.empty((M, N), device=a.device, dtype=out_dtype)
    assert a.numel() == M * K and b.numel() == N * K, (a.shape, b.shape, M, N, K)
    use_tc = gemm_uses_tc(a, M, N, K)
    if not use_tc:
        assert rowsum_out is None, 'rowsum_out needs the tcgen05 path'
        check(lib().apb_gemm_simt(_p(a), _p(b), _p(out), _p(bias), _p(aux), M, N, K, int(trans_a), int(trans_b), epilogue,
                                  dt(a), _CODES[out_dtype], _st()), 'gemm_simt')
        return (out, aux) if epilogue == EPI_GELU else out
    if (_PAIR_GEMM and not trans_a and not trans_b and epilogue == EPI_NONE and rowsum_out is None and M >= 8192
            and K >= 384 and N % 192 == 0 and out_dtype in (torch.bfloat16, torch.float32)):
        # CTA-pair kernel (cta_group::2, half the B-tile traffic per SM): opt-in (APB_GEMM_PAIR=1).  Measured in CUDA graphs
        # (tools/gemm_shapes.py) it ties the 4-stage one-CTA kernel and loses to the 5-stage one (22.0 vs 20.1 us at
        # 25088 x 384 x 1152): the main loop is bound by bytes in flight per SM, not by operand traffic
        check(lib().apb_gemm_tc_pair(_p(a), _p(b), _p(out), _p(bias), M, N, K, _CODES[out_dtype], _st()), 'gemm_tc_pair')
        return out
    split = 1
    if out_dtype == torch.float32 and epilogue == EPI_NONE and bias is None:
        split = int(lib().apb_gemm_tc_suggest_split(M, N, K))
    slots = int(lib().apb_gemm_tc_rowsum_slots(N, split)) if rowsum_out is not None else 0
    rparts = torch.empty((slots, M), device=a.device, dtype=torch.float32) if slots > 1 else rowsum_out
    if split > 1:   # deterministic split-K: fp32 partial tiles, then a fixed-order sum over the split dim
        parts = torch.empty((split, M, N), device=a.device, dtype=torch.float32)
        check(lib().apb_gemm_tc_rowsum(_p(a), _p(b), _p(parts), None, None, M, N, K, int(trans_a), int(trans_b), 0, dt(a), F32,
                                       split, _p(rparts), _st()), 'gemm_tc(split-k)')
        check(lib().apb_splitk_reduce2(_p(parts), _p(out), M * N, split, _p(rparts), _p(rowsum_out), M if slots else 0, slots,
                                       _st()), 'gemm_tc(split-k reduce)')
        return out
    check(lib().apb_gemm_tc_rowsum(_p(a), _p(b), _p(out), _p(bias), _p(aux), M, N, K, int(trans_a), int(trans_b), epilogue, dt(a),
                                   _CODES[out_dtype], 1, _p(rparts), _st()), 'gemm_tc')
    if slots > 1:
        check(lib().apb_splitk_reduce(_p(rparts), _p(rowsum_out), slots, M, _st()), 'gemm_tc(row-sum reduce)')
    return (out, aux) if epilogue == EPI_GELU else out


_PAIR_GEMM = os.environ.get('APB_GEMM_PAIR', '0') == '1'   # off: the 5-stage one-CTA kernel is faster on every shape (profiles/r2_kernels.md)


def gemm_uses_tc(a: torch.Tensor, M: int, N: int, K: int) -> bool:
    return a.dtype == torch.bfloat16 and not _FORCE_SIMT and tc_supported(M, N, K)


def tc_supported(M: int, N: int, K: int) -> bool:
    """Shape envelope of the tcgen05 kernel: TMA needs 16-byte aligned row pitches (bf16: multiples of 8)."""
    return M % 8 == 0 and N % 8 == 0 and K % 8 == 0 and M >= 8 and N >= 8 and K >= 8


# ------------------------------------------------------------------ attention cores
def mhsa_fwd(qkv: torch.Tensor, heads: int, scale: float):
    B, N, C3 = qkv.shape
    D = C3 // (3 * heads)
    out = torch.empty((B, N, heads * D), device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty((B, heads, N), device=qkv.device, dtype=torch.float32)
    check(lib().apb_mhsa_fwd(_p(qkv), _p(out), _p(lse), B, N, heads, D, scale, dt(qkv), _st()), 'mhsa_fwd')
    return out, lse


def mhsa_bwd(qkv, out, dout, lse, heads: int, scale: float):
    B, N, C3 = qkv.shape
    D = C3 // (3 * heads)
    dqkv = torch.empty_like(qkv)
    ws = torch.empty(B * heads * N, device=qkv.device, dtype=torch.float32)
    check(lib().apb_mhsa_bwd(_p(qkv), _p(out), _p(dout), _p(lse), _p(dqkv), _p(ws), B, N, heads, D, scale, dt(qkv), _st()),
          'mhsa_bwd')
    return dqkv


def class_attn_fwd(q, kv, heads: int, scale: float):
    B, N, C2 = kv.shape
    D = C2 // (2 * heads)
    out = torch.empty_like(q)
    check(lib().apb_class_attn_fwd(_p(q), _p(kv), _p(out), B, N, heads, D, scale, dt(q), _st()), 'class_attn_fwd')
    return out


def class_attn_bwd(q, kv, dout, heads: int, scale: float):
    B, N, C2 = kv.shape
    D = C2 // (2 * heads)
    dq = torch.empty_like(q)
    dkv = torch.empty_like(kv)
    check(lib().apb_class_attn_bwd(_p(q), _p(kv), _p(dout), _p(dq), _p(dkv), B, N, heads, D, scale, dt(q), _st()),
          'class_attn_bwd')
    return dq, dkv


def class_attn_fwd_split(q, kv_cls, kv_tok, heads: int, scale: float):
    """Keys in two buffers: kv_cls [B, 2C] (the class token, key 0) and kv_tok [B, N-1, 2C] (the patch tokens)."""
    B, Nt, C2 = kv_tok.shape
    D = C2 // (2 * heads)
    out = torch.empty_like(q)
    check(lib().apb_class_attn_fwd_split(_p(q), _p(kv_cls), _p(kv_tok), _p(out), B, Nt + 1, heads, D, scale, dt(q), _st()),
          'class_attn_fwd_split')
    return out


def class_attn_bwd_split(q, kv_cls, kv_tok, dout, heads: int, scale: float):
    B, Nt, C2 = kv_tok.shape
    D = C2 // (2 * heads)
    dq = torch.empty_like(q)
    dkv_cls = torch.empty_like(kv_cls)
    dkv_tok = torch.empty_like(kv_tok)
    check(lib().apb_class_attn_bwd_split(_p(q), _p(kv_cls), _p(kv_tok), _p(dout), _p(dq), _p(dkv_cls), _p(dkv_tok), B, Nt + 1,
                                         heads, D, scale, dt(q), _st()), 'class_attn_bwd_split')
    return dq, dkv_cls, dkv_tok


# ------------------------------------------------------------------ elementwise / layout
def avgpool2_fwd(x):
    B, H, W, Cc = x.shape
    y = torch.empty((B, (H + 1) // 2, (W + 1) // 2, Cc), device=x.device, dtype=x.dtype)
    check(lib().apb_avgpool2_fwd(_p(x), _p(y), B, H, W, Cc, dt(x), _st()), 'avgpool2_fwd')
    return y


def avgpool2_bwd(dy, H: int, W: int, accumulate_into=None):
    B, _, _, Cc = dy.shape
    dx = accumulate_into if accumulate_into is not None else torch.empty((B, H, W, Cc), device=dy.device, dtype=dy.dtype)
    check(lib().apb_avgpool2_bwd(_p(dy), _p(dx), B, H, W, Cc, int(accumulate_into is not None), dt(dy), _st()), 'avgpool2_bwd')
    return dx


def flip_in_box(x, box: Sequence[int]):
    """x [B,H,W,C]; box = (r0, c0, r1, c1) on the (H, W) grid."""
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    r0, c0, r1, c1 = [int(b) for b in box]
    check(lib().apb_flip_in_box(_p(x), _p(y), B, H, W, Cc, r0, c0, r1, c1, dt(x), _st()), 'flip_in_box')
    return y


def flip_in_box_dev(x, box_dev, scale: int):
    """Same as flip_in_box with the box (r0,c0,r1,c1) read from a device int32[4] tensor and multiplied by `scale`."""
    B, H, W, Cc = x.shape
    y = torch.empty_like(x)
    assert box_dev.dtype == torch.int32 and box_dev.numel() == 4
    check(lib().apb_flip_in_box_dev(_p(x), _p(y), B, H, W, Cc, _p(box_dev), int(scale), dt(x), _st()), 'flip_in_box_dev')
    return y


def patchify(x, p: int):
    B, H, W, Cc = x.shape
    rows = torch.empty((B * (H // p) * (W // p), p * p * Cc), device=x.device, dtype=x.dtype)
    check(lib().apb_patchify(_p(x), _p(rows), B, H, W, Cc, p, dt(x), _st()), 'patchify')
    return rows


def unpatchify(rows, B: int, H: int, W: int, Cc: int, p: int):
    x = torch.empty((B, H, W, Cc), device=rows.device, dtype=rows.dtype)
    check(lib().apb_unpatchify(_p(rows), _p(x), B, H, W, Cc, p, dt(rows), _st()), 'unpatchify')
    return x


def im2col(x, KH: int, KW: int, stride: int, pad: int, Kpad: int):
    """x: [B, C, H, W] (any strides, fp32 / bf16) -> bf16 [B*OH*OW, Kpad], column k = c*KH*KW + ky*KW + kx."""
    B, Cc, H, W = x.shape
    OH, OW = (H + 2 * pad - KH) // stride + 1, (W + 2 * pad - KW) // stride + 1
    col = torch.empty((B * OH * OW, Kpad), device=x.device, dtype=torch.bfloat16)
    sb, sc, sh, sw = x.stride()
    if not x.is_cuda:
        raise RuntimeError('autoprog_b200 kernels need CUDA tensors (there is no CPU fallback)')
    check(lib().apb_im2col(x.data_ptr(), _p(col), B, Cc, H, W, KH, KW, stride, pad, Kpad, sb, sc, sh, sw, dt(x), _st()), 'im2col')
    return col, OH, OW


def bilinear_resize(x: torch.Tensor, OH: int, OW: int, out_dtype=None) -> torch.Tensor:
    """[B, C, H, W] fp32 -> [B, C, OH, OW] (fp32 or bf16): F.interpolate(mode='bilinear', align_corners=False)."""
    B, Cc, H, W = x.shape
    assert x.dtype == torch.float32, x.dtype
    out_dtype = out_dtype or x.dtype
    out = torch.empty((B, Cc, OH, OW), device=x.device, dtype=out_dtype)
    check(lib().apb_bilinear_resize(_p(x), _p(out), B * Cc, H, W, OH, OW, _CODES[out_dtype], _st()), 'bilinear_resize')
    return out


def bicubic_resize(src, h0: int, w0: int):
    h, w, Cc = src.shape
    dst = torch.empty((h0, w0, Cc), device=src.device, dtype=torch.float32)
    check(lib().apb_bicubic_resize(_p(src), _p(dst), h, w, h0, w0, Cc, _st()), 'bicubic_resize')
    return dst


def bicubic_resize_bwd(ddst, h: int, w: int):
    h0, w0, Cc = ddst.shape
    dsrc = torch.empty((h, w, Cc), device=ddst.device, dtype=torch.float32)
    check(lib().apb_bicubic_resize_bwd(_p(ddst), _p(dsrc), h, w, h0, w0, Cc, _st()), 'bicubic_resize_bwd')
    return dsrc


def add_bcast(x, p, out_dtype=None):
    """x [B, ...] + p [...] (fp32) broadcast over the batch dim; output dtype may differ from x's."""
    out_dtype = out_dtype or x.dtype
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    inner = p.numel()
    check(lib().apb_add_bcast(_p(x), _p(p), _p(out), x.numel() // inner, inner, dt(x), _CODES[out_dtype], _st()), 'add_bcast')
    return out


def scale_cast(x, out_dtype, rs=None):
    """out[b] = x[b] * rs[b] converted to out_dtype (rs None = plain cast)."""
    if rs is None and x.dtype == out_dtype:
        return x
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    batch = x.shape[0] if rs is not None else 1
    check(lib().apb_scale_cast(_p(x), _p(rs), _p(out), batch, x.numel() // batch, dt(x), _CODES[out_dtype], _st()), 'scale_cast')
    return out


def residual_add(x, r, rs=None, out_dtype=None):
    """out = x + rs[b] * r  (x: stream dtype, r: compute dtype)."""
    out_dtype = out_dtype or x.dtype
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    batch = x.shape[0] if rs is not None else 1
    check(lib().apb_residual_add(_p(x), _p(r), _p(rs), _p(out), batch, x.numel() // batch, dt(x), dt(r), _CODES[out_dtype],
                                 _st()), 'residual_add')
    return out


def cast(x, dtype):
    if x.dtype == dtype:
        return x
    out = torch.empty(x.shape, device=x.device, dtype=dtype)
    check(lib().apb_cast(_p(x), _p(out), x.numel(), dt(x), _CODES[dtype], _st()), 'cast')
    return out


def add(a, b):
    out = torch.empty_like(a)
    check(lib().apb_add(_p(a), _p(b), _p(out), a.numel(), dt(a), _st()), 'add')
    return out


def gelu_fwd(x):
    y = torch.empty_like(x)
    check(lib().apb_gelu_fwd(_p(x), _p(y), x.numel(), dt(x), _st()), 'gelu_fwd')
    return y


def gelu_bwd(x, dy):
    dx = torch.empty_like(x)
    check(lib().apb_gelu_bwd(_p(x), _p(dy), _p(dx), x.numel(), dt(x), _st()), 'gelu_bwd')
    return dx


def bn_supported(C: int) -> bool:
    return C % 8 == 0 and C <= 2048 and 256 % (C // 8) == 0


def bn_relu_fwd(x, gamma, beta, running_mean, running_var, momentum: float, eps: float, training: bool):
    """x [.., C] channels-last flattened.  Returns (y, mean, invstd)."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    y = torch.empty_like(x)
    if training:
        mean = torch.empty(Cc, device=x.device, dtype=torch.float32)
        invstd = torch.empty(Cc, device=x.device, dtype=torch.float32)
    else:
        mean = running_mean.detach().float().contiguous()
        invstd = torch.rsqrt(running_var.detach().float() + eps).contiguous()
    ws = torch.empty(int(lib().apb_bn_workspace_floats(rows, Cc)), device=x.device, dtype=torch.float32)
    check(lib().apb_bn_relu_fwd(_p(x), _p(y), _p(gamma), _p(beta), _p(mean), _p(invstd), _p(running_mean) if training else None,
                                _p(running_var) if training else None, momentum, eps, int(training), _p(ws), rows, Cc, dt(x),
                                _st()), 'bn_relu_fwd')
    return y, mean, invstd


def bn_relu_bwd(x, y, dy, gamma, mean, invstd, beta=None):
    """beta given: the ReLU mask is recomputed from x (y may be None and is not read); else it is read from y."""
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    dx = torch.empty_like(x)
    dg = torch.empty(Cc, device=x.device, dtype=torch.float32)
    db = torch.empty(Cc, device=x.device, dtype=torch.float32)
    ws = torch.empty(int(lib().apb_bn_workspace_floats(rows, Cc)), device=x.device, dtype=torch.float32)
    check(lib().apb_bn_relu_bwd(_p(x), _p(y), _p(dy), _p(gamma), _p(beta), _p(mean), _p(invstd), _p(dx), _p(dg), _p(db), _p(ws),
                                rows, Cc, dt(x), _st()), 'bn_relu_bwd')
    return dx, dg, db


def adamw_ema(p, g, m, v, hyper, beta1, beta2, eps, wd, emas=(), decays=(), shadow=None):
    """hyper: device fp32 tensor [lr, 1-beta1^t, sqrt(1-beta2^t)]."""
    n = p.numel()
    k = len(emas)
    ptrs = (C.c_void_p * max(k, 1))(*[e.data_ptr() for e in emas])
    dec = (C.c_float * max(k, 1))(*[float(d) for d in decays])
    check(lib().apb_adamw_ema(_p(p), _p(g), _p(m), _p(v), n, _p(hyper), beta1, beta2, eps, wd,
                              C.cast(ptrs, C.POINTER(C.c_void_p)), C.cast(dec, C.POINTER(C.c_float)), k, _p(shadow), _st()),
          'adamw_ema')


def launch_count() -> int:
    return int(lib().apb_launch_count())


def set_pdl(on: bool) -> None:
    """Programmatic dependent launch of the hot kernels (off by default; see include/autoprog_b200.h)."""
    lib().apb_set_pdl(int(bool(on)))


def get_pdl() -> bool:
    return bool(lib().apb_get_pdl())


def debug_gemm_switches(dbg: int = 0, five_stage: bool = True) -> None:
    """Diagnostics for tools/gemm_bound.py: bit 0 no TMA loads, bit 1 no MMAs, bit 2 no stores; five_stage=False times the
    four-stage 128 x 192 kernels."""
    lib().apb_debug_gemm_switches(int(dbg), int(bool(five_stage)))


def fallback_count() -> int:
    """bf16 calls that a tensor-core kernel declined and a CUDA-core kernel served (each one is also logged on stderr)."""
    return int(lib().apb_fallback_count())
