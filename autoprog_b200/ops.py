"""Autograd layer over the C-ABI kernels: per-op Functions plus fused per-block Functions.

dtype policy (the `prog/scaler.py` AMP path, re-targeted at bf16):
  * parameters, LayerNorm/softmax/loss statistics, optimizer state and the RESIDUAL STREAM are fp32;
  * under `torch.autocast('cuda', dtype=torch.bfloat16)` (or `autoprog_b200.autocast()`) every GEMM / attention /
    branch activation is bf16 with fp32 accumulation; without autocast everything is fp32 on the CUDA-core
    parity kernels (1e-5 mode).
Backward never consults autocast state: each Function records its compute dtype at forward time.
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch

from . import kernels as K

BF16, F32 = torch.bfloat16, torch.float32


def compute_dtype(x: Optional[torch.Tensor] = None) -> torch.dtype:
    if torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == BF16:
        return BF16
    if x is not None and x.dtype == BF16:
        return BF16
    return F32


class autocast(torch.autocast):
    """`with autoprog_b200.autocast():` == torch.autocast('cuda', dtype=torch.bfloat16) (drop-in for amp_autocast)."""

    def __init__(self, enabled: bool = True):
        super().__init__('cuda', dtype=BF16, enabled=enabled)


# ---------------------------------------------------------------------------------------------
# bf16 shadow copies of fp32 parameters, refreshed when the parameter is modified in place
# ---------------------------------------------------------------------------------------------
_WCACHE = {}


def wcast(p: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    src = p.detach()
    if not src.is_contiguous():
        src = src.contiguous()
    if src.dtype == dtype:
        return src
    key = id(p)
    ent = _WCACHE.get(key)
    if ent is not None and ent[1] == p.data_ptr() and ent[2].shape == p.shape and ent[2].dtype == dtype:
        if ent[0] == p._version:
            return ent[2]
        if ent[3]:
            # optimizer-owned shadow (the fused step keeps it fresh without bumping `_version`), but the parameter was
            # written in place since (load_state_dict, MoGrow loaders, resume): re-cast INTO the owner's buffer
            K.check(K.lib().apb_cast(src.data_ptr(), ent[2].data_ptr(), src.numel(), K.dt(src), K._CODES[dtype],
                                     torch.cuda.current_stream().cuda_stream), 'cast(shadow refresh)')
            _WCACHE[key] = (p._version, ent[1], ent[2], True)
            return ent[2]
    t = K.cast(src, dtype)
    if ent is None:
        weakref.finalize(p, _WCACHE.pop, key, None)
    _WCACHE[key] = (p._version, p.data_ptr(), t, False)
    return t


def register_shadow(p: torch.Tensor, shadow: torch.Tensor) -> None:
    """Let a fused optimizer publish the bf16 copy it wrote (skips the cast kernel on the next forward)."""
    if id(p) not in _WCACHE:
        weakref.finalize(p, _WCACHE.pop, id(p), None)
    _WCACHE[id(p)] = (p._version, p.data_ptr(), shadow, True)   # True: owner-refreshed; stale only after an in-place write


def invalidate_derived_caches() -> None:
    """Called by optimizers that update parameters through raw pointers (no autograd version bump): drops every cached
    re-laid-out / cast copy except the shadows the optimizer itself keeps fresh."""
    _CONVW.clear()
    _PADW.clear()
    for k in [k for k, v in _WCACHE.items() if not v[3]]:
        del _WCACHE[k]


def _c(t: torch.Tensor) -> torch.Tensor:
    return t if t.is_contiguous() else t.contiguous()


# ---------------------------------------------------------------------------------------------
# raw (non-autograd) linear helpers shared by the Functions
# ---------------------------------------------------------------------------------------------
def _lin_fwd(x2d, w, bias, epilogue=K.EPI_NONE):
    """x2d [M,K] (compute dtype), w [N,K] same dtype, bias fp32 or None."""
    M, Kd = x2d.shape
    return K.gemm(x2d, w, M, w.shape[0], Kd, bias=bias, epilogue=epilogue)


def grad_dest(p):
    """Flat-gradient view a kernel may write this parameter's gradient into (see flat.FlatGroup): only when the
    parameter has no `.grad` yet -- i.e. not while gradients are being accumulated over several backward passes -- and
    only ONCE per backward: a parameter used twice in one forward (e.g. ClassAttention.kv on the class token and on the
    patch tokens) gets its second contribution as a fresh tensor that autograd adds to the first (`.grad` is still None
    while the engine collects both, so without the claim the second kernel would overwrite the first one's output)."""
    if p is None or p.grad is not None:
        return None
    v = getattr(p, '_apb_grad_view', None)
    if v is None or getattr(p, '_apb_grad_claimed', False):
        return None
    p._apb_grad_claimed = True
    return v


# The weight-gradient GEMM (+ its split-K reduce) of a Linear does not depend on the input-gradient GEMM: with
# SIDE_WGRAD they are enqueued on two streams (fork on dY, join right after both are enqueued), so the CTAs of one fill
# the SMs the other's last wave leaves idle and one launch gap per pair disappears.  The join precedes every later
# kernel, so no tensor outlives its producer across streams; under CUDA-graph capture the fork becomes two branches.
SIDE_WGRAD = os.environ.get('APB_SIDE_WGRAD', '1') == '1'
_side_streams = {}


def _side_stream(device):
    s = _side_streams.get(device)
    if s is None:
        s = _side_streams[device] = torch.cuda.Stream(device=device)
    return s


def _lin_bwd(dy2d, x2d, w, need_dx=True, need_dw=True, need_db=True, dgelu_aux=None, dw_out=None, db_out=None):
    """Returns (dx [M,K] compute dtype, dw [N,K] fp32, db [N] fp32).  dw_out / db_out: in-place destinations."""
    if SIDE_WGRAD and need_dx and need_dw and dy2d.is_cuda and dy2d.dtype == BF16:
        main = torch.cuda.current_stream(dy2d.device)
        side = _side_stream(dy2d.device)
        side.wait_stream(main)                       # dY (and everything before it) is ready
        dx, _, _ = _lin_bwd_impl(dy2d, x2d, w, True, False, False, dgelu_aux, None, None)
        with torch.cuda.stream(side):
            _, dw, db = _lin_bwd_impl(dy2d, x2d, w, False, True, need_db, None, dw_out, db_out)
        main.wait_stream(side)
        return dx, dw, db
    return _lin_bwd_impl(dy2d, x2d, w, need_dx, need_dw, need_db, dgelu_aux, dw_out, db_out)


def _lin_bwd_impl(dy2d, x2d, w, need_dx, need_dw, need_db, dgelu_aux, dw_out, db_out):
    M, N = dy2d.shape
    Kd = w.shape[1]
    dx = dw = db = None
    if need_dx:
        if dgelu_aux is not None:
            dx = K.gemm(dy2d, w, M, Kd, N, trans_b=True, epilogue=K.EPI_DGELU, aux=dgelu_aux)
        else:
            dx = K.gemm(dy2d, w, M, Kd, N, trans_b=True)
    if need_dw:
        if dw_out is not None and tuple(dw_out.shape) != (N, Kd):
            dw_out = None
        if need_db and db_out is not None and db_out.numel() != N:
            db_out = None
        # bf16: the bias gradient (column sums of dY) is accumulated by the wgrad GEMM itself on the tensor pipe
        fused_db = need_db and K.gemm_uses_tc(dy2d, N, Kd, M)
        if fused_db:
            db = db_out if db_out is not None else torch.empty(N, device=dy2d.device, dtype=F32)
        dw = K.gemm(dy2d, x2d, N, Kd, M, trans_a=True, trans_b=True, out_dtype=F32, out=dw_out,
                    rowsum_out=db if fused_db else None)
        if dw_out is not None:
            dw = dw.detach()       # fresh alias: autograd adopts it as `.grad` without cloning (sole reference)
        if fused_db:
            if db_out is not None:
                db = db.detach()
            need_db = False
    if need_db:
        if db_out is not None and db_out.numel() != N:
            db_out = None
        db = K.colsum(dy2d, N, out=db_out)
        if db_out is not None:
            db = db.detach()
    return dx, dw, db


# ---------------------------------------------------------------------------------------------
# per-op Functions
# ---------------------------------------------------------------------------------------------
class LinearFn(torch.autograd.Function):
    """y = x W^T + b (nn.Linear, e.g. models/volo.py:67-71,156-158,180-182,253-256,547-554)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        cdt = compute_dtype(x)
        xc = _c(K.cast(_c(x), cdt)).reshape(-1, x.shape[-1])
        w = wcast(weight, cdt)
        y = _lin_fwd(xc, w, None if bias is None else bias.detach())
        ctx.save_for_backward(xc, w)
        ctx.has_bias = bias is not None
        ctx.x_dtype = x.dtype
        ctx.x_shape = x.shape
        ctx.params = (weight, bias)
        return y.reshape(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xc, w = ctx.saved_tensors
        dy2 = _c(K.cast(_c(dy), xc.dtype)).reshape(-1, w.shape[0])
        dx, dw, db = _lin_bwd(dy2, xc, w, ctx.needs_input_grad[0], ctx.needs_input_grad[1],
                              ctx.has_bias and ctx.needs_input_grad[2], dw_out=grad_dest(ctx.params[0]),
                              db_out=grad_dest(ctx.params[1]))
        if dx is not None:
            dx = K.cast(dx, ctx.x_dtype).reshape(ctx.x_shape)
        return dx, dw, db


class LayerNormFn(torch.autograd.Function):
    """y = LN(x) in the compute dtype; x may be the fp32 stream (nn.LayerNorm, models/volo.py:472)."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        cdt = compute_dtype(x)
        xc = _c(x)
        if xc.dtype == BF16 and cdt == F32:
            xc = K.cast(xc, F32)
        _, y, mean, rstd = K.ln_fwd(xc, weight.detach(), bias.detach(), eps, cdt)
        ctx.save_for_backward(xc, mean, rstd, weight.detach())
        ctx.x_dtype = x.dtype
        ctx.cdt = cdt
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, mean, rstd, w = ctx.saved_tensors
        dyc = K.cast(_c(dy), ctx.cdt)
        dxs, _, dg, db = K.ln_bwd(dyc, xc, mean, rstd, w)
        return K.cast(dxs, ctx.x_dtype), dg, db, None


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        xc = _c(x)
        ctx.save_for_backward(xc)
        return K.gelu_fwd(xc)

    @staticmethod
    def backward(ctx, dy):
        (xc,) = ctx.saved_tensors
        return K.gelu_bwd(xc, K.cast(_c(dy), xc.dtype))


class OutlookCoreFn(torch.autograd.Function):
    """unfold -> softmax -> attn@v -> fold as one kernel per direction (models/volo.py:83-98)."""

    @staticmethod
    def forward(ctx, v, logits, heads, scale, simt=False):
        v, logits = _c(v), _c(logits)
        ctx.save_for_backward(v, logits)
        ctx.cfg = (heads, scale, simt)
        return K.outlook_fwd(v, logits, heads, scale, simt)

    @staticmethod
    def backward(ctx, dy):
        v, logits = ctx.saved_tensors
        heads, scale, simt = ctx.cfg
        dv, dl = K.outlook_bwd(v, logits, K.cast(_c(dy), v.dtype), heads, scale, simt)
        return dv, dl, None, None, None


class MhsaCoreFn(torch.autograd.Function):
    """softmax(q k^T scale) v on packed qkv [B,N,3C] (models/volo.py:188-197)."""

    @staticmethod
    def forward(ctx, qkv, heads, scale):
        qkv = _c(qkv)
        out, lse = K.mhsa_fwd(qkv, heads, scale)
        ctx.save_for_backward(qkv, out, lse)
        ctx.cfg = (heads, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        heads, scale = ctx.cfg
        return K.mhsa_bwd(qkv, out, K.cast(_c(dout), qkv.dtype), lse, heads, scale), None, None


class ClassAttnCoreFn(torch.autograd.Function):
    """cls-query attention (models/volo.py:264-275): q [B,C], kv [B,N,2C] -> [B,C]."""

    @staticmethod
    def forward(ctx, q, kv, heads, scale):
        q, kv = _c(q), _c(kv)
        ctx.save_for_backward(q, kv)
        ctx.cfg = (heads, scale)
        return K.class_attn_fwd(q, kv, heads, scale)

    @staticmethod
    def backward(ctx, dout):
        q, kv = ctx.saved_tensors
        heads, scale = ctx.cfg
        dq, dkv = K.class_attn_bwd(q, kv, K.cast(_c(dout), q.dtype), heads, scale)
        return dq, dkv, None, None


class ClassAttnCoreSplitFn(torch.autograd.Function):
    """cls-query attention with the keys kept in two buffers: q [B,C], kv_cls [B,2C] (key 0), kv_tok [B,N-1,2C] -> [B,C]."""

    @staticmethod
    def forward(ctx, q, kv_cls, kv_tok, heads, scale):
        q, kv_cls, kv_tok = _c(q), _c(kv_cls), _c(kv_tok)
        ctx.save_for_backward(q, kv_cls, kv_tok)
        ctx.cfg = (heads, scale)
        return K.class_attn_fwd_split(q, kv_cls, kv_tok, heads, scale)

    @staticmethod
    def backward(ctx, dout):
        q, kv_cls, kv_tok = ctx.saved_tensors
        heads, scale = ctx.cfg
        dq, dkc, dkt = K.class_attn_bwd_split(q, kv_cls, kv_tok, K.cast(_c(dout), q.dtype), heads, scale)
        return dq, dkc, dkt, None, None


class AvgPool2Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.hw = (x.shape[1], x.shape[2])
        return K.avgpool2_fwd(_c(x))

    @staticmethod
    def backward(ctx, dy):
        return K.avgpool2_bwd(_c(dy), *ctx.hw)


class FlipInBoxFn(torch.autograd.Function):
    """mix-token / un-mix (models/volo.py:655-658, 687-689); self-inverse permutation."""

    @staticmethod
    def forward(ctx, x, box):
        ctx.box = tuple(int(b) for b in box)
        return K.flip_in_box(_c(x), ctx.box)

    @staticmethod
    def backward(ctx, dy):
        return K.flip_in_box(_c(dy), ctx.box), None


class DevBox(tuple):
    """(bbx1, bby1, bbx2, bby2) whose live values sit in a device int32[4] tensor (`.dev`); the tuple entries are the
    values seen when the step was captured into a CUDA graph.  Returned by VOLO.forward in graph mode."""

    def __new__(cls, values, dev):
        obj = super().__new__(cls, values)
        obj.dev = dev
        return obj


class FlipInBoxDevFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, box_dev, scale):
        ctx.box, ctx.scale = box_dev, scale
        return K.flip_in_box_dev(_c(x), box_dev, scale)

    @staticmethod
    def backward(ctx, dy):
        return K.flip_in_box_dev(_c(dy), ctx.box, ctx.scale), None, None


class PatchConvFn(torch.autograd.Function):
    """Conv2d(kernel=p, stride=p) on an NHWC tensor as patchify + GEMM (models/volo.py:370-373, 389)."""

    @staticmethod
    def forward(ctx, x, weight, bias, p):
        cdt = compute_dtype(x)
        B, H, W, Cin = x.shape
        xc = K.cast(_c(x), cdt)
        rows = K.patchify(xc, p)
        # conv weight [Cout, Cin, p, p] -> GEMM weight [Cout, (kh, kw, cin)]
        w2 = wcast_conv(weight, cdt)
        y = _lin_fwd(rows, w2, None if bias is None else bias.detach())
        ctx.save_for_backward(rows, w2)
        ctx.meta = (B, H, W, Cin, p, x.dtype, weight.shape, bias is not None)
        return y.reshape(B, H // p, W // p, weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        rows, w2 = ctx.saved_tensors
        B, H, W, Cin, p, xdt, wshape, has_bias = ctx.meta
        dy2 = K.cast(_c(dy), rows.dtype).reshape(-1, wshape[0])
        drows, dw2, db = _lin_bwd(dy2, rows, w2, ctx.needs_input_grad[0], True, has_bias)
        dx = None
        if drows is not None:
            dx = K.cast(K.unpatchify(drows, B, H, W, Cin, p), xdt)
        dw = dw2.reshape(wshape[0], p, p, Cin).permute(0, 3, 1, 2).contiguous()
        return dx, dw, db, None


_CONVW = {}


def wcast_conv(weight: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    """[Cout,Cin,p,p] conv weight -> cached [Cout, p*p*Cin] GEMM weight in K order (kh,kw,cin)."""
    key = id(weight)
    ent = _CONVW.get(key)
    if ent is not None and ent[0] == weight._version and ent[1] == weight.data_ptr() and ent[3] == dtype:
        return ent[2]
    co = weight.shape[0]
    w2 = K.cast(weight.detach().permute(0, 2, 3, 1).reshape(co, -1).contiguous(), dtype)
    if ent is None:
        weakref.finalize(weight, _CONVW.pop, key, None)
    _CONVW[key] = (weight._version, weight.data_ptr(), w2, dtype)
    return w2


def wcast_conv_kpad(weight: torch.Tensor, Kpad: int) -> torch.Tensor:
    """[Cout,Cin,KH,KW] conv weight -> cached bf16 [Cout, Kpad] GEMM weight, K order (cin,kh,kw), zero-padded columns."""
    key = (id(weight), Kpad)
    ent = _CONVW.get(key)
    if ent is not None and ent[0] == weight._version and ent[1] == weight.data_ptr():
        return ent[2]
    co = weight.shape[0]
    w2 = torch.zeros((co, Kpad), device=weight.device, dtype=BF16)
    w2[:, :weight[0].numel()] = weight.detach().reshape(co, -1)
    if ent is None:
        weakref.finalize(weight, _CONVW.pop, key, None)
    _CONVW[key] = (weight._version, weight.data_ptr(), w2, BF16)
    return w2


class StemConvFn(torch.autograd.Function):
    """The 7x7 stride-2 stem convolution (Cin = 3; models/volo.py:352-353) in bf16 mode: im2col kernel + one tcgen05
    GEMM; weight gradient = one split-K GEMM over the saved im2col rows.  NCHW (any strides) in, NHWC bf16 out.
    The image itself gets no gradient through this path (callers that need d/dx use the library convolution)."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad):
        co, ci, kh, kw = weight.shape
        kpad = (ci * kh * kw + 7) // 8 * 8
        col, OH, OW = K.im2col(x, kh, kw, stride, pad, kpad)
        w2 = wcast_conv_kpad(weight, kpad)
        y = K.gemm(col, w2, col.shape[0], co, kpad)
        ctx.save_for_backward(col)
        ctx.meta = (tuple(weight.shape), kpad)
        return y.reshape(x.shape[0], OH, OW, co)

    @staticmethod
    def backward(ctx, dy):
        (col,) = ctx.saved_tensors
        wshape, kpad = ctx.meta
        co = wshape[0]
        dy2 = K.cast(_c(dy), BF16).reshape(-1, co)
        dw2 = K.gemm(dy2, col, co, kpad, col.shape[0], trans_a=True, trans_b=True, out_dtype=F32)
        dw = dw2[:, :wshape[1] * wshape[2] * wshape[3]].reshape(wshape)
        return None, dw, None, None


class ParityConvFn(torch.autograd.Function):
    """fp32 PARITY mode stem convolution (models/volo.py:355-368; library convolution, TF32 off): the weight gradient is
    a sum over B*H*W positions of products that almost cancel behind a train-mode BatchNorm (condition number ~1e4 at
    224 px: torch's own fp32 result is off by 1e-3 against exact arithmetic, cuDNN's serial fp32 accumulation by 2e-3), so
    parity mode accumulates it in fp64.  Data gradient and forward stay fp32.  Not used by the bf16 training path."""

    @staticmethod
    def forward(ctx, x, weight, stride, padding):
        ctx.save_for_backward(x, weight)
        ctx.cfg = (stride, padding)
        return torch.nn.functional.conv2d(x, weight, None, stride, padding)

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        stride, padding = ctx.cfg
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.nn.grad.conv2d_input(x.shape, weight, dy, stride, padding)
        if ctx.needs_input_grad[1]:
            dw = torch.nn.grad.conv2d_weight(x.double(), weight.shape, dy.double(), stride, padding).to(weight.dtype)
        return dx, dw, None, None


class PosEmbedAddFn(torch.autograd.Function):
    """x + bicubic_resize(pos_embed) (models/volo.py:580-596, 627-628); output joins the fp32 residual stream."""

    @staticmethod
    def forward(ctx, x, pos):
        B, h0, w0, Cc = x.shape
        _, h, w, _ = pos.shape
        p = pos.detach().reshape(h, w, Cc)
        p = _c(p) if (h == h0 and w == w0) else K.bicubic_resize(_c(p), h0, w0)
        ctx.meta = (h, w, h0, w0, Cc, x.dtype)
        out_dtype = F32 if x.dtype in (F32, BF16) else x.dtype
        return K.add_bcast(_c(x), p, out_dtype=out_dtype)

    @staticmethod
    def backward(ctx, dout):
        h, w, h0, w0, Cc, xdt = ctx.meta
        d = _c(dout)
        dp = K.colsum(d, h0 * w0 * Cc).reshape(h0, w0, Cc)
        if not (h == h0 and w == w0):
            dp = K.bicubic_resize_bwd(dp, h, w)
        return K.cast(d, xdt), dp.reshape(1, h, w, Cc)


class CastFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, dtype):
        ctx.src = x.dtype
        return K.cast(_c(x), dtype)

    @staticmethod
    def backward(ctx, dy):
        return K.cast(_c(dy), ctx.src), None


class BNReLUFn(torch.autograd.Function):
    """BatchNorm2d (train: batch statistics + running-stat update; eval: running stats) + ReLU on an NHWC tensor
    (PatchEmbed stem, models/volo.py:358-368)."""

    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, momentum, eps, training):
        xc = _c(x)
        y, mean, invstd = K.bn_relu_fwd(xc, weight.detach().float(), bias.detach().float(), running_mean, running_var,
                                        momentum, eps, training)
        # backward recomputes the ReLU mask from x (kernel reads x and dy only): y is not kept for it
        ctx.save_for_backward(xc, weight.detach().float(), bias.detach().float(), mean, invstd)
        ctx.training = training
        return y

    @staticmethod
    def backward(ctx, dy):
        xc, w, b, mean, invstd = ctx.saved_tensors
        if not ctx.training:
            raise RuntimeError('BNReLUFn: backward through eval-mode BatchNorm is not supported')
        dx, dg, db = K.bn_relu_bwd(xc, None, K.cast(_c(dy), xc.dtype), w, mean, invstd, beta=b)
        return dx, dg, db, None, None, None, None, None


class ResidualAddFn(torch.autograd.Function):
    """out = x + rs[b] * r : x fp32 stream, r compute dtype (DropPath scale rs optional, not differentiated)."""

    @staticmethod
    def forward(ctx, x, r, rs, out_dtype):
        ctx.meta = (x.dtype, r.dtype)
        ctx.rs = rs
        return K.residual_add(_c(x), _c(r), rs, out_dtype)

    @staticmethod
    def backward(ctx, dout):
        xdt, rdt = ctx.meta
        d = _c(dout)
        return K.scale_cast(d, xdt), K.scale_cast(d, rdt, ctx.rs), None, None


class TokenLabelCEFn(torch.autograd.Function):
    """Fused TokenLabelCrossEntropy forward+gradient (loss/cross_entropy.py:136-156)."""

    @staticmethod
    def forward(ctx, x_cls, x_aux, target, box_area, w_cls, w_dense, box_dev=None):
        loss, d_cls, d_aux = K.tlce_fwd_bwd(_c(x_cls), _c(x_aux), _c(target), box_area, w_cls, w_dense, box_dev)
        applied = torch.ones(1, device=loss.device, dtype=F32)     # upstream-gradient factor already folded into d_*
        ctx.save_for_backward(d_cls, d_aux, applied)
        return loss

    @staticmethod
    def backward(ctx, g):
        # the kernel produced the gradients for upstream gradient 1 -- the case of every training step (the loss is the
        # root of backward): the lazy scale is then a no-op launch instead of a read + write pass over [B, N, C]
        d_cls, d_aux, applied = ctx.saved_tensors
        K.scale_lazy(d_cls, d_aux, g, applied)
        return d_cls, d_aux, None, None, None, None, None


# ---------------------------------------------------------------------------------------------
# fused block Functions: one autograd node per Outlooker / Transformer block.
#
# The residual stream is carried as (x, r, rs): x fp32, r = the previous block's MLP output still pending
# (compute dtype), rs = its per-sample DropPath scale.  A block's first LayerNorm kernel performs
# xs = x + rs*r, so no residual add ever runs as its own pass.
# ---------------------------------------------------------------------------------------------
def _flat(t):
    return t.reshape(-1, t.shape[-1])


class _BlockBase(torch.autograd.Function):
    @staticmethod
    def _mlp_fwd(n2, w1, b1, w2, b2):
        u_h = K.gemm(n2, w1, n2.shape[0], w1.shape[0], n2.shape[1], bias=b1, epilogue=K.EPI_GELU)
        hdn, u = u_h
        z = K.gemm(hdn, w2, hdn.shape[0], w2.shape[0], hdn.shape[1], bias=b2)
        return u, hdn, z

    @staticmethod
    def _mlp_bwd(dz, n2, u, hdn, w1, w2, P):
        du, dw2, db2 = _lin_bwd(dz, hdn, w2, dgelu_aux=u, dw_out=grad_dest(P['w2']), db_out=grad_dest(P['b2']))
        dn2, dw1, db1 = _lin_bwd(du, n2, w1, dw_out=grad_dest(P['w1']), db_out=grad_dest(P['b1']))
        return dn2, dw1, db1, dw2, db2


_PADW = {}


def wcast_pad_rows(p: torch.Tensor, dtype: torch.dtype, rows: int) -> torch.Tensor:
    """cast copy of a [N, K] weight (or [N] bias) zero-padded to `rows` rows; refreshed when the parameter changes."""
    if rows == p.shape[0]:
        return wcast(p, dtype) if p.dim() == 2 else p.detach()
    key = (id(p), rows, dtype)
    ent = _PADW.get(key)
    if ent is not None and ent[0] == p._version and ent[1] == p.data_ptr():
        return ent[2]
    t = torch.zeros((rows,) + tuple(p.shape[1:]), device=p.device, dtype=dtype)
    K.check(K.lib().apb_cast(p.detach().contiguous().data_ptr(), t.data_ptr(), p.numel(), K.dt(p), K._CODES[dtype],
                             torch.cuda.current_stream().cuda_stream), 'cast(pad)')
    if ent is None:
        weakref.finalize(p, _PADW.pop, key, None)
    _PADW[key] = (p._version, p.data_ptr(), t)
    return t


class OutlookerFn(_BlockBase):
    """Outlooker.forward (models/volo.py:140-144) with OutlookAttention.forward (:77-103) and Mlp.forward (:161-167)."""

    @staticmethod
    def forward(ctx, x, r_in, rs_in, rs_blk, heads, eps, n1w, n1b, wv, wa, ba, wp, bp, n2w, n2b, w1, b1, w2, b2):
        cdt = compute_dtype()
        B, H, W, Cc = x.shape
        rps = H * W
        scale = (Cc // heads) ** -0.5
        xs, n1, mu1, rstd1 = K.ln_fwd(_c(x), n1w.detach(), n1b.detach(), eps, cdt, r=None if r_in is None else _c(r_in),
                                      rs=rs_in, rows_per_sample=rps)
        xs = xs if xs is not None else _c(x)
        cv, cp, c1, c2 = (wcast(t, cdt) for t in (wv, wp, w1, w2))
        # bf16: pad the 81*heads logit columns to a multiple of 8 so the GEMMs around the core stay on the TMA path
        nl = wa.shape[0]
        nlp = (nl + 7) // 8 * 8 if cdt == BF16 else nl
        ca = wcast_pad_rows(wa, cdt, nlp)
        ba_p = wcast_pad_rows(ba, F32, nlp)
        n1f = _flat(n1)
        v = _lin_fwd(n1f, cv, None).reshape(B, H, W, Cc)
        pooled = K.avgpool2_fwd(n1)
        lg = _lin_fwd(_flat(pooled), ca, ba_p).reshape(B, pooled.shape[1], pooled.shape[2], -1)
        y = K.outlook_fwd(v, lg, heads, scale)
        o = _lin_fwd(_flat(y), cp, bp.detach()).reshape(B, H, W, Cc)
        x1, n2, mu2, rstd2 = K.ln_fwd(xs, n2w.detach(), n2b.detach(), eps, cdt, r=o, rs=rs_blk, rows_per_sample=rps)
        n2f = _flat(n2)
        u, hdn, z = _BlockBase._mlp_fwd(n2f, c1, b1.detach(), c2, b2.detach())
        ctx.save_for_backward(xs, mu1, rstd1, n1, v, pooled, lg, y, x1, mu2, rstd2, n2, u, hdn, cv, ca, cp, c1, c2,
                              n1w.detach(), n2w.detach(), rs_in if rs_in is not None else x.new_empty(0),
                              rs_blk if rs_blk is not None else x.new_empty(0))
        ctx.meta = (heads, scale, r_in is not None, rs_in is not None, rs_blk is not None, nl)
        ctx.P = dict(n1w=n1w, n1b=n1b, wv=wv, wp=wp, bp=bp, n2w=n2w, n2b=n2b, w1=w1, b1=b1, w2=w2, b2=b2)
        return x1, z.reshape(B, H, W, Cc)

    @staticmethod
    def backward(ctx, dx1_out, dz):
        (xs, mu1, rstd1, n1, v, pooled, lg, y, x1, mu2, rstd2, n2, u, hdn, cv, ca, cp, c1, c2, n1w, n2w, rs_in,
         rs_blk) = ctx.saved_tensors
        heads, scale, has_r, has_rs_in, has_rs_blk, nl = ctx.meta
        B, H, W, Cc = xs.shape
        rps = H * W
        cdt = n1.dtype
        P = ctx.P
        dz2 = _flat(K.cast(_c(dz), cdt))
        dn2, dw1, db1, dw2, db2 = _BlockBase._mlp_bwd(dz2, _flat(n2), u, hdn, c1, c2, P)
        dx1, do, dn2w, dn2b = K.ln_bwd(dn2.reshape(xs.shape), x1, mu2, rstd2, n2w, dres=_c(dx1_out), want_dr=True,
                                       rs=rs_blk if has_rs_blk else None, rows_per_sample=rps,
                                       dg_out=grad_dest(P['n2w']), db_out=grad_dest(P['n2b']))
        do2 = _flat(do)
        dy, dwp, dbp = _lin_bwd(do2, _flat(y), cp, dw_out=grad_dest(P['wp']), db_out=grad_dest(P['bp']))
        dv, dlg = K.outlook_bwd(v, lg, dy.reshape(xs.shape), heads, scale)
        dn1, dwv, _ = _lin_bwd(_flat(dv), _flat(n1), cv, need_db=False, dw_out=grad_dest(P['wv']))
        dpooled, dwa, dba = _lin_bwd(_flat(dlg), _flat(pooled), ca)
        dwa, dba = dwa[:nl], dba[:nl]                      # drop the padding rows
        dn1 = K.avgpool2_bwd(dpooled.reshape(pooled.shape), H, W, accumulate_into=dn1.reshape(xs.shape))
        dxs, dr, dn1w, dn1b = K.ln_bwd(dn1, xs, mu1, rstd1, n1w, dres=dx1, want_dr=has_r,
                                       rs=rs_in if has_rs_in else None, rows_per_sample=rps,
                                       dg_out=grad_dest(P['n1w']), db_out=grad_dest(P['n1b']))
        return (dxs, dr, None, None, None, None, dn1w, dn1b, dwv, dwa, dba, dwp, dbp, dn2w, dn2b, dw1, db1, dw2, db2)


class TransformerFn(_BlockBase):
    """Transformer.forward (models/volo.py:230-234) with Attention.forward (:185-201) and Mlp.forward (:161-167)."""

    @staticmethod
    def forward(ctx, x, r_in, rs_in, rs_blk, heads, eps, n1w, n1b, wqkv, bqkv, wp, bp, n2w, n2b, w1, b1, w2, b2):
        cdt = compute_dtype()
        shp = x.shape
        Cc = shp[-1]
        B = shp[0]
        N = x.numel() // (B * Cc)
        scale = (Cc // heads) ** -0.5
        xs, n1, mu1, rstd1 = K.ln_fwd(_c(x), n1w.detach(), n1b.detach(), eps, cdt, r=None if r_in is None else _c(r_in),
                                      rs=rs_in, rows_per_sample=N)
        xs = xs if xs is not None else _c(x)
        cqkv, cp, c1, c2 = (wcast(t, cdt) for t in (wqkv, wp, w1, w2))
        qkv = _lin_fwd(_flat(n1), cqkv, None if bqkv is None else bqkv.detach()).reshape(B, N, 3 * Cc)
        a, lse = K.mhsa_fwd(qkv, heads, scale)
        o = _lin_fwd(_flat(a), cp, bp.detach()).reshape(shp)
        x1, n2, mu2, rstd2 = K.ln_fwd(xs, n2w.detach(), n2b.detach(), eps, cdt, r=o, rs=rs_blk, rows_per_sample=N)
        u, hdn, z = _BlockBase._mlp_fwd(_flat(n2), c1, b1.detach(), c2, b2.detach())
        ctx.save_for_backward(xs, mu1, rstd1, n1, qkv, a, lse, x1, mu2, rstd2, n2, u, hdn, cqkv, cp, c1, c2,
                              n1w.detach(), n2w.detach(), rs_in if rs_in is not None else x.new_empty(0),
                              rs_blk if rs_blk is not None else x.new_empty(0))
        ctx.meta = (heads, scale, r_in is not None, rs_in is not None, rs_blk is not None, bqkv is not None, N)
        ctx.P = dict(n1w=n1w, n1b=n1b, wqkv=wqkv, bqkv=bqkv, wp=wp, bp=bp, n2w=n2w, n2b=n2b, w1=w1, b1=b1, w2=w2, b2=b2)
        return x1, z.reshape(shp)

    @staticmethod
    def backward(ctx, dx1_out, dz):
        (xs, mu1, rstd1, n1, qkv, a, lse, x1, mu2, rstd2, n2, u, hdn, cqkv, cp, c1, c2, n1w, n2w, rs_in,
         rs_blk) = ctx.saved_tensors
        heads, scale, has_r, has_rs_in, has_rs_blk, has_bqkv, N = ctx.meta
        cdt = n1.dtype
        P = ctx.P
        dz2 = _flat(K.cast(_c(dz), cdt))
        dn2, dw1, db1, dw2, db2 = _BlockBase._mlp_bwd(dz2, _flat(n2), u, hdn, c1, c2, P)
        dx1, do, dn2w, dn2b = K.ln_bwd(dn2.reshape(xs.shape), x1, mu2, rstd2, n2w, dres=_c(dx1_out), want_dr=True,
                                       rs=rs_blk if has_rs_blk else None, rows_per_sample=N,
                                       dg_out=grad_dest(P['n2w']), db_out=grad_dest(P['n2b']))
        da, dwp, dbp = _lin_bwd(_flat(do), _flat(a), cp, dw_out=grad_dest(P['wp']), db_out=grad_dest(P['bp']))
        dqkv = K.mhsa_bwd(qkv, a, da.reshape(a.shape), lse, heads, scale)
        dn1, dwqkv, dbqkv = _lin_bwd(_flat(dqkv), _flat(n1), cqkv, need_db=has_bqkv, dw_out=grad_dest(P['wqkv']),
                                     db_out=grad_dest(P['bqkv']))
        dxs, dr, dn1w, dn1b = K.ln_bwd(dn1.reshape(xs.shape), xs, mu1, rstd1, n1w, dres=dx1, want_dr=has_r,
                                       rs=rs_in if has_rs_in else None, rows_per_sample=N,
                                       dg_out=grad_dest(P['n1w']), db_out=grad_dest(P['n1b']))
        return (dxs, dr, None, None, None, None, dn1w, dn1b, dwqkv, dbqkv, dwp, dbp, dn2w, dn2b, dw1, db1, dw2, db2)
