"""ctypes loader for the C-ABI kernel library (`_apb.so`, declared in include/autoprog_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_apb.so')

_vp, _i, _f, _ll = C.c_void_p, C.c_int, C.c_float, C.c_longlong

# name -> (restype, argtypes); mirrors include/autoprog_b200.h one to one
SIGNATURES = {
    'apb_last_error': (C.c_char_p, []),
    'apb_abi_version': (_i, []),
    'apb_outlook_fwd_simt': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    'apb_outlook_bwd_simt': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    'apb_outlook_fwd_fma': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_outlook_fwd_mma': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_outlook_bwd_fma': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_outlook_bwd_mma': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_outlook_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    'apb_outlook_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _i, _vp]),
    'apb_tlce_workspace_floats': (_ll, [_i, _i]),
    'apb_tlce_fwd_bwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _f, _f, _vp, _vp, _vp, _vp, _vp, _i, _vp]),
    'apb_token_label_target': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_onehot_smooth': (_i, [_vp, _vp, _i, _i, _f, _vp]),
    'apb_scale_lazy': (_i, [_vp, _ll, _vp, _ll, _vp, _vp, _i, _vp]),
    'apb_scale_by_scalar': (_i, [_vp, _vp, _ll, _vp, _i, _vp]),
    'apb_ln_fwd': (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _f, _i, _i, _vp]),
    'apb_ln_bwd_workspace_floats': (_ll, [_i]),
    'apb_ln_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _ll, _i, _i, _i, _vp]),
    'apb_colsum_workspace_floats': (_ll, [_ll, _i]),
    'apb_colsum': (_i, [_vp, _ll, _i, _vp, _i, _vp, _i, _vp]),
    'apb_gemm_simt': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_gemm_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_gemm_tc_suggest_split': (_i, [_i, _i, _i]),
    'apb_gemm_tc_pair': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    'apb_splitk_reduce': (_i, [_vp, _vp, _i, _ll, _vp]),
    'apb_gemm_tc_rowsum': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'apb_gemm_tc_rowsum_slots': (_i, [_i, _i]),
    'apb_splitk_reduce2': (_i, [_vp, _vp, _ll, _i, _vp, _vp, _ll, _i, _vp]),
    'apb_mhsa_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_mhsa_fwd_tc': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'apb_mhsa_bwd_tc': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'apb_mhsa_fwd_mma': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'apb_mhsa_bwd_mma': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp]),
    'apb_mhsa_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_mhsa_fwd_simt': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_mhsa_bwd_simt': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_class_attn_fwd': (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_class_attn_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_class_attn_fwd_split': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_class_attn_bwd_split': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _i, _vp]),
    'apb_avgpool2_fwd': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'apb_avgpool2_bwd': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_flip_in_box': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_flip_in_box_dev': (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    'apb_patchify': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_unpatchify': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    'apb_im2col': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _ll, _ll, _ll, _ll, _i, _vp]),
    'apb_bilinear_resize': (_i, [_vp, _vp, _ll, _i, _i, _i, _i, _i, _vp]),
    'apb_bicubic_resize': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'apb_bicubic_resize_bwd': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp]),
    'apb_add_bcast': (_i, [_vp, _vp, _vp, _ll, _ll, _i, _i, _vp]),
    'apb_scale_cast': (_i, [_vp, _vp, _vp, _ll, _ll, _i, _i, _vp]),
    'apb_residual_add': (_i, [_vp, _vp, _vp, _vp, _ll, _ll, _i, _i, _i, _vp]),
    'apb_cast': (_i, [_vp, _vp, _ll, _i, _i, _vp]),
    'apb_add': (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    'apb_gelu_fwd': (_i, [_vp, _vp, _ll, _i, _vp]),
    'apb_gelu_bwd': (_i, [_vp, _vp, _vp, _ll, _i, _vp]),
    'apb_bn_workspace_floats': (_ll, [_ll, _i]),
    'apb_bn_relu_fwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f, _i, _vp, _ll, _i, _i, _vp]),
    'apb_bn_relu_bwd': (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    'apb_adamw_ema': (_i, [_vp, _vp, _vp, _vp, _ll, _vp, _f, _f, _f, _f, C.POINTER(_vp), C.POINTER(_f), _i, _vp, _vp]),
    'apb_launch_count': (_ll, []),
    'apb_fallback_count': (_ll, []),
    'apb_set_pdl': (None, [_i]),
    'apb_get_pdl': (_i, []),
    'apb_debug_gemm_switches': (None, [_i, _i]),
    'apb_debug_umma_probe': (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    'apb_debug_umma_timing': (_i, [_vp, _i, _i, _i, _vp]),
    'apb_debug_mhsa_trace': (_i, [_vp]),
}

_lib = None


class KernelError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load (once) and return the kernel library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise KernelError(
                f'{LIB_PATH} is missing: the CUDA extension has not been built (run `python -m autoprog_b200.build` or '
                f'`__graft_entry__.build()`); autoprog_b200 has no CPU / eager fallback.')
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the header and the library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().apb_last_error()
        raise KernelError(f'{what} failed with code {code}: {msg.decode() if msg else ""}')
