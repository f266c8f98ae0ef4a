"""autoprog_b200: the AutoProg (VOLO / DeiT + token-labeling) training hot path on hand-written sm_100a kernels.

    from autoprog_b200 import create_model, TokenLabelCrossEntropy, autocast
    model = create_model('volo_d1', img_size=224).cuda()
    with autocast():                              # bf16 tensor-core path; omit for the fp32 parity path
        loss = TokenLabelCrossEntropy(dense_weight=0.5)(model(images), token_labels)

The kernels live in `_apb.so` (C ABI: include/autoprog_b200.h), built by `python -m autoprog_b200.build`.
There is no CPU or eager fallback.
"""
from .registry import create_model, register_model, list_models, is_model  # noqa: F401
from .ops import autocast  # noqa: F401
from . import volo, submodels, deit  # noqa: F401  (registers volo_d1..d5, model_variant, deit_*)
from .volo import VOLO  # noqa: F401
from .cross_entropy import (SoftTargetCrossEntropy, TokenLabelCrossEntropy, TokenLabelGTCrossEntropy,  # noqa: F401
                            TokenLabelSoftTargetCrossEntropy)
from .token_label import create_token_label_target  # noqa: F401
from .progressive import progressive_schedule, make_divisible  # noqa: F401
from .helpers import new_idx, get_new_layer_idx  # noqa: F401
from .scaler import NoScaler, Bf16Scaler, NativeScaler, ApexScaler  # noqa: F401

__version__ = '0.1.0'
