"""`model_variant`: the elastic VOLO family `volo_h{h}_l{l}` (reference models/submodels.py:9-41).

Unlike the reference (whose assert only admits 'volo_h12_l18' although main_prog.py:387 asks for e.g. 'volo_h12_l9',
SURVEY.md gotcha 1) any even h and any l is accepted; the arithmetic of the body is the reference's.
"""
from .progressive import make_divisible
from .registry import register_model
from .volo import build_volo


def parse_variant(variant: str):
    parts = variant.split('_')
    if len(parts) != 3 or parts[0] != 'volo' or not parts[1].startswith('h') or not parts[2].startswith('l'):
        raise ValueError(f"variant must look like 'volo_h12_l18', got {variant!r}")
    h, l = int(parts[1][1:]), int(parts[2][1:])
    assert h % 2 == 0, 'h must be divisible by 2'
    return h, l


def variant_layers(l: int):
    if l > 2:
        l0 = make_divisible(l * 0.23, 2)
        return [l0, l - l0, 0, 0]
    print('Warning: layer too small, set to 2')
    return [1, 1, 0, 0]


@register_model
def model_variant(variant='', pretrained=False, **kwargs):
    h, l = parse_variant(variant)
    return build_volo(variant_layers(l), [h * 16, h * 32, h * 32, h * 32], [h // 2, h, h, h], 3, **kwargs)
