"""timm-compatible model registry (`register_model` / `create_model`).

The reference obtains both from timm 0.4.5 (models/volo.py:23, main_prog.py:411-426), which is not installed
here.  Same behaviour as that version's `create_model` for the kwargs the trainer passes: kwargs whose value is
None are dropped, `bn_tf/bn_momentum/bn_eps` are popped, `drop_connect_rate` maps to `drop_path_rate`, the factory
is called with `pretrained=` + the rest.  If timm IS importable the factories are registered there as well, so
`timm.create_model('volo_d1')` resolves to this package.
"""
from __future__ import annotations

_MODELS = {}


def register_model(fn):
    _MODELS[fn.__name__] = fn
    try:  # pragma: no cover - timm is absent in the build image
        from timm.models.registry import register_model as _timm_register
        _timm_register(fn)
    except Exception:
        pass
    return fn


def list_models():
    return sorted(_MODELS)


def is_model(name):
    return name in _MODELS


def create_model(model_name, pretrained=False, checkpoint_path='', scriptable=None, exportable=None, no_jit=None, **kwargs):
    if model_name not in _MODELS:
        raise RuntimeError(f'Unknown model ({model_name}); registered: {list_models()}')
    for k in ('bn_tf', 'bn_momentum', 'bn_eps'):
        kwargs.pop(k, None)
    dcr = kwargs.pop('drop_connect_rate', None)
    if dcr is not None and kwargs.get('drop_path_rate', None) is None:
        kwargs['drop_path_rate'] = dcr
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    model = _MODELS[model_name](pretrained=pretrained, **kwargs)
    if checkpoint_path:
        import torch
        sd = torch.load(checkpoint_path, map_location='cpu')
        model.load_state_dict(sd.get('state_dict', sd.get('model', sd)))
    return model
