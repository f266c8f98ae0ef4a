"""Layer-index mapping used by elastic depth and MoGrow (reference prog/helpers.py:254-262)."""


def new_idx(idx, prev_l, new_l):
    """Old-model layer that layer `idx` of the grown model inherits from."""
    span = new_l // prev_l * prev_l
    untouched = prev_l - new_l % prev_l
    base = idx * prev_l // span
    if base < untouched:
        return base
    return (idx + untouched) * prev_l // (span + prev_l)


def get_new_layer_idx(prev_l, new_l):
    """Indices of the layers that are NEW in the grown model (they repeat their predecessor's source layer)."""
    return [i for i in range(new_l) if new_idx(i, prev_l, new_l) == new_idx(i - 1, prev_l, new_l)]


# ---------------------------------------------------------------------------------------------------------------
# MoGrow / slice weight inheritance for DEPTH growth at fixed width (the only case `auto_grow` supports:
# main_prog.py:1561 asserts len(h_list) == 1).  The reference's `prog/helpers.py` loaders (load_slice_clone_ema & co.)
# also run unmodified on autoprog_b200 modules -- see tests/test_mogrow.py -- this is the same mapping expressed
# over state_dict names.
# ---------------------------------------------------------------------------------------------------------------
_STAGE_PREFIXES = ('network.0.', 'network.2.', 'network.3.', 'network.4.')


def _unwrap(m):
    return m.module if hasattr(m, 'module') else m


def source_name(name: str, new_model, old_model) -> str:
    """Name of the old-model tensor that `name` of the grown model inherits from (prog/helpers.py:622-627)."""
    for pre in _STAGE_PREFIXES:
        if name.startswith(pre):
            parts = name.split('.')
            stage = int(parts[1])
            n_new, n_old = len(new_model.network[stage]), len(old_model.network[stage])
            if n_new > n_old:
                parts[2] = str(new_idx(int(parts[2]), n_old, n_new))
            return '.'.join(parts)
    return name


def load_slice_clone_ema(model, checkpoint_model, ema_model_list=None, debug=False):
    """Grow `model` (deeper, same width) from `checkpoint_model` (the slowest EMA in the reference's call,
    main_prog.py:1378-1382): layer i takes layer new_idx(i) of the source.  Parameters only: like the reference
    (prog/helpers.py:665-668, copy commented out) BatchNorm running statistics are NOT inherited."""
    import torch
    model_u, src = _unwrap(model), _unwrap(checkpoint_model)
    src_params = dict(src.named_parameters())
    with torch.no_grad():
        for name, p in model_u.named_parameters():
            sname = source_name(name, model_u, src)
            if sname not in src_params:
                if debug:
                    print(f"no parameter '{sname}' in slice.")
                continue
            q = src_params[sname]
            if q.shape != p.shape:
                raise NotImplementedError(f'width growth is not supported ({name}: {tuple(q.shape)} -> {tuple(p.shape)})')
            p.copy_(q)
    return model, checkpoint_model


load_slice_clone = load_slice_clone_ema
