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
