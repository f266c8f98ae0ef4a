class VisionTransformer:  # placeholder: DeiT arithmetic lives in real timm, absent here
    def __init__(self, *a, **k):
        raise RuntimeError('timm VisionTransformer is not available in the shim')


def _cfg(url='', **kwargs):
    return dict(url=url, **kwargs)
