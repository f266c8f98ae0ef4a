_REG = {}


def register_model(fn):
    _REG[fn.__name__] = fn
    return fn


def create_model(name, **kw):
    kw = {k: v for k, v in kw.items() if v is not None}
    return _REG[name](**kw)
