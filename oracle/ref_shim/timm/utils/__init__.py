def unwrap_model(model):
    return model.module if hasattr(model, 'module') else model
