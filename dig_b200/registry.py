"""Minimal stand-in for timm 0.3.2's model registry (`timm.models.registry.register_model`,
`timm.models.create_model`), which the reference runner uses to build the model
(run_mae_pretraining_moco.py:278-294).  timm is not a dependency of dig_b200: when it is installed the
factories are registered with it as well, so `timm.models.create_model(args.model, ...)` keeps working.
"""
_MODELS = {}


def register_model(fn):
    _MODELS[fn.__name__] = fn
    try:  # also expose through a real timm, if present
        from timm.models.registry import register_model as _timm_register
        _timm_register(fn)
    except Exception:
        pass
    return fn


def create_model(model_name, pretrained=False, **kwargs):
    """timm 0.3.2 semantics: kwargs whose value is None are dropped before the factory call."""
    if model_name not in _MODELS:
        raise RuntimeError("Unknown model (%s)" % model_name)
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _MODELS[model_name](pretrained=pretrained, **kwargs)


def list_models():
    return sorted(_MODELS)
