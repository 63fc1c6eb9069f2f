"""TEST INFRASTRUCTURE ONLY -- live-reference loader (authoring container only).

Imports the *unmodified* reference from /root/reference under the import stubs listed in
SURVEY.md section 8(c) so that golden fixtures can be generated from the reference's own
forward / train_one_epoch.  /root/reference does not exist on the GPU box, therefore nothing
that runs there (gpu tests, smoke(), bench.py) may import this module; they use the committed
fixtures under tests/golden/ and the restatement in oracle/restatement.py instead.

Nothing under dig_b200/ imports this file.
"""
import math
import os
import sys
import types

import torch
import torch.nn as nn

REFERENCE_ROOT = os.environ.get("DIG_REFERENCE_ROOT", "/root/reference")

_MODEL_REGISTRY = {}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modeling_pretrain_moco_mim_ori.py"))


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _register_model(fn):
    _MODEL_REGISTRY[fn.__name__] = fn
    return fn


def _create_model(name, pretrained=False, **kwargs):
    # timm 0.3.2 create_model drops kwargs whose value is None (R:280-292 passes drop_block_rate=None)
    kwargs = {k: v for k, v in kwargs.items() if v is not None}
    return _MODEL_REGISTRY[name](pretrained=pretrained, **kwargs)


def _drop_path(x, drop_prob=0.0, training=False):
    if drop_prob == 0.0 or not training:
        return x
    keep = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    rnd = keep + torch.rand(shape, dtype=x.dtype, device=x.device)
    rnd.floor_()
    return x.div(keep) * rnd


def _to_2tuple(x):
    return tuple(x) if isinstance(x, (tuple, list)) else (x, x)


def install_stubs():
    """Inject the eight missing names into sys.modules."""
    if "timm" in sys.modules and getattr(sys.modules["timm"], "_dig_stub", False):
        return
    timm = _mod("timm", _dig_stub=True)
    models = _mod("timm.models", create_model=_create_model)
    registry = _mod("timm.models.registry", register_model=_register_model)
    layers = _mod("timm.models.layers", trunc_normal_=nn.init.trunc_normal_, drop_path=_drop_path,
                  to_2tuple=_to_2tuple)
    data = _mod("timm.data")
    constants = _mod("timm.data.constants", IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406),
                     IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225))
    tutils = _mod("timm.utils", get_state_dict=lambda m, *a, **k: m.state_dict())
    timm.models, timm.data, timm.utils = models, data, tutils
    models.registry, models.layers = registry, layers
    data.constants = constants
    _mod("torch._six", inf=math.inf)

    class _SummaryWriter:  # tensorboardX is absent
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass

        def flush(self):
            pass

    _mod("tensorboardX", SummaryWriter=_SummaryWriter)

    class _Turtle:  # E:10 imports turtle (tkinter absent); never used
        pass

    _mod("turtle", Turtle=_Turtle)


_REF_CACHE = None
_SHADOWED = ("modeling_finetune", "modeling_pretrain_vit", "modeling_pretrain_moco_mim_ori", "engine_for_pretraining_moco")


def import_reference():
    """Returns (modeling module M, engine module E, utils module U, custom_optim AdamW class) of the REFERENCE.

    The repo root holds drop-in modules with the reference's names (modeling_pretrain_moco_mim_ori.py,
    engine_for_pretraining_moco.py); they are moved aside in sys.modules while the reference's own files are imported from
    REFERENCE_ROOT and put back afterwards, so both can live in one process."""
    global _REF_CACHE
    if _REF_CACHE is not None:
        return _REF_CACHE
    if not reference_available():
        raise RuntimeError("reference sources not found at %s" % REFERENCE_ROOT)
    install_stubs()
    saved = {n: sys.modules.pop(n) for n in _SHADOWED if n in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        M = importlib.import_module("modeling_pretrain_moco_mim_ori")
        E = importlib.import_module("engine_for_pretraining_moco")
        U = importlib.import_module("utils.utils")
        AdamW = importlib.import_module("custom_optim.adamw").AdamW
        assert os.path.abspath(M.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), M.__file__
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for n in _SHADOWED:
            sys.modules.pop(n, None)
        sys.modules.update(saved)
    _REF_CACHE = (M, E, U, AdamW)
    return _REF_CACHE


def create_reference_model(name="pretrain_simmim_moco_ori_vit_small_patch4_32x128", seed=0, **over):
    """Factory call exactly as run_mae_pretraining_moco.get_model makes it (R:278-294)."""
    import_reference()
    torch.manual_seed(seed)
    kw = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2,
              num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
    kw.update(over)
    model = _create_model(name, **kw)
    model.train()
    return model


def ensure_cpu_process_group():
    """contrastive_loss needs an initialised group even at W=1 (M:449-453)."""
    import torch.distributed as dist
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("gloo", rank=0, world_size=1)


class cpu_patches:
    """Context manager: make Tensor.cuda / cuda.synchronize no-ops for a CPU run."""

    def __enter__(self):
        self._cuda = torch.Tensor.cuda
        self._sync = torch.cuda.synchronize
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.cuda.synchronize = lambda *a, **k: None
        return self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._cuda
        torch.cuda.synchronize = self._sync
        return False
