"""TEST INFRASTRUCTURE ONLY -- live-reference loader.

Imports the *unmodified* reference under the import stubs listed in SURVEY.md section 8(c): from /root/reference in the
authoring container (golden fixtures, oracle pinning), or from the staged copy baseline/_ref that
`__graft_entry__.build()` makes there (git-ignored, shipped to the GPU box by gpurun), which is what the reference arm of
bench.py (`--impl reference`, `cpu_baseline`, `gpu_eager_baseline`) and tests/test_gpu_runner.py use on the box.
Nothing that runs on the box reads /root/reference.  Nothing under dig_b200/ imports this file.
"""
import math
import os
import sys
import types

import torch
import torch.nn as nn

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED_ROOT = os.path.join(_REPO_ROOT, "baseline", "_ref")     # git-ignored copy made by __graft_entry__.build() (travels to the GPU box)


def _pick_root():
    for cand in (os.environ.get("DIG_REFERENCE_ROOT"), "/root/reference", STAGED_ROOT):
        if cand and os.path.isfile(os.path.join(cand, "modeling_pretrain_moco_mim_ori.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _pick_root()


def stage_reference(src="/root/reference", dst=STAGED_ROOT):
    """Copy the reference's Python files (unmodified) into the git-ignored baseline/_ref so that the reference arm of bench.py and the
    runner test can import them on the GPU box, where /root/reference does not exist.  Returns the number of files staged."""
    import shutil
    if not os.path.isfile(os.path.join(src, "modeling_pretrain_moco_mim_ori.py")):
        return 0
    n = 0
    for base, dirs, files in os.walk(src):
        dirs[:] = [d for d in dirs if not d.startswith(".") and d not in ("output", "__pycache__")]
        for f in files:
            if f.endswith(".py"):
                rel = os.path.relpath(os.path.join(base, f), src)
                out = os.path.join(dst, rel)
                os.makedirs(os.path.dirname(out), exist_ok=True)
                if not os.path.isfile(out) or os.path.getmtime(out) < os.path.getmtime(os.path.join(base, f)):
                    shutil.copyfile(os.path.join(base, f), out)
                n += 1
    return n

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


_REF_FT_CACHE = None


def import_reference_finetune():
    """Returns (models.model_builder module, SeqCrossEntropyLoss class) of the REFERENCE (fine-tuning path)."""
    global _REF_FT_CACHE
    if _REF_FT_CACHE is not None:
        return _REF_FT_CACHE
    import_reference()
    saved = {n: sys.modules.pop(n) for n in _SHADOWED + ("models", "engine_for_finetuning") if n in sys.modules}
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        import importlib
        importlib.import_module("modeling_pretrain_vit")          # registers simmim_vit_*_patch4_32x128 (V:114-136)
        mb = importlib.import_module("models.model_builder")
        crit = importlib.import_module("loss.seqCrossEntropyLoss").SeqCrossEntropyLoss
        assert os.path.abspath(mb.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), mb.__file__
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for n in _SHADOWED:
            sys.modules.pop(n, None)
        sys.modules.update({k: v for k, v in saved.items() if k in _SHADOWED})
    _REF_FT_CACHE = (mb, crit)
    return _REF_FT_CACHE


def create_reference_model(name="pretrain_simmim_moco_ori_vit_small_patch4_32x128", seed=0, **over):
    """Factory call exactly as run_mae_pretraining_moco.get_model makes it (R:278-294)."""
    M = import_reference()[0]
    torch.manual_seed(seed)
    kw = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2,
              num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
    kw.update(over)
    # the reference's own factory function (the stub registry is shared with dig_b200's drop-in factories of the same names)
    model = getattr(M, name)(**{k: v for k, v in kw.items() if v is not None})
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


# ----------------------------------------------------------------------------------------------------------------------
# The reference RUNNER (run_mae_pretraining_moco.py) driven through this repo's drop-in modules
# ----------------------------------------------------------------------------------------------------------------------
class SyntheticCrops(torch.utils.data.Dataset):
    """Stand-in for AloneImageLmdb (dataset/dataset_image.py:126-158): item = ((img, aug_img, vis_mask), ones(1), ones(1)) with two
    normalised 3x32x128 views and a float64 [num_view, 256] mask holding int(mask_ratio*256) ones per view (masking_generator.py:20-46)."""

    def __init__(self, n, mask_ratio=0.7, num_view=2, seed=1234):
        self.n, self.num_view, self.seed = n, num_view, seed
        self.n_mask = int(mask_ratio * 256)

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        import numpy as np
        g = torch.Generator().manual_seed(self.seed + i)
        img = torch.rand(3, 32, 128, generator=g) * 2 - 1
        aug = torch.rand(3, 32, 128, generator=g) * 2 - 1
        mask = torch.zeros(self.num_view, 256, dtype=torch.float64)
        for v in range(self.num_view):
            mask[v, torch.randperm(256, generator=g)[:self.n_mask]] = 1.0
        return (img, aug, mask), np.ones(1), np.ones(1)


def import_runner(dataset_len=24):
    """Import the reference's run_mae_pretraining_moco.py so that its `import modeling_pretrain_moco_mim_ori` / `from
    engine_for_pretraining_moco import train_one_epoch` (R:34,36) bind THIS repo's drop-in modules (repo root first on sys.path), while
    `optim_factory`, `utils.*` and `custom_optim` are the reference's own files.  The LMDB dataset package (needs lmdb / imgaug / cv2) is
    replaced by SyntheticCrops; the ten timm.optim classes optim_factory imports but the README configuration never constructs are stubs."""
    if not reference_available():
        raise RuntimeError("reference sources not found at %s (run __graft_entry__.build() where /root/reference exists)" % REFERENCE_ROOT)
    install_stubs()

    class _Unavailable:
        def __init__(self, *a, **k):
            raise RuntimeError("timm.optim is stubbed: only --opt adamw (custom_optim.AdamW) is available")

    _mod("timm.optim")
    for mod, cls in (("adafactor", "Adafactor"), ("adahessian", "Adahessian"), ("adamp", "AdamP"), ("lookahead", "Lookahead"),
                     ("nadam", "Nadam"), ("novograd", "NovoGrad"), ("nvnovograd", "NvNovoGrad"), ("radam", "RAdam"),
                     ("rmsprop_tf", "RMSpropTF"), ("sgdp", "SGDP")):
        _mod("timm.optim." + mod, **{cls: _Unavailable})
    _mod("dataset")
    _mod("dataset.datasets", build_pretraining_aloneimage_dataset=lambda args: SyntheticCrops(
        dataset_len * args.batch_size, args.mask_ratio, args.num_view))
    _mod("dataset.dist_multisrc_batch_sampler", DistributedMultiSrcBatchWiseSampler=object)
    _mod("dataset.concatdatasets", ConcatDataset=object)
    import importlib
    for n in ("run_mae_pretraining_moco", "optim_factory", "utils", "utils.utils", "utils.logging", "custom_optim") + _SHADOWED:
        sys.modules.pop(n, None)
    for k in [k for k in sys.modules if k.startswith("custom_optim.") or k.startswith("utils.")]:
        sys.modules.pop(k, None)
    saved_path = list(sys.path)
    sys.path[:] = [_REPO_ROOT, REFERENCE_ROOT] + [p for p in saved_path if p not in (_REPO_ROOT, REFERENCE_ROOT)]
    try:
        runner = importlib.import_module("run_mae_pretraining_moco")
        eng = sys.modules["engine_for_pretraining_moco"]
        assert os.path.abspath(eng.__file__).startswith(_REPO_ROOT) and not os.path.abspath(eng.__file__).startswith(
            os.path.abspath(REFERENCE_ROOT)), eng.__file__
        assert os.path.abspath(runner.__file__).startswith(os.path.abspath(REFERENCE_ROOT)), runner.__file__
        import dig_b200.registry as dig_registry
        import dig_b200.modeling  # noqa: F401
        _MODEL_REGISTRY.update(dig_registry._MODELS)      # timm.models.create_model (stub) -> this repo's drop-in factories

        class _Tee:      # utils/logging.py:27-67 Logger closes sys.stdout in __del__ (it would take pytest's capture file with it)
            def __init__(self, fpath=None):
                self.console = sys.stdout
                self.file = open(fpath, "w") if fpath else None

            def write(self, msg):
                self.console.write(msg)
                if self.file is not None:
                    self.file.write(msg)

            def flush(self):
                self.console.flush()
                if self.file is not None:
                    self.file.flush()

            def close(self):
                if self.file is not None:
                    self.file.close()
                    self.file = None

        runner.Logger = _Tee
    finally:
        sys.path[:] = saved_path
    return runner
