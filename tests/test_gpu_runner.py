"""-m gpu: the reference's OWN runner (run_mae_pretraining_moco.py, unmodified, imported from the staged baseline/_ref) driven
through this repo's drop-in modules -- VERDICT r1 "drive the real runner".

What is the reference's and what is this repo's in these tests:
  reference : main() / get_args() / get_model() via timm.models.create_model (R:278-294, stub registry), optim_factory.create_optimizer
              -> custom_optim.AdamW (per-tensor torch ops), utils.NativeScalerWithGradNormCount (fp16 GradScaler, loss scale 65536,
              U:477-498), cosine schedules, DataLoader + DistributedSampler, save_model / auto_load_model (U:546-669), log.txt
  this repo : `import modeling_pretrain_moco_mim_ori` (factories -> DigMoCoViT on the sm_100a kernels) and
              `from engine_for_pretraining_moco import train_one_epoch`
"""
import json
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
MODEL = "pretrain_simmim_moco_ori_vit_tiny_patch4_32x128"


def _argv(out_dir, epochs, batch=8):
    return ["run_mae_pretraining_moco.py", "--batch_size", str(batch), "--epochs", str(epochs), "--model", MODEL, "--mask_ratio", "0.7",
            "--num_view", "2", "--moco_dim", "256", "--moco_mlp_dim", "4096", "--moco_m", "0.99", "--moco_t", "0.2", "--num_windows", "4",
            "--patchnet_name", "no_patchtrans", "--loss_weight_pixel", "1.0", "--loss_weight_contrast", "0.1", "--only_mim_on_ori_img",
            "--contrast_warmup_steps", "0", "--warmup_epochs", "0", "--opt", "adamw", "--opt_betas", "0.9", "0.999", "--lr", "1.5e-4",
            "--weight_decay", "0.05", "--drop_path", "0.0", "--num_workers", "0", "--output_dir", out_dir, "--log_dir", "", "--device", "cuda",
            "--save_ckpt_freq", "1", "--seed", "0"]


def _run_main(runner, argv):
    old_argv, old_stdout, old_load = sys.argv, sys.stdout, torch.load
    sys.argv = argv
    # the reference's auto_load_model (U:581-669) calls torch.load(path, map_location='cpu') on a checkpoint that holds its argparse
    # Namespace; torch >= 2.6 defaults to weights_only=True.  One more version shim of the kind SURVEY.md 8(c) lists.
    torch.load = lambda *a, **k: old_load(*a, **{**k, "weights_only": k.get("weights_only", False)})
    try:
        args = runner.get_args()
        if not args.log_dir:
            args.log_dir = None
        os.makedirs(args.output_dir, exist_ok=True)
        runner.main(args)
    finally:
        sys.argv, sys.stdout, torch.load = old_argv, old_stdout, old_load
    return args


def _runner():
    import __graft_entry__ as ge
    ge.build()
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference not staged (baseline/_ref is created by __graft_entry__.build() where /root/reference exists)")
    return ref_shims.import_runner(dataset_len=4)


def test_unmodified_runner_trains_saves_and_resumes(tmp_path):
    runner = _runner()
    out = str(tmp_path / "run")
    _run_main(runner, _argv(out, epochs=1))
    lines = [json.loads(l) for l in open(os.path.join(out, "log.txt"))]
    assert len(lines) == 1 and lines[0]["epoch"] == 0
    st = lines[0]
    for k in ("train_loss", "train_loss_pixel", "train_loss_contrast", "train_grad_norm", "train_loss_scale", "train_lr", "train_moco_m"):
        assert k in st and st[k] == st[k], k
    assert st["train_loss_scale"] == 65536.0                      # the REFERENCE's fp16 GradScaler drove our backward (U:481)
    assert 0.0 < st["train_loss"] < 5.0 and st["train_grad_norm"] > 0.0
    ck = os.path.join(out, "checkpoint-0.pth")                     # written by the reference's save_model (U:546-573)
    assert os.path.isfile(ck)
    sd = torch.load(ck, map_location="cpu", weights_only=False)
    assert {"model", "optimizer", "epoch", "scaler"} <= set(sd) and "encoder.blocks.0.attn.qkv.weight" in sd["model"]
    # resume: the reference's auto_load_model finds checkpoint-0.pth, loads model + custom_optim.AdamW + GradScaler state, runs epoch 1
    _run_main(runner, _argv(out, epochs=2))
    lines = [json.loads(l) for l in open(os.path.join(out, "log.txt"))]
    assert [l["epoch"] for l in lines] == [0, 1]
    assert lines[1]["train_loss_pixel"] < lines[0]["train_loss_pixel"]      # it keeps learning from the resumed state


def test_reference_scaler_and_optimizer_match_the_repos_own_path(tmp_path):
    """Same data, same init: (reference GradScaler x65536 + custom_optim.AdamW) vs (dig_b200 scaler + FusedAdamW): epoch-average losses
    agree to 1e-3 relative (the bf16 gradient rounding is scale-invariant; the two AdamW implementations compute the same update)."""
    runner = _runner()
    out = str(tmp_path / "ref")
    args = _run_main(runner, _argv(out, epochs=1))
    ref = json.loads(open(os.path.join(out, "log.txt")).readline())

    import numpy as np
    import random
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from dig_b200.engine import train_one_epoch
    from dig_b200.optim import FusedAdamW
    from dig_b200.utils import NativeScalerWithGradNormCount, cosine_scheduler
    from oracle.ref_shims import SyntheticCrops
    torch.manual_seed(0); np.random.seed(0); random.seed(0)
    model = dig_b200.create_model(MODEL, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2,
                                  num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans").cuda()
    ds = SyntheticCrops(4 * 8, 0.7, 2)
    sampler = torch.utils.data.DistributedSampler(ds, num_replicas=1, rank=0, shuffle=True)
    sampler.set_epoch(0)
    loader = torch.utils.data.DataLoader(ds, sampler=sampler, batch_size=8, num_workers=0, drop_last=True)
    decay = [p for n, p in model.named_parameters() if p.requires_grad and not (p.dim() == 1 or n.endswith(".bias"))]
    nodecay = [p for n, p in model.named_parameters() if p.requires_grad and (p.dim() == 1 or n.endswith(".bias"))]
    lr = 1.5e-4 * 8 / 256
    opt = FusedAdamW([{"params": nodecay, "weight_decay": 0.0, "lr_scale": 1.0}, {"params": decay, "weight_decay": 0.05, "lr_scale": 1.0}],
                     lr=lr, betas=(0.9, 0.999))
    n_it = len(loader)
    lr_s = cosine_scheduler(lr, args.min_lr, 1, n_it, warmup_epochs=0, warmup_steps=args.warmup_steps)
    wd_s = cosine_scheduler(0.05, 0.05, 1, n_it)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        own = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 0, NativeScalerWithGradNormCount(), None,
                              start_steps=0, lr_schedule_values=lr_s, wd_schedule_values=wd_s, patch_size=4, normlize_target=False, args=args)
    assert own["loss"] == pytest.approx(ref["train_loss"], rel=1e-3)
    assert own["loss_pixel"] == pytest.approx(ref["train_loss_pixel"], rel=1e-3)
    assert own["loss_contrast"] == pytest.approx(ref["train_loss_contrast"], rel=2e-3)
    assert own["grad_norm"] == pytest.approx(ref["train_grad_norm"], rel=2e-2)
