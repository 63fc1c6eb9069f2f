"""Input stage: the numpy oracle against torchvision / PIL and the reference's RandomMaskingGenerator on CPU, and
(-m gpu) the CUDA kernels against the oracle bit-for-bit."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import input_stage_ref as O


def test_oracle_normalize_matches_torchvision():
    tv = pytest.importorskip("torchvision.transforms")
    from PIL import Image
    rng = np.random.RandomState(0)
    x = rng.randint(0, 256, size=(3, 32, 128, 3), dtype=np.uint8)
    tf = tv.Compose([tv.ToTensor(), tv.Normalize(mean=torch.tensor(0.5), std=torch.tensor(0.5))])     # datasets.py:30-37
    ref = torch.stack([tf(Image.fromarray(x[i])) for i in range(3)]).numpy()
    assert np.array_equal(O.normalize_view(x), ref)                                                 # bit-exact
    gray = np.stack([np.asarray(tv.Grayscale(num_output_channels=3)(Image.fromarray(x[i]))) for i in range(3)])   # RandomGrayscale's conversion
    assert np.array_equal(O.to_gray_u8(x), gray)


def test_oracle_masks_have_the_reference_distribution():
    from oracle import ref_shims
    if not ref_shims.reference_available():
        pytest.skip("reference not available")
    sys.path.insert(0, ref_shims.REFERENCE_ROOT)
    try:
        import importlib
        mg = importlib.import_module("masking_generator")
    finally:
        sys.path.remove(ref_shims.REFERENCE_ROOT)
    gen = mg.RandomMaskingGenerator((8, 32), 0.7, num_view=2)
    np.random.seed(0)
    ref = np.stack([gen() for _ in range(400)])                      # [400, 2, 256] float64
    ours = O.random_masks(400, 2, gen.num_mask, seed=7, step=3)
    assert gen.num_mask == 179
    assert (ref.sum(-1) == 179).all() and (ours.sum(-1) == 179).all()          # exact count per (sample, view)
    # per-position masking frequency: binomial(400*2, 0.699) -> sigma = 0.016; both generators within 5 sigma of 179/256 everywhere
    for m in (ref, ours.astype(np.float64)):
        f = m.mean(axis=(0, 1))
        assert abs(f - 179 / 256).max() < 0.085
    # the two views of a sample are independent draws: overlap count ~ hypergeometric mean 179*179/256 = 125.2
    for m in (ref, ours.astype(np.float64)):
        ov = (m[:, 0] * m[:, 1]).sum(-1).mean()
        assert abs(ov - 179 * 179 / 256) < 1.5
    # pure function of (seed, step, global sample index, view): independent of how the batch is split
    a = O.random_masks(6, 2, 179, seed=7, step=3)
    b = O.random_masks(3, 2, 179, seed=7, step=3, sample0=3)
    assert np.array_equal(a[3:], b) and not np.array_equal(a[:3], b)
    assert not np.array_equal(O.random_masks(2, 2, 179, 7, 4), O.random_masks(2, 2, 179, 7, 3))


@pytest.mark.gpu
def test_kernels_match_the_oracle_bit_for_bit():
    import __graft_entry__ as ge
    ge.build()
    from dig_b200.input_stage import GpuInputStage
    rng = np.random.RandomState(1)
    B = 37
    img = rng.randint(0, 256, size=(B, 32, 128, 3), dtype=np.uint8)
    aug = rng.randint(0, 256, size=(B, 32, 128, 3), dtype=np.uint8)
    img[0, 0, :4] = [[0, 0, 0], [255, 255, 255], [1, 2, 3], [254, 128, 127]]
    stage = GpuInputStage(mask_ratio=0.7, num_view=2, gray_p=0.2, seed=11)
    x, y, m = stage(torch.from_numpy(img).cuda(), torch.from_numpy(aug).cuda(), sample0=5, step=9)
    rx, ry = O.normalize_views(img, aug, 0.2, 11, 9, sample0=5)
    assert np.array_equal(x.cpu().numpy(), rx) and np.array_equal(y.cpu().numpy(), ry)
    assert sum(bool(O.gray_decision(11, 9, 5 + b, 0.2)) for b in range(B)) > 0          # the grayscale branch was exercised
    assert m.dtype == torch.bool and tuple(m.shape) == (B, 2, 256)
    assert np.array_equal(m.cpu().numpy().astype(np.uint8), O.random_masks(B, 2, 179, 11, 9, sample0=5))
    # the loader's float64 layout and a pinned-host input
    from dig_b200.ops import call
    mf = torch.empty(B, 2, 256, dtype=torch.float64, device="cuda")
    call("dig_random_masks", None, mf, B, 2, 179, 11, 9, 5)
    assert torch.equal(mf.bool(), m)
    x2, _, _ = stage(torch.from_numpy(img).pin_memory(), torch.from_numpy(aug).pin_memory(), sample0=5, step=9)
    assert torch.equal(x2, x)


@pytest.mark.gpu
def test_engine_accepts_uint8_batches_through_the_input_stage():
    """train_one_epoch with a loader that ships uint8 views and no masks (args.gpu_input_stage): same meters, finite, learns."""
    import types
    import __graft_entry__ as ge
    ge.build()
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    from dig_b200.engine import train_one_epoch
    from dig_b200.input_stage import GpuInputStage
    from dig_b200.optim import FusedAdamW
    from dig_b200.utils import NativeScalerWithGradNormCount
    torch.manual_seed(0)
    model = dig_b200.create_model("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", pretrained=False, drop_path_rate=0.0, drop_block_rate=None,
                                  mlp_dim=512, dim=64, T=0.2, num_windows=4, encoder_type="vit", queue_size=8, patchnet_name="no_patchtrans").cuda()
    opt = FusedAdamW([{"params": [p for p in model.parameters() if p.requires_grad], "weight_decay": 0.05, "lr_scale": 1.0}], lr=1e-3)
    args = types.SimpleNamespace(num_view=2, moco_m=0.99, use_moco_m_cos=1, epochs=2, contrast_start_epoch=0, contrast_warmup_steps=0,
                                 loss_weight_contrast=0.1, loss_weight_pixel=1.0, only_mim_on_ori_img=True, eval_freq=10 ** 9, output_dir=None,
                                 gpu_input_stage=GpuInputStage(0.7, 2, gray_p=0.2, seed=3))
    rng = np.random.RandomState(2)
    u8 = torch.from_numpy(rng.randint(0, 256, size=(8, 32, 128, 3), dtype=np.uint8)).pin_memory()
    u8b = torch.from_numpy(rng.randint(0, 256, size=(8, 32, 128, 3), dtype=np.uint8)).pin_memory()
    loader = [([u8, u8b], None, None)] * 10
    st0 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 0, NativeScalerWithGradNormCount(), max_norm=None,
                          patch_size=4, normlize_target=False, start_steps=0, args=args)
    st1 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 1, NativeScalerWithGradNormCount(), max_norm=None,
                          patch_size=4, normlize_target=False, start_steps=10, args=args)
    assert all(v == v for v in st1.values()) and st1["loss_pixel"] < st0["loss_pixel"]
