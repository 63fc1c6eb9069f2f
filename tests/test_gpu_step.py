"""-m gpu: the whole pre-training step through the drop-in boundary (factory -> model.forward -> backward, and train_one_epoch)
against the reference's golden fixtures (tests/golden/, written from the unmodified reference) and the CPU oracle.

Tolerances (bf16 tensor-core operands, fp32 accumulation/statistics, compared with the fp32 reference):
  * masked-pixel loss and weighted total loss: 1e-3 relative  (BASELINE.json north_star)
  * InfoNCE loss: 1e-3 relative at B >= 8; 5e-3 at B = 2, where every BatchNorm of the contrastive heads normalises over
    only 16 rows and amplifies operand rounding (the same deviation appears when the fp32 oracle itself rounds its GEMM
    operands to bf16 -- see DESIGN.md "Numerics")
  * encoder activations: 8e-2 absolute on values of mean |x| = 2.2
"""
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")


def make(name):
    import __graft_entry__ as ge
    ge.build()
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    torch.manual_seed(0)
    return dig_b200.create_model(name, **KW).train()


def run_step(model, g):
    from oracle import restatement as R
    from dig_b200.engine import masked_pixel_mse
    img, aug, mask = R.synthetic_batch(g["B"], seed=g["seed_data"])
    mk = mask.clone()
    mk[:, 1, :] = False
    model.cuda()
    out = model(img.cuda(), aug.cuda(), mk.cuda(), g["m"], True)
    lpix = masked_pixel_mse(out["vis_out"][0], img.cuda(), mk[:, 0].cuda())
    loss = out["contra_loss"] * 0.1 + lpix
    loss.backward()
    torch.cuda.synchronize()
    return out, lpix, loss


def test_step_with_both_views_masked_matches_reference_golden():
    """--only_mim_on_ori_img 0 (M:571-575, E:137-141): the second view keeps its mask, the pixel head decodes the masked rows of both views
    and both are compared with the ORIGINAL image's patches, 1/2 each -- against the fixture of the unmodified reference."""
    from oracle import restatement as R
    from dig_b200.engine import masked_pixel_mse
    g = torch.load(os.path.join(GOLD, "ref_step_tiny_b4_bothviews.pt"), weights_only=False)
    model = make(g["model"]).cuda()
    img, aug, mask = R.synthetic_batch(g["B"], seed=g["seed_data"])
    mk = mask.bool().cuda()
    out = model(img.cuda(), aug.cuda(), mk, g["m"], False)
    assert len(out["vis_out"]) == 2
    lpix = sum(masked_pixel_mse(v, img.cuda(), mk[:, i]) for i, v in enumerate(out["vis_out"])) * 0.5
    loss = out["contra_loss"] * 0.1 + lpix
    loss.backward()
    # tolerances: the 192-wide tiny encoder with 70 % of BOTH views replaced by the mask token measured 1.2e-3 on the pixel loss (bf16
    # operands; the small / base fixtures of the README configuration stay within 1e-3 in test_step_matches_reference_golden); the
    # contrastive heads normalise over 16 rows per view here (small-batch BatchNorm amplification)
    assert float(lpix.detach()) == pytest.approx(g["loss_pixel"], rel=2e-3) and float(loss.detach()) == pytest.approx(g["loss"], rel=2e-3)
    assert float(out["contra_loss"].detach()) == pytest.approx(g["contra_loss"], rel=1e-2)
    for o, ref in zip(out["vis_out"], g["vis_out_all"]):
        assert o.shape == ref.shape and torch.allclose(o.detach().cpu(), ref, atol=3e-2)
    named = dict(model.named_parameters())
    for n in ("pix_decoder.4.weight", "pix_decoder.0.weight", "encoder.mask_token"):
        assert float(named[n].grad.norm()) == pytest.approx(g["grad_norms"][n], rel=6e-2), n
    tot = sum(float(p.grad.float().pow(2).sum()) for p in model.parameters() if p.grad is not None) ** 0.5
    assert tot == pytest.approx(sum(v ** 2 for v in g["grad_norms"].values()) ** 0.5, rel=8e-2)
    # each view separately against the fp32 oracle on this GPU (same weights, same inputs): neither view is systematically off
    sd = {k: v.detach().clone().float() for k, v in make(g["model"]).cuda().state_dict().items()}
    with torch.no_grad():
        _, o_ref, _ = R.step_losses(sd, img.cuda(), aug.cuda(), mask.bool().cuda(), g["m"], model.encoder.num_heads, only_mim_on_ori_img=False)
    for i in range(2):
        rel = float((out["vis_out"][i].detach() - o_ref["vis_out"][i]).norm() / o_ref["vis_out"][i].norm())
        assert rel < 2e-2, (i, rel)


@pytest.mark.parametrize("tag,tol_contra", [("small_b2", 5e-3), ("small_b8", 1e-3), ("base_b2", 5e-3)])
def test_step_matches_reference_golden(tag, tol_contra):
    g = torch.load(os.path.join(GOLD, "ref_step_%s.pt" % tag), weights_only=False)
    model = make(g["model"])
    out, lpix, loss = run_step(model, g)
    assert float(lpix) == pytest.approx(g["loss_pixel"], rel=1e-3)
    assert float(loss) == pytest.approx(g["loss"], rel=1e-3)
    assert float(out["contra_loss"]) == pytest.approx(g["contra_loss"], rel=tol_contra)
    assert torch.allclose(out["vis_out"][0].cpu(), g["vis_out"], atol=3e-2)
    assert out["vis_out"][0].shape == g["vis_out"].shape and out["contra_loss"].dim() == 0
    for k in ("q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5"):
        assert out[k].shape == (1,)
    # gradients reach ordinary leaf parameters; pixel-path dominated tensors agree closely, global norm within 5 %
    named = dict(model.named_parameters())
    for n in ("pix_decoder.4.weight", "pix_decoder.0.weight"):
        assert float(named[n].grad.norm()) == pytest.approx(g["grad_norms"][n], rel=3e-2), n
    tot = sum(float(p.grad.float().pow(2).sum()) for p in model.parameters() if p.grad is not None) ** 0.5
    ref_tot = sum(v ** 2 for v in g["grad_norms"].values()) ** 0.5
    assert tot == pytest.approx(ref_tot, rel=5e-2)
    assert all(p.grad is None for n, p in named.items() if not p.requires_grad)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for n, p in named.items() if p.requires_grad)
    # EMA (M:428-442) and BatchNorm running statistics
    sd = model.state_dict()
    for k, ref in g["momentum_after"].items():
        assert torch.allclose(sd[k].flatten()[:64].cpu(), ref, atol=1e-6), k
    for k, ref in g["bn_after"].items():
        assert torch.allclose(sd[k].cpu(), ref, atol=5e-3, rtol=5e-2), k


def test_pixel_path_gradients_match_oracle():
    """With the contrastive weight at 0 the gradient flows encoder -> pix_decoder only (no BatchNorm): every encoder tensor
    must agree with the fp32 oracle to bf16 accuracy (cosine > 0.999, relative L2 < 3e-2)."""
    from oracle import restatement as R
    from dig_b200.engine import masked_pixel_mse
    model = make("pretrain_simmim_moco_ori_vit_small_patch4_32x128")
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():      # non-trivial biases / mask token so that their gradients and uses are exercised
        for n, p in model.named_parameters():
            if p.requires_grad and (p.dim() == 1 or n.endswith("mask_token")):
                p.add_(torch.randn(p.shape, generator=gen) * 0.05)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    names = R.trainable_names(sd)
    for n in names:
        sd[n] = sd[n].requires_grad_(True)
    img, aug, mask = R.synthetic_batch(2, seed=1)
    loss_o, _, _ = R.step_losses(sd, img, aug, mask, 0.99, model.encoder.num_heads, w_contrast=0.0, w_pixel=1.0)
    grads_o = dict(zip(names, torch.autograd.grad(loss_o, [sd[n] for n in names], allow_unused=True)))
    mk = mask.clone()
    mk[:, 1, :] = False
    model.cuda()
    out = model(img.cuda(), aug.cuda(), mk.cuda(), 0.99, True)
    masked_pixel_mse(out["vis_out"][0], img.cuda(), mk[:, 0].cuda()).backward()
    checked = 0
    for n, p in model.named_parameters():
        go = grads_o.get(n)
        if go is None or float(go.norm()) == 0.0:
            continue
        gd = p.grad.float().cpu()
        rel = float((gd - go).norm() / go.norm())
        cos = float(torch.nn.functional.cosine_similarity(gd.flatten(), go.flatten(), dim=0))
        assert rel < 3e-2 and cos > 0.999, (n, rel, cos)
        checked += 1
    assert checked >= 150


def test_train_one_epoch_contract_and_learning():
    from dig_b200.engine import train_one_epoch
    from dig_b200.optim import FusedAdamW
    from dig_b200.utils import NativeScalerWithGradNormCount
    from bench import synthetic_batch
    model = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128").cuda()
    decay = [p for n, p in model.named_parameters() if p.requires_grad and p.dim() > 1]
    nodecay = [p for n, p in model.named_parameters() if p.requires_grad and p.dim() <= 1]
    opt = FusedAdamW([{"params": decay, "weight_decay": 0.05, "lr_scale": 1.0}, {"params": nodecay, "weight_decay": 0.0, "lr_scale": 1.0}], lr=1e-3)
    args = types.SimpleNamespace(num_view=2, moco_m=0.99, use_moco_m_cos=1, epochs=2, contrast_start_epoch=0, contrast_warmup_steps=2,
                                 loss_weight_contrast=0.1, loss_weight_pixel=1.0, only_mim_on_ori_img=True, eval_freq=10 ** 9, output_dir=None)
    batch = synthetic_batch(8, 3)
    loader = [([batch[0], batch[1], batch[2]], None, None)] * 12
    lr_sched = [1e-3] * 24
    stats0 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 0, NativeScalerWithGradNormCount(), max_norm=None,
                             patch_size=4, normlize_target=False, start_steps=0, lr_schedule_values=lr_sched, args=args)
    assert set(stats0) == {"lr", "min_lr", "moco_m", "loss_contrast", "q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5", "loss_pixel", "loss",
                           "loss_scale", "weight_decay", "grad_norm"}                                # E:204 meter set
    stats1 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 1, NativeScalerWithGradNormCount(), max_norm=3.0,
                             patch_size=4, normlize_target=False, start_steps=12, lr_schedule_values=lr_sched, args=args)
    assert stats1["loss_pixel"] < stats0["loss_pixel"]          # the same batch is being fitted
    assert all(v == v for v in stats1.values())
    # normlize_target=True (E:89-94): per-patch standardised pixel targets (unit scale instead of ~0.08: a larger, still falling loss)
    stats2 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 1, NativeScalerWithGradNormCount(), max_norm=None,
                             patch_size=4, normlize_target=True, start_steps=12, lr_schedule_values=lr_sched, args=args)
    stats3 = train_one_epoch(model, None, None, loader, None, opt, torch.device("cuda"), 1, NativeScalerWithGradNormCount(), max_norm=None,
                             patch_size=4, normlize_target=True, start_steps=12, lr_schedule_values=lr_sched, args=args)
    assert stats2["loss_pixel"] > 5 * stats1["loss_pixel"] and stats3["loss_pixel"] < stats2["loss_pixel"]
    assert all(v == v for v in stats3.values())


def test_state_dict_survives_device_move_and_reload():
    model = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    from bench import synthetic_batch
    img, aug, m = synthetic_batch(2, 1)
    mk = m.bool()
    mk[:, 1] = False
    model.cuda()
    with torch.no_grad():
        o1 = model(img.cuda(), aug.cuda(), mk.cuda(), 1.0, True)["contra_loss"].item()     # m = 1: momentum weights unchanged
    model2 = make("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    model2.load_state_dict(sd)
    model2.cuda()
    with torch.no_grad():
        o2 = model2(img.cuda(), aug.cuda(), mk.cuda(), 1.0, True)["contra_loss"].item()
    assert o1 == pytest.approx(o2, rel=2e-3)      # BatchNorm statistics are summed with atomics: run-to-run order differs


def test_two_stream_and_single_stream_steps_agree_and_stale_forward_is_refused(monkeypatch):
    """The two-stream schedule (momentum branch / weight gradients on the side stream) must not change results beyond the run-to-run
    noise floor of the single-stream schedule (fp32 atomics in the column sums and split-K accumulation are order-dependent, and the
    small-batch BatchNorm heads amplify that: scripts/stream_determinism.py); and a forward whose activations were overwritten by a later
    forward cannot be back-propagated."""
    from oracle import restatement as R
    from dig_b200.engine import masked_pixel_mse
    from dig_b200.ops import DigError
    img, aug, mask = R.synthetic_batch(8, seed=3)
    mk = mask.clone()
    mk[:, 1, :] = False
    img, aug, mk = img.cuda(), aug.cuda(), mk.cuda()

    def run(flag):
        monkeypatch.setenv("DIG_TWO_STREAMS", flag)
        model = make("pretrain_simmim_moco_ori_vit_small_patch4_32x128").cuda()
        out = model(img, aug, mk, 0.99, True)
        lpix = masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
        loss = out["contra_loss"] * 0.1 + lpix
        loss.backward()
        torch.cuda.synchronize()
        return model, float(loss), float(lpix), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}

    _, l0, p0, g0 = run("0")
    _, l0b, p0b, g0b = run("0")
    model, l1, p1, g1 = run("1")
    assert p1 == pytest.approx(p0, rel=1e-5)                                   # the pixel path has no BatchNorm: near bit-stable
    assert abs(l1 - l0) <= 3 * abs(l0b - l0) + 2e-4 * abs(l0)

    def rel(x, y):
        return float((x - y).norm()) / (float(y.norm()) + 1e-12)
    for n in g1:
        noise = rel(g0b[n], g0[n])
        assert rel(g1[n], g0[n]) <= 3 * noise + 2e-2, (n, rel(g1[n], g0[n]), noise)
    # stale forward
    out1 = model(img, aug, mk, 0.99, True)
    model(img, aug, mk, 0.99, True)
    with pytest.raises(DigError):
        out1["contra_loss"].backward()
