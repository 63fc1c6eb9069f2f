"""The CPU oracle (oracle/restatement.py) against the reference: golden fixtures written by oracle/make_golden.py from the
unmodified reference, and -- where /root/reference exists -- the live reference itself."""
import os

import pytest
import torch

from oracle import restatement as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")


def holder_state(name):
    import dig_b200
    from dig_b200 import modeling  # noqa: F401
    torch.manual_seed(0)
    m = dig_b200.create_model(name, **KW)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}, m.encoder.num_heads


@pytest.mark.parametrize("tag", ["small_b2", "small_b8", "base_b2"])
def test_oracle_reproduces_reference_golden(tag):
    g = torch.load(os.path.join(GOLD, "ref_step_%s.pt" % tag), weights_only=False)
    sd, heads = holder_state(g["model"])
    names = R.trainable_names(sd)
    for n in names:
        sd[n] = sd[n].requires_grad_(True)
    img, aug, mask = R.synthetic_batch(g["B"], seed=g["seed_data"])
    loss, out, lpix = R.step_losses(sd, img, aug, mask, g["m"], heads)
    assert float(out["contra_loss"]) == pytest.approx(g["contra_loss"], rel=2e-5)
    assert float(lpix) == pytest.approx(g["loss_pixel"], rel=2e-5)
    assert float(loss) == pytest.approx(g["loss"], rel=2e-5)
    assert [float(out[k]) for k in ("q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5")] == g["accs"]
    assert torch.allclose(out["vis_out"][0], g["vis_out"], atol=2e-5, rtol=1e-4)
    grads = torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)
    gd = dict(zip(names, grads))
    for n, ref in g["grad_norms"].items():
        assert float(gd[n].norm()) == pytest.approx(ref, rel=5e-3, abs=1e-7), n
    for n, ref in g["grad_samples"].items():
        assert torch.allclose(gd[n].flatten()[:64], ref, rtol=2e-2, atol=1e-6 + 1e-3 * float(ref.abs().max())), n
    for k, ref in g["momentum_after"].items():
        assert torch.allclose(sd[k].detach().flatten()[:64], ref, rtol=1e-5, atol=1e-7), k
    for k, ref in g["bn_after"].items():
        assert torch.allclose(sd[k].detach(), ref, rtol=1e-3, atol=1e-5), k


def test_oracle_reproduces_reference_golden_with_both_views_masked():
    """--only_mim_on_ori_img 0 (M:571-575, E:137-141): the pixel head on both views, both compared with the original image's patches."""
    g = torch.load(os.path.join(GOLD, "ref_step_tiny_b4_bothviews.pt"), weights_only=False)
    assert g["only_mim"] is False and len(g["vis_out_all"]) == 2
    sd, heads = holder_state(g["model"])
    names = R.trainable_names(sd)
    for n in names:
        sd[n] = sd[n].requires_grad_(True)
    img, aug, mask = R.synthetic_batch(g["B"], seed=g["seed_data"])
    loss, out, lpix = R.step_losses(sd, img, aug, mask, g["m"], heads, only_mim_on_ori_img=False)
    assert float(out["contra_loss"]) == pytest.approx(g["contra_loss"], rel=2e-5)
    assert float(lpix) == pytest.approx(g["loss_pixel"], rel=2e-5) and float(loss) == pytest.approx(g["loss"], rel=2e-5)
    for o, ref in zip(out["vis_out"], g["vis_out_all"]):
        assert torch.allclose(o, ref, atol=2e-5, rtol=1e-4)
    gd = dict(zip(names, torch.autograd.grad(loss, [sd[n] for n in names], allow_unused=True)))
    for n, ref in g["grad_norms"].items():
        assert float(gd[n].norm()) == pytest.approx(ref, rel=5e-3, abs=1e-7), n


def test_known_answer_of_survey():
    """contra 1.6791600, pixel 0.3883830, total 0.5562990 for ViT-S, B=2, seeds (0, 1), m=0.99."""
    g = torch.load(os.path.join(GOLD, "ref_step_small_b2.pt"), weights_only=False)
    assert g["contra_loss"] == pytest.approx(1.6791600, abs=2e-6)
    assert g["loss_pixel"] == pytest.approx(0.3883830, abs=2e-6)
    assert g["loss"] == pytest.approx(0.5562990, abs=2e-6)
    assert g["accs"] == [12.5, 50.0, 0.0, 87.5]


def test_decoder_on_gathered_rows_equals_reference_order():
    """The product path decodes only the masked rows; the reference decodes all rows, then gathers (M:561-570)."""
    sd, heads = holder_state("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128")
    x = torch.randn(2, 256, 192)
    mask = torch.zeros(2, 256, dtype=torch.bool)
    mask[:, ::3] = True
    full = R.pix_decoder(sd, x)[mask]
    first = R.pix_decoder(sd, x[mask])
    assert torch.allclose(full, first, atol=1e-6)


def test_normalised_targets_equal_the_reference_statement():
    """E:89-94 (`normlize_target=True`) written literally with einops, against R.build_targets(normalize_target=True)."""
    from einops import rearrange
    torch.manual_seed(4)
    images = torch.rand(3, 3, 32, 128) * 2 - 1
    mask = torch.zeros(3, 2, 256, dtype=torch.bool)
    for b in range(3):
        mask[b, 0, torch.randperm(256)[:179]] = True
    unnorm = images * 0.5 + 0.5
    sq = rearrange(unnorm, "b c (h p1) (w p2) -> b (h w) (p1 p2) c", p1=4, p2=4)
    norm = (sq - sq.mean(dim=-2, keepdim=True)) / (sq.var(dim=-2, unbiased=True, keepdim=True).sqrt() + 1e-6)
    want = rearrange(norm, "b n p c -> b n (p c)")[mask[:, 0]].reshape(3, -1, 48)
    got = R.build_targets(images, mask, True, normalize_target=True)[0]
    assert torch.equal(got, want)
    plain = rearrange(unnorm, "b c (h p1) (w p2) -> b (h w) (p1 p2 c)", p1=4, p2=4)[mask[:, 0]].reshape(3, -1, 48)
    assert torch.equal(R.build_targets(images, mask, True)[0], plain)


def test_adamw_matches_torch():
    """custom_optim/_functional.py:115-140 restated == torch.optim.AdamW (same decoupled rule)."""
    torch.manual_seed(0)
    p = torch.randn(37, 5)
    g = torch.randn(37, 5)
    q = p.clone().requires_grad_(True)
    opt = torch.optim.AdamW([q], lr=1e-2, weight_decay=0.1, betas=(0.9, 0.999), eps=1e-8)
    ea, eas = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(1, 4):
        q.grad = g.clone()
        opt.step()
        R.adamw_step(p, g, ea, eas, t, 1e-2, 0.1)
    assert torch.allclose(p, q.detach(), atol=1e-6)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference not present on this machine")
def test_oracle_against_live_reference_tiny():
    from oracle import ref_shims
    ref_shims.ensure_cpu_process_group()
    model = ref_shims.create_reference_model("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", seed=3)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    img, aug, mask = R.synthetic_batch(3, seed=7)
    mk = mask.clone()
    mk[:, 1, :] = False
    with ref_shims.cpu_patches(), torch.no_grad():
        out = model(img, aug, mk, 0.9, True)
    with torch.no_grad():
        o = R.moco_vit_forward(sd, img, aug, mk, 0.9, 3)
    assert float(o["contra_loss"]) == pytest.approx(float(out["contra_loss"]), rel=1e-5)
    assert torch.allclose(o["vis_out"][0], out["vis_out"][0], atol=1e-5)
    for k, v in model.state_dict().items():     # EMA parameters and BN running statistics after the forward
        assert torch.allclose(sd[k].float(), v.float(), atol=1e-5, rtol=1e-4), k
