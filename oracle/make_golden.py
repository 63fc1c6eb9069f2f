"""TEST INFRASTRUCTURE ONLY -- writes golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the authoring container (needs /root/reference):  python oracle/make_golden.py
The reference's `MoCo_ViT.forward` + the engine's target/loss lines are executed on CPU in fp32 under the import
stubs of oracle/ref_shims.py, for BASELINE config 1 (ViT-S, B=2), a B=8 case and the d=512 "base" variant.  Fixtures
are small (.pt, a few hundred KB) and are what the `-m "not gpu"` oracle tests and the `-m gpu` parity tests compare
against on machines where /root/reference does not exist.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402
from oracle import restatement as R  # noqa: E402


def reference_step(model_name, B, seed_model=0, seed_data=1, m=0.99, w_contrast=0.1, w_pixel=1.0, only_mim=True):
    """only_mim=False: --only_mim_on_ori_img 0 -- the mask of the second view stays, the pixel head is applied to both views and BOTH are
    compared with patches of the ORIGINAL image (E:83-111 builds every view's labels from `images`), each weighted 1/num_view (E:137-141)."""
    ref_shims.ensure_cpu_process_group()
    model = ref_shims.create_reference_model(model_name, seed=seed_model)
    sd0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    img, aug, mask = R.synthetic_batch(B, seed=seed_data)
    mk = mask.clone()
    if only_mim:
        mk[:, 1, :] = False                                              # engine_for_pretraining_moco.py:103-104
    labels = R.build_targets(img, mk, only_mim)
    with ref_shims.cpu_patches():
        out = model(img, aug, mk, m, only_mim)
    loss_pixel = sum(torch.nn.functional.mse_loss(o, l) for o, l in zip(out["vis_out"], labels)) / len(labels)
    loss = out["contra_loss"] * w_contrast + loss_pixel * w_pixel
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    sd1 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return dict(model=model_name, B=B, m=m, seed_model=seed_model, seed_data=seed_data,
                contra_loss=float(out["contra_loss"]), loss_pixel=float(loss_pixel), loss=float(loss),
                accs=[float(out[k]) for k in ("q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5")],
                vis_out=out["vis_out"][0].detach().clone(), vis_out_all=[o.detach().clone() for o in out["vis_out"]], only_mim=only_mim,
                state_keys=[(k, tuple(v.shape), str(v.dtype)) for k, v in sd0.items()],
                grad_norms={n: float(g.norm()) for n, g in grads.items()},
                grad_samples={n: grads[n].flatten()[:64].clone() for n in
                              ("encoder.blocks.0.attn.qkv.weight", "encoder.blocks.11.mlp.fc2.weight", "encoder.patch_embed.proj.weight",
                               "pix_decoder.4.weight", "predictor.3.weight", "encoder_projection_layer.0.weight", "encoder.mask_token")},
                momentum_after={k: sd1[k].flatten()[:64].clone() for k in
                                ("momentum_encoder.blocks.3.mlp.fc1.weight", "momentum_projection_layer.0.weight", "pix_projector_m.3.weight")},
                bn_after={k: sd1[k].clone() for k in ("predictor.1.running_mean", "predictor.1.running_var", "pix_projector.7.running_var",
                                                     "momentum_projection_layer.7.running_mean")},
                param_checksum={k: float(v.double().sum()) for k, v in sd0.items() if v.dtype.is_floating_point},
                trainable=sum(p.numel() for p in model.parameters() if p.requires_grad),
                frozen=sum(p.numel() for p in model.parameters() if not p.requires_grad))


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for name, B, tag in (("pretrain_simmim_moco_ori_vit_small_patch4_32x128", 2, "small_b2"),
                         ("pretrain_simmim_moco_ori_vit_small_patch4_32x128", 8, "small_b8"),
                         ("pretrain_simmim_moco_ori_vit_base_patch4_32x128", 2, "base_b2")):
        g = reference_step(name, B)
        torch.save(g, os.path.join(out_dir, "ref_step_%s.pt" % tag))
        print(tag, "contra %.7f pixel %.7f total %.7f accs %s" % (g["contra_loss"], g["loss_pixel"], g["loss"], g["accs"]))
    # --only_mim_on_ori_img 0 (M:571-575, E:137-141): masked-pixel head on both views
    g = reference_step("pretrain_simmim_moco_ori_vit_tiny_patch4_32x128", 4, only_mim=False)
    torch.save(g, os.path.join(out_dir, "ref_step_tiny_b4_bothviews.pt"))
    print("tiny_b4_bothviews", "contra %.7f pixel %.7f total %.7f" % (g["contra_loss"], g["loss_pixel"], g["loss"]))


if __name__ == "__main__":
    main()
