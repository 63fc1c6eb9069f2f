"""Whole-step parity on the GPU: dig_b200 model (CUDA kernels) vs the fp32 CPU oracle (oracle/restatement.py)."""
import sys, time
import torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # registers factories
from oracle import restatement as R

name = sys.argv[1] if len(sys.argv) > 1 else "pretrain_simmim_moco_ori_vit_small_patch4_32x128"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
torch.manual_seed(0)
model = dig_b200.create_model(name, pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2,
                              num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
model.train()
heads = model.encoder.num_heads
# make biases / mask token / LN non-trivial so their gradients and uses are exercised
g = torch.Generator().manual_seed(5)
with torch.no_grad():
    for n, p in model.named_parameters():
        if p.requires_grad and (p.dim() == 1 or n.endswith("mask_token")):
            p.add_(torch.randn(p.shape, generator=g) * 0.05)
    model._init_momentum(model.encoder, model.momentum_encoder)
    model._init_momentum(model.encoder_projection_layer, model.momentum_projection_layer)
    model._init_momentum(model.pix_projector, model.pix_projector_m)
sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
img, aug, mask = R.synthetic_batch(B, seed=1)
m = 0.99
import os
WC = float(os.environ.get("WC", "0.1")); WP = float(os.environ.get("WP", "1.0"))

# ---- oracle (CPU fp32) ----
names = R.trainable_names(sd)
for n in names:
    sd[n] = sd[n].requires_grad_(True)
taps = {}
t0 = time.time()
loss_o, out_o, lpix_o = R.step_losses(sd, img, aug, mask, m, heads, w_contrast=WC, w_pixel=WP, taps=taps)
grads_o = torch.autograd.grad(loss_o, [sd[n] for n in names], allow_unused=True)
print("oracle: contra %.7f pixel %.7f total %.7f (%.1fs)" % (out_o["contra_loss"].item(), lpix_o.item(), loss_o.item(), time.time() - t0))

# ---- dig_b200 (GPU) ----
model.cuda()
mk = mask.clone(); mk[:, 1, :] = False
labels = R.build_targets(img, mk)[0].cuda()
out = model(img.cuda(), aug.cuda(), mk.cuda(), m)
lpix = torch.nn.functional.mse_loss(out["vis_out"][0], labels)
loss = out["contra_loss"] * WC + lpix * WP
loss.backward()
torch.cuda.synchronize()
print("dig   : contra %.7f pixel %.7f total %.7f" % (out["contra_loss"].item(), lpix.item(), loss.item()))
rel = lambda a, b: abs(a - b) / max(abs(b), 1e-12)
print("rel err: contra %.3e pixel %.3e total %.3e" % (rel(out["contra_loss"].item(), out_o["contra_loss"].item()),
      rel(lpix.item(), lpix_o.item()), rel(loss.item(), loss_o.item())))
print("acc dig", [out[k].item() for k in ("q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5")], "oracle",
      [out_o[k].item() for k in ("q1_acc1", "q1_acc5", "q2_acc1", "q2_acc5")])
vo = out_o["vis_out"][0]
print("vis_out max abs err %.4e (ref absmax %.3f)" % ((out["vis_out"][0].cpu() - vo).abs().max().item(), vo.abs().max().item()))
step = model._step
enc = step.bufs.d["o.x12"].cpu().view(2 * B, 256, -1)
print("encoder out: max abs err %.4e, mean|x| %.4f" % ((enc - taps["enc"]).abs().max().item(), taps["enc"].abs().mean().item()))
# momentum parameter after EMA and BN buffers
msd = model.state_dict()
worst = 0
for k in sd:
    if k.startswith(("momentum_", "pix_projector_m")) or R.is_buffer(k):
        e = (msd[k].float().cpu() - sd[k].detach().float()).abs().max().item()
        s_ = sd[k].detach().float().abs().max().item()
        if e / max(s_, 1e-6) > 1e-3:
            print("   buffer/ema mismatch", k, e, s_)
        worst = max(worst, e / max(s_, 1e-6))
print("EMA params / BN buffers worst rel err %.3e" % worst)
bad = 0
rows = []
for n, go in zip(names, grads_o):
    p = dict(model.named_parameters())[n]
    gd = p.grad
    if go is None:
        go = torch.zeros_like(sd[n])
    if gd is None:
        print("MISSING grad", n); bad += 1; continue
    gd = gd.float().cpu()
    num = (gd - go).norm().item(); den = go.norm().item()
    cos = torch.nn.functional.cosine_similarity(gd.flatten(), go.flatten(), dim=0).item() if den > 0 else 1.0
    rows.append((num / max(den, 1e-12), cos, n, den))
rows.sort(reverse=True)
for r in rows[:25]:
    print("grad relL2 %.3e cos %.5f |g| %.3e  %s" % (r[0], r[1], r[3], r[2]))
import statistics
print("grad relL2 median %.3e ; min cos %.5f ; n=%d" % (statistics.median(r[0] for r in rows), min(r[1] for r in rows), len(rows)))
tot_o = torch.sqrt(sum((g_ ** 2).sum() for g_ in grads_o if g_ is not None)).item()
tot_d = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in model.parameters() if p.grad is not None)).item()
print("grad norm oracle %.6f dig %.6f" % (tot_o, tot_d))
