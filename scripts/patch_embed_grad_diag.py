"""Where does the patch-embed weight-gradient error come from (cos 0.983 vs the fp32 oracle at bs=128)?  Recompute it in fp32 torch from
dig_b200's own fp32 residual gradient and the fp32 images, and from bf16-rounded operands, and compare all with the oracle."""
import sys, torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa
from dig_b200.engine import masked_pixel_mse
from oracle import restatement as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")
torch.manual_seed(0)
model = dig_b200.create_model("pretrain_simmim_moco_ori_vit_small_patch4_32x128", **KW).train()
sd = {k: v.detach().clone().cuda() for k, v in model.state_dict().items()}
names = R.trainable_names(sd)
for n in names: sd[n].requires_grad_(True)
img, aug, mask = [t.cuda() for t in R.synthetic_batch(B, seed=1)]
taps = []
loss_o, out_o, _ = R.step_losses(sd, img, aug, mask, 0.99, 6)
go = torch.autograd.grad(loss_o, sd["encoder.patch_embed.proj.weight"])[0].reshape(384, 48)
mk = mask.clone(); mk[:, 1, :] = False
model.cuda()
out = model(img, aug, mk, 0.99, True)
loss = out["contra_loss"] * 0.1 + masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
loss.backward(); torch.cuda.synchronize()
gd = model.encoder.patch_embed.proj.weight.grad.reshape(384, 48).float()
st = model._step
g = st.bufs.d["bw.g"].clone()                    # fp32 gradient w.r.t. the patch-embed output (after the mask-token mix)
m8 = st.bufs.d["mask"].bool()
gz = g.clone(); gz[m8] = 0
allimg = torch.cat([img, aug])
cols = torch.nn.functional.unfold(allimg, 4, stride=4).transpose(1, 2).reshape(-1, 48)      # (c, kh, kw) order
def rel(a, b): return float((a - b).norm() / b.norm()), float(torch.nn.functional.cosine_similarity(a.flatten(), b.flatten(), dim=0))
print("dig kernel grad vs oracle           rel %.4f cos %.5f" % rel(gd, go))
print("fp32 GEMM of dig's fp32 g vs oracle rel %.4f cos %.5f" % rel(gz.t() @ cols, go))
print("bf16-rounded operands, fp32 acc     rel %.4f cos %.5f" % rel(gz.bfloat16().float().t() @ cols.bfloat16().float(), go))
print("|grad| oracle %.4e; sum_t |g_t||x_t| scale %.4e" % (float(go.norm()), float((gz.norm(dim=1) * cols.norm(dim=1)).sum())))
