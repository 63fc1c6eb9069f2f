"""Run-to-run noise floor (single stream twice) vs two-stream schedule: loss and per-tensor gradient relative L2 differences."""
import os, sys, torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa
from dig_b200.engine import masked_pixel_mse
from oracle import restatement as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")
img, aug, mask = R.synthetic_batch(B, seed=3)
mk = mask.clone(); mk[:, 1, :] = False
img, aug, mk = img.cuda(), aug.cuda(), mk.cuda()
def run(flag):
    os.environ["DIG_TWO_STREAMS"] = flag
    torch.manual_seed(0)
    model = dig_b200.create_model("pretrain_simmim_moco_ori_vit_small_patch4_32x128", **KW).train().cuda()
    out = model(img, aug, mk, 0.99, True)
    lp = masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
    loss = out["contra_loss"] * 0.1 + lp
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), float(out["contra_loss"]), float(lp), {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
a, b, c = run("0"), run("0"), run("1")
def cmp(x, y, tag):
    worst = sorted(((float((x[3][n] - y[3][n]).norm()) / (float(y[3][n].norm()) + 1e-12), n) for n in x[3]), reverse=True)[:4]
    print("%s: loss %.3e contra %.3e pixel %.3e | worst grads %s" % (tag, abs(x[0] - y[0]) / abs(y[0]), abs(x[1] - y[1]) / abs(y[1]),
          abs(x[2] - y[2]) / abs(y[2]), ", ".join("%s %.2e" % (n.replace("encoder.", "e."), e) for e, n in worst)))
cmp(a, b, "single vs single")
cmp(c, a, "two-stream vs single")
cmp(c, b, "two-stream vs single'")
