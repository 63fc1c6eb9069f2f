"""One resident-input pre-training step under the CUDA profiler API (use with `ncu --profile-from-start off`)."""
import sys
import torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa
from dig_b200.engine import masked_pixel_mse
from dig_b200.optim import FusedAdamW
from dig_b200.utils import NativeScalerWithGradNormCount
from bench import synthetic_batch

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda", 0)
torch.manual_seed(0)
model = dig_b200.create_model("pretrain_simmim_moco_ori_vit_small_patch4_32x128", pretrained=False, drop_path_rate=0.0,
                              drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit", queue_size=65536,
                              patchnet_name="no_patchtrans").to(dev).train()
opt = FusedAdamW([{"params": [p for p in model.parameters() if p.requires_grad], "weight_decay": 0.05, "lr_scale": 1.0}], lr=1.5e-4)
scaler = NativeScalerWithGradNormCount()
img, aug, maskf = synthetic_batch(B, 1)
img_d, aug_d = img.to(dev), aug.to(dev)
mask_d = maskf.to(dev).flatten(1).to(torch.bool).view(B, 2, -1)
mask_d[:, 1, :] = False


def step():
    out = model(img_d, aug_d, mask_d, 0.99, True)
    lp = masked_pixel_mse(out["vis_out"][0], img_d, mask_d[:, 0])
    loss = out["contra_loss"] * 0.1 + lp
    opt.zero_grad()
    scaler(loss, opt, clip_grad=None, parameters=model.parameters())


for _ in range(warm):
    step()
torch.cuda.synchronize()
torch.cuda.profiler.start()
step()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
