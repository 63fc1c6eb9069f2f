"""GPU idle time inside a step: torch.profiler (CUPTI) kernel timeline of 3 resident-input steps -> busy vs span, idle by gap size."""
import sys, torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa
from dig_b200.engine import masked_pixel_mse
from dig_b200.optim import FusedAdamW
from dig_b200.utils import NativeScalerWithGradNormCount
from bench import synthetic_batch
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
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
for _ in range(4): step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N): step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
span = (ev[-1].time_range.end - ev[0].time_range.start) / N
busy = sum(e.time_range.end - e.time_range.start for e in ev) / N
gaps = [ev[i + 1].time_range.start - ev[i].time_range.end for i in range(len(ev) - 1)]
print("kernels/step %d  span %.3f ms  busy %.3f ms  idle %.3f ms" % (len(ev) // N, span / 1e3, busy / 1e3, (span - busy) / 1e3))
for lo, hi in ((0, 2), (2, 5), (5, 10), (10, 20), (20, 50), (50, 1e9)):
    g = [x for x in gaps if lo <= x < hi]
    print("  gaps %3g-%-5g us: %5d  total %.3f ms/step" % (lo, hi, len(g) // N, sum(g) / N / 1e3))
import collections
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:60]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:22]:
    print("  %8.3f ms %4d x %7.1f us  %s" % (t / N / 1e3, n // N, t / n, k))
print("gaps > 20 us (previous kernel -> next kernel), last profiled step:")
start = len(ev) - len(ev) // N
for i in range(start, len(ev) - 1):
    g = ev[i + 1].time_range.start - ev[i].time_range.end
    if g > 20:
        print("  %7.1f us  after #%d %s  -> %s" % (g, i - start, ev[i].name[:50], ev[i + 1].name[:50]))
