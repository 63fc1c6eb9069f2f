"""Context number (not a bench line): the oracle port of the reference step (oracle/restatement.py, plain PyTorch ops, the reference's own
module structure) run EAGERLY on the B200 under bf16 autocast -- i.e. torch 2.11 + cuBLASLt/ATen sm_100 kernels, the practical bar SURVEY.md
names for a reference that ships no kernels of its own.  Same synthetic batch, forward + backward + grad-norm + AdamW."""
import sys, time, torch
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa: F401 (parameter holder: the reference's init and state-dict keys)
from oracle import restatement as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
torch.manual_seed(0)
model = dig_b200.create_model("pretrain_simmim_moco_ori_vit_small_patch4_32x128", pretrained=False, drop_path_rate=0.0, drop_block_rate=None,
                              mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit", queue_size=65536, patchnet_name="no_patchtrans")
sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
heads = model.encoder.num_heads
del model
tr = R.OracleTrainer(sd, heads, lr=1.5e-4 * B / 256, weight_decay=0.05)
img, aug, mask = R.synthetic_batch(B, seed=1)
img, aug, mask = img.to(dev), aug.to(dev), mask.to(dev).bool()
for dtype in (torch.bfloat16, None):
    def one():
        if dtype is None:
            return tr.step(img, aug, mask, 0.99)
        with torch.autocast("cuda", dtype=dtype):
            return tr.step(img, aug, mask, 0.99)
    for _ in range(3): one()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps): st = one()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    print("torch eager %s: %.2f ms/step  %.0f crops/s  (loss %.4f)  peak mem %.1f GB" % (
        "bf16 autocast" if dtype is not None else "fp32 (TF32 off)", dt * 1e3, B / dt, st[0]["loss"], torch.cuda.max_memory_allocated() / 2**30))
    sys.stdout.flush()
