"""torchrun --nproc-per-node 2 scripts/dp_check.py: DigDataParallel vs torch DistributedDataParallel on one step -- gradients must be
identical across ranks (every segment was averaged) and agree with DDP's up to the run-to-run noise floor (see stream_determinism.py)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
import dig_b200
from dig_b200 import modeling  # noqa
from dig_b200.engine import masked_pixel_mse
from dig_b200.parallel import DigDataParallel
from oracle import restatement as R
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local))
KW = dict(pretrained=False, drop_path_rate=0.0, drop_block_rate=None, mlp_dim=4096, dim=256, T=0.2, num_windows=4, encoder_type="vit",
          queue_size=65536, patchnet_name="no_patchtrans")
img, aug, mask = R.synthetic_batch(8, seed=10 + rank)
mk = mask.clone(); mk[:, 1, :] = False
img, aug, mk = img.cuda(), aug.cuda(), mk.cuda()
def run(kind):
    torch.manual_seed(0)
    m = dig_b200.create_model("pretrain_simmim_moco_ori_vit_small_patch4_32x128", **KW).train().cuda()
    net = torch.nn.SyncBatchNorm.convert_sync_batchnorm(m)
    net = DigDataParallel(net) if kind == "dig" else torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], find_unused_parameters=True)
    out = net(img, aug, mk, 0.99, True)
    loss = out["contra_loss"] * 0.1 + masked_pixel_mse(out["vis_out"][0], img, mk[:, 0])
    loss.backward()
    torch.cuda.synchronize()
    return float(loss), {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
la, ga = run("dig")
lb, gb = run("torch")
flat = torch.cat([g.flatten() for g in ga.values()])
other = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(other, flat)
cross = max(float((o - other[0]).abs().max()) for o in other)
worst = sorted(((float((ga[n] - gb[n]).norm()) / (float(gb[n].norm()) + 1e-12), n) for n in ga), reverse=True)[:3]
tot = float((flat - torch.cat([g.flatten() for g in gb.values()])).norm() / torch.cat([g.flatten() for g in gb.values()]).norm())
if rank == 0:
    print("loss dig %.6f torch-DDP %.6f | max |grad_rank_i - grad_rank_0| = %.3e (must be 0) | rel L2 vs DDP: total %.3e, worst %s" % (
        la, lb, cross, tot, ", ".join("%s %.2e" % (n, e) for e, n in worst)))
dist.destroy_process_group()
