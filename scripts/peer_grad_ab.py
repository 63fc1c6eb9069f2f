"""Micro-benchmark of the peer-memory gradient all-reduce (csrc/peer.cu) under torchrun: bandwidth of back-to-back exchanges of the
step's flat gradient buffer (43.6 M floats) for several grid sizes, NCCL's all_reduce of the same buffer beside it, and a correctness
check.  torchrun --nproc-per-node N scripts/peer_grad_ab.py"""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, ".")
from dig_b200 import ops, peer
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(os.environ.get("AB_N", 43_600_000)) // 4 * 4
comm = peer.get(dev, 1 << 20)
assert comm is not None, "peer workspaces unavailable"
buf, table = peer.shared_float_buffer(comm, n)
ref = torch.empty(n, device=dev)


def fill(seed):
    g = torch.Generator(device=dev); g.manual_seed(seed * 131 + rank)
    buf.copy_(torch.randn(n, device=dev, generator=g))


def peer_ar(blocks, small=0):
    ops.call("dig_peer_grad_allreduce", comm.bases, table, world, rank, peer.CH_GRADS, comm.next_epoch(peer.CH_GRADS), 0, n, blocks, small)


# correctness: average over ranks
fill(1)
ref.copy_(buf)
dist.all_reduce(ref, op=dist.ReduceOp.AVG)
peer_ar(0)
torch.cuda.synchronize()
err = float((buf - ref).abs().max())
chk = torch.tensor([float(buf.double().sum())], device=dev, dtype=torch.float64)
both = [torch.zeros_like(chk) for _ in range(world)]
dist.all_gather(both, chk)
if rank == 0:
    print("max |peer - nccl AVG| = %.3e; identical on every rank: %s" % (err, all(torch.equal(b, both[0]) for b in both)))


def timeit(name, f, iters=10):
    for _ in range(2): f()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        moved = 2.0 * (world - 1) / world * n * 4     # bytes each rank sends (and receives) per all-reduce
        print("%-44s %8.1f us   %6.0f GB/s per direction" % (name, float(t) * 1e3, moved / (float(t) * 1e-3) / 1e9)); sys.stdout.flush()


for blocks in (148, 64, 32, 16, 8):
    timeit("peer grad all-reduce, %3d blocks" % blocks, lambda b=blocks: peer_ar(b))
for blocks in (148, 296, 592):
    timeit("peer grad all-reduce, %3d small blocks (128 thr)" % blocks, lambda b=blocks: peer_ar(b, 1))
timeit("NCCL all_reduce (one call, 174 MB)", lambda: dist.all_reduce(buf, op=dist.ReduceOp.AVG))
seg = n // 14 // 4 * 4
timeit("NCCL all_reduce (14 segments)", lambda: [dist.all_reduce(buf[i * seg:(i + 1) * seg], op=dist.ReduceOp.AVG) for i in range(14)])
comm.check()
dist.destroy_process_group()
