"""CUPTI kernel timeline of the data-parallel step under torchrun (rank 0 reports): where do the extra milliseconds of the N-GPU step go?
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/timeline_multi.py [batch]
Prints span / busy / idle per step, time per kernel family (NCCL kernels, the NVLink peer-memory exchange kernels of csrc/peer.cu, GEMMs,
attention, ...), and the longest waits.  Run it at N=1 as well (plain python) for the baseline column."""
import collections
import os
import sys

import torch
import torch.distributed as dist
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, ".")
from bench import MODEL, Workload, sync_all  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
wl = Workload(MODEL, B, dev, rank, world, os.environ.get("DIG_BENCH_DDP", "dig"))
for _ in range(5):
    wl.step()
sync_all(world)
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N):
        wl.step()
    torch.cuda.synchronize()
sync_all(world)
if rank == 0:
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
    ev.sort(key=lambda e: e.time_range.start)
    span = (ev[-1].time_range.end - ev[0].time_range.start) / N
    # union of busy intervals (two streams overlap)
    busy, cur_end = 0.0, None
    for e in ev:
        s, t = e.time_range.start, e.time_range.end
        if cur_end is None or s > cur_end:
            busy += t - s
            cur_end = t
        elif t > cur_end:
            busy += t - cur_end
            cur_end = t
    busy /= N
    print("world %d  peer-memory exchanges %s  kernels/step %d  span %.3f ms  GPU busy (union of streams) %.3f ms  idle %.3f ms" % (
        world, "on" if (os.environ.get("DIG_PEER", "1") != "0" and world > 1) else "off", len(ev) // N, span / 1e3, busy / 1e3, (span - busy) / 1e3))

    def family(n):
        if "nccl" in n.lower():
            return "NCCL: " + n[:60]
        if "peer_" in n or "colsum_kernel<float, (bool)1, (bool)1>" in n or "bn_bwd_stats_kernel<(bool)1>" in n:
            return "peer exchange (csrc/peer.cu): " + n[:60]
        if "gemm" in n:
            return "tcgen05 GEMMs"
        if "attn" in n:
            return "attention"
        if "layernorm" in n:
            return "LayerNorm"
        if "mt_" in n:
            return "multi-tensor optimizer / EMA"
        if "dig::" in n:
            return "other dig kernels"
        return "torch: " + n[:50]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in ev:
        a = agg[family(e.name)]
        a[0] += 1
        a[1] += e.time_range.end - e.time_range.start
    print("time per family (sum of kernel durations per step; streams overlap, so the sum exceeds the span):")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:24]:
        print("  %8.3f ms %5d x %8.1f us  %s" % (t / N / 1e3, n // N, t / n, k))
if world > 1:
    dist.destroy_process_group()
