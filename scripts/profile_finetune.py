"""CUPTI kernel table of the fine-tuning step (BASELINE configs[4], bs=256): time per kernel, 3 profiled steps."""
import collections, sys, torch
sys.path.insert(0, ".")
from torch.profiler import ProfilerActivity, profile
from bench import FinetuneWorkload
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
dev = torch.device("cuda", 0)
wl = FinetuneWorkload(B, dev, 0, 1)
for _ in range(4): wl.step()
torch.cuda.synchronize()
N = 3
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(N): wl.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
ev.sort(key=lambda e: e.time_range.start)
span = (ev[-1].time_range.end - ev[0].time_range.start) / N
print("kernels/step %d  span %.3f ms" % (len(ev) // N, span / 1e3))
agg = collections.defaultdict(lambda: [0, 0.0])
for e in ev:
    a = agg[e.name[:70]]; a[0] += 1; a[1] += e.time_range.end - e.time_range.start
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print("  %8.3f ms %4d x %7.1f us  %s" % (t / N / 1e3, n // N, t / n, k))
