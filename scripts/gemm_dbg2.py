import sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
def bench(name, f, flops, iters=20):
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/iters
    print("%-44s %.3f ms  %.0f TFLOP/s" % (name, ms, flops*1e-9/ms)); sys.stdout.flush()
for (M,N,K) in ((65536,1152,384),(65536,1536,384),(65536,384,1536),(65536,384,384)):
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda")*0.05).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bench("fwd LINEAR bf16 %dx%dx%d" % (M,N,K), lambda: ops.gemm(a, w, out), 2.0*M*N*K)
