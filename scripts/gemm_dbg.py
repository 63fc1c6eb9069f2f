import sys, torch, ctypes
sys.path.insert(0, ".")
from dig_b200 import ops
def bench(M,N,K,mode,iters=20):
    a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    aux = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.gemm(a, b, out, epilogue=mode, aux=aux if mode else None)
    for _ in range(3): f()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)/iters
    print("M%d N%d K%d mode 0x%x: %.3f ms  %.0f TFLOP/s" % (M,N,K,mode,ms,2e-9*M*N*K/ms)); sys.stdout.flush()
for shape in ((65536,1152,384),(65536,384,1536),(65536,1536,384)):
    for mode in (0, 0x100, 0x200, 0x300):
        bench(*shape, mode)
