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
    print("%-40s %.3f ms  %.0f TFLOP/s" % (name, ms, flops*1e-9/ms)); sys.stdout.flush()
M, N, K = 65536, 1536, 384
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(K, N, device="cuda")*0.05).bfloat16()  # [K,N] MN-major B
wk = w.t().contiguous()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); aux = torch.randn(M, N, device="cuda").bfloat16()
cs = torch.zeros(N, device="cuda"); bias = torch.randn(N, device="cuda")
fl = 2.0*M*N*K
bench("dgrad LINEAR bf16", lambda: ops.gemm(a, w, out, b_mn_major=True), fl)
bench("dgrad GELU_BWD", lambda: ops.gemm(a, w, out, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=aux), fl)
bench("dgrad GELU_BWD + colsum", lambda: ops.gemm(a, w, out, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=aux, colsum=cs), fl)
bench("fwd LINEAR bf16 + bias", lambda: ops.gemm(a, wk, out, bias=bias), fl)
bench("fwd GELU + bias", lambda: ops.gemm(a, wk, out, bias=bias, epilogue=ops.EPI_GELU, aux=aux), fl)
o32 = torch.empty(M, K, device="cuda"); res = torch.randn(M, K, device="cuda"); h = torch.randn(M, N, device="cuda").bfloat16(); w2 = torch.randn(K, N, device="cuda").bfloat16()
bench("fc2 fwd fp32 out + bias + residual", lambda: ops.gemm(h, w2, o32, bias=bias[:K].contiguous(), residual=res), fl)
