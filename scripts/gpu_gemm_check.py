"""GPU bring-up check for dig_gemm: every operand-major combination, tails, epilogues, split-K, timing."""
import sys, time
import torch
sys.path.insert(0, ".")
from dig_b200 import ops

torch.manual_seed(0)
dev = "cuda"
print("sm", ops.load().dig_sm())


def ref(a, b, amn, bmn):
    A = a.float().t() if amn else a.float()
    B = b.float() if bmn else b.float().t()
    return A @ B


def run(M, N, K, amn, bmn, out_dtype=torch.float32, split_k=1, **kw):
    a = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=out_dtype)
    ops.gemm(a, b, out, a_mn_major=amn, b_mn_major=bmn, split_k=split_k, **kw)
    torch.cuda.synchronize()
    r = ref(a, b, amn, bmn)
    err = (out.float() - r).abs().max().item()
    rel = err / r.abs().max().item()
    print("M%d N%d K%d amn%d bmn%d %s split%d: maxerr %.4g rel %.3g %s" % (M, N, K, amn, bmn, str(out_dtype)[6:], split_k, err, rel,
          "OK" if rel < (1e-2 if out_dtype == torch.bfloat16 else 1e-4) else "FAIL"))
    sys.stdout.flush()


for amn in (False, True):
    for bmn in (False, True):
        run(128, 128, 64, amn, bmn)
        run(256, 384, 384, amn, bmn)
        run(1024, 1152, 384, amn, bmn, torch.bfloat16)
        Mt = 360 if amn else 358           # MN-major A needs lda % 8 == 0
        run(Mt, 192, 384, amn, bmn)        # M tail, BN=64 path
        run(Mt, 48 + 16, 192, amn, bmn)
        run(384, 1152, 4096, amn, bmn, split_k=8)
run(65536, 1152, 384, False, False, torch.bfloat16)
run(1152, 384, 65536, True, True, split_k=16)

# epilogues
M, N, K = 512, 384, 384
a = torch.randn(M, K, device=dev).bfloat16(); b = (torch.randn(N, K, device=dev) * 0.05).bfloat16()
bias = torch.randn(N, device=dev); res = torch.randn(M, N, device=dev)
out = torch.empty(M, N, device=dev)
ops.gemm(a, b, out, bias=bias, residual=res)
r = a.float() @ b.float().t() + bias + res
print("bias+res", (out - r).abs().max().item())
aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16); outb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
ops.gemm(a, b, outb, bias=bias, epilogue=ops.EPI_GELU, aux=aux)
pre = a.float() @ b.float().t() + bias
print("gelu", (outb.float() - torch.nn.functional.gelu(pre)).abs().max().item(), "pre", (aux.float() - pre).abs().max().item())
bt = b.t().contiguous()
cs = torch.zeros(N, device=dev)
ops.gemm(a, bt, outb, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=aux, colsum=cs)
x = aux.float().requires_grad_(True); torch.nn.functional.gelu(x).sum().backward()
refb = (a.float() @ b.float().t()) * x.grad
print("gelu_bwd", (outb.float() - refb).abs().max().item(), "colsum rel", ((cs - refb.sum(0)).abs().max() / refb.sum(0).abs().max()).item())
mask = (torch.rand(M, device=dev) < 0.5).to(torch.uint8); mval = torch.randn(N, device=dev); pos = torch.randn(256, N, device=dev)
ops.gemm(a, b, out, bias=bias, residual=pos, res_row_mod=256, row_mask=mask, row_mask_value=mval)
r = torch.where(mask.bool()[:, None], mval[None, :].expand(M, N), a.float() @ b.float().t() + bias) + pos.repeat(M // 256, 1)
print("patch-epi", (out - r).abs().max().item())

# timing of the encoder shapes
def bench(M, N, K, amn=False, bmn=False, out_dtype=torch.bfloat16, split_k=1, iters=20, **kw):
    a = torch.randn((K, M) if amn else (M, K), device=dev).bfloat16()
    b = torch.randn((K, N) if bmn else (N, K), device=dev).bfloat16()
    out = torch.zeros(M, N, device=dev, dtype=out_dtype)
    for _ in range(3):
        ops.gemm(a, b, out, a_mn_major=amn, b_mn_major=bmn, split_k=split_k, **kw)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        ops.gemm(a, b, out, a_mn_major=amn, b_mn_major=bmn, split_k=split_k, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    A = a.t() if amn else a; B = b if bmn else b.t()
    for _ in range(3): torch.matmul(A, B)
    e0.record()
    for _ in range(iters): torch.matmul(A, B)
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    print("bench M%d N%d K%d amn%d bmn%d split%d: %.3f ms %.1f TFLOP/s | cublas %.3f ms %.1f TFLOP/s" % (
        M, N, K, amn, bmn, split_k, ms, 2e-9 * M * N * K / ms, ms2, 2e-9 * M * N * K / ms2))
    sys.stdout.flush()

bench(65536, 1152, 384)
bench(65536, 384, 384)
bench(65536, 1536, 384)
bench(65536, 384, 1536)
bench(65536, 384, 1152, bmn=True)
bench(1152, 384, 65536, True, True, torch.float32, split_k=16)
bench(1536, 384, 65536, True, True, torch.float32, split_k=12)
bench(1024, 4096, 4096)
