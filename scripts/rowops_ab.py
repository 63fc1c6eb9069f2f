"""LayerNorm forward / backward micro-benchmark at the bs=128 step's shape (65 536 tokens x 384), three rotating buffer sets (working set
exceeds L2), correctness of the bf16-residual backward against torch autograd on a slice.  Use with DIG_B200_LIB for A/B comparisons."""
import os, sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
call = ops.call
M, d = int(os.environ.get("AB_M", 65536)), int(os.environ.get("AB_D", 384))
torch.manual_seed(0)
dev = "cuda"
sets = []
for _ in range(3):
    x = torch.randn(M, d, device=dev) * 2 + 0.3
    sets.append(dict(x=x, y=torch.empty(M, d, device=dev, dtype=torch.bfloat16), mean=torch.empty(M, device=dev), rstd=torch.empty(M, device=dev),
                     dy=torch.randn(M, d, device=dev).bfloat16(), gres=torch.randn(M, d, device=dev).bfloat16(),
                     gout=torch.empty(M, d, device=dev, dtype=torch.bfloat16)))
gamma = torch.rand(d, device=dev) + 0.5
beta = torch.randn(d, device=dev) * 0.1
dgamma = torch.zeros(d, device=dev); dbeta = torch.zeros(d, device=dev); dxsum = torch.zeros(d, device=dev)


def timeit(name, f, nbytes, iters=30):
    for i in range(3): f(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(iters): f(i)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    print("%-34s %8.1f us  %6.0f GB/s" % (name, best * 1e3, nbytes * 1e-6 / best)); sys.stdout.flush()


def fwd(i):
    s = sets[i % 3]
    call("dig_layernorm_fwd", s["x"], gamma, beta, s["y"], s["mean"], s["rstd"], M, d, 1e-5, 0)


def bwd(i):
    s = sets[i % 3]
    call("dig_layernorm_bwd_bf16res", s["dy"], s["x"], s["mean"], s["rstd"], gamma, s["gres"], s["gout"], dgamma, dbeta, dxsum, M, d)


timeit("layernorm fwd", fwd, M * d * 6.0)
for i in range(3): fwd(i)
timeit("layernorm bwd (bf16 residual)", bwd, M * d * 10.0)

# correctness on the first 2048 rows of set 0
s = sets[0]
n = 2048
dgamma.zero_(); dbeta.zero_(); dxsum.zero_()
call("dig_layernorm_bwd_bf16res", s["dy"], s["x"], s["mean"], s["rstd"], gamma, s["gres"], s["gout"], dgamma, dbeta, dxsum, n, d)
xr = s["x"][:n].clone().requires_grad_(True)
g = gamma.clone().requires_grad_(True); b = beta.clone().requires_grad_(True)
y = torch.nn.functional.layer_norm(xr, (d,), g, b, 1e-5)
y.backward(s["dy"][:n].float())
ref = xr.grad + s["gres"][:n].float()
print("bwd dx max err %.4f (ref absmax %.2f)  dgamma rel %.2e  dbeta rel %.2e  dxsum rel %.2e" % (
    (s["gout"][:n].float() - ref).abs().max().item(), ref.abs().max().item(),
    ((dgamma - g.grad).norm() / g.grad.norm()).item(), ((dbeta - b.grad).norm() / b.grad.norm()).item(),
    ((dxsum - ref.sum(0)).norm() / ref.sum(0).norm()).item()))
