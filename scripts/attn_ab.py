"""Attention forward/backward micro-benchmark at the bs=128 step's shape (S=256 sequences x 6 heads); correctness vs torch on a slice."""
import os, sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
S, h = int(os.environ.get("AB_S", 256)), 6
d, scale = h * 64, 64 ** -0.5
torch.manual_seed(0)
sets = []
for _ in range(3):
    qkv = (torch.randn(S * 256, 3 * d, device="cuda") * 1.5).bfloat16()
    sets.append((qkv, torch.empty(S * 256, d, device="cuda", dtype=torch.bfloat16), torch.empty(S, h, 256, device="cuda"),
                 torch.randn(S * 256, d, device="cuda").bfloat16(), torch.empty_like(qkv)))
def timeit(name, f, flops, iters=18):
    for i in range(3): f(i)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): f(i)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    for _ in range(2):   # best of 3 timed rounds
        e0.record()
        for i in range(iters): f(i)
        e1.record(); torch.cuda.synchronize()
        ms = min(ms, e0.elapsed_time(e1) / iters)
    print("%-28s %8.1f us  %6.0f TFLOP/s" % (name, ms * 1e3, flops * 1e-9 / ms)); sys.stdout.flush()
def fwd(i):
    q, o, l, do, dq = sets[i % 3]
    ops.attention_fwd(q, o, l, h, scale)
def bwd(i):
    q, o, l, do, dq = sets[i % 3]
    ops.attention_bwd(q, o, do, l, dq, h, scale)
dsums = [(st[3].float() * st[1].float()).view(S * 256, h, 64).sum(-1).contiguous() for st in sets]
def bwd_d(i):
    q, o, l, do, dq = sets[i % 3]
    ops.attention_bwd_d(q, do, l, dsums[i % 3], dq, h, scale)
timeit("attention fwd", fwd, 4.0 * S * h * 256 * 256 * 64)
for i in range(3):   # D of the freshly computed outputs
    dsums[i] = (sets[i][3].float() * sets[i][1].float()).view(S * 256, h, 64).sum(-1).contiguous()
timeit("attention bwd (one-shot)", bwd, 10.0 * S * h * 256 * 256 * 64)
dq_ref = sets[0][4].clone()
timeit("attention bwd (persistent)", bwd_d, 10.0 * S * h * 256 * 256 * 64)
print("persistent vs one-shot dqkv max diff %.4f" % (sets[0][4].float() - dq_ref.float()).abs().max().item())
# correctness on the first 4 sequences of set 0
q, o, l, do, dq = sets[0]
n = 4 * 256
x = q[:n].float().requires_grad_(True)
qq, kk, vv = x.view(4, 256, 3, h, 64).permute(2, 0, 3, 1, 4)
s = (qq * scale) @ kk.transpose(-1, -2)
ref = (s.softmax(-1) @ vv).transpose(1, 2).reshape(n, d)
ref.backward(do[:n].float())
print("fwd max err %.4f  lse err %.2e  bwd max err %.4f (ref absmax %.2f / %.2f)" % (
    (o[:n].float() - ref).abs().max().item(), (l[:4] - torch.logsumexp(s, -1)).abs().max().item(),
    (dq[:n].float() - x.grad).abs().max().item(), ref.abs().max().item(), x.grad.abs().max().item()))
