import sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
M, N, K = (int(x) for x in sys.argv[1:4])
a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(a, b, out)
torch.cuda.synchronize()
