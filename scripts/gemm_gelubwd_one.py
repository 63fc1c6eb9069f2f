import sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
M, N, K = 65536, 1536, 384
a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(K, N, device="cuda")*0.05).bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16); aux = torch.randn(M, N, device="cuda").bfloat16()
for _ in range(3):
    ops.gemm(a, w, out, b_mn_major=True, epilogue=ops.EPI_GELU_BWD, aux=aux)
torch.cuda.synchronize()
