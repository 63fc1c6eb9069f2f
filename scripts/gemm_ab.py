"""Micro-benchmark of every GEMM shape of the ViT-S bs=128 step (CUDA events, 20 launches each, operands rotate over 3 buffer sets so the
working set exceeds L2).  Use with DIG_B200_LIB / DIG_GEMM_* environment switches for A/B comparisons."""
import os, sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
M = int(os.environ.get("AB_M", 65536)); d = 384
only = sys.argv[1:] 
def bench(name, mk, flops, iters=18):
    if only and not any(o in name for o in only): return
    sets = [mk() for _ in range(3)]
    for f in sets: f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters): sets[i % 3]()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    print("%-34s %8.1f us  %6.0f TFLOP/s" % (name, ms * 1e3, flops * 1e-9 / ms)); sys.stdout.flush()
bf = lambda *s: (torch.randn(*s, device="cuda") * 0.5).bfloat16()
f32 = lambda *s: torch.randn(*s, device="cuda")
def split_k(m, n, k, bn=128):
    tiles = ((m + 127) // 128) * ((n + bn - 1) // bn); kb = (k + 63) // 64
    return max(1, min(kb, 148 // max(tiles, 1)))
def fwd(N, K, out_f32, res, epi=ops.EPI_LINEAR, with_aux=True, q8=False):
    def mk():
        a, w, b = bf(M, K), bf(N, K), f32(N)
        out = torch.empty(M, N, device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
        r = f32(M, N) if res else None
        aux = torch.empty(M, N, device="cuda", dtype=torch.uint8 if q8 else torch.bfloat16) if (epi != ops.EPI_LINEAR and with_aux) else None
        return lambda: ops.gemm(a, w, out, bias=b, residual=r, epilogue=epi, aux=aux)
    return mk
def dgrad(N, K, epi=ops.EPI_LINEAR, colsum=False, q8=False):
    def mk():
        a, w = bf(M, K), bf(K, N)
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        aux = bf(M, N) if epi != ops.EPI_LINEAR else None
        if q8: aux = torch.randint(0, 256, (M, N), device="cuda", dtype=torch.uint8)
        cs = torch.zeros(N, device="cuda") if colsum else None
        return lambda: ops.gemm(a, w, out, b_mn_major=True, epilogue=epi, aux=aux, colsum=cs)
    return mk
def wgrad(Mo, No):
    def mk():
        a, b = bf(M, Mo), bf(M, No)
        out = torch.zeros(Mo, No, device="cuda")
        sk = int(os.environ.get("AB_SPLITK", 0)) or -1
        return lambda: ops.gemm(a, b, out, a_mn_major=True, b_mn_major=True, split_k=sk)
    return mk
fl = lambda n, k: 2.0 * M * n * k
bench("qkv fwd 1152x384 bf16", fwd(3 * d, d, False, False), fl(3 * d, d))
bench("proj fwd 384x384 f32+res", fwd(d, d, True, True), fl(d, d))
bench("fc1 fwd GELU 1536x384", fwd(4 * d, d, False, False, ops.EPI_GELU), fl(4 * d, d))
bench("fc1 fwd GELU 1536x384 q8 aux", fwd(4 * d, d, False, False, ops.EPI_GELU, True, True), fl(4 * d, d))
bench("fc1 fwd GELU no-aux (momentum)", fwd(4 * d, d, False, False, ops.EPI_GELU, False), fl(4 * d, d))
bench("fc1 fwd plain 1536x384 bf16", fwd(4 * d, d, False, False), fl(4 * d, d))
bench("fc2 fwd 384x1536 f32+res", fwd(d, 4 * d, True, True), fl(d, 4 * d))
bench("fc2 dgrad GELU_BWD+colsum", dgrad(4 * d, d, ops.EPI_GELU_BWD, True), fl(4 * d, d))
bench("fc2 dgrad GELU_BWD+colsum q8 aux", dgrad(4 * d, d, ops.EPI_GELU_BWD, True, True), fl(4 * d, d))
bench("fc2 dgrad GELU_BWD no-colsum", dgrad(4 * d, d, ops.EPI_GELU_BWD, False), fl(4 * d, d))
bench("fc2 dgrad plain (N=1536,K=384)", dgrad(4 * d, d), fl(4 * d, d))
bench("fc1 dgrad (N=384,K=1536)", dgrad(d, 4 * d), fl(d, 4 * d))
bench("proj dgrad (N=384,K=384)", dgrad(d, d), fl(d, d))
bench("qkv dgrad (N=384,K=1152)", dgrad(d, 3 * d), fl(d, 3 * d))
bench("fc2 wgrad 384x1536", wgrad(d, 4 * d), fl(d, 4 * d))
bench("fc1 wgrad 1536x384", wgrad(4 * d, d), fl(d, 4 * d))
bench("proj wgrad 384x384", wgrad(d, d), fl(d, d))
bench("qkv wgrad 1152x384", wgrad(3 * d, d), fl(d, 3 * d))
