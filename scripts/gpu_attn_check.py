"""GPU bring-up check for dig_attention_fwd / _bwd against torch autograd (fp32 math on bf16-rounded inputs)."""
import sys
import torch
sys.path.insert(0, ".")
from dig_b200 import ops

torch.manual_seed(0)
dev = "cuda"
S, h = 8, 6
d = h * 64
scale = 64 ** -0.5
qkv = (torch.randn(S * 256, 3 * d, device=dev) * 1.5).bfloat16()


def ref(qkv32):
    q, k, v = qkv32.view(S, 256, 3, h, 64).permute(2, 0, 3, 1, 4)
    s = (q * scale) @ k.transpose(-1, -2)
    p = s.softmax(-1)
    o = (p @ v).transpose(1, 2).reshape(S * 256, d)
    return o, torch.logsumexp(s, -1)


x = qkv.float().requires_grad_(True)
o_ref, lse_ref = ref(x)
for variant in (0, 1):
    out = torch.zeros(S * 256, d, device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(S, h, 256, device=dev)
    try:
        ops.attention_fwd(qkv, out, lse, h, scale, p_in_smem=bool(variant))
        torch.cuda.synchronize()
        print("fwd variant", variant, "max err", (out.float() - o_ref).abs().max().item(), "lse err", (lse - lse_ref).abs().max().item(),
              "ref absmax", o_ref.abs().max().item())
    except Exception as e:
        print("fwd variant", variant, "EXC", e)
    sys.stdout.flush()

out = torch.zeros(S * 256, d, device=dev, dtype=torch.bfloat16)
lse = torch.zeros(S, h, 256, device=dev)
ops.attention_fwd(qkv, out, lse, h, scale, p_in_smem=True)
dout = torch.randn(S * 256, d, device=dev).bfloat16()
o_ref.backward(dout.float())
dqkv = torch.zeros_like(qkv)
ops.attention_bwd(qkv, out, dout, lse, dqkv, h, scale)
torch.cuda.synchronize()
g = x.grad
for name, sl in (("dq", slice(0, d)), ("dk", slice(d, 2 * d)), ("dv", slice(2 * d, 3 * d))):
    e = (dqkv.float()[:, sl] - g[:, sl]).abs().max().item()
    print(name, "max err", e, "ref absmax", g[:, sl].abs().max().item())

# timing at the bench shape
S2 = 256
qkv2 = torch.randn(S2 * 256, 3 * d, device=dev).bfloat16()
out2 = torch.empty(S2 * 256, d, device=dev, dtype=torch.bfloat16); lse2 = torch.empty(S2, h, 256, device=dev)
dout2 = torch.randn_like(out2); dqkv2 = torch.empty_like(qkv2)
for variant in (0, 1):
    for _ in range(3): ops.attention_fwd(qkv2, out2, lse2, h, scale, p_in_smem=bool(variant))
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.attention_fwd(qkv2, out2, lse2, h, scale, p_in_smem=bool(variant))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("fwd variant %d: %.3f ms, %.1f TFLOP/s" % (variant, ms, 4 * S2 * h * 256 * 256 * 64 * 1e-9 / ms))
for _ in range(3): ops.attention_bwd(qkv2, out2, dout2, lse2, dqkv2, h, scale)
e0.record()
for _ in range(20): ops.attention_bwd(qkv2, out2, dout2, lse2, dqkv2, h, scale)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
print("bwd: %.3f ms, %.1f TFLOP/s" % (ms, 10 * S2 * h * 256 * 256 * 64 * 1e-9 / ms))
q, k, v = qkv2.view(S2, 256, 3, h, 64).permute(2, 0, 3, 1, 4)
for _ in range(3): torch.nn.functional.scaled_dot_product_attention(q, k, v)
e0.record()
for _ in range(20): torch.nn.functional.scaled_dot_product_attention(q, k, v)
e1.record(); torch.cuda.synchronize()
print("torch sdpa fwd: %.3f ms" % (e0.elapsed_time(e1) / 20))
