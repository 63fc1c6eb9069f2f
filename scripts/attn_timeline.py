"""Clock-stamp timeline of CTA 0 of the persistent attention forward (bring-up instrumentation in attention.cu)."""
import ctypes, sys, torch
sys.path.insert(0, ".")
from dig_b200 import ops
lib = ops.load()
S, h = 256, 6
d, scale = h * 64, 64 ** -0.5
qkv = (torch.randn(S * 256, 3 * d, device="cuda") * 1.5).bfloat16()
out = torch.empty(S * 256, d, device="cuda", dtype=torch.bfloat16); lse = torch.empty(S, h, 256, device="cuda")
for _ in range(3): ops.attention_fwd(qkv, out, lse, h, scale)
buf = torch.zeros(3, 12, 8, dtype=torch.int64, device="cuda")
lib.dig_attention_debug_buffer.argtypes = [ctypes.c_void_p]
lib.dig_attention_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); ops.attention_fwd(qkv, out, lse, h, scale); e1.record()
torch.cuda.synchronize()
print('kernel (events): %.1f us' % (e0.elapsed_time(e1) * 1e3))
lib.dig_attention_debug_buffer(None)
b = buf.cpu()
k = b[0, 11]
print('CTA 0: %d clocks in %d ns -> %.3f GHz' % (k[1] - k[0], k[3] - k[2], float(k[1] - k[0]) / float(k[3] - k[2])))
b[0, 11] = 0
t0 = int(b[b > 0].min())
names = ["MMA  : - | S issued slot0 | slot1 | - | PV(keys 0-127) issued slot0 | slot1 | PV(keys 128-255) issued slot0 | slot1",
         "slot0: loop top | s_full | max pass done | P half 0 arrive | P half 1 arrive | o_full | s_free arrive (O in registers) | stored",
         "slot1: (same)"]
for r in range(3):
    print(names[r])
    for n in range(11):
        print("   item %2d: " % n + " ".join("%7d" % (int(x) - t0) if x > 0 else "      -" for x in b[r, n]))

# ---- backward (persistent kernel): CTA 0, first 12 blocks (3 items x 4 blocks of 128 queries x 128 keys) ----
dout = torch.randn(S * 256, d, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
dsum = (dout.float() * out.float()).view(S * 256, h, 64).sum(-1).contiguous()
for _ in range(2): ops.attention_bwd_d(qkv, dout, lse, dsum, dqkv, h, scale)
buf.zero_()
lib.dig_attention_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
e0.record(); ops.attention_bwd_d(qkv, dout, lse, dsum, dqkv, h, scale); e1.record()
torch.cuda.synchronize()
print('backward kernel (events): %.1f us' % (e0.elapsed_time(e1) * 1e3))
lib.dig_attention_debug_buffer(None)
b = buf.cpu()
t0 = int(b[b > 0].min())
print("bwd MMA warp (per block g): S,dP of g consumed (bar_sc) seen | S,dP(g+1) issued | P,dS of g stored (bar_pds) seen | dV,dK,dQ(g) issued")
for n in range(12): print("   block %2d: " % n + " ".join("%7d" % (int(x) - t0) if x > 0 else "      -" for x in b[0, n][:4]))
print("bwd compute thread 0 (per block g): loop top | S,dP seen | phase A done | previous dV,dK,dQ done (bar_mma) | staging done | P,dS stored, arrive")
for n in range(12): print("   block %2d: " % n + " ".join("%7d" % (int(x) - t0) if x > 0 else "      -" for x in b[1, n][:6]))
