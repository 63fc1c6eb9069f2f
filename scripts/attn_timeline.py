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

# ---- backward (one-shot kernel): CTA 0 ----
dout = torch.randn(S * 256, d, device="cuda").bfloat16(); dqkv = torch.empty_like(qkv)
for _ in range(2): ops.attention_bwd(qkv, out, dout, lse, dqkv, h, scale)
buf.zero_()
lib.dig_attention_debug_buffer(ctypes.c_void_p(buf.data_ptr()))
ops.attention_bwd(qkv, out, dout, lse, dqkv, h, scale)
torch.cuda.synchronize()
lib.dig_attention_debug_buffer(None)
b = buf.cpu()
t0 = int(b[b > 0].min())
print("bwd MMA thread : row0 = start | loads issued | loads landed ; rows 1-4 (it): S,dP committed | pds seen | dV,dK,dQ committed ; row5: - | end | after sync")
for n in range(6): print("   ", " ".join("%7d" % (int(x) - t0) if x > 0 else "      -" for x in b[0, n]))
print("bwd thread 0   : row0 = start | D,lse done ; rows 1-4 (it): top | sdp seen | bufs free | pds arrive | mma seen | dV,dK stored ; row5: loop done | end | after sync")
for n in range(6): print("   ", " ".join("%7d" % (int(x) - t0) if x > 0 else "      -" for x in b[1, n]))
