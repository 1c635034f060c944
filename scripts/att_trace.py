"""Phase timeline of attention_f16_kernel<4 | 5> (MER_ATT_F16_VER) (block 0, first items): clock64 stamps written by the kernel when the
debug hook mer_debug_attention_trace() holds a buffer.  Prints, per item, cycles relative to the item's first stamp."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import _lib as L  # noqa: E402

NAMES = {0: "prod:issued", 1: "mma:kq_ready", 2: "mma:S0_issued", 3: "mma:S1_issued", 4: "mma:v_ready",
         5: "mma:pv t0 (v6+) / c0 t0", 6: "mma:pv t1 (v6+) / c0 t1", 7: "mma:pv c1 t0", 8: "mma:pv c1 t1", 9: "mma:pv c2 t0", 10: "mma:pv c2 t1",
         11: "mma:pv c3 t0", 12: "mma:pv c3 t1", 13: "sm0:sfull", 14: "sm0:pass1", 15: "sm0:p0", 16: "sm0:p1", 17: "sm0:p2",
         18: "sm0:p3", 19: "sm0:ofull", 20: "sm0:otfree", 21: "sm0:ofree", 22: "sm1:sfull", 23: "sm1:pass1", 24: "sm1:p0",
         25: "sm1:p1", 26: "sm1:p2", 27: "sm1:p3", 28: "sm1:ofull", 29: "sm1:otfree", 30: "sm1:ofree"}

dev = torch.device("cuda:0")
n_seq, S, heads = 148, 197, 12
tokens = n_seq * S
qkv = (torch.randn(tokens, 3 * heads * 64, device=dev) * 1.5).half()
vt = torch.zeros(heads * 64, (tokens + 7) // 8 * 8, dtype=torch.float16, device=dev)
vt[:, :tokens] = qkv[:, 2 * heads * 64:].T
cu = torch.arange(n_seq + 1, dtype=torch.int32, device=dev) * S
ctx = torch.empty(tokens, heads * 64, dtype=torch.float16, device=dev)
os.environ["MER_ATT_F16_VER"] = os.environ.get("MER_ATT_F16_VER", "7")
for _ in range(2):
    L.attention(qkv, ctx, cu, S, heads, vt=vt)
buf = torch.zeros(16 * 32, dtype=torch.int64, device=dev)
lib = L.lib()
lib.mer_debug_attention_trace.argtypes = [C.c_void_p]
lib.mer_debug_attention_trace(C.c_void_p(buf.data_ptr()))
L.attention(qkv, ctx, cu, S, heads, vt=vt)
torch.cuda.synchronize()
lib.mer_debug_attention_trace(None)
t = buf.cpu().view(16, 32)
t0 = int(t[t > 0].min())
for it in range(2, 8):
    row = t[it]
    ev = sorted((int(row[k]) - t0, NAMES[k]) for k in NAMES if int(row[k]) > 0)
    print(f"item {it}: " + "  ".join(f"{n}@{c}" for c, n in ev))
print("item period (sm0:otfree):", [int(t[i + 1][20] - t[i][20]) for i in range(2, 10)])
