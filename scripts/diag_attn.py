"""Diagnostic of the tcgen05 attention kernel stages (MER_ATTENTION_DEBUG=1|2|3)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import _lib as L
dbg = int(os.environ.get("MER_ATTENTION_DEBUG", "0"))
heads, lens = 12, [197, 80, 64]
torch.manual_seed(0)
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32, device="cuda")
tot = sum(lens)
x = torch.randn(tot, 3 * heads * 64, device="cuda") * 1.5
i = x.view(torch.int32); qkv = ((i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF).view(torch.float32)
ctx = torch.full((tot, heads * 64), float("nan"), device="cuda")
ld = (tot + 3) // 4 * 4
vt = torch.zeros(heads * 64, ld, device="cuda"); vt[:, :tot] = qkv[:, 2 * heads * 64:].t()
L.attention(qkv, ctx, cu, max(lens), heads, vt=vt)
torch.cuda.synchronize()
s0 = 0
for n in lens:
    q, k, v = qkv[s0:s0 + n].double().view(n, 3, heads, 64).permute(1, 2, 0, 3)   # [h, n, 64]
    s = q @ k.transpose(-1, -2)
    mx = (s / 8).max(-1, keepdim=True).values
    p = torch.exp(s / 8 - mx)
    if dbg == 1: ref = s[..., :64]
    elif dbg == 2: ref = p[..., :64]
    elif dbg == 3: ref = p[..., :32] @ k[:, :64, :32].transpose(-1, -2)
    else: ref = (p / p.sum(-1, keepdim=True)) @ v
    if dbg in (1, 2) and ref.shape[-1] < 64:
        ref = torch.nn.functional.pad(ref, (0, 64 - ref.shape[-1]))
    nk = min(n, 64) if dbg in (1, 2) else 64
    got = ctx[s0:s0 + n].double().view(n, heads, 64).permute(1, 0, 2)[..., :nk]
    ref = ref[..., :nk]
    err = (got - ref).abs()
    print(f"debug {dbg} len {n}: max err {err.max().item():.3e} (ref max {ref.abs().max().item():.2e}) nan {int(torch.isnan(got).sum())}",
          "worst head", int(err.amax((1, 2)).argmax()), "row", int(err.amax((0, 2)).argmax()), "col", int(err.amax((0, 1)).argmax()))
    if err.max() > 1e-2 * ref.abs().max():
        print("   got", got[0, 0, :6].tolist()); print("   ref", ref[0, 0, :6].tolist())
    s0 += n
