"""Stand-alone timing of the attention kernels at the bench shapes (CUDA events, L2 flushed by the operand size):
ViT: 2,048 sequences x 197 tokens x 12 heads (fp16 operands); HuBERT: 256 x 249 x 12 (TF32 operands).
One JSON line per (kernel, version); TFLOP/s = 4 S^2 64 flop per (sequence, head)."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import _lib as L  # noqa: E402


def run(dtype, n_seq, S, heads, env, ver, iters, name=None):
    dev = torch.device("cuda:0")
    tokens = n_seq * S
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = (torch.randn(tokens, 3 * heads * 64, generator=g, device=dev) * 1.5).to(dtype)
    if dtype == torch.float32:
        L.round_tf32_(qkv)
    align = 8 if dtype == torch.float16 else 4
    ld = (tokens + align - 1) // align * align
    vt = torch.zeros(heads * 64, ld, dtype=dtype, device=dev)
    vt[:, :tokens] = qkv[:, 2 * heads * 64:].T
    cu = torch.arange(n_seq + 1, dtype=torch.int32, device=dev) * S
    ctx = torch.empty(tokens, heads * 64, dtype=dtype, device=dev)
    os.environ[env] = str(ver)
    for _ in range(3):
        L.attention(qkv, ctx, cu, S, heads, vt=vt)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.attention(qkv, ctx, cu, S, heads, vt=vt)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 4.0 * S * S * 64 * n_seq * heads
    return dict(kernel=name or ("attention_f16" if dtype == torch.float16 else "attention_tc"), ver=ver, n_seq=n_seq, S=S,
                ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1))


def run_legacy(n_seq, S, heads, iters):
    """The mma.sync flash kernel of attention.cu on TF32-rounded fp32 operands (what rows > 253 tokens took in round 1)."""
    dev = torch.device("cuda:0")
    tokens = n_seq * S
    g = torch.Generator(device=dev).manual_seed(3)
    qkv = torch.randn(tokens, 3 * heads * 64, generator=g, device=dev) * 1.5
    L.round_tf32_(qkv)
    cu = torch.arange(n_seq + 1, dtype=torch.int32, device=dev) * S
    ctx = torch.empty(tokens, heads * 64, device=dev)
    for _ in range(2):
        L.attention(qkv, ctx, cu, S, heads, round_out=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        L.attention(qkv, ctx, cu, S, heads, round_out=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    return dict(kernel="attention_kernel (mma.sync, tf32 operands)", n_seq=n_seq, S=S, ms=round(ms, 4),
                tflops=round(4.0 * S * S * 64 * n_seq * heads / ms / 1e9, 1))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--f16-vers", type=int, nargs="*", default=[3, 4])
    ap.add_argument("--tc-vers", type=int, nargs="*", default=[2])
    ap.add_argument("--poly", type=int, nargs="*", default=[None], help="MER_ATT_F16_POLY values to sweep")
    ap.add_argument("--n-seq", type=int, nargs="*", default=[2048], help="ViT sequences (37 = 3 items per SM, L2-resident)")
    ap.add_argument("--only-f16", action="store_true", help="skip the long-key / legacy / TF32 kernels")
    a = ap.parse_args()
    for n in a.n_seq:
        for v in a.f16_vers:
            for pl in a.poly:
                if pl is not None:
                    os.environ["MER_ATT_F16_POLY"] = str(pl)
                r = run(torch.float16, n, 197, 12, "MER_ATT_F16_VER", v, a.iters)
                r["poly"] = pl
                print(json.dumps(r), flush=True)
    if a.only_f16:
        sys.exit(0)
    # attention_f16_long.cu: audio rows of 10 s (499 frames), 7 s (349), CLIP L/14 (257 tokens, 16 heads); beside them the
    # round-1 path for such rows (MER_ATT_F16_LONG=0 is read by the stacks, not here: the TF32-operand mma.sync kernel
    # is what an fp32 qkv without V^T gets)
    for n, S, heads in ((256, 499, 12), (256, 349, 12), (512, 257, 16)):
        print(json.dumps(run(torch.float16, n, S, heads, "MER_ATT_F16_POLY", os.environ.get("MER_ATT_F16_POLY", "1"),
                             a.iters, name="attention_f16_long")), flush=True)
    print(json.dumps(run_legacy(256, 499, 12, max(2, a.iters // 4))), flush=True)
    for v in a.tc_vers:
        print(json.dumps(run(torch.float32, 256, 249, 12, "MER_ATT_TC_VER", v, a.iters)), flush=True)
