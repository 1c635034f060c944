"""Stand-alone A/B timing of the HBM-side kernels and of the ViT layer's four GEMMs at the bench shapes (CUDA events on
the launching stream, operands larger than L2).  One JSON line per measurement.

    python scripts/bench_ab.py ln      # LayerNorm: MER_LN_VER=1 (round 1) against 2, ViT / HuBERT / BERT row counts
    python scripts/bench_ab.py conv0   # conv0 + GroupNorm + GELU through a 1-layer HuBERT forward (class-14 event timers)
    python scripts/bench_ab.py gemm    # qkv / out-proj / fc1 / fc2 of one ViT layer at 403,456 rows
    python scripts/bench_ab.py fc2     # two fc2 launches (for ncu --metrics dram__bytes_*.sum -k regex:gemm_kernel)
"""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import _lib as L  # noqa: E402

DEV = torch.device("cuda:0")


def timed(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_ln():
    g = torch.randn(768, device=DEV)
    b = torch.randn(768, device=DEV)
    for name, rows, f16, split16 in (("vit", 403456, True, False), ("hubert", 63744, False, True), ("bert", 8192, False, True)):
        x = torch.randn(rows, 768, device=DEV)
        y = torch.empty(rows, 768, device=DEV, dtype=torch.float16 if f16 else torch.float32)
        ys = torch.empty(rows, 768, device=DEV, dtype=torch.float16) if split16 else None
        flags = (L.MER_LN_OUT_F16 if f16 else 0) | (32 if split16 else 0)  # 32 = MER_LN_SPLIT_F16
        nbytes = rows * 768 * (4 + (2 if f16 else 4) + (2 if split16 else 0))
        outs = {}
        for ver in (1, 2, 1, 2):
            os.environ["MER_LN_VER"] = str(ver)
            ms = timed(lambda: L.layernorm(x, g, b, y, eps=1e-12, y_split=ys, flags=flags), 30)
            outs[ver] = y.clone()
            print(json.dumps({"kernel": "layernorm", "shape": name, "rows": rows, "ver": ver, "ms": round(ms, 4),
                              "GBps": round(nbytes / ms / 1e6, 1)}), flush=True)
        assert torch.equal(outs[1], outs[2]), "LayerNorm v2 differs from v1"
    os.environ.pop("MER_LN_VER", None)


def bench_conv0():
    from mertools_b200 import synthetic as S
    from mertools_b200.encoders import HubertEncoder
    sd = S.hubert_state_dict(seed=1, layers=4)  # (the last-four readout needs four layers)
    enc = HubertEncoder(sd, device=DEV)
    wave = torch.randn(256, 80000, device=DEV) * 0.1
    ref = None
    for packed in ("1", "0", "1"):
        os.environ["MER_CONV0_PACKED"] = packed
        enc.forward(wave)
        torch.cuda.synchronize()
        L.lib().mer_profile_enable(1)
        for _ in range(5):
            utt, _ = enc.forward(wave)
        torch.cuda.synchronize()
        t, w, n = C.c_double(), C.c_double(), C.c_int()
        L.lib().mer_profile_collect(14, C.byref(t), C.byref(w), C.byref(n))
        L.lib().mer_profile_enable(0)
        line = {"kernel": "conv0 (moments + coef + apply)", "packed": packed, "launch_groups": n.value,
                "ms": round(t.value / max(1, n.value), 4), "GBps": round(w.value / t.value / 1e6, 1)}
        if ref is None:
            ref = utt.clone()
        else:
            line["max_rel_vs_first"] = float((utt - ref).abs().max() / ref.abs().max())
        print(json.dumps(line), flush=True)
    os.environ.pop("MER_CONV0_PACKED", None)


def vit_gemms(rows=403456):
    g = torch.Generator(device=DEV).manual_seed(1)
    h = lambda *s: (torch.randn(*s, generator=g, device=DEV) * 0.5).half()  # noqa: E731
    x = torch.randn(rows, 768, generator=g, device=DEV)
    xn, hh = h(rows, 768), h(rows, 3072)
    out32 = torch.empty(rows, 768, device=DEV)
    qkv = torch.empty(rows, 2304, device=DEV, dtype=torch.float16)
    h_out = torch.empty(rows, 3072, device=DEV, dtype=torch.float16)
    w_qkv, w_o, w_1, w_2 = h(2304, 768) * 0.1, h(768, 768) * 0.1, h(3072, 768) * 0.1, h(768, 3072) * 0.1
    bias = lambda n: torch.randn(n, generator=g, device=DEV) * 0.1  # noqa: E731
    b_qkv, b_o, b_1, b_2 = bias(2304), bias(768), bias(3072), bias(768)
    F = L.MER_GEMM_F16
    return {
        "qkv": (lambda: L.gemm(xn, w_qkv, qkv, bias=b_qkv, mode=F, f16_out=True), 2.0 * rows * 768 * 2304),
        "out-proj": (lambda: L.gemm(xn, w_o, out32, bias=b_o, res=x, mode=F), 2.0 * rows * 768 * 768),
        "fc1": (lambda: L.gemm(xn, w_1, h_out, bias=b_1, gelu=True, mode=F, f16_out=True), 2.0 * rows * 768 * 3072),
        "fc2": (lambda: L.gemm(hh, w_2, out32, bias=b_2, res=x, mode=F), 2.0 * rows * 768 * 3072),
    }


def bench_gemm():
    """The four GEMMs of a ViT layer, twice (the box drifts along its power cap: only neighbouring lines compare)."""
    gm = vit_gemms()
    for rep in range(2):
        for name, (fn, flops) in gm.items():
            ms = timed(fn, 10)
            print(json.dumps({"kernel": "gemm F16 2SM", "launch": name, "rep": rep, "ms": round(ms, 4),
                              "tflops": round(flops / ms / 1e9, 1)}), flush=True)


def one_fc2():
    """Two fc2 launches for `ncu --metrics dram__bytes_read.sum,... -k regex:gemm_kernel`: as in the step (fp32 residual)
    and without a residual operand -- which operand accounts for the DRAM reads above the algorithmic 3.7 GB."""
    rows = 403456
    g = torch.Generator(device=DEV).manual_seed(1)
    hh = (torch.randn(rows, 3072, generator=g, device=DEV) * 0.5).half()
    w = (torch.randn(768, 3072, generator=g, device=DEV) * 0.05).half()
    x = torch.randn(rows, 768, generator=g, device=DEV)
    bias = torch.randn(768, generator=g, device=DEV) * 0.1
    out = torch.empty(rows, 768, device=DEV)
    torch.cuda.synchronize()
    L.gemm(hh, w, out, bias=bias, res=x, mode=L.MER_GEMM_F16)
    L.gemm(hh, w, out, bias=bias, mode=L.MER_GEMM_F16)
    torch.cuda.synchronize()
    print("fc2 done")


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["ln", "conv0", "gemm", "fc2"])
    a = ap.parse_args()
    {"ln": bench_ln, "conv0": bench_conv0, "gemm": bench_gemm, "fc2": one_fc2}[a.what]()
