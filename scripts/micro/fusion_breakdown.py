"""Where the fusion step's time goes: eval forward (row kernel, forward only), train-mode forward, backward (row kernel
both directions + weight-gradient kernel), full step; CUDA events over back-to-back launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mertools_b200 import synthetic as S  # noqa: E402
from mertools_b200.fusion import FusionNet  # noqa: E402

dev = "cuda:0"


def timeit(fn, iters=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / iters


for B in (4, 32, 256):
    net = FusionNet(dropout=0.3, device=dev, seed=7).load_state_dict(S.fusion_state_dict(seed=3))
    a, t, v, emo, val = S.synth_fusion_features(B, seed=4)
    T = torch.from_numpy
    d = [T(a).to(dev), T(t).to(dev), T(v).to(dev), T(emo).to(dev), T(val).view(-1, 1).to(dev)]
    batch = {"audios": d[0], "texts": d[1], "videos": d[2]}
    up = [torch.randn(B, 128, device=dev), torch.randn(B, 6, device=dev), torch.randn(B, 1, device=dev)]
    print(f"B={B}: eval fwd {timeit(lambda: net.forward(batch)):.1f} us | train fwd {timeit(lambda: net.forward_train(*d[:3])):.1f} us | "
          f"backward (rows + wgrad) {timeit(lambda: net.backward(*d[:3], *up)):.1f} us | "
          f"step eager {timeit(lambda: net.train_step(*d, weight_decay=1e-5, use_graph=False)):.1f} us | "
          f"step graph {timeit(lambda: net.train_step(*d, weight_decay=1e-5)):.1f} us", flush=True)
