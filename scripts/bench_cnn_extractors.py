#!/usr/bin/env python
"""Frames per second of the frame-level CNN extractors (ResNet-18, FER+ ResNet-50 / SENet-50, MA-Net, EmoNet) and
examples per second of VGGish on synthetic inputs resident in HBM.  One JSON line per extractor.

    python scripts/bench_cnn_extractors.py [--frames 256] [--reps 3]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import encoders as En  # noqa: E402
from mertools_b200 import synthetic as S  # noqa: E402


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=256)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    rng = np.random.default_rng(0)
    frames = torch.from_numpy(rng.integers(0, 256, (args.frames, 224, 224, 3), dtype=np.uint8)).to(dev)
    frames256 = torch.from_numpy(rng.integers(0, 256, (args.frames, 256, 256, 3), dtype=np.uint8)).to(dev)
    cases = [("resnet18", lambda: En.ResNet18Encoder(S.resnet18_state_dict(), device=dev), frames, 64),
             ("ferplus_resnet50", lambda: En.FerplusResnet50Encoder(S.ferplus_resnet50_state_dict(), device=dev), frames256, 64),
             ("ferplus_senet50", lambda: En.FerplusResnet50Encoder(S.ferplus_resnet50_state_dict(se=True), device=dev), frames256, 64),
             ("manet", lambda: En.ManetEncoder(S.manet_state_dict(), device=dev), frames, 64),
             ("emonet", lambda: En.EmonetEncoder(S.emonet_state_dict(), device=dev), frames256, 8)]
    for name, make, x, chunk in cases:
        enc = make()
        s = timed(lambda: enc.frame_features(x, max_frames=chunk), args.reps)
        print(json.dumps({"extractor": name, "frames": int(x.shape[0]), "frames_per_s": x.shape[0] / s, "ms": s * 1e3}))
        del enc
        torch.cuda.empty_cache()
    ex = torch.from_numpy(rng.normal(-2.0, 2.0, (args.frames, 96, 64)).astype(np.float32)).to(dev)
    enc = En.VggishEncoder(S.vggish_state_dict(), device=dev)
    s = timed(lambda: enc.embeddings(ex), args.reps)
    print(json.dumps({"extractor": "vggish", "examples": int(ex.shape[0]), "examples_per_s": ex.shape[0] / s, "ms": s * 1e3}))


if __name__ == "__main__":
    main()
