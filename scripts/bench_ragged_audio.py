#!/usr/bin/env python
"""Audio extraction throughput on a corpus-like mix of clip lengths (uniform 1.5 .. 9.5 s, no two alike), through
``AudioExtractor.extract_waves``: the default grouping (one device pass per distinct length = one per clip here)
against ``ragged=True`` (``mer_hubert_forward_ragged``: sorted clips share passes).  One JSON line per mode.

    python scripts/bench_ragged_audio.py [--clips 256] [--layers 12] [--large]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import synthetic as S  # noqa: E402
from mertools_b200.extract.audio import AudioExtractor  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=256)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--large", action="store_true")
    args = ap.parse_args()
    rng = np.random.default_rng(0)
    lens = sorted(set(int(n) for n in rng.integers(24000, 152000, size=args.clips)))
    waves = [rng.standard_normal(n).astype(np.float64) * 0.1 for n in lens]
    sd = S.hubert_state_dict(seed=1, layers=args.layers, large=args.large)
    ref = None
    for ragged in (False, True):
        ext = AudioExtractor(sd, device="cuda:0", ragged=ragged)
        ext.extract_waves(waves[:8], "UTTERANCE")          # warm-up (workspace, attribute setup)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = ext.extract_waves(waves, "UTTERANCE")
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line = {"mode": "ragged" if ragged else "per-length", "clips": len(waves), "seconds_of_audio": sum(lens) / 16000.0,
                "clips_per_s": len(waves) / dt, "wall_s": dt}
        if ref is None:
            ref = out
        else:
            line["max_rel_vs_per_length"] = float(max(np.abs(a - b).max() / np.abs(a).max() for a, b in zip(ref, out)))
        print(json.dumps(line))
        del ext
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
