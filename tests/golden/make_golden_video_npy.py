"""Golden outputs of the reference's ``load_video_from_npy`` (MERBench/toolkit/utils/functions.py:79-118): the source
text of that function and of ``func_opencv_to_decord`` (:70-71) is executed AS IS (the module itself imports torchaudio /
an OpenAI client, which are irrelevant here), with ``func_video_to_face`` bound to the synthetic clip.  All four
``readtype`` modes; the two random ones with ``np.random.seed`` set right before the call.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_video_npy.py
"""
import os
import sys

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mertools_b200 import synthetic as S  # noqa: E402

SRC = "/root/reference/MERBench/toolkit/utils/functions.py"
CASES = [("uniform", 8, 37, 112), ("uniform", 8, 5, 112), ("all", 0, 6, 100), ("continuous", 8, 120, 112),
         ("continuous", 8, 3, 112), ("continuous_polish", 8, 150, 96), ("uniform", 16, 1000, 64)]


def golden_clip(vlen, size, seed):
    return S.synth_frames(1, vlen, size=size, seed=seed)[0]


def main():
    lines = open(SRC).read().splitlines()
    src = "\n".join(lines[69:71] + [""] + lines[78:118])          # func_opencv_to_decord, load_video_from_npy
    assert src.lstrip().startswith("def func_opencv_to_decord") and "def load_video_from_npy" in src
    out = {"cases": np.array([f"{r}|{n}|{v}|{s}" for r, n, v, s in CASES])}
    for ci, (readtype, n_frms, vlen, size) in enumerate(CASES):
        frames = golden_clip(vlen, size, 200 + ci)
        ns = {"np": np, "cv2": cv2, "torch": torch, "func_video_to_face": lambda vname, f=frames: f}
        exec(src, ns)                                             # noqa: S102  (the reference's own source text)
        np.random.seed(1000 + ci)
        got = ns["load_video_from_npy"]("clip", n_frms=n_frms, height=224, width=224, readtype=readtype)
        out[f"shape{ci}"] = np.array(got.shape)
        x = got.numpy()
        out[f"probe{ci}"] = x[:, :, ::16, ::16].astype(np.uint8)  # values are integral (uint8 frames as float)
        out[f"sum{ci}"] = np.array([float(x.sum(dtype=np.float64))])
    np.savez_compressed(os.path.join(HERE, "video_npy_golden.npz"), **out)
    for k, v in out.items():
        print(k, v.shape if k != "cases" else list(v))


if __name__ == "__main__":
    main()
