"""Golden fixture for the frame-level data shaping: the reference's OWN functions
toolkit/utils/read_data.py:{feature_scale_compress, align_to_text, pad_to_maxlen_pre_modality} applied in
Data_Feat's order (toolkit/data/feat_data.py:33-44) to seeded ragged [T_i, 16] arrays.

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_shaping.py
Writes tests/golden/frame_shaping_golden.npz.
"""
import os
import sys

import numpy as np

REF = "/root/reference/MERBench"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED, N, DIM = 7, 9, 16


def ragged(seed):
    rng = np.random.default_rng(seed)
    lens = rng.integers(1, 60, (3, N))
    return [[rng.standard_normal((int(t), DIM)).astype(np.float32) for t in lens[m]] for m in range(3)]


def main():
    cwd = os.getcwd()
    os.chdir(REF)
    sys.path.insert(0, REF)
    try:
        import toolkit.utils.read_data as rd
    finally:
        os.chdir(cwd)
    out = {}
    for feat_type, scale in (("frm_align", 6), ("frm_unalign", 12), ("frm_unalign", 1)):
        a, t, v = ragged(SEED)
        a, t, v = rd.feature_scale_compress(a, t, v, scale)
        if feat_type == "frm_align":
            a, t, v = rd.align_to_text(a, t, v)
        a, t, v = rd.pad_to_maxlen_pre_modality(a, t, v)
        for name, x in (("a", a), ("t", t), ("v", v)):
            out[f"{feat_type}_{scale}_{name}"] = np.array(x)
    np.savez(os.path.join(OUT, "frame_shaping_golden.npz"), seed=SEED, n=N, dim=DIM, **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
