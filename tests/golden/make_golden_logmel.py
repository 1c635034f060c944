"""Golden fixture for the log-mel front-end: the reference's OWN vggish_input.waveform_to_examples
(mel_features.py / vggish_params.py, pure numpy) on seeded waveforms.

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_logmel.py
Writes tests/golden/logmel_golden.npz.  Stub: an empty `resampy` module (imported, never called at 16 kHz).
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MERBench/feature_extraction/audio"  # the module does `from vggish import mel_features`
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

LENS, SEED0, HOP_SEC = (80000, 16000, 20011), 400, 0.25


def main():
    sys.modules.setdefault("resampy", types.ModuleType("resampy"))
    sys.path.insert(0, REF)
    from vggish import vggish_input
    out = {}
    for i, n in enumerate(LENS):
        w = S.synth_waves(1, n, seed=SEED0 + i)[0].astype(np.float64) / 32768.0
        ex = vggish_input.waveform_to_examples(w, 16000, HOP_SEC)
        out[f"ex{i}"] = np.ascontiguousarray(ex[:: max(1, len(ex) // 4)]).astype(np.float32)  # a few patches
        out[f"n{i}"] = np.array(ex.shape)
    np.savez(os.path.join(OUT, "logmel_golden.npz"), lens=np.array(LENS), seed0=SEED0, hop_sec=HOP_SEC, **out)
    print({k: (v.shape if v.ndim > 1 else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
