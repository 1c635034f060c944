"""Which of HuBERT's convolutions can run on fp16 operands?  (Second step of scripts/precision_table.py.)

With the 12 transformer layers on fp16 operands (the round-2 default), the readout error is measured -- through the oracle
on the CPU, operands of every product rounded the way each scheme's MMA would, fp32 accumulation -- with fp16 operands in
conv1, conv1-2, conv1-3 and conv1-6, the other convolutions and the feature projection on bf16 (hi, lo) splits and conv0
exact (it runs on the fp32 pipe).  Four seeded checkpoints x clips, because one max-norm figure is noisy.
Writes profiles/r2_precision_conv_layers.json.  CPU only; ~3 minutes."""
import json
import os
import sys

import numpy as np
import torch
import torch.nn.functional as TF

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import precision_table as PT  # noqa: E402
from mertools_b200 import synthetic as S  # noqa: E402
from oracle import encoders as E  # noqa: E402
from oracle import pipeline as P  # noqa: E402


class ConvShim(PT.Shim):
    """fp16 (11-bit) operands in the transformer layers and in the convolutions listed in `f16_convs` (1..6); the other
    convolutions on 16-bit (hi + lo) operands; conv0 exact."""

    def __init__(self, f16_convs):
        super().__init__("f16-layers")
        self.f16_convs, self.n = set(f16_convs), 0

    def conv1d(self, x, w, b=None, **kw):
        self.in_layers = False
        if kw.get("groups", 1) > 1:   # positional conv: an fp16 GEMM in every scheme
            return TF.conv1d(PT.rnd(x, 11), PT.rnd(w, 11), b, **kw)
        k, self.n = self.n, self.n + 1
        if k == 0:
            return TF.conv1d(x, w, b, **kw)
        bits = 11 if k in self.f16_convs else 16
        return TF.conv1d(PT.rnd(x, bits), PT.rnd(w, bits), b, **kw)


class patched:
    def __init__(self, f16_convs):
        self.ctx, self.shim = PT.patched("f16-layers"), ConvShim(f16_convs)

    def __enter__(self):
        self.ctx.__enter__()
        E.F = self.shim   # (the linear layers keep precision_table's f16-layers rule)

    def __exit__(self, *a):
        self.ctx.__exit__(*a)


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    schemes = {"none": [], "conv1": [1], "conv1-2": [1, 2], "conv1-3": [1, 2, 3], "conv1-6": [1, 2, 3, 4, 5, 6]}
    res = {k: [] for k in schemes}
    for s in range(4):
        wav = (S.synth_waves(1, 80000, seed=31 + s).astype(np.float64) / 32768.0)[0]
        sd = S.hubert_state_dict(seed=1 + s, layers=12)
        with torch.no_grad():
            ref = P.audio_clip_features(sd, wav, layers=12)
            for name, convs in schemes.items():
                with patched(convs):
                    res[name].append(PT.rel(P.audio_clip_features(sd, wav, layers=12), ref))
        print(s, {k: f"{v[-1]:.2e}" for k, v in res.items()}, flush=True)
    out = {"what": __doc__.split("\n\n")[1].replace("\n", " "), "rows": res,
           "summary": {k: {"mean": float(np.mean(v)), "max": float(np.max(v))} for k, v in res.items()},
           "choice": "conv1 + conv2 on fp16 operands (77 % of the conv stack's flops)"}
    json.dump(out, open(os.path.join(ROOT, "profiles", "r2_precision_conv_layers.json"), "w"), indent=1)
    print(out["summary"])


if __name__ == "__main__":
    main()
