"""Golden fixture for the resize step of the visual path: the UNMODIFIED reference script
extract_vision_huggingface.py run on 112x112 face crops (the size OpenFace writes), which HF
ViTImageProcessor resizes to 224x224 (:137-138).

Run once in the build container (needs /root/reference + transformers; NOT on the GPU box):
    python tests/golden/make_golden_resize.py
Writes tests/golden/visual112_golden.npz.  Same stubs as make_golden.py (empty `timm`, patched `config`).
"""
import os
import runpy
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MERBench"
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

SEED, SIZE, NFRAME = 131, 112, 64


def main():
    from transformers import ViTConfig, ViTImageProcessor, ViTModel
    work = tempfile.mkdtemp(prefix="mer_golden_rs_")
    tools = os.path.join(work, "tools", "transformers")
    feats = os.path.join(work, "features")
    os.makedirs(tools)
    os.makedirs(feats)
    cfg = types.ModuleType("config")
    cfg.PATH_TO_RAW_FACE = {"MER2023": os.path.join(work, "openface_face")}
    cfg.PATH_TO_FEATURES = {"MER2023": feats}
    cfg.PATH_TO_PRETRAINED_MODELS = os.path.join(work, "tools")
    sys.modules["config"] = cfg
    sys.modules["timm"] = types.ModuleType("timm")
    vdir = os.path.join(tools, "dinov2-large")  # AutoModel dispatches on config.json: a ViT-B/16 under this name
    m = ViTModel(ViTConfig())
    m.load_state_dict({k: torch.from_numpy(v) for k, v in S.vit_state_dict(seed=0).items()}, strict=True)
    m.save_pretrained(vdir)
    ViTImageProcessor().save_pretrained(vdir)
    clips = S.synth_frames(2, 8, size=SIZE, seed=SEED)
    for i, c in enumerate(clips):
        d = os.path.join(cfg.PATH_TO_RAW_FACE["MER2023"], f"clip{i}")
        os.makedirs(d)
        np.save(os.path.join(d, f"clip{i}.npy"), c)
    for level in ("UTTERANCE", "FRAME"):
        sys.argv = ["extract_vision_huggingface.py", "--dataset=MER2023", "--model_name=dinov2-large",
                    f"--feature_level={level}", "--gpu=-1"]
        cwd = os.getcwd()
        os.chdir(os.path.join(REF, "feature_extraction", "visual"))
        try:
            runpy.run_path("extract_vision_huggingface.py", run_name="__main__")
        finally:
            os.chdir(cwd)
    out = {}
    for i in range(2):
        out[f"utt{i}"] = np.load(os.path.join(feats, "dinov2-large-UTT", f"clip{i}.npy"))
        out[f"fra{i}"] = np.load(os.path.join(feats, "dinov2-large-FRA", f"clip{i}.npy"))[:8]  # frames 0..7 (then repeats)
    np.savez(os.path.join(OUT, "visual112_golden.npz"), seed=SEED, size=SIZE, n_clips=2, nframe=NFRAME, **out)
    print("visual112:", {k: v.shape for k, v in out.items()})
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
