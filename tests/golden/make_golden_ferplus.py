"""Golden outputs of the reference's FER+ extractor (MERBench/feature_extraction/visual/extract_ferplus_embedding.py),
UNMODIFIED: its ``load_model`` builds ``resnet50_ferplus_dag`` from the reference's model definition and loads our
seeded synthetic checkpoint, ``compose_transforms(model.meta)`` + the reference ``FaceDataset`` read
``<face_dir>/<vid>/<vid>.npy``, ``get_feature(model, 'conv5_3_3x3_relu', imgs)`` produces the embeddings
(the script's own ``extract`` differs only by ``imgs.cuda()``), and the save rules of ``__main__`` (:170-194) are
applied as written there.  Stubs: ``skimage`` (imported by dataset.py, unused on this path) and ``config``.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_ferplus.py
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mertools_b200 import synthetic as S  # noqa: E402

VIS = "/root/reference/MERBench/feature_extraction/visual"
SEED = 9


def golden_clips():
    """The input videos (rebuilt from their seeds by the tests instead of being stored)."""
    return {"vidA": S.synth_frames(1, 3, size=256, seed=71)[0],      # Resize(256) is the identity here
            "vidB": S.synth_frames(1, 2, size=112, seed=72)[0],      # upscaled faces
            "vidC": np.random.default_rng(73).integers(0, 256, (1, 200, 300, 3), dtype=np.uint8)}  # non-square


def main():
    sk = types.ModuleType("skimage")
    sk.io, sk.img_as_float = types.ModuleType("skimage.io"), (lambda x: x)
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, sk.io
    sys.modules["config"] = types.ModuleType("config")
    sys.path.insert(0, VIS)
    spec = importlib.util.spec_from_file_location("extract_ferplus_embedding", os.path.join(VIS, "extract_ferplus_embedding.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    clips = golden_clips()
    out = {"seed": SEED, "names": np.array(list(clips))}
    for model_name, prefix, se in (("resnet50_ferplus_dag", "", False), ("senet50_ferplus_dag", "se_", True)):
        sd = S.ferplus_resnet50_state_dict(SEED, se=se)
        with tempfile.TemporaryDirectory() as tmp:
            torch.save({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, os.path.join(tmp, model_name + ".pth"))
            model = ref.load_model(model_name, os.path.join(VIS, "pytorch-benchmarks/model"), tmp).eval()
            transform = ref.compose_transforms(model.meta)
            for vid, frames in clips.items():
                os.makedirs(os.path.join(tmp, vid))
                np.save(os.path.join(tmp, vid, f"{vid}.npy"), frames)
                ds = ref.FaceDataset(vid, tmp, transform=transform)
                loader = torch.utils.data.DataLoader(ds, batch_size=32)
                feats, names = [], []
                with torch.no_grad():
                    for imgs, ids in loader:
                        feats.extend(ref.get_feature(model, "conv5_3_3x3_relu", imgs))
                        names.extend(ids)
                emb = np.array(feats)[np.argsort(np.array(names))]
                fra = np.array(emb).squeeze()                      # __main__ :178-184
                out[f"{prefix}fra_{vid}"] = fra[np.newaxis, :] if len(fra.shape) == 1 else fra
                utt = np.array(emb).squeeze()                      # :185-191
                out[f"{prefix}utt_{vid}"] = np.mean(utt, axis=0) if len(utt.shape) == 2 else utt
                if not se:  # preprocess probe (the two models share meta / transforms)
                    out[f"x_{vid}"] = torch.stack([ds[i][0] for i in range(len(ds))]).numpy()[:, :, ::16, ::16]
    np.savez_compressed(os.path.join(HERE, "ferplus_golden.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
