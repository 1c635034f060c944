"""Golden outputs of the reference's MA-Net extractor (MERBench/feature_extraction/visual/extract_manet_embedding.py),
UNMODIFIED pieces: ``manet(num_classes=7)`` from the reference's model definition with our seeded synthetic
checkpoint (strict ``load_state_dict``), the script's transform (Resize((224, 224)) + ToTensor, :60-61), the
reference ``FaceDataset``, ``model(images, return_embedding=True)`` as in the script's ``extract`` (:31-41, minus
``.cuda()``), and the save rules of ``__main__`` (:88-103).  Stubs: ``skimage`` (imported by dataset.py) and ``config``.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_manet.py
"""
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
SEED = 10


def golden_clips():
    """The input videos (rebuilt from their seeds by the tests instead of being stored)."""
    return {"vidA": S.synth_frames(1, 3, size=224, seed=81)[0],
            "vidB": S.synth_frames(1, 1, size=112, seed=82)[0]}     # one upscaled frame


def main():
    sk = types.ModuleType("skimage")
    sk.io, sk.img_as_float = types.ModuleType("skimage.io"), (lambda x: x)
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, sk.io
    sys.modules["config"] = types.ModuleType("config")
    sys.path.insert(0, VIS)
    import torchvision.transforms as transforms
    from dataset import FaceDataset
    from manet.model.manet import manet

    sd = S.manet_state_dict(SEED)
    model = manet(num_classes=7)
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})   # strict: same names / shapes
    model.eval()
    transform = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor()])
    out = {"seed": SEED, "names": np.array(list(golden_clips()))}
    with tempfile.TemporaryDirectory() as tmp:
        for vid, frames in golden_clips().items():
            os.makedirs(os.path.join(tmp, vid))
            np.save(os.path.join(tmp, vid, f"{vid}.npy"), frames)
            loader = torch.utils.data.DataLoader(FaceDataset(vid, tmp, transform=transform), batch_size=32)
            feats, names = [], []
            with torch.no_grad():
                for images, ids in loader:
                    feats.append(model(images, return_embedding=True).cpu().detach().numpy())
                    names.extend(ids)
            emb = np.vstack(feats)[np.argsort(np.array(names))]
            fra = np.array(emb).squeeze()
            out[f"fra_{vid}"] = fra[np.newaxis, :] if len(fra.shape) == 1 else fra
            utt = np.array(emb).squeeze()
            out[f"utt_{vid}"] = np.mean(utt, axis=0) if len(utt.shape) == 2 else utt
    np.savez_compressed(os.path.join(HERE, "manet_golden.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
