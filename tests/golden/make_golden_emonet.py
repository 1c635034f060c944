"""Golden outputs of the reference's EmoNet extractor (MERBench/feature_extraction/visual/extract_emonet_embedding.py),
UNMODIFIED pieces: ``EmoNet()`` from the reference's model definition with our seeded synthetic checkpoint (strict
``load_state_dict``), ``DataAugmentor(256, 256)`` + ``ToTensor`` (:53-54) through the reference
``FaceDatasetForEmoNet``, ``model(images, return_embedding=True)`` as in the script's ``extract`` (:22-33, minus
``.cuda()``), the save rules of ``__main__`` (:76-94).  Stubs: ``skimage`` (imported by dataset.py) and ``config``.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_emonet.py
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
SEED = 11


def golden_clips():
    return {"vidA": S.synth_frames(1, 2, size=256, seed=95)[0],       # cv2.resize is the identity here
            "vidB": S.synth_frames(1, 2, size=112, seed=96)[0],       # upscaled faces
            "vidC": S.synth_frames(1, 1, size=300, seed=97)[0]}       # downscaled


def main():
    sk = types.ModuleType("skimage")
    sk.io, sk.img_as_float = types.ModuleType("skimage.io"), (lambda x: x)
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, sk.io
    sys.modules["config"] = types.ModuleType("config")
    sys.path.insert(0, VIS)
    import torchvision.transforms as transforms
    from dataset import FaceDatasetForEmoNet
    from emonet.data_augmentation import DataAugmentor
    from emonet.models.emonet import EmoNet

    sd = S.emonet_state_dict(SEED)
    model = EmoNet()
    model.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})   # strict
    model.eval()
    augmentor = DataAugmentor(256, 256)
    transform = transforms.Compose([transforms.ToTensor()])
    out = {"seed": SEED, "names": np.array(list(golden_clips()))}
    with tempfile.TemporaryDirectory() as tmp:
        for vid, frames in golden_clips().items():
            os.makedirs(os.path.join(tmp, vid))
            np.save(os.path.join(tmp, vid, f"{vid}.npy"), frames)
            ds = FaceDatasetForEmoNet(vid, tmp, transform=transform, augmentor=augmentor)
            loader = torch.utils.data.DataLoader(ds, batch_size=32)
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
            out[f"x_{vid}"] = (ds[0][0].numpy() * 255.0).round().astype(np.uint8)[:, ::8, ::8]   # probe of augmented frame 0 (RGB)
    np.savez_compressed(os.path.join(HERE, "emonet_golden.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
