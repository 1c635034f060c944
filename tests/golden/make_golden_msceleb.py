"""Golden outputs of the reference's MS-Celeb extractor (MERBench/feature_extraction/visual/extract_msceleb_embedding.py),
UNMODIFIED pieces: its own ``ResNet`` / ``BasicBlock`` classes (:21-117) with our seeded synthetic checkpoint loaded as the
script does (``load_state_dict(..., strict=False)``, fc dropped by ``nn.Sequential(*children[:-1])``, :144-149), its
transform (:152-154), the reference ``FaceDataset`` and the ``squeeze`` / save rules (:125-199).
Stubs: ``skimage`` (imported by dataset.py) and ``config``.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_msceleb.py
"""
import importlib.util
import os
import sys
import tempfile
import types

import numpy as np
import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from mertools_b200 import synthetic as S  # noqa: E402

VIS = "/root/reference/MERBench/feature_extraction/visual"
SEED = 6


def golden_clips():
    return {"vidA": S.synth_frames(1, 3, seed=91)[0], "vidB": S.synth_frames(1, 1, size=112, seed=92)[0]}


def main():
    sk = types.ModuleType("skimage")
    sk.io, sk.img_as_float = types.ModuleType("skimage.io"), (lambda x: x)
    sys.modules["skimage"], sys.modules["skimage.io"] = sk, sk.io
    sys.modules["config"] = types.ModuleType("config")
    sys.path.insert(0, VIS)
    spec = importlib.util.spec_from_file_location("extract_msceleb_embedding", os.path.join(VIS, "extract_msceleb_embedding.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    from torchvision import transforms

    sd = S.resnet18_state_dict(SEED)
    model = ref.ResNet(block=ref.BasicBlock, n_blocks=[2, 2, 2, 2], channels=[64, 128, 256, 512], output_dim=1000)
    missing = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert sorted(missing.missing_keys) == ["fc.bias", "fc.weight"] and not missing.unexpected_keys, missing
    model = nn.Sequential(*list(model.children())[:-1]).eval()
    transform = transforms.Compose([transforms.Resize((224, 224)), transforms.ToTensor(),
                                    transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
    out = {"seed": SEED, "names": np.array(list(golden_clips()))}
    with tempfile.TemporaryDirectory() as tmp:
        for vid, frames in golden_clips().items():
            os.makedirs(os.path.join(tmp, vid))
            np.save(os.path.join(tmp, vid, f"{vid}.npy"), frames)
            loader = torch.utils.data.DataLoader(ref.FaceDataset(vid, tmp, transform=transform), batch_size=32)
            feats, names = [], []
            with torch.no_grad():
                for images, ids in loader:
                    feats.append(model(images).squeeze().cpu().detach().numpy())       # extract() :125-131
                    names.extend(ids)
            emb = np.vstack(feats)[np.argsort(np.array(names))]
            fra = np.array(emb).squeeze()
            out[f"fra_{vid}"] = fra[np.newaxis, :] if len(fra.shape) == 1 else fra
            utt = np.array(emb).squeeze()
            out[f"utt_{vid}"] = np.mean(utt, axis=0) if len(utt.shape) == 2 else utt
    np.savez_compressed(os.path.join(HERE, "msceleb_golden.npz"), **out)
    for k, v in out.items():
        print(k, getattr(v, "shape", v))


if __name__ == "__main__":
    main()
