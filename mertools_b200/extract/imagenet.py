"""ImageNet CNN extractor: mirror of MERBench/feature_extraction/visual/extract_imagenet_embedding.py.

Same flags (``--dataset --feature_level --gpu``), same input layout (``<face_dir>/<vid>/<vid>.npy`` read by
FaceDataset, dataset.py:12-47), same output directory ``imagenet_<UTT|FRA>`` and save rules (:73-94); the
torchvision resnet18 forward runs in libmer_b200.so.  The reference downloads the ImageNet weights through
torchvision (``resnet18(True)``); there is no network here, so the checkpoint is read from
``$MER_RESNET18_CKPT`` or torch hub's cache (``~/.cache/torch/hub/checkpoints/resnet18-*.pth``).
"""
from __future__ import annotations

import argparse
import glob
import os

import numpy as np
import torch

from . import common
from .visual import VisualExtractor, func_read_frames


def load_resnet18_state_dict():
    path = os.environ.get("MER_RESNET18_CKPT")
    if not path:
        hits = sorted(glob.glob(os.path.expanduser("~/.cache/torch/hub/checkpoints/resnet18-*.pth")))
        assert hits, "no resnet18 checkpoint: set MER_RESNET18_CKPT (torchvision's resnet18-f37072fd.pth)"
        path = hits[0]
    sd = torch.load(path, map_location="cpu")
    return {k: v.float().numpy() for k, v in sd.items() if not k.startswith("fc.")}


def main(params, config=None, state_dict=None, clips_per_launch=32, name="imagenet"):
    """``name``: "imagenet" (this script) or "msceleb" (extract/msceleb.py: the same network and transform with the
    MS-Celeb checkpoint, written to ``msceleb_<UTT|FRA>``)."""
    if config is None:
        from .. import config as config  # noqa: PLW0127
    print(f"==> Extracting {name} embedding...")
    face_dir = config.PATH_TO_RAW_FACE[params.dataset]
    save_dir = os.path.join(config.PATH_TO_FEATURES[params.dataset], f"{name}_{params.feature_level[:3]}")
    os.makedirs(save_dir, exist_ok=True)
    gpu = int(params.gpu)
    assert gpu != -1, "mertools_b200 has no CPU path"
    ext = VisualExtractor(state_dict if state_dict is not None else load_resnet18_state_dict(), device=f"cuda:{gpu}")
    vids = os.listdir(face_dir)
    print(f'Find total "{len(vids)}" videos.')
    for s in range(0, len(vids), clips_per_launch):
        chunk = vids[s:s + clips_per_launch]
        clips = [func_read_frames(face_dir, vid) for vid in chunk]
        files = [os.path.join(save_dir, f"{vid}.npy") for vid in chunk]
        ext.extract_clips(clips, params.feature_level, nframe=None, save_files=files)
        print(f"Processed {min(s + clips_per_launch, len(vids))}/{len(vids)} videos")


def build_parser():
    parser = argparse.ArgumentParser(description="Run.")
    parser.add_argument("--dataset", type=str, default="BoxOfLies", help="input dataset")
    parser.add_argument("--feature_level", type=str, default="UTTERANCE", help="feature level [FRAME or UTTERANCE]")
    parser.add_argument("--gpu", type=str, default="1", help="gpu id")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
