"""EmoNet CNN extractor: mirror of MERBench/feature_extraction/visual/extract_emonet_embedding.py.

Same flags (``--dataset --feature_level --gpu``, :36-41), input layout (``<face_dir>/<vid>/<vid>.npy`` through
FaceDatasetForEmoNet, dataset.py:50-86), output directory ``emonet_<UTT|FRA>`` (:45) and save rules (:76-94).
``DataAugmentor(256, 256)`` without augmentation (= cv2.resize to 256 x 256, bit-exact kernel), ``ToTensor`` and
``model(images, return_embedding=True)`` run in libmer_b200.so.  The checkpoint is the reference's
``<PRETRAINED>/emonet/emonet_8.pth`` (``module.`` prefixes stripped as in :50-51).
"""
from __future__ import annotations

import argparse
import os

import torch

from ..encoders import EmonetEncoder
from .ferplus import extract_video
from .visual import func_read_frames

CHECKPOINT = "emonet/emonet_8.pth"


def load_emonet_state_dict(path):
    checkpoint = torch.load(path, map_location="cpu")
    return {k.replace("module.", ""): (v.float().numpy() if v.is_floating_point() else v.numpy())
            for k, v in checkpoint.items()}


def main(params, config=None, state_dict=None):
    if config is None:
        from .. import config as config  # noqa: PLW0127
    print("==> Extracting emonet embedding...")
    face_dir = config.PATH_TO_RAW_FACE[params.dataset]
    save_dir = os.path.join(config.PATH_TO_FEATURES[params.dataset], f"emonet_{params.feature_level[:3]}")
    if not os.path.exists(save_dir):
        os.makedirs(save_dir)
    if state_dict is None:
        state_dict = load_emonet_state_dict(os.path.join(config.PATH_TO_PRETRAINED_MODELS, CHECKPOINT))
    enc = EmonetEncoder(state_dict, device=f"cuda:{int(str(params.gpu).split(',')[0])}")
    vids = os.listdir(face_dir)
    print(f'Find total "{len(vids)}" videos.')
    for i, vid in enumerate(vids, 1):
        print(f"Processing video '{vid}' ({i}/{len(vids)})...")
        extract_video(enc, func_read_frames(face_dir, vid), params.feature_level, os.path.join(save_dir, f"{vid}.npy"),
                      frames_per_launch=8)


def build_parser():
    parser = argparse.ArgumentParser(description="Run.")
    parser.add_argument("--dataset", type=str, default="MER2023", help="input dataset")
    parser.add_argument("--feature_level", type=str, default="UTTERANCE", help="feature level [FRAME or UTTERANCE]")
    parser.add_argument("--gpu", type=str, default="0", help="gpu id")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
