"""MA-Net CNN extractor: mirror of MERBench/feature_extraction/visual/extract_manet_embedding.py.

Same flags (``--dataset --feature_level --gpu``, :43-48), input layout (``<face_dir>/<vid>/<vid>.npy`` through
FaceDataset, dataset.py:12-47), output directory ``manet_<UTT|FRA>`` (:52) and save rules (:88-103).  The transform
(Resize((224, 224)) + ToTensor, :60-61) and ``model(images, return_embedding=True)`` run in libmer_b200.so.  The
checkpoint is the reference's ``<PRETRAINED>/manet/[02-08]-[21-19]-model_best-acc88.33.pth`` (``state_dict`` entry,
``module.`` prefixes stripped as in :56-58).
"""
from __future__ import annotations

import argparse
import os

import torch

from ..encoders import ManetEncoder
from .ferplus import extract_video
from .visual import func_read_frames

CHECKPOINT = "manet/[02-08]-[21-19]-model_best-acc88.33.pth"


def load_manet_state_dict(path):
    checkpoint = torch.load(path, map_location="cpu")
    return {k.replace("module.", ""): (v.float().numpy() if v.is_floating_point() else v.numpy())
            for k, v in checkpoint["state_dict"].items()}


def main(params, config=None, state_dict=None):
    if config is None:
        from .. import config as config  # noqa: PLW0127
    print("==> Extracting manet embedding...")
    face_dir = config.PATH_TO_RAW_FACE[params.dataset]
    save_dir = os.path.join(config.PATH_TO_FEATURES[params.dataset], f"manet_{params.feature_level[:3]}")
    if not os.path.exists(save_dir):
        os.makedirs(save_dir)
    if state_dict is None:
        state_dict = load_manet_state_dict(os.path.join(config.PATH_TO_PRETRAINED_MODELS, CHECKPOINT))
    enc = ManetEncoder(state_dict, device=f"cuda:{int(str(params.gpu).split(',')[0])}")
    vids = os.listdir(face_dir)
    print(f'Find total "{len(vids)}" videos.')
    for i, vid in enumerate(vids, 1):
        print(f"Processing video '{vid}' ({i}/{len(vids)})...")
        extract_video(enc, func_read_frames(face_dir, vid), params.feature_level, os.path.join(save_dir, f"{vid}.npy"))


def build_parser():
    parser = argparse.ArgumentParser(description="Run.")
    parser.add_argument("--dataset", type=str, default="BoxOfLies", help="input dataset")
    parser.add_argument("--feature_level", type=str, default="UTTERANCE", help="feature level [FRAME or UTTERANCE]")
    parser.add_argument("--gpu", type=str, default="1", help="gpu id")
    return parser


if __name__ == "__main__":
    main(build_parser().parse_args())
