"""MS-Celeb CNN extractor: mirror of MERBench/feature_extraction/visual/extract_msceleb_embedding.py.

The script defines its own ``ResNet`` / ``BasicBlock`` (:21-117), which is torchvision's resnet18 under the same
parameter names, drops the fc layer (:149) and applies the ImageNet transform (:152-154): the network and the
preprocessing are those of the ImageNet extractor (``ResNet18Encoder``, libmer_b200.so), only the checkpoint
(``<PRETRAINED>/msceleb/resnet18_msceleb.pth``, entry ``state_dict``, loaded with ``strict=False``) and the output
directory ``msceleb_<UTT|FRA>`` (:139) differ.  Same flags (``--dataset --feature_level --gpu``, :135-138).
"""
from __future__ import annotations

import os

import torch

from . import imagenet

CHECKPOINT = "msceleb/resnet18_msceleb.pth"


def load_msceleb_state_dict(path):
    sd = torch.load(path, map_location="cpu")["state_dict"]
    return {k: v.float().numpy() for k, v in sd.items()
            if v.is_floating_point() and not k.startswith("fc.") and (k.startswith(("conv1.", "bn1.", "layer")))}


def main(params, config=None, state_dict=None, clips_per_launch=32):
    if config is None:
        from .. import config as config  # noqa: PLW0127
    if state_dict is None:
        state_dict = load_msceleb_state_dict(os.path.join(config.PATH_TO_PRETRAINED_MODELS, CHECKPOINT))
    imagenet.main(params, config=config, state_dict=state_dict, clips_per_launch=clips_per_launch, name="msceleb")


build_parser = imagenet.build_parser

if __name__ == "__main__":
    main(build_parser().parse_args())
