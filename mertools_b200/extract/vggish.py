"""VGGish audio extractor: mirror of MERBench/feature_extraction/audio/extract_vggish_embedding.py.

Same ``extract(audio_files, save_dir, feature_level, batch_size=2048)`` (:23-62), flags (``--gpu --feature_level
--dataset``, :65-70), output directory ``vggish_<FRA|UTT>`` (:82) and save rules (:52-59).  Log-mel examples come
from ``mer_logmel`` (vggish_input mirror), the network (vggish_slim.py) runs in ``mer_vggish_forward``.

Weights: the reference restores the TF-slim checkpoint ``<PRETRAINED>/vggish/vggish_model.ckpt`` with a TF
session; TensorFlow is not part of this stack, so the same variables are read from
``<PRETRAINED>/vggish/vggish_model.npz`` (keys = TF variable names ``vggish/conv1/weights`` ...; one-off export:
``r = tf.train.load_checkpoint(ckpt); np.savez(out, **{n: r.get_tensor(n) for n, _ in tf.train.list_variables(ckpt)})``)
or from the torchvggish port's ``vggish-10086976.pth`` in the same directory.
"""
from __future__ import annotations

import argparse
import glob
import os
import time

import numpy as np
import torch

from ..encoders import VggishEncoder
from . import vggish_input


def load_vggish_state_dict(model_dir):
    npz = os.path.join(model_dir, "vggish_model.npz")
    if os.path.exists(npz):
        with np.load(npz) as z:
            return {k: z[k] for k in z.files}
    pth = sorted(glob.glob(os.path.join(model_dir, "vggish*.pth")))
    assert pth, f"no VGGish weights in {model_dir}: expected vggish_model.npz (TF variable names) or vggish-*.pth"
    return {k: v.float().numpy() for k, v in torch.load(pth[0], map_location="cpu").items()}


def save_embeddings(csv_file, embeddings, feature_level):
    """:52-59."""
    if feature_level == "UTTERANCE":
        embeddings = np.array(embeddings).squeeze()
        if len(embeddings.shape) != 1:
            embeddings = np.mean(embeddings, axis=0)
    if csv_file is not None:
        np.save(csv_file, embeddings)
    return embeddings


def extract(audio_files, save_dir, feature_level, batch_size=2048, config=None, state_dict=None, device="cuda:0"):
    start_time = time.time()
    if feature_level == "FRAME":
        label_interval = 50.0
    if feature_level == "UTTERANCE":
        label_interval = 500.0
    if state_dict is None:
        if config is None:
            from .. import config as config  # noqa: PLW0127
        state_dict = load_vggish_state_dict(os.path.join(config.PATH_TO_PRETRAINED_MODELS, "vggish"))
    enc = VggishEncoder(state_dict, device=device)
    for i, audio_file in enumerate(audio_files, 1):
        print(f'Processing "{os.path.basename(audio_file)}" ({i}/{len(audio_files)})...')
        vid = os.path.basename(audio_file)[:-4]
        samples = vggish_input.wavfile_to_examples(audio_file, label_interval / 1000.0, device=device)
        if samples.shape[0] == 0:  # the reference's np.row_stack([]) raises here too
            raise ValueError("need at least one array to concatenate")
        examples = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float32)).to(device)
        embeddings = enc.embeddings(examples, max_examples=min(batch_size, 256)).cpu().numpy()  # (segment_num, 128)
        save_embeddings(os.path.join(save_dir, f"{vid}.npy"), embeddings, feature_level)
    print(f"Total time used: {time.time() - start_time:.1f}s.")


def build_parser():
    parser = argparse.ArgumentParser(description="Run.")
    parser.add_argument("--gpu", type=int, default=0, help="index of gpu")
    parser.add_argument("--feature_level", type=str, default="FRAME", help="feature_level: FRAME or UTTERANCE")
    parser.add_argument("--dataset", type=str, default="MER2023", help="input dataset")
    return parser


def main(args, config=None, state_dict=None):
    if config is None:
        from .. import config as config  # noqa: PLW0127
    audio_dir = config.PATH_TO_RAW_AUDIO[args.dataset]
    save_dir = config.PATH_TO_FEATURES[args.dataset]
    audio_files = glob.glob(os.path.join(audio_dir, "*.wav"))
    print(f'Find total "{len(audio_files)}" audio files.')
    save_dir = os.path.join(save_dir, f"vggish_{args.feature_level[:3]}")
    if not os.path.exists(save_dir):
        os.makedirs(save_dir)
    extract(audio_files, save_dir, args.feature_level, config=config, state_dict=state_dict, device=f"cuda:{args.gpu}")


if __name__ == "__main__":
    main(build_parser().parse_args())
