"""Shared host-side pieces of the extractors: checkpoint loading and the .npy feature contract."""
from __future__ import annotations

import os

import numpy as np


def load_hf_state_dict(model_dir):
    """HF checkpoint directory -> {name: numpy fp32} without instantiating the HF model
    (the reference does ``AutoModel.from_pretrained(model_dir)``, e.g.
    extract_vision_huggingface.py:85-90; only the tensors are needed here)."""
    st = os.path.join(model_dir, "model.safetensors")
    if os.path.exists(st):
        from safetensors.numpy import load_file
        sd = load_file(st)
    else:
        import torch
        pt = os.path.join(model_dir, "pytorch_model.bin")
        assert os.path.exists(pt), f"no model.safetensors / pytorch_model.bin under {model_dir}"
        sd = {k: v.float().numpy() for k, v in torch.load(pt, map_location="cpu").items()}
    out = {}
    for k, v in sd.items():
        # AutoModel strips the task-model prefix ("vit.", "hubert.", "bert.", "roberta.", ...)
        for pre in ("vit.", "hubert.", "bert.", "roberta.", "wav2vec2.", "data2vec_audio.", "electra.", "videomae.", "wavlm.", "data2vec_vision.", "dinov2."):
            if k.startswith(pre):
                k = k[len(pre):]
                break
        out[k] = np.asarray(v, dtype=np.float32)
    return out


def save_feature(save_file, embeddings, feature_level, feature_dim):
    """The on-disk contract shared by the three scripts (extract_vision_huggingface.py:175-189,
    extract_text_huggingface.py:235-249): UTTERANCE -> 1-D [D] (mean over rows when 2-D),
    FRAME -> 2-D [T, D]; empty -> zeros (float64, as np.zeros in the reference)."""
    emb = np.array(embeddings).squeeze()
    if feature_level == "FRAME":
        if emb.size == 0:
            emb = np.zeros((1, feature_dim))
        elif emb.ndim == 1:
            emb = emb[np.newaxis, :]
    else:
        if emb.size == 0:
            emb = np.zeros((feature_dim,))
        elif emb.ndim == 2:
            emb = np.mean(emb, axis=0)
    if save_file is not None:
        np.save(save_file, emb)
    return emb
