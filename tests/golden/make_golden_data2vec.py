"""Golden fixture for data2vec-audio: the UNMODIFIED reference function extract_audio_huggingface.py:extract run on a
`data2vec-audio-base-960h`-style checkpoint (Data2VecAudioModel: LayerNorm after every bias-free conv, a chain of
five k = 19 positional convs, post-LN layers; 4 layers to keep the run short).

Run once in the build container (needs /root/reference + transformers; NOT on the GPU box):
    python tests/golden/make_golden_data2vec.py
Writes tests/golden/audio_data2vec_golden.npz.  Same stubs as make_golden.py (`soundfile.read` via scipy,
patched `config`).
"""
import importlib.util
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MERBench"
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

LAYERS, SEED, SEED0 = 4, 12, 400
LENS = (80000, 36000)  # 5 s and 2.25 s


def main():
    import scipy.io.wavfile as wavfile
    from transformers import Data2VecAudioConfig, Data2VecAudioModel, Wav2Vec2FeatureExtractor
    work = tempfile.mkdtemp(prefix="mer_golden_d2v_")
    tools = os.path.join(work, "tools", "transformers")
    feats = os.path.join(work, "features")
    os.makedirs(tools)
    os.makedirs(feats)
    cfg = types.ModuleType("config")
    cfg.PATH_TO_RAW_AUDIO = {"MER2023": os.path.join(work, "audio")}
    cfg.PATH_TO_FEATURES = {"MER2023": feats}
    cfg.PATH_TO_PRETRAINED_MODELS = os.path.join(work, "tools")
    sys.modules["config"] = cfg
    sf = types.ModuleType("soundfile")

    def sf_read(path):
        sr, x = wavfile.read(path)
        return x.astype(np.float64) / 32768.0, sr
    sf.read = sf_read
    sys.modules["soundfile"] = sf
    name = "data2vec-audio-base-960h"
    adir = os.path.join(tools, name)
    m = Data2VecAudioModel(Data2VecAudioConfig(num_hidden_layers=LAYERS))
    sd = {k: torch.from_numpy(v) for k, v in S.hubert_state_dict(seed=SEED, layers=LAYERS, data2vec=True).items()}
    m.load_state_dict(sd, strict=True)
    m.save_pretrained(adir)
    Wav2Vec2FeatureExtractor(do_normalize=True).save_pretrained(adir)
    os.makedirs(cfg.PATH_TO_RAW_AUDIO["MER2023"])
    files = []
    for i, n in enumerate(LENS):
        f = os.path.join(cfg.PATH_TO_RAW_AUDIO["MER2023"], f"wav{i}.wav")
        wavfile.write(f, 16000, S.synth_waves(1, n, seed=SEED0 + i)[0])
        files.append(f)
    spec = importlib.util.spec_from_file_location(
        "ref_audio", os.path.join(REF, "feature_extraction", "audio", "extract_audio_huggingface.py"))
    ref_audio = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_audio)
    out = {}
    for level in ("UTTERANCE", "FRAME"):
        d = os.path.join(feats, f"{name}-{level[:3]}")
        os.makedirs(d, exist_ok=True)
        ref_audio.extract(name, files, d, level, gpu=-1)
        for i in range(len(LENS)):
            x = np.load(os.path.join(d, f"wav{i}.npy"))
            out[f"{level[:3].lower()}{i}"] = x if x.ndim == 1 else x[::16]  # FRAME: every 16th row (fixture size)
    np.savez(os.path.join(OUT, "audio_data2vec_golden.npz"), lens=np.array(LENS), seed=SEED, seed0=SEED0,
             layers=LAYERS, **out)
    print("audio data2vec:", {k: v.shape for k, v in out.items()})
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
