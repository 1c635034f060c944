"""Golden fixture for the Whisper branch: the UNMODIFIED reference function extract_audio_huggingface.py:extract run on a
`whisper-base`-shaped checkpoint (WhisperModel + WhisperFeatureExtractor; 2 + 2 layers and a 64-token vocabulary to keep
the fixture light).

Run once in the build container (needs /root/reference + transformers; NOT on the GPU box):
    python tests/golden/make_golden_whisper.py
Writes tests/golden/audio_whisper_golden.npz.  Same stubs as make_golden.py (`soundfile.read` via scipy,
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

LAYERS, SEED, SEED0, START = 2, 13, 500, 5
LENS = (80000, 36000)  # 5 s and 2.25 s


def main():
    import scipy.io.wavfile as wavfile
    from transformers import WhisperConfig, WhisperFeatureExtractor, WhisperModel
    work = tempfile.mkdtemp(prefix="mer_golden_wh_")
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
    name = "whisper-base"
    adir = os.path.join(tools, name)
    m = WhisperModel(WhisperConfig(vocab_size=64, d_model=512, encoder_layers=LAYERS, decoder_layers=LAYERS,
                                   encoder_attention_heads=8, decoder_attention_heads=8, encoder_ffn_dim=2048,
                                   decoder_ffn_dim=2048, decoder_start_token_id=START, pad_token_id=0, bos_token_id=1,
                                   eos_token_id=2))
    sd = {k: torch.from_numpy(v) for k, v in S.whisper_state_dict(seed=SEED, enc_layers=LAYERS, dec_layers=LAYERS).items()}
    m.load_state_dict(sd, strict=True)
    m.save_pretrained(adir)
    WhisperFeatureExtractor().save_pretrained(adir)
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
            out[f"{level[:3].lower()}{i}"] = x
    np.savez(os.path.join(OUT, "audio_whisper_golden.npz"), lens=np.array(LENS), seed=SEED, seed0=SEED0,
             layers=LAYERS, start=START, **out)
    print("audio whisper:", {k: v.shape for k, v in out.items()})
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
