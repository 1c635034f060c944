"""Generate the golden fixtures by running the UNMODIFIED reference code from /root/reference.

Run once in the build container (needs /root/reference + transformers; NOT on the GPU box):
    python tests/golden/make_golden.py
Writes tests/golden/{visual,audio,text,fusion}_golden.npz (a few KB each).  Inputs and weights are
regenerated deterministically by mertools_b200/synthetic.py, so only OUTPUTS (and token ids) are
stored.  Stubs, exactly those listed in SURVEY.md §8c: an empty `timm` module, `soundfile.read` via
scipy, a `config` module with patched paths.  No reference source is copied.
"""
import os
import runpy
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


def t(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


def main():
    import transformers
    from transformers import (BertConfig, BertModel, BertTokenizer, HubertConfig, HubertModel, ViTConfig,
                              ViTImageProcessor, ViTModel, Wav2Vec2FeatureExtractor)
    work = tempfile.mkdtemp(prefix="mer_golden_")
    tools = os.path.join(work, "tools", "transformers")
    feats = os.path.join(work, "features")
    os.makedirs(tools)
    os.makedirs(feats)

    # ---- config module the reference scripts import (`import config`) ----
    cfg = types.ModuleType("config")
    cfg.PATH_TO_RAW_FACE = {"MER2023": os.path.join(work, "openface_face")}
    cfg.PATH_TO_RAW_AUDIO = {"MER2023": os.path.join(work, "audio")}
    cfg.PATH_TO_TRANSCRIPTIONS = {"MER2023": os.path.join(work, "transcription.csv")}
    cfg.PATH_TO_FEATURES = {"MER2023": feats}
    cfg.PATH_TO_PRETRAINED_MODELS = os.path.join(work, "tools")
    sys.modules["config"] = cfg
    sys.modules["timm"] = types.ModuleType("timm")

    # =============================== visual ===============================
    vdir = os.path.join(tools, "dinov2-large")  # AutoModel dispatches on config.json: a ViT-B/16 under this name
    m = ViTModel(ViTConfig())
    m.load_state_dict(t(S.vit_state_dict(seed=0)), strict=True)
    m.save_pretrained(vdir)
    ViTImageProcessor().save_pretrained(vdir)
    clips = S.synth_frames(2, 8, seed=101)
    for i, c in enumerate(clips):
        d = os.path.join(cfg.PATH_TO_RAW_FACE["MER2023"], f"clip{i}")
        os.makedirs(d)
        np.save(os.path.join(d, f"clip{i}.npy"), c)
    for level in ("UTTERANCE", "FRAME"):
        sys.argv = ["extract_vision_huggingface.py", "--dataset=MER2023", "--model_name=dinov2-large",
                    f"--feature_level={level}", "--gpu=-1"]
        cwd = os.getcwd()
        os.chdir(os.path.join(REF, "feature_extraction", "visual"))
        try:
            runpy.run_path("extract_vision_huggingface.py", run_name="__main__")
        finally:
            os.chdir(cwd)
    vis = {}
    for i in range(2):
        vis[f"utt{i}"] = np.load(os.path.join(feats, "dinov2-large-UTT", f"clip{i}.npy"))
        vis[f"fra{i}"] = np.load(os.path.join(feats, "dinov2-large-FRA", f"clip{i}.npy"))
    np.savez(os.path.join(OUT, "visual_golden.npz"), seed=101, n_clips=2, nframe=64, **vis)
    print("visual:", {k: v.shape for k, v in vis.items()})

    # =============================== audio ===============================
    import scipy.io.wavfile as wavfile
    sf = types.ModuleType("soundfile")

    def sf_read(path):
        sr, x = wavfile.read(path)
        return x.astype(np.float64) / 32768.0, sr
    sf.read = sf_read
    sys.modules["soundfile"] = sf
    adir = os.path.join(tools, "chinese-hubert-base")
    m = HubertModel(HubertConfig())
    m.load_state_dict(t(S.hubert_state_dict(seed=1)), strict=True)
    m.save_pretrained(adir)
    Wav2Vec2FeatureExtractor(do_normalize=True).save_pretrained(adir)
    os.makedirs(cfg.PATH_TO_RAW_AUDIO["MER2023"])
    lens = (80000, 48000, 170000)  # 5 s, 3 s, and one > 10 s clip (split_into_batch path)
    files = []
    for i, n in enumerate(lens):
        w = S.synth_waves(1, n, seed=200 + i)[0]
        f = os.path.join(cfg.PATH_TO_RAW_AUDIO["MER2023"], f"wav{i}.wav")
        wavfile.write(f, 16000, w)
        files.append(f)
    spec = __import__("importlib.util").util.spec_from_file_location(
        "ref_audio", os.path.join(REF, "feature_extraction", "audio", "extract_audio_huggingface.py"))
    ref_audio = __import__("importlib.util").util.module_from_spec(spec)
    spec.loader.exec_module(ref_audio)
    aud = {}
    for level in ("UTTERANCE", "FRAME"):
        sd = os.path.join(feats, f"chinese-hubert-base-{level[:3]}")
        os.makedirs(sd, exist_ok=True)
        ref_audio.extract("chinese-hubert-base", files, sd, level, gpu=-1)
        for i in range(len(lens)):
            x = np.load(os.path.join(sd, f"wav{i}.npy"))
            aud[f"{level[:3].lower()}{i}"] = x if x.ndim == 1 else x[::8]  # FRAME: every 8th row (fixture size)
    np.savez(os.path.join(OUT, "audio_golden.npz"), lens=np.array(lens), seed0=200, **aud)
    print("audio:", {k: v.shape for k, v in aud.items()})

    # =============================== text ===============================
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, "dataset", "mer2023-dataset-process", "transcription-engchi-polish.csv"))
    chars = sorted(set("".join(str(s) for s in df["chinese"] if isinstance(s, str))))
    vocab = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + chars
    tdir = os.path.join(tools, "chinese-roberta-wwm-ext")
    os.makedirs(tdir)
    with open(os.path.join(tdir, "vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    tok = BertTokenizer(os.path.join(tdir, "vocab.txt"))
    tok.save_pretrained(tdir)
    m = BertModel(BertConfig(vocab_size=len(vocab)))
    m.load_state_dict(t(S.bert_state_dict(len(vocab), seed=2)), strict=True)
    m.save_pretrained(tdir)
    rows = df.iloc[[0, 1, 2, 3, 4, 5, 6, 7]].copy()
    rows.loc[rows.index[3], "chinese"] = np.nan  # the reference's empty-sentence branch
    rows.to_csv(cfg.PATH_TO_TRANSCRIPTIONS["MER2023"], index=False)
    spec = __import__("importlib.util").util.spec_from_file_location(
        "ref_text", os.path.join(REF, "feature_extraction", "text", "extract_text_huggingface.py"))
    ref_text = __import__("importlib.util").util.module_from_spec(spec)
    spec.loader.exec_module(ref_text)
    txt = {}
    for level in ("UTTERANCE", "FRAME"):
        ref_text.extract_embedding("chinese-roberta-wwm-ext", cfg.PATH_TO_TRANSCRIPTIONS["MER2023"], feats,
                                   level, gpu=-1)
        sd = os.path.join(feats, f"chinese-roberta-wwm-ext-{level[:3]}")
        for i, name in enumerate(rows["name"]):
            txt[f"{level[:3].lower()}{i}"] = np.load(os.path.join(sd, f"{name}.npy"))
    ids = {}
    for i, s in enumerate(rows["chinese"]):
        if isinstance(s, str):
            ids[f"ids{i}"] = np.array(tok(s)["input_ids"], dtype=np.int64)
    start, end = ref_text.find_start_end_pos(tok)
    np.savez(os.path.join(OUT, "text_golden.npz"), vocab_size=len(vocab), start=start, end=end,
             sentences=np.array([s if isinstance(s, str) else "" for s in rows["chinese"]]), **ids, **txt)
    with open(os.path.join(OUT, "text_vocab.txt"), "w", encoding="utf-8") as f:
        f.write("\n".join(vocab) + "\n")
    print("text:", len(vocab), "vocab;", {k: v.shape for k, v in txt.items()})

    # =============================== fusion ===============================
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from toolkit.models.attention import Attention
        from toolkit.utils.loss import CELoss, MSELoss
    finally:
        os.chdir(cwd)
    args = types.SimpleNamespace(text_dim=768, audio_dim=768, video_dim=768, output_dim1=6, output_dim2=1,
                                 dropout=0.0, hidden_dim=128, grad_clip=-1.0, feat_type="utt")
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self  # CPU-only container: `torch.tensor(0).cuda()` in forward
    try:
        net = Attention(args)
        net.load_state_dict(t(S.fusion_state_dict(seed=3)), strict=True)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
        a, tt, v, emo, val = S.synth_fusion_features(32, seed=7)
        batch = dict(audios=torch.from_numpy(a), texts=torch.from_numpy(tt), videos=torch.from_numpy(v))
        ce_l, mse_l = CELoss(), MSELoss()
        losses, first_grads = [], None
        for step in range(20):
            opt.zero_grad()
            feat, eo, vo, inter = net(batch)
            loss = inter + ce_l(eo, torch.from_numpy(emo)) + mse_l(vo, torch.from_numpy(val))
            loss.backward()
            if step == 0:
                first_grads = {k: p.grad.detach().numpy().copy() for k, p in net.named_parameters()}
                first_out = (feat.detach().numpy().copy(), eo.detach().numpy().copy(), vo.detach().numpy().copy())
            opt.step()
            losses.append(float(loss.detach()))
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez(os.path.join(OUT, "fusion_golden.npz"), losses=np.array(losses), feat0=first_out[0],
             emos0=first_out[1], vals0=first_out[2],
             grad_fc_att_w=first_grads["fc_att.weight"], grad_audio_l1_b=first_grads["audio_encoder.linear_1.bias"],
             final_fc_out_1_w=net.fc_out_1.weight.detach().numpy(),
             final_audio_l1_w_row0=net.audio_encoder.linear_1.weight.detach().numpy()[0])
    print("fusion losses:", losses[:3], "...", losses[-1])
    shutil.rmtree(work, ignore_errors=True)
    print("transformers", transformers.__version__, "torch", torch.__version__)


if __name__ == "__main__":
    main()
