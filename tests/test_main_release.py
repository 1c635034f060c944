"""main_release mirror: host-side data helpers (CPU) and an end-to-end 5-fold run on a synthetic
MER2023-format corpus (GPU)."""
import os
import random
import types

import numpy as np
import pytest
import torch

from mertools_b200 import main_release as MR


def _make_corpus(root, n_train=32, n_test=8, seed=0):
    rng = np.random.default_rng(seed)
    corp = {}
    for split, n in (("train", n_train), ("test1", n_test), ("test2", n_test), ("test3", n_test)):
        corp[f"{split}_corpus"] = {
            f"{split}_{i:04d}": {"emo": MR.EMOS_MER[int(rng.integers(0, 6))],
                                 "val": float(rng.uniform(-3, 3)) if i % 7 else ""}
            for i in range(n)}
    np.savez(os.path.join(root, "label-6way.npz"), **corp)
    feats = os.path.join(root, "features")
    for fname, frame_level in (("synA-UTT", False), ("synT-UTT", False), ("synV-FRA", True), ("synA-FRA", True),
                               ("synT-FRA", True)):
        os.makedirs(os.path.join(feats, fname))
        for split in corp.values():
            for name in split:
                hi = 40 if fname == "synA-FRA" else 6
                x = rng.standard_normal((int(rng.integers(2, hi)), 768) if frame_level else (768,)).astype(np.float32)
                np.save(os.path.join(feats, fname, name + ".npy"), x)
    cfg = types.SimpleNamespace(PATH_TO_LABEL={"MER2023": os.path.join(root, "label-6way.npz")},
                                PATH_TO_FEATURES={"MER2023": feats})
    return cfg


def test_label_and_feature_readers(tmp_path):
    cfg = _make_corpus(str(tmp_path))
    names, labels = MR.read_names_labels(cfg.PATH_TO_LABEL["MER2023"], "train")
    assert len(names) == 32 and set(labels[0]) == {"emo", "val"}
    assert labels[0]["val"] == -10 and 0 <= labels[1]["emo"] < 6          # '' -> -10 (mer2023.py:96-99)
    f = MR.read_utt_feature(os.path.join(cfg.PATH_TO_FEATURES["MER2023"], "synV-FRA"), names[0])
    raw = np.load(os.path.join(cfg.PATH_TO_FEATURES["MER2023"], "synV-FRA", names[0] + ".npy"))
    np.testing.assert_allclose(f, raw.mean(0))
    with pytest.raises(Exception, match="feature path or dir do not exist"):
        MR.read_utt_feature(cfg.PATH_TO_FEATURES["MER2023"], "nope")


def test_five_fold_split_matches_reference_rule():
    random.seed(3)
    folds = MR.random_split_indexes(32, 5)
    assert [len(e) for _, e in folds] == [6, 6, 6, 6, 8] and [len(t) for t, _ in folds] == [26, 26, 26, 26, 24]
    for tr, ev in folds:
        assert sorted(list(tr) + list(ev)) == list(range(32))
    random.seed(3)
    idx = np.arange(32)
    random.shuffle(idx)
    assert list(folds[0][1]) == list(idx[:6])                              # same RNG stream as the reference


@pytest.mark.gpu
def test_main_release_end_to_end(cuda, tmp_path):
    cfg = _make_corpus(str(tmp_path))
    hyper = tmp_path / "hyper.yaml"
    hyper.write_text("attention:\n  hidden_dim: 128\n  dropout: 0.3\n  grad_clip: -1.0\n  lr: 0.001\n")
    args = MR.build_parser().parse_args([
        "--audio_feature=synA-UTT", "--text_feature=synT-UTT", "--video_feature=synV-FRA",
        f"--hyper_path={hyper}", "--epochs=3", f"--save_root={tmp_path}/saved", "--gpu=0"])
    torch.manual_seed(0)
    random.seed(0)
    saved = MR.main(args, config=cfg)
    assert len(saved) == 4 and all(os.path.exists(p) for p in saved)
    assert os.path.basename(saved[0]).startswith("cv_features:synA-UTT+synT-UTT+synV-FRA_dataset:MER2023_model:attention+utt+None_f1:")
    stored = np.load(saved[0], allow_pickle=True)["args"].item()
    assert stored.hidden_dim == 128 and stored.audio_dim == 768 and stored.duration > 0


@pytest.mark.gpu
@pytest.mark.parametrize("feat_type", ["frm_align", "frm_unalign"])
def test_main_release_frame_level(cuda, tmp_path, feat_type):
    """--feat_type frm_align / frm_unalign: FRA features, feat_scale pooling / pre-padding, LSTM encoders."""
    cfg = _make_corpus(str(tmp_path))
    hyper = tmp_path / "hyper.yaml"
    hyper.write_text("attention:\n  hidden_dim: 64\n  dropout: 0.2\n  grad_clip: -1.0\n  lr: 0.001\n")
    args = MR.build_parser().parse_args([
        "--audio_feature=synA-FRA", "--text_feature=synT-FRA", "--video_feature=synV-FRA", f"--feat_type={feat_type}",
        f"--hyper_path={hyper}", "--epochs=2", f"--save_root={tmp_path}/saved", "--gpu=0"])
    torch.manual_seed(0)
    random.seed(0)
    saved = MR.main(args, config=cfg)
    assert len(saved) == 4 and all(os.path.exists(p) for p in saved)
    assert f"model:attention+{feat_type}+None_f1:" in os.path.basename(saved[0])
    stored = np.load(saved[0], allow_pickle=True)["args"].item()
    assert stored.feat_scale == (6 if feat_type == "frm_align" else 12) and stored.hidden_dim == 64
