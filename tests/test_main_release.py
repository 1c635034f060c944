"""main_release mirror: host-side data helpers (CPU) and an end-to-end 5-fold run on a synthetic
MER2023-format corpus (GPU)."""
import os
import random
import types

import numpy as np
import pytest
import torch

from mertools_b200 import main_release as MR


def _make_corpus(root, seed=0):
    from mertools_b200 import synthetic as S
    label_path, feats = S.write_mer2023_corpus(root, seed=seed, frame_level=True)
    return types.SimpleNamespace(PATH_TO_LABEL={"MER2023": label_path}, PATH_TO_FEATURES={"MER2023": feats})


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


def _metrics_of(name):
    """'..._f1:0.1533_acc:0.3167_val:18.1225' -> (prefix, [f1, acc, val])"""
    head, f1, acc, val = name.rsplit("_", 3)
    return head, [float(x.split(":")[1]) for x in (f1, acc, val)]


@pytest.mark.gpu
def test_main_release_reproduces_the_unmodified_reference_run(cuda, tmp_path):
    """tests/golden/main_release_golden.npz was recorded from ``runpy`` of the
    UNMODIFIED /root/reference/MERBench/main-release.py (tests/golden/make_golden_main_release.py: config C1, 32 + 3x8
    clips, 5 folds x 3 epochs, dropout 0, torch / random seeded with 0).  The mirror, seeded the same way, must make
    the same fold split, visit the samples in the same order (same sampler classes on the same generator), start
    from the same initial weights (torch's own nn.Linear constructors in the reference's order) and land on the same
    losses, best epochs, predictions and result-file names.  Floating point: 1e-3 relative (north_star)."""
    from mertools_b200 import synthetic as S
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "main_release_golden.npz"))
    label_path, feats = S.write_mer2023_corpus(str(tmp_path), seed=int(g["seed"]))
    cfg = types.SimpleNamespace(PATH_TO_LABEL={"MER2023": label_path}, PATH_TO_FEATURES={"MER2023": feats})
    hyper = tmp_path / "hyper.yaml"
    hyper.write_text(str(g["hyper"]))
    args = MR.build_parser().parse_args([
        "--audio_feature=synA-UTT", "--text_feature=synT-UTT", "--video_feature=synV-UTT",
        f"--hyper_path={hyper}", f"--epochs={int(g['epochs'])}", f"--save_root={tmp_path}/saved", "--gpu=0"])
    torch.manual_seed(int(g["seed"]))
    random.seed(int(g["seed"]))
    np.random.seed(int(g["seed"]))
    log = {}
    saved = MR.main(args, config=cfg, log=log)
    E = int(g["epochs"])
    # bit-exact parts: who is in which fold, and in which order the loaders visited them
    for f in range(5):
        assert log["eval_names"][f * E] == list(g[f"fold{f}_eval_names_epoch0"]), f"fold {f}: eval split / order"
        assert log["train_names"][f * E] == list(g[f"fold{f}_train_names_epoch0"]), f"fold {f}: train order"
    assert log["best_index"] == list(g["best_index"])
    # losses of every epoch of every fold
    rel = lambda a, b: np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max()  # noqa: E731
    tr, ev = np.array(log["train_loss"]).reshape(5, E), np.array(log["eval_loss"]).reshape(5, E)
    ts = np.array(log["test_loss"]).reshape(5, E, 3).transpose(2, 0, 1)
    print(f"main-release golden: train loss rel {rel(tr, g['train_loss']):.2e}, eval {rel(ev, g['eval_loss']):.2e}, "
          f"test {rel(ts, g['test_loss']):.2e}")
    assert rel(tr, g["train_loss"]) < 1e-3 and rel(ev, g["eval_loss"]) < 1e-3 and rel(ts, g["test_loss"]) < 1e-3
    # best-epoch predictions per fold
    for f, best in enumerate(log["folder_save"]):
        assert best["eval_names"] == list(g[f"fold{f}_eval_names"])
        for k in ("eval_emoprobs", "eval_valpreds", "test1_emoprobs", "test1_valpreds", "test3_emoprobs"):
            assert rel(best[k], g[f"fold{f}_{k}"]) < 1e-3, (f, k)
        assert abs(best["eval_emofscore"] - float(g[f"fold{f}_eval_emofscore"])) < 1e-12
        assert abs(best["eval_emoacc"] - float(g[f"fold{f}_eval_emoacc"])) < 1e-12
        assert abs(best["eval_valmse"] - float(g[f"fold{f}_eval_valmse"])) < 1e-3 * float(g[f"fold{f}_eval_valmse"])
    # result files: same names up to the wall-clock suffix (the val figure may differ in its last printed digit)
    assert len(saved) == 4
    for path, ref in zip(saved, g["saved"]):
        name = os.path.basename(path).rsplit("_", 1)[0]
        (h1, m1), (h2, m2) = _metrics_of(name), _metrics_of(str(ref))
        assert h1 == h2 and m1[:2] == m2[:2] and abs(m1[2] - m2[2]) <= 2e-4 * max(1.0, m2[2]), (name, str(ref))


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
