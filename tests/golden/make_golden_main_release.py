"""Golden of the UNMODIFIED trainer: ``runpy`` of /root/reference/MERBench/main-release.py on config C1
(32 train clips + 3 x 8 test clips, random 768-d A/T/V features, --model attention --feat_type utt,
hidden 128, lr 1e-3, l2 1e-5, batch 32, 5 folds) for 3 epochs with dropout 0 (the only RNG-free setting: the
reference draws its dropout masks from torch's generator).

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_main_release.py
Writes tests/golden/main_release_golden.npz: fold membership (eval names per fold, in the order the reference's
SubsetRandomSampler visited them), train / eval / test loss of every epoch of every fold, the printed metric
lines, the best epoch of each fold with its eval / test predictions, and the result-file names.

Harness only -- no reference source is copied or edited.  Stubs (CPU-only container): ``omegaconf``
(OmegaConf.load = yaml.safe_load), ``thop`` (unused import), ``torch.cuda.set_device / empty_cache``, ``Tensor.cuda``,
``Module.cuda`` as no-ops; ``multiprocessing.Pool`` -> a thread pool (the reference reads features with Pool(8);
forking a process that holds torch threads can hang).  Two pass-through observers record what the script computes:
``toolkit.utils.metric.gain_metric_from_results`` (called once per epoch with the train and eval results) and
``toolkit.utils.functions.func_update_storage`` (called with eval and test results).
"""
import contextlib
import io
import os
import random
import runpy
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MERBench"
OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 0
EPOCHS = 3
HYPER = "attention:\n  hidden_dim: 128\n  dropout: 0.0\n  grad_clip: -1.0\n  lr: 0.001\n"

from mertools_b200 import synthetic as S  # noqa: E402


def main():
    import yaml
    work = tempfile.mkdtemp(prefix="mer_golden_mr_")
    label_path, feats = S.write_mer2023_corpus(work, seed=SEED)
    hyper = os.path.join(work, "hyper.yaml")
    open(hyper, "w").write(HYPER)

    # ---- stubs for the CPU-only container ----
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = types.SimpleNamespace(load=lambda p: yaml.safe_load(open(p)))
    sys.modules["omegaconf"] = oc
    thop = types.ModuleType("thop")
    thop.profile = lambda *a, **k: (0, 0)
    sys.modules["thop"] = thop
    torch.cuda.set_device = lambda *a, **k: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    import multiprocessing
    import multiprocessing.pool
    multiprocessing.Pool = lambda processes=None: multiprocessing.pool.ThreadPool(processes)

    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    import config as ref_config  # the reference's own config module (paths patched below)
    ref_config.PATH_TO_LABEL["MER2023"] = label_path
    ref_config.PATH_TO_FEATURES["MER2023"] = feats

    # ---- observers ----
    import toolkit.utils.functions as F
    import toolkit.utils.metric as M
    seen = dict(train_loss=[], eval_loss=[], test_loss=[[], [], []], eval_names=[], train_names=[])
    gain, update = M.gain_metric_from_results, F.func_update_storage
    calls = [0]

    def gain_observer(res, metric_name="emoval"):
        key = "train" if calls[0] % 2 == 0 else "eval"
        calls[0] += 1
        seen[f"{key}_loss"].append(float(res["loss"]))
        seen[f"{key}_names"].append(list(res["names"]))
        return gain(res, metric_name)

    def update_observer(inputs, prefix, outputs):
        if prefix.startswith("test"):
            seen["test_loss"][int(prefix[4:]) - 1].append(float(inputs["loss"]))
        return update(inputs, prefix, outputs)

    M.gain_metric_from_results = gain_observer
    F.func_update_storage = update_observer

    sys.argv = ["main-release.py", "--model=attention", "--feat_type=utt", "--dataset=MER2023",
                "--audio_feature=synA-UTT", "--text_feature=synT-UTT", "--video_feature=synV-UTT",
                f"--hyper_path={hyper}", f"--epochs={EPOCHS}", f"--save_root={work}/saved", "--gpu=0"]
    torch.manual_seed(SEED)
    random.seed(SEED)
    np.random.seed(SEED)
    buf = io.StringIO()
    try:
        with contextlib.redirect_stdout(buf):
            g = runpy.run_path("main-release.py", run_name="__main__")
    finally:
        os.chdir(cwd)
    lines = buf.getvalue().splitlines()
    epoch_lines = [l for l in lines if l.startswith("epoch:")]
    saved = [os.path.basename(l.split("save results in ")[1]) for l in lines if l.startswith("save results in ")]
    name_time = str(g["name_time"])
    saved = [s.replace("_" + name_time + ".npz", "") for s in saved]  # drop the wall-clock suffix
    assert len(epoch_lines) == 5 * EPOCHS and len(saved) == 4, (len(epoch_lines), saved)

    out = dict(seed=SEED, epochs=EPOCHS, hyper=HYPER,
               train_loss=np.array(seen["train_loss"], np.float64).reshape(5, EPOCHS),
               eval_loss=np.array(seen["eval_loss"], np.float64).reshape(5, EPOCHS),
               test_loss=np.array(seen["test_loss"], np.float64).reshape(3, 5, EPOCHS),
               epoch_lines=np.array(epoch_lines), saved=np.array(saved),
               best_index=np.array([int(l.split("best_index: ")[1].split(",")[0]) for l in lines
                                    if "best_index:" in l]))
    for f in range(5):
        out[f"fold{f}_train_names_epoch0"] = np.array(seen["train_names"][f * EPOCHS])
        out[f"fold{f}_eval_names_epoch0"] = np.array(seen["eval_names"][f * EPOCHS])
        best = g["folder_save"][f]
        out[f"fold{f}_eval_names"] = np.array(best["eval_names"])
        for k in ("eval_emoprobs", "eval_valpreds", "test1_emoprobs", "test1_valpreds", "test3_emoprobs"):
            out[f"fold{f}_{k}"] = np.asarray(best[k], np.float32)
        for k in ("eval_emofscore", "eval_emoacc", "eval_valmse", "eval_loss"):
            out[f"fold{f}_{k}"] = float(best[k])
    np.savez_compressed(os.path.join(OUT, "main_release_golden.npz"), **out)
    print("\n".join(epoch_lines))
    print("\n".join(saved))
    print("train_loss", out["train_loss"])


if __name__ == "__main__":
    main()
