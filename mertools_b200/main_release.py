"""Attention-fusion training CLI — B200 mirror of MERBench/main-release.py for
``--model attention --feat_type utt --dataset MER2023`` (the configuration of SURVEY.md §8 rows
a9-a12).  Same flags (:93-124), same hyper-parameter handling (model-tune.yaml ``attention`` grid or
``--hyper_path``, :159-165), same label / feature file formats (toolkit/dataloader/mer2023.py:82-104,
toolkit/utils/read_data.py:15-41,92-97), same 5-fold protocol (mer2023.py:108-134), metrics
(toolkit/utils/metric.py) and result files ``cv_*.npz`` / ``test{j}_*.npz`` (:256-272).

What changes is where the work happens: all features live on the GPU once, batches are index-selected
on the device, and every training step is one fused FusionNet.train_step (forward + CE/MSE + backward
+ Adam); evaluation passes run whole splits in one eval forward.
"""
from __future__ import annotations

import argparse
import os
import random
import time

import numpy as np
import torch

from .fusion import Adam, FusionNet

EMOS_MER = ["neutral", "angry", "happy", "sad", "worried", "surprise"]      # toolkit/globals.py:2
EMO2IDX = {e: i for i, e in enumerate(EMOS_MER)}
ATTENTION_GRID = dict(hidden_dim=[64, 128, 256], dropout=[0.2, 0.3, 0.4, 0.5], grad_clip=[-1.0],
                      lr=[1e-3, 1e-4])                                       # toolkit/model-tune.yaml:76-80


# ---- data (host side, same file formats as the reference) ---------------------------------------
def read_names_labels(label_path, data_type, debug=False):
    """mer2023.py:82-104."""
    assert data_type in ("train", "test1", "test2", "test3")
    corpus = np.load(label_path, allow_pickle=True)[f"{data_type}_corpus"].tolist()
    names, labels = [], []
    for name, label in corpus.items():
        names.append(name)
        val = label["val"] if ("val" in label and label["val"] != "") else -10
        labels.append({"emo": EMO2IDX[label["emo"]], "val": val})
    if debug:
        names, labels = names[:100], labels[:100]
    return names, labels


def read_utt_feature(feature_root, name):
    """read_data.py:15-41 + align_to_utt (:92-97): one clip -> [D] (mean over time when 2-D)."""
    path = os.path.join(feature_root, name + ".npy")
    d = os.path.join(feature_root, name)
    if os.path.exists(path):
        feat = np.load(path).squeeze()
    elif os.path.isdir(d):
        feat = np.array([np.load(os.path.join(d, f)) for f in sorted(os.listdir(d))]).squeeze()
    else:
        raise Exception("feature path or dir do not exist!")
    if feat.ndim == 1:
        feat = feat[np.newaxis, :]
    return np.mean(feat, axis=0)


def read_frm_feature(feature_root, name):
    """read_data.py:15-41 (func_read_one_feat): one clip -> [T, D] (a single vector becomes [1, D])."""
    path = os.path.join(feature_root, name + ".npy")
    d = os.path.join(feature_root, name)
    if os.path.exists(path):
        feat = np.load(path).squeeze()
    elif os.path.isdir(d):
        feat = np.array([np.load(os.path.join(d, f)) for f in sorted(os.listdir(d))]).squeeze()
    else:
        raise Exception("feature path or dir do not exist!")
    if feat.ndim == 1:
        feat = feat[np.newaxis, :]
    return feat


def random_split_indexes(whole_num, num_folder):
    """mer2023.py:108-134 (python ``random`` shuffle, last fold takes the remainder)."""
    indices = np.arange(whole_num)
    random.shuffle(indices)
    each = int(whole_num / num_folder)
    folds = [indices[each * i: each * (i + 1)] for i in range(num_folder - 1)] + [indices[each * (num_folder - 1):]]
    out = []
    for i in range(num_folder):
        train = [x for j in range(num_folder) if j != i for x in folds[j]]
        out.append([train, list(folds[i])])
    return out


class DeviceSplit:
    """One corpus split resident on the GPU: A/T/V fp32 -- [N, D] for feat_type 'utt', [N, T_m, D] for the
    frame-level types (shaped exactly as Data_Feat does, feat_data.py:33-44) -- emo int64, val fp32."""

    def __init__(self, args, names, labels, config, device):
        root = config.PATH_TO_FEATURES[args.dataset]
        feats = []
        if args.feat_type == "utt":
            for fname in (args.audio_feature, args.text_feature, args.video_feature):
                fr = os.path.join(root, fname)
                feats.append(np.stack([read_utt_feature(fr, n) for n in names]).astype(np.float32))
        else:
            from . import frame_features as FF
            raw = [[read_frm_feature(os.path.join(root, fname), n) for n in names]
                   for fname in (args.audio_feature, args.text_feature, args.video_feature)]
            shaped = FF.shape_split(raw[0], raw[1], raw[2], args.feat_type, args.feat_scale)
            feats = [np.array(x).astype(np.float32) for x in shaped]  # torch.FloatTensor(np.array(...)) in the collater
        self.names = names
        self.a, self.t, self.v = (torch.from_numpy(f).to(device) for f in feats)
        self.emo = torch.tensor([l["emo"] for l in labels], dtype=torch.int64, device=device)
        self.val = torch.tensor([l["val"] for l in labels], dtype=torch.float32, device=device).view(-1, 1)
        self.dims = tuple(int(f.shape[-1]) for f in feats)

    def __len__(self):
        return len(self.names)


def calculate_results(emo_probs, emo_labels, val_preds, val_labels):
    """mer2023.py:137-155."""
    from sklearn.metrics import accuracy_score, f1_score, mean_squared_error
    emo_preds = np.argmax(emo_probs, 1)
    acc = accuracy_score(emo_labels, emo_preds)
    f1 = f1_score(emo_labels, emo_preds, average="weighted")
    mse = mean_squared_error(val_labels, val_preds)
    res = dict(emoprobs=emo_probs, emolabels=emo_labels, emoacc=acc, emofscore=f1, valpreds=val_preds,
               vallabels=val_labels, valmse=mse)
    return res, f"f1:{f1:.4f}_acc:{acc:.4f}_val:{mse:.4f}"


def gain_metric_from_results(res, metric_name="emoval"):
    """toolkit/utils/metric.py:15-32."""
    if metric_name == "emoval":
        return res["emofscore"] - 0.25 * res["valmse"]
    if metric_name == "emo":
        return res["emofscore"]
    if metric_name == "val":
        return -res["valmse"]
    return -res["loss"]


# ---- one pass over a split (main-release.py:17-87) ------------------------------------------------
def run_split(args, net, split, idxs, optimizer=None, train=False, world_size=1):
    """idxs: sample indices of this pass.  Training draws them in the order a SubsetRandomSampler would
    (torch.randperm on the global generator), batch by batch, one fused step each."""
    dev = split.a.device
    idxs = torch.as_tensor(idxs, dtype=torch.int64)
    if train:
        idxs = idxs[torch.randperm(len(idxs))]
    names = [split.names[i] for i in idxs.tolist()]
    idxs = idxs.to(dev)
    emo_probs, val_preds, losses = [], [], []
    net.train(train)
    for s in range(0, len(idxs), args.batch_size):
        b = idxs[s:s + args.batch_size]
        a, t, v = split.a.index_select(0, b), split.t.index_select(0, b), split.v.index_select(0, b)
        emo, val = split.emo.index_select(0, b), split.val.index_select(0, b)
        if train:
            loss3, eo, vo = net.train_step(a, t, v, emo, val, lr=optimizer.lr, betas=optimizer.betas,
                                           eps=optimizer.eps, weight_decay=optimizer.weight_decay,
                                           world_size=world_size)
            losses.append(loss3[2:3].clone())
        else:
            _, eo, vo, _ = net({"audios": a, "texts": t, "videos": v})
            ce = torch.nn.functional.cross_entropy(eo, emo, reduction="sum") / len(eo)
            mse = torch.nn.functional.mse_loss(vo, val, reduction="sum") / len(vo)
            losses.append((ce + mse).view(1))
        emo_probs.append(eo.clone())
        val_preds.append(vo.clone())
    emo_probs = torch.cat(emo_probs).cpu().numpy()
    val_preds = torch.cat(val_preds).cpu().numpy()
    emo_labels = split.emo.index_select(0, idxs).cpu().numpy()
    val_labels = split.val.index_select(0, idxs).cpu().numpy()
    res, _ = calculate_results(emo_probs, emo_labels, val_preds, val_labels)
    return dict(names=names, loss=float(torch.cat(losses).mean().cpu()), **res)


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--dataset", type=str, default="MER2023")
    p.add_argument("--save_root", type=str, default="./saved")
    p.add_argument("--debug", action="store_true", default=False)
    p.add_argument("--audio_feature", type=str, default=None)
    p.add_argument("--text_feature", type=str, default=None)
    p.add_argument("--video_feature", type=str, default=None)
    p.add_argument("--feat_type", type=str, default="utt")
    p.add_argument("--feat_scale", type=int, default=None)
    p.add_argument("--e2e_name", type=str, default=None)
    p.add_argument("--hyper_path", type=str, default=None)
    p.add_argument("--model", type=str, default="attention")
    p.add_argument("--lr", type=float, default=None)
    p.add_argument("--l2", type=float, default=0.00001)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--num_workers", type=int, default=0)
    p.add_argument("--epochs", type=int, default=100)
    p.add_argument("--print_iters", type=int, default=1e8)
    p.add_argument("--gpu", default=0, type=int)
    return p


def main(args, config=None):
    if config is None:
        from . import config as config  # noqa: PLW0127
    assert args.model == "attention" and args.dataset == "MER2023", \
        "the B200 path covers --model attention --dataset MER2023 (SURVEY.md §8)"
    torch.cuda.set_device(args.gpu)
    device = torch.device("cuda", args.gpu)
    # pre-compression of the frame-level types (main-release.py:131-142)
    if args.feat_type == "utt":
        args.feat_scale = 1
    else:
        assert args.feat_type in ("frm_align", "frm_unalign"), args.feat_type
        for f in (args.audio_feature, args.text_feature, args.video_feature):
            assert f.endswith("FRA"), f"feat_type {args.feat_type} needs frame-level features, got {f}"
        args.feat_scale = 6 if args.feat_type == "frm_align" else 12
    feats = [f for f in (args.audio_feature, args.text_feature, args.video_feature) if f is not None]
    args.save_root = f"{args.save_root}-" + {0: "others", 1: "unimodal", 2: "bimodal", 3: "trimodal"}[len(set(feats))]
    if args.hyper_path is None:
        cfg = {k: v[random.randint(0, len(v) - 1)] for k, v in ATTENTION_GRID.items()}   # func_random_select
    else:
        import yaml
        cfg = yaml.safe_load(open(args.hyper_path))[args.model]
    for k, v in cfg.items():                                                               # merge_args_config
        if getattr(args, k, None) is None:
            setattr(args, k, v)
    args.output_dim1, args.output_dim2, args.metric_name = 6, 1, "emoval"
    print("args: ", args)
    save_resroot = os.path.join(args.save_root, "result")
    os.makedirs(save_resroot, exist_ok=True)
    os.makedirs(os.path.join(args.save_root, "model"), exist_ok=True)
    feature_name = "+".join(sorted(set(feats)))
    prefix_name = f"features:{feature_name}_dataset:{args.dataset}_model:{args.model}+{args.feat_type}+{args.e2e_name}"

    print("====== Reading Data =======")
    label_path = config.PATH_TO_LABEL[args.dataset]
    names, labels = read_names_labels(label_path, "train", args.debug)
    print(f"train: sample number {len(names)}")
    train_split = DeviceSplit(args, names, labels, config, device)
    folds = random_split_indexes(len(names), 5)
    tests = []
    for dt in ("test1", "test2", "test3"):
        n, l = read_names_labels(label_path, dt, args.debug)
        print(f"{dt}: sample number {len(n)}")
        tests.append(DeviceSplit(args, n, l, config, device))
    args.audio_dim, args.text_dim, args.video_dim = train_split.dims

    print("====== Training and Evaluation =======")
    folder_save, folder_duration = [], []
    name_time = time.time()
    for ii, (train_idxs, eval_idxs) in enumerate(folds):
        print(f">>>>> Cross-validation: training on the {ii + 1} folder >>>>>")
        start_time = name_time = time.time()
        net = FusionNet(args.audio_dim, args.text_dim, args.video_dim, args.hidden_dim, 6, 1,
                        dropout=args.dropout, grad_clip=args.grad_clip, device=device,
                        seed=random.randint(0, 2 ** 31 - 1), feat_type=args.feat_type)
        net.load_state_dict(default_init(net))
        optimizer = Adam(lr=args.lr, weight_decay=args.l2)
        whole_store, whole_metrics = [], []
        for epoch in range(args.epochs):
            epoch_store = {}
            train_res = run_split(args, net, train_split, train_idxs, optimizer, train=True)
            eval_res = run_split(args, net, train_split, eval_idxs)
            for k, v in eval_res.items():
                epoch_store[f"eval_{k}"] = v
            tm, em = (gain_metric_from_results(r, args.metric_name) for r in (train_res, eval_res))
            whole_metrics.append(em)
            print("epoch:%d; metric:%s; train results:%.4f; eval results:%.4f" % (epoch + 1, args.metric_name, tm, em))
            for jj, ts in enumerate(tests):
                res = run_split(args, net, ts, range(len(ts)))
                for k, v in res.items():
                    epoch_store[f"test{jj + 1}_{k}"] = v
            whole_store.append(epoch_store)
        best_index = int(np.argmax(np.array(whole_metrics)))
        folder_save.append(whole_store[best_index])
        folder_duration.append(time.time() - start_time)
        print(f">>>>> Finish: training on the {ii + 1}-th folder, best_index: {best_index}, "
              f"duration: {folder_duration[-1]} >>>>>")

    print("====== Prediction and Saving =======")
    args.duration = float(np.sum(folder_duration))
    f1 = np.mean([e["eval_emofscore"] for e in folder_save])
    acc = np.mean([e["eval_emoacc"] for e in folder_save])
    mse = np.mean([e["eval_valmse"] for e in folder_save])
    cv_result = f"f1:{f1:.4f}_acc:{acc:.4f}_val:{mse:.4f}"
    saved = []
    path = f"{save_resroot}/cv_{prefix_name}_{cv_result}_{name_time}.npz"
    np.savez_compressed(path, args=np.array(args, dtype=object))
    saved.append(path)
    for jj in range(len(tests)):
        emo_labels = folder_save[0][f"test{jj + 1}_emolabels"]
        emo_probs = np.mean(np.array([f[f"test{jj + 1}_emoprobs"] for f in folder_save]), axis=0)
        val_labels = folder_save[0][f"test{jj + 1}_vallabels"]
        val_preds = np.mean(np.array([f[f"test{jj + 1}_valpreds"] for f in folder_save]), axis=0)
        _, test_result = calculate_results(emo_probs, emo_labels, val_preds, val_labels)
        path = f"{save_resroot}/test{jj + 1}_{prefix_name}_{test_result}_{name_time}.npz"
        np.savez_compressed(path, args=np.array(args, dtype=object))
        saved.append(path)
    for p in saved:
        print(f"save results in {p}")
    return saved


def default_init(net):
    """nn.Linear default init (kaiming-uniform a=sqrt(5) == U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for both
    weight and bias), drawn from torch's global CPU generator as ``get_models(args)`` would."""
    sd = {}
    for name, shape in net.shapes.items():
        if ".rnn." in name:  # nn.LSTM.reset_parameters: every tensor U(-1/sqrt(hidden), 1/sqrt(hidden))
            fan_in = net.dims.hidden
        else:
            fan_in = shape[1] if len(shape) == 2 else net.shapes[name.replace(".bias", ".weight")][1]
        bound = 1.0 / np.sqrt(fan_in)
        sd[name] = (torch.rand(shape) * 2 - 1) * bound
    return sd


if __name__ == "__main__":
    main(build_parser().parse_args())
