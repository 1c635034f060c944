"""Latency of one Attention-fusion training step ('latency-bound; report us/step & clips/s').

CUDA-graph replay of FusionNet.train_step (forward + CE/MSE + backward + Adam = fus_rows_kernel + fus_wgrad_kernel),
timed with CUDA events over `--iters` back-to-back replays; the reference's CPU step (the oracle trainer: the torch
ops main-release.py executes) beside it.  One JSON line per batch size."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def fusion_step_us(B, iters=200, dropout=0.3, hidden=128, device="cuda:0", profile=False):
    from mertools_b200 import synthetic as S
    from mertools_b200.fusion import FusionNet
    net = FusionNet(hidden_dim=hidden, dropout=dropout, device=device, seed=7)
    net.load_state_dict(S.fusion_state_dict(seed=3, hidden=hidden))
    a, t, v, emo, val = S.synth_fusion_features(B, seed=4)
    T = torch.from_numpy
    dev = [T(a).to(device), T(t).to(device), T(v).to(device), T(emo).to(device), T(val).view(-1, 1).to(device)]
    for _ in range(5):
        net.train_step(*dev, lr=1e-3, weight_decay=1e-5)
    key = next(iter(net._graphs))
    graph = net._graphs[key][0]
    torch.cuda.synchronize()
    if profile:  # `ncu --profile-from-start off`: exactly one replay (fus_rows_fast_kernel + fus_wgrad_kernel) is captured
        torch.cuda.cudart().cudaProfilerStart()
        graph.replay()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        graph.replay()
    e1.record()
    torch.cuda.synchronize()
    replay = e0.elapsed_time(e1) * 1e3 / iters
    e0.record()
    for _ in range(iters):
        net.train_step(*dev, lr=1e-3, weight_decay=1e-5)  # + the five input copies into the static buffers
    e1.record()
    torch.cuda.synchronize()
    return dict(batch=B, hidden=hidden, graph_replay_us=round(replay, 2),
                train_step_call_us=round(e0.elapsed_time(e1) * 1e3 / iters, 2),
                clips_per_s=round(B / (replay * 1e-6)), kernels_per_step=net._graphs[key][2])


def reference_cpu_step_us(B, iters=20):
    from mertools_b200 import synthetic as S
    from oracle import fusion as OF
    tr = OF.Trainer(S.fusion_state_dict(seed=3), lr=1e-3, l2=1e-5, dropout=0.0)
    a, t, v, emo, val = S.synth_fusion_features(B, seed=4)
    T = torch.from_numpy
    args = (T(a), T(t), T(v), T(emo), T(val).view(-1, 1))
    tr.step(*args)
    t0 = time.perf_counter()
    for _ in range(iters):
        tr.step(*args)
    return (time.perf_counter() - t0) / iters * 1e6


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, nargs="+", default=[32, 256])
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--profile", action="store_true", help="bracket one graph replay with cudaProfilerStart / Stop")
    a = ap.parse_args()
    for B in a.batches:
        r = fusion_step_us(B, a.iters, profile=a.profile)
        if not a.no_cpu:
            r["reference_cpu_step_us"] = round(reference_cpu_step_us(B), 1)
        print(json.dumps(r), flush=True)
