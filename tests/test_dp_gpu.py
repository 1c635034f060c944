"""Data-parallel fusion training on the real kernels: two ranks against one.

Two processes, NCCL when two GPUs are visible, otherwise gloo with both ranks driving cuda:0 (the collective then goes
through the host; the kernels, the rank-strided batch split and the step are the same).  W = 2 on rows rank::2 of a
26-row batch must equal W = 1 on the whole batch: same loss (the all-reduced one), same parameters after 4 steps."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, backend, out):
    import torch.distributed as dist

    from mertools_b200 import synthetic as S
    from mertools_b200.fusion import FusionNet
    dev = torch.device("cuda", rank if torch.cuda.device_count() >= world else 0)
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    sd = S.fusion_state_dict(seed=3)
    a, t, v, emo, val = S.synth_fusion_features(26, seed=12)
    T = torch.from_numpy
    full = [T(a).to(dev), T(t).to(dev), T(v).to(dev), T(emo).to(dev), T(val).view(-1, 1).to(dev)]
    mine = [x[rank::world].contiguous() for x in full]
    net = FusionNet(dropout=0.0, device=dev)
    if rank == 0:
        net.load_state_dict(sd)
    net.broadcast_from(0)                      # rank 1 starts from zeros: the broadcast makes it a replica
    losses = []
    for _ in range(4):
        loss, eo, _ = net.train_step(*mine, lr=1e-3, weight_decay=1e-5, world_size=world, global_batch=26)
        losses.append(loss.clone())
    torch.cuda.synchronize()
    sig = torch.stack([net.params.double().sum(), net.params.double().abs().sum()])
    lo, hi = sig.clone(), sig.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_replicas = bool(torch.equal(lo, hi))
    if rank == 0:
        ref = FusionNet(dropout=0.0, device=dev).load_state_dict(sd)
        ref_losses = []
        for _ in range(4):
            l, _, _ = ref.train_step(*full, lr=1e-3, weight_decay=1e-5, use_graph=False)
            ref_losses.append(l.clone())
        dl = max(float((x - y).abs().max()) for x, y in zip(losses, ref_losses))
        dp = float((net.params - ref.params).norm() / ref.params.norm())
        out["loss_diff"], out["param_rel"], out["same_replicas"] = dl, dp, same_replicas
        out["loss0"] = float(ref_losses[0][2])
    # a rank that holds different parameters must be caught
    if rank == 1:
        net.params[0] += 1.0
    net._replicas_checked = False
    try:
        net.check_replicas(world)
        caught = False
    except RuntimeError:
        caught = True
    out[f"caught{rank}"] = caught
    dist.destroy_process_group()


def test_two_ranks_equal_one_rank_on_the_concatenated_batch(cuda):
    import torch.multiprocessing as mp
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, backend, out), nprocs=2, join=True)
    out = dict(out)
    print(f"data-parallel fusion ({backend}): {out}")
    assert out["same_replicas"] and out["caught0"] and out["caught1"]
    assert out["loss_diff"] <= 1e-5 * max(1.0, out["loss0"]) and out["param_rel"] <= 1e-5
