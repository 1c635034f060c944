"""clock64 stamps of fus_rows_fast_kernel (block 0, thread 0): where the row kernel's time goes."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mertools_b200 import _lib as L  # noqa: E402
from mertools_b200 import synthetic as S  # noqa: E402
from mertools_b200.fusion import FusionNet  # noqa: E402

NAMES = ["entry", "inputs staged", "cluster.sync#0", "layer 1 done (9 chunks)", "sync#1", "resident tiles waited", "layer 2 done",
         "sync#2", "layers 3, a1, a2, a3 done (4 syncs inside)", "sync#6", "head done", "first scatter done", "sync#7",
         "end (4 more exchanges)"]
dev = "cuda:0"
for B in (4, 32):
    net = FusionNet(dropout=0.3, device=dev, seed=7).load_state_dict(S.fusion_state_dict(seed=3))
    a, t, v, emo, val = S.synth_fusion_features(B, seed=4)
    T = torch.from_numpy
    d = [T(a).to(dev), T(t).to(dev), T(v).to(dev), T(emo).to(dev), T(val).view(-1, 1).to(dev)]
    for _ in range(3):
        net.train_step(*d, weight_decay=1e-5, use_graph=False)
    buf = torch.zeros(32, dtype=torch.int64, device=dev)
    lib = L.lib()
    lib.mer_debug_fusion_trace.argtypes = [C.c_void_p]
    lib.mer_debug_fusion_trace(C.c_void_p(buf.data_ptr()))
    net.train_step(*d, weight_decay=1e-5, use_graph=False)
    torch.cuda.synchronize()
    lib.mer_debug_fusion_trace(None)
    t_ = buf.cpu().tolist()
    print(f"B={B}: " + "  ".join(f"{n}@{t_[i] - t_[0]}" for i, n in enumerate(NAMES) if t_[i]))
    print("   head: " + "  ".join(f"{n}@{t_[16 + i] - t_[0]}" for i, n in enumerate(["att", "outputs", "loss grads", "d fused + datt"]) if t_[16 + i]))
