"""Diagnostic: where do FusionNet and the oracle trainer diverge?  (GPU box)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mertools_b200 import synthetic as S
from mertools_b200.fusion import FusionNet, param_names
from oracle import fusion as OF

sd = S.fusion_state_dict(seed=3)
net = FusionNet(device="cuda").load_state_dict(sd)
ref = OF.Trainer(sd, lr=1e-3, l2=1e-5)
a, t, v, emo, val = S.synth_fusion_features(32, seed=7)
T = torch.from_numpy
ca = (T(a), T(t), T(v), T(emo), T(val).view(-1, 1))
da = [x.cuda() for x in ca]
for step in range(30):
    pre = {k: p.detach().clone() for k, p in ref.sd.items()}
    ce, mse, tot, eo, vo, grads = ref.step(*ca)
    loss3, _, _ = net.train_step(*da, lr=1e-3, weight_decay=1e-5, use_graph=False)
    gv, pv = net.named_views(net.grads), net.named_views()
    worst = None
    for n in param_names():
        g, r = gv[n].cpu().double(), grads[n].double()
        rel = ((g - r).abs() / (r.abs() + 1e-12))
        sig = r.abs() > 1e-9
        relmax = rel[sig].max().item() if sig.any() else 0.0
        pd = (pv[n].cpu() - ref.sd[n].detach()).abs().max().item()
        if worst is None or pd > worst[1]:
            worst = (n, pd, relmax)
    if step in (0, 1, 2, 5, 10, 29):
        n = worst[0]
        g, r = gv[n].cpu().double(), grads[n].double()
        d = (pv[n].cpu() - ref.sd[n].detach()).abs()
        idx = np.unravel_index(int(d.argmax()), d.shape)
        print(f"step {step} loss {loss3[2].item():.6f}/{tot:.6f} worst param {n} maxdiff {worst[1]:.3e} "
              f"grad relmax(|r|>1e-9) {worst[2]:.3e} at {idx}: g={g[idx].item():.4e} r={r[idx].item():.4e} "
              f"p={pv[n].cpu()[idx].item():.6f} pr={ref.sd[n].detach()[idx].item():.6f}")
        if r.dim() == 2:
            row = idx[0]
            print("   row grads mine", g[row, :4].tolist(), "ref", r[row, :4].tolist(),
                  " |r| row max", r[row].abs().max().item())
