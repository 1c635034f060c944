"""Golden fixture for Attention_TOPN: the reference's OWN class
MER2026/MER2026_Track1/toolkit/models/attention_topn.py:Attention_TOPN + the toolkit's losses + torch.optim.Adam,
15 training steps on seeded features of five different widths, dropout off.

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_topn.py
Writes tests/golden/fusion_topn_golden.npz.  The three reference files (attention_topn.py, modules/encoder.py,
utils/loss.py) are loaded by path (the toolkit's package __init__ imports matplotlib, absent here); stub:
`Tensor.cuda` as identity (CPU-only container).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MER2026/MER2026_Track1"
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

DIMS, SEED, B, DATA_SEED = (768, 1024, 512, 768, 128), 8, 20, 33


def data():
    rng = np.random.default_rng(5000 + DATA_SEED)
    feats = [rng.standard_normal((B, d), dtype=np.float32) for d in DIMS]
    return feats, rng.integers(0, 6, B).astype(np.int64), rng.uniform(-3, 3, B).astype(np.float32)


def load(name, path, package=None):
    import importlib.util
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=None)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    # import the three reference files directly (the package's __init__ chain pulls in matplotlib etc.)
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "toolkit", "models")]
    sys.modules["refmodels"] = pkg
    sub = types.ModuleType("refmodels.modules")
    sub.__path__ = [os.path.join(REF, "toolkit", "models", "modules")]
    sys.modules["refmodels.modules"] = sub
    load("refmodels.modules.encoder", os.path.join(REF, "toolkit", "models", "modules", "encoder.py"))
    Attention_TOPN = load("refmodels.attention_topn", os.path.join(REF, "toolkit", "models", "attention_topn.py")).Attention_TOPN
    lossmod = load("ref_loss", os.path.join(REF, "toolkit", "utils", "loss.py"))
    CELoss, MSELoss = lossmod.CELoss, lossmod.MSELoss
    args = types.SimpleNamespace(audio_dim=list(DIMS), output_dim1=6, output_dim2=1, dropout=0.0, hidden_dim=128,
                                 grad_clip=-1.0)
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        net = Attention_TOPN(args)
        sd = {k: torch.from_numpy(v) for k, v in S.fusion_topn_state_dict(DIMS, seed=SEED).items()}
        assert list(sd) == list(net.state_dict()), "state_dict order / names differ from the reference"
        net.load_state_dict(sd, strict=True)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
        feats, emo, val = data()
        batch = {f"feat{i}": torch.from_numpy(f) for i, f in enumerate(feats)}
        ce_l, mse_l = CELoss(), MSELoss()
        losses = []
        for step in range(15):
            opt.zero_grad()
            feat, eo, vo, inter = net(batch)
            loss = inter + ce_l(eo, torch.from_numpy(emo)) + mse_l(vo, torch.from_numpy(val))
            loss.backward()
            if step == 0:
                g0 = {k: p.grad.detach().numpy().copy() for k, p in net.named_parameters()}
                out0 = (feat.detach().numpy().copy(), eo.detach().numpy().copy())
            opt.step()
            losses.append(float(loss.detach()))
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez(os.path.join(OUT, "fusion_topn_golden.npz"), dims=np.array(DIMS), seed=SEED, batch=B, data_seed=DATA_SEED,
             losses=np.array(losses), feat0=out0[0], emos0=out0[1], grad_fc_att_w=g0["fc_att.weight"],
             grad_enc3_l1_b=g0["encoder3.linear_1.bias"], grad_attmlp_l1_w_row0=g0["attention_mlp.linear_1.weight"][0])
    print("topn losses:", losses[:3], "...", losses[-1])


if __name__ == "__main__":
    main()
