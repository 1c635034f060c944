"""Golden fixture for the frame-level fusion path (feat_type = frm_align): the reference's OWN classes
toolkit/models/attention.py:Attention (with modules/encoder.py:LSTMEncoder), toolkit/utils/loss.py and
torch.optim.Adam, 20 training steps on a seeded [B, T, 768] batch (zero pre-padding as
read_data.py:pad_to_maxlen_pre_modality produces), dropout off.

Run once in the build container (needs /root/reference; NOT on the GPU box):
    python tests/golden/make_golden_frm.py
Writes tests/golden/fusion_frm_golden.npz.  Stub: `Tensor.cuda` as identity (CPU-only container).
"""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MERBench"
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

SEED, B, LENS, DATA_SEED = 5, 24, (9, 5, 12), 17


def main():
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        from toolkit.models.attention import Attention
        from toolkit.utils.loss import CELoss, MSELoss
    finally:
        os.chdir(cwd)
    args = types.SimpleNamespace(text_dim=768, audio_dim=768, video_dim=768, output_dim1=6, output_dim2=1,
                                 dropout=0.0, hidden_dim=128, grad_clip=-1.0, feat_type="frm_align")
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        net = Attention(args)
        sd = {k: torch.from_numpy(v) for k, v in S.fusion_state_dict(seed=SEED, feat_type="frm_align").items()}
        assert list(sd) == list(net.state_dict()), "state_dict order / names differ from the reference"
        net.load_state_dict(sd, strict=True)
        net.train()
        opt = torch.optim.Adam(net.parameters(), lr=1e-3, weight_decay=1e-5)
        a, t, v, emo, val = S.synth_fusion_sequences(B, lens=LENS, seed=DATA_SEED)
        batch = dict(audios=torch.from_numpy(a), texts=torch.from_numpy(t), videos=torch.from_numpy(v))
        ce_l, mse_l = CELoss(), MSELoss()
        losses = []
        for step in range(20):
            opt.zero_grad()
            feat, eo, vo, inter = net(batch)
            loss = inter + ce_l(eo, torch.from_numpy(emo)) + mse_l(vo, torch.from_numpy(val))
            loss.backward()
            if step == 0:
                g0 = {k: p.grad.detach().numpy().copy() for k, p in net.named_parameters()}
                out0 = (feat.detach().numpy().copy(), eo.detach().numpy().copy(), vo.detach().numpy().copy())
            opt.step()
            losses.append(float(loss.detach()))
    finally:
        torch.Tensor.cuda = orig_cuda
    np.savez(os.path.join(OUT, "fusion_frm_golden.npz"), seed=SEED, batch=B, lens=np.array(LENS), data_seed=DATA_SEED,
             losses=np.array(losses), feat0=out0[0], emos0=out0[1], vals0=out0[2],
             grad_audio_whh_row0=g0["audio_encoder.rnn.weight_hh_l0"][0],
             grad_text_bih=g0["text_encoder.rnn.bias_ih_l0"], grad_video_wih_row5=g0["video_encoder.rnn.weight_ih_l0"][5],
             grad_fc_att_w=g0["fc_att.weight"], final_fc_out_1_w=net.fc_out_1.weight.detach().numpy())
    print("frm fusion losses:", losses[:3], "...", losses[-1])


if __name__ == "__main__":
    main()
