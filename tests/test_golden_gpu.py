"""CUDA hot path vs the golden outputs of the unmodified reference scripts (tests/golden/*.npz),
through the public extractor API.  Tolerance 1e-3 relative (north_star)."""
import os

import numpy as np
import pytest
import torch

from mertools_b200 import synthetic as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / np.abs(b).max())


def test_visual_extractor_vs_reference_golden(cuda):
    from mertools_b200.extract.visual import VisualExtractor
    g = np.load(os.path.join(G, "visual_golden.npz"))
    clips = S.synth_frames(int(g["n_clips"]), 8, seed=int(g["seed"]))
    ext = VisualExtractor(S.vit_state_dict(seed=0), device=cuda)
    utt = ext.extract_clips(list(clips), "UTTERANCE", nframe=int(g["nframe"]))
    fra = ext.extract_clips(list(clips), "FRAME", nframe=int(g["nframe"]))
    for i in range(len(clips)):
        assert utt[i].shape == g[f"utt{i}"].shape and utt[i].dtype == g[f"utt{i}"].dtype
        assert _rel(utt[i], g[f"utt{i}"]) < TOL
        assert fra[i].shape == g[f"fra{i}"].shape and _rel(fra[i], g[f"fra{i}"]) < TOL


def test_visual_extractor_vs_reference_golden_112px(cuda):
    """112x112 face crops: resized on the device (bit-exact PIL bilinear), then the same path."""
    from mertools_b200.extract.visual import VisualExtractor
    g = np.load(os.path.join(G, "visual112_golden.npz"))
    clips = S.synth_frames(int(g["n_clips"]), 8, size=int(g["size"]), seed=int(g["seed"]))
    ext = VisualExtractor(S.vit_state_dict(seed=0), device=cuda)
    utt = ext.extract_clips(list(clips), "UTTERANCE", nframe=int(g["nframe"]))
    fra = ext.extract_clips(list(clips), "FRAME", nframe=None)
    for i in range(len(clips)):
        assert utt[i].shape == (768,) and utt[i].dtype == g[f"utt{i}"].dtype
        assert _rel(utt[i], g[f"utt{i}"]) < TOL
        assert fra[i].shape == (8, 768) and _rel(fra[i], g[f"fra{i}"]) < TOL


def test_audio_extractor_vs_reference_golden(cuda):
    from mertools_b200.extract.audio import AudioExtractor
    g = np.load(os.path.join(G, "audio_golden.npz"))
    waves = [S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
             for i, n in enumerate(g["lens"])]
    ext = AudioExtractor(S.hubert_state_dict(seed=1), device=cuda)
    utt = ext.extract_waves(waves, "UTTERANCE")
    fra = ext.extract_waves(waves, "FRAME")
    for i in range(len(waves)):
        assert utt[i].shape == (768,) and utt[i].dtype == g[f"utt{i}"].dtype
        assert _rel(utt[i], g[f"utt{i}"]) < TOL, f"clip {i}"
        assert _rel(fra[i][::8], g[f"fra{i}"]) < 2 * TOL, f"clip {i} frames"


def test_audio_extractor_vs_reference_golden_hubert_large_family(cuda):
    from mertools_b200.extract.audio import AudioExtractor
    g = np.load(os.path.join(G, "audio_large_golden.npz"))
    waves = [S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
             for i, n in enumerate(g["lens"])]
    ext = AudioExtractor(S.hubert_state_dict(seed=int(g["seed"]), layers=int(g["layers"]), large=True), device=cuda)
    utt = ext.extract_waves(waves, "UTTERANCE")
    fra = ext.extract_waves(waves, "FRAME")
    for i in range(len(waves)):
        assert utt[i].shape == (1024,) and utt[i].dtype == g[f"utt{i}"].dtype
        assert _rel(utt[i], g[f"utt{i}"]) < TOL, f"clip {i}"
        assert _rel(fra[i][::16], g[f"fra{i}"]) < 2 * TOL, f"clip {i} frames"


def test_text_extractor_vs_reference_golden(cuda):
    transformers = pytest.importorskip("transformers")
    from mertools_b200.extract.text import TextExtractor
    g = np.load(os.path.join(G, "text_golden.npz"))
    tok = transformers.BertTokenizer(os.path.join(G, "text_vocab.txt"))
    ext = TextExtractor(S.bert_state_dict(int(g["vocab_size"]), seed=2), tok, device=cuda)
    assert (ext.start, ext.end) == (int(g["start"]), int(g["end"]))
    sents = [str(s) if str(s) else None for s in g["sentences"]]
    for i, s in enumerate(sents):
        if s:
            np.testing.assert_array_equal(np.array(ext.tokenize(s)), g[f"ids{i}"])  # bit-exact ids
    utt = ext.extract_sentences(sents, "UTTERANCE")
    fra = ext.extract_sentences(sents, "FRAME")
    for i, s in enumerate(sents):
        assert utt[i].shape == g[f"utt{i}"].shape and utt[i].dtype == g[f"utt{i}"].dtype
        assert fra[i].shape == g[f"fra{i}"].shape
        if s:
            assert _rel(utt[i], g[f"utt{i}"]) < TOL and _rel(fra[i], g[f"fra{i}"]) < 2 * TOL
        else:
            assert not utt[i].any() and not fra[i].any()


def test_fusion_vs_reference_golden(cuda):
    """north_star: fusion-train step matching reference loss to 1e-3 (20 Adam steps, dropout 0)."""
    from mertools_b200.fusion import FusionNet
    g = np.load(os.path.join(G, "fusion_golden.npz"))
    net = FusionNet(device=cuda).load_state_dict(S.fusion_state_dict(seed=3))
    a, t, v, emo, val = S.synth_fusion_features(32, seed=7)
    dev = [torch.from_numpy(x).to(cuda) for x in (a, t, v, emo)] + [torch.from_numpy(val).view(-1, 1).to(cuda)]
    for step, ref in enumerate(g["losses"]):
        loss, eo, vo = net.train_step(*dev, lr=1e-3, weight_decay=1e-5)
        assert abs(float(loss[2].cpu()) - ref) <= 1e-3, f"step {step}: {float(loss[2])} vs {ref}"
        if step == 0:
            assert _rel(eo.cpu().numpy(), g["emos0"]) < 1e-4
            gv = net.named_views(net.grads)
            assert _rel(gv["fc_att.weight"].cpu().numpy(), g["grad_fc_att_w"]) < 1e-3
    assert _rel(net.named_views()["fc_out_1.weight"].cpu().numpy(), g["final_fc_out_1_w"]) < 1e-2


def test_frame_level_fusion_vs_reference_golden(cuda):
    """LSTM-encoder fusion (feat_type = frm_align) against the reference's own classes, 20 steps."""
    from mertools_b200.fusion import FusionNet
    g = np.load(os.path.join(G, "fusion_frm_golden.npz"))
    net = FusionNet(device=cuda, feat_type="frm_align").load_state_dict(
        S.fusion_state_dict(seed=int(g["seed"]), feat_type="frm_align"))
    a, t, v, emo, val = S.synth_fusion_sequences(int(g["batch"]), lens=tuple(int(x) for x in g["lens"]),
                                                 seed=int(g["data_seed"]))
    dev = [torch.from_numpy(x).to(cuda) for x in (a, t, v, emo, val.reshape(-1, 1))]
    for step, ref in enumerate(g["losses"]):
        loss3, eo, vo = net.train_step(*dev, lr=1e-3, weight_decay=1e-5, use_graph=False)
        got = float(loss3[2])
        assert abs(got - ref) <= 1e-3 * max(1.0, abs(ref)), f"step {step}: {got} vs {ref}"
        if step == 0:
            assert _rel(eo.cpu().numpy(), g["emos0"]) < 1e-4
            gv = net.named_views(net.grads)
            assert _rel(gv["audio_encoder.rnn.weight_hh_l0"][0].cpu().numpy(), g["grad_audio_whh_row0"]) < 1e-3
            assert _rel(gv["text_encoder.rnn.bias_ih_l0"].cpu().numpy(), g["grad_text_bih"]) < 1e-3


def test_logmel_examples_vs_reference_golden(cuda):
    """mer_logmel through the vggish_input mirror against the reference's numpy front-end."""
    from mertools_b200.extract import vggish_input as VI
    g = np.load(os.path.join(G, "logmel_golden.npz"))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        ex = VI.waveform_to_examples(w, 16000, float(g["hop_sec"]), device=cuda)
        assert tuple(ex.shape) == tuple(g[f"n{i}"]) and ex.dtype == np.float32
        sub = ex[:: max(1, len(ex) // 4)]
        ref = g[f"ex{i}"]
        assert np.abs(sub - ref).max() <= TOL * np.abs(ref).max(), f"clip {i}: {np.abs(sub - ref).max():.2e}"


def test_logmel_batch_matches_oracle(cuda):
    from mertools_b200.extract import vggish_input as VI
    from oracle import pipeline as P
    w = (S.synth_waves(3, 24000, seed=77).astype(np.float64) / 32768.0)
    got = VI.log_mel_spectrogram(torch.from_numpy(w.astype(np.float32)).to(cuda)).cpu().numpy()
    for r in range(3):
        ref = P.log_mel_spectrogram(w[r].astype(np.float32))
        assert got[r].shape == ref.shape == (148, 64)
        assert np.abs(got[r] - ref).max() <= TOL * np.abs(ref).max()


def test_attention_topn_vs_reference_golden(cuda):
    """Attention_TOPN (five features of different widths) against the reference's own class, 15 Adam steps."""
    from mertools_b200.fusion import TopnFusionNet
    g = np.load(os.path.join(G, "fusion_topn_golden.npz"))
    rng = np.random.default_rng(5000 + int(g["data_seed"]))
    dims, B = [int(d) for d in g["dims"]], int(g["batch"])
    feats = [rng.standard_normal((B, d), dtype=np.float32) for d in dims]
    emo, val = rng.integers(0, 6, B).astype(np.int64), rng.uniform(-3, 3, B).astype(np.float32)
    net = TopnFusionNet(dims, device=cuda).load_state_dict(S.fusion_topn_state_dict(dims, seed=int(g["seed"])))
    dev = [torch.from_numpy(f).to(cuda) for f in feats]
    demo, dval = torch.from_numpy(emo).to(cuda), torch.from_numpy(val.reshape(-1, 1)).to(cuda)
    for step, ref in enumerate(g["losses"]):
        loss3, eo, vo = net.train_step(dev, demo, dval, lr=1e-3, weight_decay=1e-5)
        got = float(loss3[2])
        assert abs(got - ref) <= 1e-3 * max(1.0, abs(ref)), f"step {step}: {got} vs {ref}"
        if step == 0:
            assert _rel(eo.cpu().numpy(), g["emos0"]) < 1e-4
            gv = net.named_views(net.grads)
            assert _rel(gv["fc_att.weight"].cpu().numpy(), g["grad_fc_att_w"]) < 1e-3
            assert _rel(gv["encoder3.linear_1.bias"].cpu().numpy(), g["grad_enc3_l1_b"]) < 1e-3
            # row 0 of attention_mlp.linear_1 is a dead ReLU unit in the reference run: its gradient is exactly 0
            assert np.abs(gv["attention_mlp.linear_1.weight"][0].cpu().numpy() - g["grad_attmlp_l1_w_row0"]).max() <= 1e-7
    f, e, v, inter = net.eval()({f"feat{i}": d for i, d in enumerate(dev)})
    assert f.shape == (B, 128) and e.shape == (B, 6) and int(inter) == 0
