"""GPU parity of the audio (HuBERT) and text (BERT/RoBERTa) hot paths against the oracle.

Tolerance: 1e-3 relative (max-abs / max-ref and relative L2), TF32 tensor-core products with fp32
accumulation on the CUDA side, fp32 torch CPU on the oracle side."""
import numpy as np
import pytest
import torch

from mertools_b200 import synthetic as S
from oracle import encoders as E
from oracle import pipeline as P

pytestmark = pytest.mark.gpu
TOL = 1e-3


def rel(got, ref):
    got, ref = got.double(), ref.double()
    return float((got - ref).abs().max() / ref.abs().max()), float((got - ref).norm() / ref.norm())


@pytest.mark.parametrize("precision", ["f16", "bf16x3"])
@pytest.mark.parametrize("layers,n_samples,batch", [(4, 16000, 3), (12, 80000, 2), (4, 5000, 1), (4, 96000, 2)])
def test_hubert_hidden_states_and_readout(cuda, layers, n_samples, batch, precision):
    """Both operand formats of the 12 layers: fp16 (default; 96,000 samples = 299 frames takes the fp32-operand
    attention between fp16 GEMMs) and the bf16 (hi, lo) split."""
    from mertools_b200.encoders import HubertEncoder
    sd = S.hubert_state_dict(seed=1, layers=layers)
    wav = (S.synth_waves(batch, n_samples, seed=21).astype(np.float64) / 32768.0).astype(np.float32)
    enc = HubertEncoder(sd, device=cuda, stack_precision=precision)
    assert enc.stack_precision == precision and HubertEncoder(sd, device=cuda).stack_precision == "f16"
    assert enc.conv_precision == precision  # conv1 / conv2 follow the layers' operand format by default
    utt, frames, hidden = enc.forward(torch.from_numpy(wav).to(cuda), normalize=True,
                                      want_frames=True, return_hidden=True)
    torch.cuda.synchronize()
    iv = torch.from_numpy(np.stack([P.wav2vec2_normalize(w) for w in wav]))
    ref_hs = E.hubert_hidden_states(sd, iv, layers=layers)
    assert hidden.shape[2] == ref_hs[0].shape[1] == E.hubert_num_frames(n_samples)
    for l in range(layers + 1):
        m, l2 = rel(hidden[l].cpu(), ref_hs[l])
        assert m < 4 * TOL, f"hidden state {l}: max-rel {m:.2e} l2-rel {l2:.2e}"
    ref_sum = torch.stack(ref_hs)[[-4, -3, -2, -1]].sum(dim=0)
    m, l2 = rel(frames.cpu(), ref_sum)
    assert m < 2 * TOL and l2 < TOL, f"frame readout: max-rel {m:.2e} l2-rel {l2:.2e}"
    m, l2 = rel(utt.cpu(), ref_sum.mean(dim=1))
    print(f"HuBERT-base {precision} {layers} layers: utterance readout max-rel {m:.2e}")
    assert m < TOL and l2 < TOL, f"utterance readout: max-rel {m:.2e} l2-rel {l2:.2e}"


@pytest.mark.parametrize("precision", ["bf16x3", "f16"])
@pytest.mark.parametrize("layers,n_samples,batch", [(4, 16000, 2), (6, 80000, 2), (5, 100000, 1)])
def test_hubert_large_family_hidden_states_and_readout(cuda, layers, n_samples, batch, precision):
    """hubert-large / chinese-hubert-large style checkpoint: hidden 1024, 16 heads, LayerNorm after every
    conv, conv biases, stable (pre-LN) encoder; 100,000 samples = 312 frames exercises the long-sequence
    attention path."""
    from mertools_b200.encoders import HubertEncoder
    sd = S.hubert_state_dict(seed=7, layers=layers, large=True)
    wav = (S.synth_waves(batch, n_samples, seed=23).astype(np.float64) / 32768.0).astype(np.float32)
    enc = HubertEncoder(sd, device=cuda, stack_precision=precision)  # f16: opt-in fp16 stack for <= 249 frames
    assert enc.hidden == 1024
    utt, frames, hidden = enc.forward(torch.from_numpy(wav).to(cuda), normalize=True,
                                      want_frames=True, return_hidden=True)
    torch.cuda.synchronize()
    iv = torch.from_numpy(np.stack([P.wav2vec2_normalize(w) for w in wav]))
    ref_hs = E.hubert_hidden_states(sd, iv, layers=layers, heads=16)
    assert hidden.shape[2] == ref_hs[0].shape[1] == E.hubert_num_frames(n_samples)
    for l in range(layers + 1):
        m, l2 = rel(hidden[l].cpu(), ref_hs[l])
        assert m < 4 * TOL, f"hidden state {l}: max-rel {m:.2e} l2-rel {l2:.2e}"
    ref_sum = torch.stack(ref_hs)[[-4, -3, -2, -1]].sum(dim=0)
    m, l2 = rel(frames.cpu(), ref_sum)
    assert m < 2 * TOL and l2 < TOL, f"frame readout: max-rel {m:.2e} l2-rel {l2:.2e}"
    m, l2 = rel(utt.cpu(), ref_sum.mean(dim=1))
    assert m < TOL and l2 < TOL, f"utterance readout: max-rel {m:.2e} l2-rel {l2:.2e}"


def test_audio_extractor_matches_oracle_pipeline(cuda):
    """Public API, including a >10 s clip that the reference splits into 10 s rows."""
    from mertools_b200.extract import audio
    sd = S.hubert_state_dict(seed=1, layers=4)
    ext = audio.AudioExtractor(sd, device=cuda)
    waves = [S.synth_waves(1, n, seed=30 + i)[0].astype(np.float64) / 32768.0
             for i, n in enumerate((80000, 80000, 40000, 170000))]
    got = ext.extract_waves(waves, feature_level="UTTERANCE")
    for g, w in zip(got, waves):
        ref = P.audio_clip_features(sd, w, layers=4)
        assert g.shape == (768,) and g.dtype == np.float32
        m = np.abs(g - ref).max() / np.abs(ref).max()
        assert m < TOL, f"utterance feature max-rel {m:.2e}"
    gotf = ext.extract_waves(waves[2:], feature_level="FRAME")
    for g, w in zip(gotf, waves[2:]):
        ref = P.audio_clip_features(sd, w, layers=4, feature_level="FRAME")
        assert g.shape == ref.shape
        assert np.abs(g - ref).max() / np.abs(ref).max() < 2 * TOL


@pytest.mark.parametrize("precision", ["f16", "bf16x3"])
@pytest.mark.parametrize("roberta", [False, True])
def test_bert_hidden_states_and_readout(cuda, roberta, precision):
    from mertools_b200.encoders import BertEncoder
    layers, vocab = (12 if precision == "f16" else 4), 400
    if roberta:
        sd = S.bert_state_dict(vocab, seed=2, layers=layers, max_pos=514, type_vocab=1)
        kw = dict(ln_eps=1e-5, position_offset=2)
    else:
        sd = S.bert_state_dict(vocab, seed=2, layers=layers)
        kw = dict(ln_eps=1e-12, position_offset=0)
    rng = np.random.default_rng(5)
    sents = [rng.integers(0, vocab, n).tolist() for n in (3, 9, 64, 65, 17, 130, 5)]
    enc = BertEncoder(sd, device=cuda, precision=precision, **kw)
    utt, toks, hidden, cu = enc.forward(sents, start=1, end=-1, want_tokens=True, return_hidden=True)
    torch.cuda.synchronize()
    for i, s in enumerate(sents):
        ref_hs = E.bert_hidden_states(sd, s, layers=layers, eps=kw["ln_eps"],
                                      position_offset=kw["position_offset"])
        for l in range(layers + 1):
            m, l2 = rel(hidden[l, cu[i]:cu[i + 1]].cpu(), ref_hs[l][0])
            assert m < 4 * TOL, f"sentence {i} hidden {l}: max-rel {m:.2e}"
        ref = P.text_clip_features(sd, s, 1, -1, layers=layers, eps=kw["ln_eps"],
                                   position_offset=kw["position_offset"])
        m = np.abs(utt[i].cpu().numpy() - ref).max() / np.abs(ref).max()
        assert m < TOL, f"sentence {i} utterance feature max-rel {m:.2e}"
