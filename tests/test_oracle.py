"""Pin the oracle (oracle/) against the HF classes the reference extractors instantiate.

CPU only.  The reference reaches its arithmetic through transformers.AutoModel; the golden-vector
tests (test_golden.py) additionally pin the oracle to outputs of the unmodified reference scripts.
"""
import numpy as np
import pytest
import torch

from mertools_b200 import synthetic as S
from oracle import encoders as E
from oracle import pipeline as P

transformers = pytest.importorskip("transformers")


def _load(model, sd):
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    return model.eval()


def _maxrel(a, b):
    return float((a - b).abs().max() / b.abs().max())


def test_vit_oracle_matches_hf():
    sd = S.vit_state_dict(seed=0, layers=2, scale=2.0)
    m = _load(transformers.ViTModel(transformers.ViTConfig(num_hidden_layers=2)), sd)
    x = P.vit_preprocess(S.synth_frames(1, 2, seed=3)[0])
    with torch.no_grad():
        ref = m(x, output_hidden_states=True).hidden_states
    got = E.vit_hidden_states(sd, x, layers=2)
    assert len(ref) == len(got) == 3
    for r, g in zip(ref, got):
        assert _maxrel(g, r) < 2e-5


def test_hubert_oracle_matches_hf():
    sd = S.hubert_state_dict(seed=1, layers=2, scale=2.0)
    m = _load(transformers.HubertModel(transformers.HubertConfig(num_hidden_layers=2)), sd)
    wav = S.synth_waves(2, 16000, seed=4).astype(np.float64) / 32768.0
    iv = torch.from_numpy(np.stack([P.wav2vec2_normalize(w) for w in wav]))
    with torch.no_grad():
        ref = m(iv, output_hidden_states=True).hidden_states
    got = E.hubert_hidden_states(sd, iv, layers=2)
    assert ref[0].shape[1] == E.hubert_num_frames(16000) == 49
    for r, g in zip(ref, got):
        assert _maxrel(g, r) < 5e-5


def test_wav2vec2_normalize_matches_hf():
    fe = transformers.Wav2Vec2FeatureExtractor(do_normalize=True)
    wav = S.synth_waves(1, 8000, seed=5)[0].astype(np.float64) / 32768.0
    ref = fe(wav, sampling_rate=16000, return_tensors="pt").input_values[0].numpy()
    got = P.wav2vec2_normalize(wav)
    assert ref.dtype == got.dtype == np.float32
    np.testing.assert_array_equal(ref, got)


def test_bert_oracle_matches_hf():
    sd = S.bert_state_dict(300, seed=2, layers=2, scale=2.0)
    m = _load(transformers.BertModel(transformers.BertConfig(num_hidden_layers=2, vocab_size=300)), sd)
    ids = torch.tensor([[5, 17, 250, 3, 99, 42, 7]])
    with torch.no_grad():
        ref = m(input_ids=ids, output_hidden_states=True).hidden_states
    got = E.bert_hidden_states(sd, ids, layers=2)
    for r, g in zip(ref, got):
        assert _maxrel(g, r) < 2e-5


def test_roberta_oracle_matches_hf():
    sd = S.bert_state_dict(300, seed=2, layers=2, max_pos=514, type_vocab=1)
    cfg = transformers.RobertaConfig(num_hidden_layers=2, vocab_size=300, max_position_embeddings=514,
                                     type_vocab_size=1, layer_norm_eps=1e-5)
    m = _load(transformers.RobertaModel(cfg), sd)
    ids = torch.tensor([[0, 17, 250, 3, 99, 42, 2]])
    with torch.no_grad():
        ref = m(input_ids=ids, output_hidden_states=True).hidden_states
    got = E.bert_hidden_states(sd, ids, layers=2, eps=1e-5, position_offset=2)
    for r, g in zip(ref, got):
        assert _maxrel(g, r) < 2e-5


def test_vit_image_processor_matches_oracle_preprocess():
    proc = transformers.ViTImageProcessor()
    frames = S.synth_frames(1, 2, seed=6)[0]
    from PIL import Image
    pil = [Image.fromarray(np.ascontiguousarray(f[..., ::-1])) for f in frames]
    ref = proc(images=pil, return_tensors="pt")["pixel_values"]
    got = P.vit_preprocess(frames)
    assert float((ref - got).abs().max()) <= 2.4e-7


@pytest.mark.parametrize("hw", [(112, 112), (300, 260), (57, 91), (448, 448), (225, 223), (1, 1), (500, 30)])
def test_pil_resize_restatement_is_bit_exact(hw):
    """oracle/pipeline.py:pil_resize_bilinear_u8 against Pillow itself (the resize inside HF ViTImageProcessor)."""
    from PIL import Image
    from oracle import pipeline as P
    rng = np.random.default_rng(hw[0] * 1000 + hw[1])
    img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((224, 224), resample=Image.BILINEAR))
    assert np.array_equal(P.pil_resize_bilinear_u8(img), ref)


def test_vit_preprocess_matches_hf_processor_on_upscaled_faces():
    """112x112 -> 224x224: the oracle's preprocessing equals HF ViTImageProcessor (5.5.0 here) to fp32 rounding."""
    from PIL import Image
    from transformers import ViTImageProcessor
    from oracle import pipeline as P
    rng = np.random.default_rng(5)
    bgr = rng.integers(0, 256, (2, 112, 112, 3), dtype=np.uint8)
    ref = ViTImageProcessor()(images=[Image.fromarray(f[..., ::-1].copy()) for f in bgr], return_tensors="pt")["pixel_values"]
    got = P.vit_preprocess(bgr)
    assert float((got - ref).abs().max()) < 2e-7


def test_oracle_audio_restatement_also_is_wav2vec2_base():
    """wav2vec2-base-960h / chinese-wav2vec2-base (extract_audio_huggingface.py:20,28; group-norm feature
    extractor, post-LN encoder) have HuBERT-base's parameter names and arithmetic: the same restatement
    -- hence the same CUDA path -- serves both model families."""
    from transformers import HubertConfig, HubertModel, Wav2Vec2Config, Wav2Vec2Model
    torch.manual_seed(3)
    cfg = Wav2Vec2Config(num_hidden_layers=2)
    assert cfg.feat_extract_norm == "group" and not cfg.do_stable_layer_norm and not cfg.conv_bias
    m = Wav2Vec2Model(cfg).eval()
    sd = dict(m.state_dict())
    assert set(sd) == set(HubertModel(HubertConfig(num_hidden_layers=2)).state_dict())
    x = torch.randn(2, 8000)
    with torch.no_grad():
        ref = m(x, output_hidden_states=True).hidden_states
        got = E.hubert_hidden_states(sd, x, layers=2)
    for a, b in zip(got, ref):
        assert float((a - b).abs().max() / b.abs().max()) < 5e-5


def test_oracle_hubert_large_family_matches_hf():
    """feat_extract_norm="layer", conv_bias, do_stable_layer_norm (hubert-large / chinese-hubert-large,
    extract_audio_huggingface.py:27): the restatement vs HF HubertModel, every hidden state."""
    from transformers import HubertConfig, HubertModel
    cfg = HubertConfig(hidden_size=1024, num_hidden_layers=3, num_attention_heads=16, intermediate_size=4096,
                       feat_extract_norm="layer", do_stable_layer_norm=True, conv_bias=True)
    m = HubertModel(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in S.hubert_state_dict(seed=5, layers=3, large=True).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all("masked_spec_embed" in k or "parametrizations" in k for k in missing), (missing, unexpected)
    x = torch.randn(2, 4000, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        ref = m(x, output_hidden_states=True).hidden_states
        got = E.hubert_hidden_states(sd, x, layers=3, heads=16)
    assert len(got) == len(ref) == 4
    for a, b in zip(got, ref):
        assert float((a - b).abs().max() / b.abs().max()) < 5e-5


def test_oracle_lstm_encoder_matches_torch_lstm():
    """oracle/fusion.py:_lstm_encoder vs the module the reference builds (modules/encoder.py:45-72:
    nn.LSTM(batch_first) -> final hidden -> dropout -> linear_1), eval mode."""
    from oracle import fusion as OF
    torch.manual_seed(0)
    D, H, B, T = 40, 32, 5, 7
    rnn = torch.nn.LSTM(D, H, num_layers=1, batch_first=True)
    lin = torch.nn.Linear(H, H)
    sd = {"e.rnn." + k: v.detach() for k, v in rnn.state_dict().items()}
    sd.update({"e.linear_1.weight": lin.weight.detach(), "e.linear_1.bias": lin.bias.detach()})
    x = torch.randn(B, T, D)
    with torch.no_grad():
        _, (h, _) = rnn(x)
        ref = lin(h.squeeze(0))
        got = OF._lstm_encoder(sd, "e", x, None, 0.0)
    assert float((got - ref).abs().max()) < 1e-6


@pytest.mark.parametrize("variant", ["b32", "l14"])
def test_oracle_clip_image_features_match_hf(variant):
    """oracle/encoders.py:clip_image_features vs HF CLIPModel.get_image_features (the call of the reference's
    CLIP branch, extract_vision_huggingface.py:121), 2 layers of each released shape."""
    from transformers import CLIPConfig, CLIPModel
    c = S.CLIP_CFGS[variant]
    cfg = CLIPConfig(vision_config=dict(hidden_size=c["hidden"], intermediate_size=c["ffn"], num_hidden_layers=2,
                                        num_attention_heads=c["heads"], patch_size=c["patch"], image_size=224,
                                        projection_dim=c["proj"]),
                     text_config=dict(num_hidden_layers=1, hidden_size=64, intermediate_size=64, num_attention_heads=2,
                                      projection_dim=c["proj"]), projection_dim=c["proj"])
    m = CLIPModel(cfg).eval()
    sd = {k: torch.from_numpy(v) for k, v in S.clip_vision_state_dict(seed=4, variant=variant, layers=2).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(not k.startswith(("vision_model", "visual_projection")) for k in missing)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = m.get_image_features(x)
        ref = getattr(ref, "pooler_output", ref)
        got, _ = E.clip_image_features(sd, x, layers=2, heads=c["heads"])
    assert got.shape == ref.shape == (2, c["proj"])
    assert float((got - ref).abs().max() / ref.abs().max()) < 5e-5


def test_clip_preprocess_restatement():
    """Bicubic resize restatement == Pillow bit for bit; the whole CLIP preprocessing vs the container's HF
    processor to within the one-uint8-level resize difference documented in oracle/pipeline.py."""
    from PIL import Image
    from transformers import CLIPImageProcessor
    rng = np.random.default_rng(9)
    for hw in [(112, 112), (300, 260), (57, 91)]:
        img = rng.integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
        ref = np.asarray(Image.fromarray(img).resize((224, 224), resample=Image.BICUBIC))
        assert np.array_equal(P.pil_resize_bilinear_u8(img, 224, 224, filter="bicubic"), ref)
    bgr = rng.integers(0, 256, (2, 112, 112, 3), dtype=np.uint8)
    hf = CLIPImageProcessor()(images=[Image.fromarray(f[..., ::-1].copy()) for f in bgr], return_tensors="pt")["pixel_values"]
    got = P.clip_preprocess(bgr)
    assert got.shape == hf.shape
    assert float((got - hf).abs().max()) <= 1.01 / 255.0 / min(P.CLIP_STD)
    non_square = rng.integers(0, 256, (1, 150, 100, 3), dtype=np.uint8)  # shorter edge -> 224, center crop
    assert tuple(P.clip_preprocess(non_square).shape) == (1, 3, 224, 224)


def test_oracle_resnet18_matches_torchvision():
    """oracle/encoders.py:resnet18_features vs torchvision.models.resnet18 minus fc (the module the reference
    builds at extract_imagenet_embedding.py:47-49); ToTensor/Normalize restatement vs torchvision transforms."""
    torchvision = pytest.importorskip("torchvision")
    from PIL import Image
    m = torchvision.models.resnet18(weights=None).eval()
    sd = {k: torch.from_numpy(v) for k, v in S.resnet18_state_dict(seed=6).items()}
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.startswith("fc.") or k.endswith("num_batches_tracked") for k in missing), missing
    feat = torch.nn.Sequential(*list(m.children())[:-1])
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        ref = feat(x).squeeze()
        got = E.resnet18_features(sd, x)
    assert got.shape == ref.shape == (3, 512)
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5
    tf = torchvision.transforms.Compose([torchvision.transforms.Resize((224, 224)), torchvision.transforms.ToTensor(),
                                         torchvision.transforms.Normalize(mean=[0.485, 0.456, 0.406],
                                                                          std=[0.229, 0.224, 0.225])])
    bgr = np.random.default_rng(3).integers(0, 256, (2, 112, 112, 3), dtype=np.uint8)
    ref_in = torch.stack([tf(Image.fromarray(f[..., ::-1].copy())) for f in bgr])
    assert float((P.imagenet_preprocess(bgr) - ref_in).abs().max()) < 1e-6


def test_oracle_vggish_matches_the_torch_port_graph():
    """The VGGish restatement (TF variable names, HWIO kernels, NHWC flatten) against the graph of the public
    PyTorch port (torchvggish: nn.Sequential features + embeddings, NCHW convs, two transposes before the
    flatten), weights converted by ``encoders.vggish_tf_names``.  TensorFlow is absent, so this pins the layout
    conventions to a second, independently written form of the same network, not to the reference graph."""
    from torch import nn

    from mertools_b200.encoders import vggish_tf_names
    torch.manual_seed(5)
    feats = nn.Sequential(
        nn.Conv2d(1, 64, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2), nn.Conv2d(64, 128, 3, padding=1), nn.ReLU(),
        nn.MaxPool2d(2), nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(), nn.Conv2d(256, 256, 3, padding=1), nn.ReLU(),
        nn.MaxPool2d(2), nn.Conv2d(256, 512, 3, padding=1), nn.ReLU(), nn.Conv2d(512, 512, 3, padding=1), nn.ReLU(),
        nn.MaxPool2d(2))
    emb = nn.Sequential(nn.Linear(512 * 4 * 6, 4096), nn.ReLU(), nn.Linear(4096, 4096), nn.ReLU(),
                        nn.Linear(4096, 128), nn.ReLU())
    sd = {f"features.{k}": v for k, v in feats.state_dict().items()}
    sd.update({f"embeddings.{k}": v for k, v in emb.state_dict().items()})
    x = torch.randn(2, 96, 64) * 2.0 - 1.0
    with torch.no_grad():
        y = feats(x[:, None])
        y = torch.transpose(torch.transpose(y, 1, 3), 1, 2).contiguous().view(2, -1)
        ref = emb(y)
        tf_sd = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in vggish_tf_names(sd).items()}
        got = E.vggish_embeddings(tf_sd, x)
    assert tf_sd["vggish/conv2/weights"].shape == (3, 3, 64, 128) and tf_sd["vggish/fc1/fc1_1/weights"].shape == (12288, 4096)
    assert got.shape == (2, 128) and float((got - ref).abs().max()) < 1e-5 * float(ref.abs().max())


@pytest.mark.parametrize("hw", [(112, 112), (300, 300), (224, 224), (256, 256), (100, 180), (512, 512), (31, 47), (513, 257)])
def test_cv2_resize_restatement_is_bit_exact(hw):
    """oracle.pipeline.cv2_resize_linear_u8 against cv2.resize itself (OpenCV is part of this image)."""
    cv2 = pytest.importorskip("cv2")
    img = np.random.default_rng(hw[0]).integers(0, 256, (hw[0], hw[1], 3), dtype=np.uint8)
    assert np.array_equal(P.cv2_resize_linear_u8(img, 256, 256), cv2.resize(img, (256, 256)))


def test_oracle_audio_restatement_covers_wav2vec2_large_960h_and_data2vec():
    """Two more configurations of the reference's audio model list (extract_audio_huggingface.py:18-29) pinned to HF:
    wav2vec2-large-960h (hidden 1024 on the GroupNorm feature extractor, post-LN) and data2vec-audio-base
    (LayerNorm convs without biases, a chain of five k = 19 positional convs)."""
    from transformers import Data2VecAudioConfig, Data2VecAudioModel, Wav2Vec2Config, Wav2Vec2Model
    x = torch.randn(2, 12000, generator=torch.Generator().manual_seed(3))
    cases = [(S.hubert_state_dict(seed=5, layers=2, large=True, group_norm=True), 16,
              Wav2Vec2Model(Wav2Vec2Config(hidden_size=1024, num_hidden_layers=2, num_attention_heads=16,
                                           intermediate_size=4096, feat_extract_norm="group",
                                           do_stable_layer_norm=False, conv_bias=False))),
             (S.hubert_state_dict(seed=3, layers=2, data2vec=True), 12,
              Data2VecAudioModel(Data2VecAudioConfig(num_hidden_layers=2))),
             # data2vec-audio-large: the same graph at hidden 1024 / 16 heads / 4096
             (S.hubert_state_dict(seed=4, layers=2, data2vec=True, large=True), 16,
              Data2VecAudioModel(Data2VecAudioConfig(num_hidden_layers=2, hidden_size=1024, num_attention_heads=16,
                                                     intermediate_size=4096)))]
    for sd, heads, model in cases:
        model.eval()
        res = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        with torch.no_grad():
            ref = model(x, output_hidden_states=True).hidden_states
        got = E.hubert_hidden_states({k: torch.from_numpy(v) for k, v in sd.items()}, x, layers=2, heads=heads)
        for a, b in zip(got, ref):
            assert float((a - b).abs().max() / b.abs().max()) < 2e-5


def test_bert_oracle_matches_hf_large_configuration():
    """bert-large / roberta-large shape (1024 / 16 heads / 4096): the same restatement at other dims, pinned to HF."""
    from transformers import BertConfig, BertModel
    sd = S.bert_state_dict(300, seed=4, layers=2, large=True)
    m = BertModel(BertConfig(vocab_size=300, hidden_size=1024, num_hidden_layers=2, num_attention_heads=16,
                             intermediate_size=4096), add_pooling_layer=False).eval()
    res = m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not res.missing_keys
    ids = torch.tensor([[2, 17, 250, 99, 42, 7, 3]])
    with torch.no_grad():
        ref = m(ids, output_hidden_states=True).hidden_states
    got = E.bert_hidden_states({k: torch.from_numpy(v) for k, v in sd.items()}, ids, layers=2, heads=16)
    for a, b in zip(got, ref):
        assert float((a - b).abs().max() / b.abs().max()) < 2e-5


def test_whisper_restatements_match_hf():
    """Log-mel front-end and mel filter bank against WhisperFeatureExtractor, encoder + decoder against WhisperModel."""
    from transformers import WhisperConfig, WhisperFeatureExtractor, WhisperModel
    fe = WhisperFeatureExtractor()
    assert np.array_equal(P.whisper_mel_filters(), fe.mel_filters)
    rng = np.random.default_rng(0)
    for n in (30000, 16000 * 7 + 123):
        x = rng.standard_normal(n) * 0.1
        ref = fe(x, sampling_rate=16000, return_tensors="np").input_features[0]
        assert np.abs(P.whisper_log_mel(x) - ref).max() < 1e-5
    sd = S.whisper_state_dict(seed=13, enc_layers=2, dec_layers=2)
    m = WhisperModel(WhisperConfig(vocab_size=64, d_model=512, encoder_layers=2, decoder_layers=2, encoder_attention_heads=8,
                                   decoder_attention_heads=8, encoder_ffn_dim=2048, decoder_ffn_dim=2048,
                                   decoder_start_token_id=5, pad_token_id=0, bos_token_id=1, eos_token_id=2)).eval()
    m.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    f = torch.randn(2, 80, 3000, generator=torch.Generator().manual_seed(1)) * 0.5
    ids = torch.tensor([[5, 5], [5, 5]])
    with torch.no_grad():
        ref = m(f, decoder_input_ids=ids).last_hidden_state
    got = E.whisper_last_hidden_state({k: torch.from_numpy(v) for k, v in sd.items()}, f, ids)
    assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5


def test_electra_base_is_the_bert_graph_under_the_same_names():
    """electra-base / -large discriminators (extract_text_huggingface.py:27-28,44-46): embedding_size == hidden_size, so
    ElectraModel is BertModel's graph with identical parameter names — the BERT restatement applies as is."""
    from transformers import ElectraConfig, ElectraModel
    m = ElectraModel(ElectraConfig(vocab_size=300, num_hidden_layers=2, embedding_size=768, hidden_size=768,
                                   num_attention_heads=12, intermediate_size=3072)).eval()
    sd = m.state_dict()
    assert not any("embeddings_project" in k for k in sd)
    ids = torch.tensor([[2, 17, 250, 99, 42, 7, 3]])
    with torch.no_grad():
        ref = m(ids, output_hidden_states=True).hidden_states
    for a, b in zip(E.bert_hidden_states(sd, ids, layers=2), ref):
        assert float((a - b).abs().max() / b.abs().max()) < 2e-5


@pytest.mark.parametrize("large", [False, True])
def test_oracle_audio_restatement_covers_wavlm(large):
    """wavlm-base / wavlm-large (extract_audio_huggingface.py:36-37) pinned to HF WavLMModel: bucketed relative position
    bias from layer 0's embedding, gated per layer / head / query; post-LN base, stable-layer-norm large."""
    from transformers import WavLMConfig, WavLMModel
    x = torch.randn(2, 12000, generator=torch.Generator().manual_seed(3))
    sd = S.hubert_state_dict(seed=6, layers=3, wavlm=True, large=large)
    kw = dict(hidden_size=1024, num_attention_heads=16, intermediate_size=4096, feat_extract_norm="layer",
              do_stable_layer_norm=True, conv_bias=False) if large else {}
    model = WavLMModel(WavLMConfig(num_hidden_layers=3, **kw)).eval()
    res = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with torch.no_grad():
        ref = model(x, output_hidden_states=True).hidden_states
    got = E.hubert_hidden_states({k: torch.from_numpy(v) for k, v in sd.items()}, x, layers=3, heads=16 if large else 12)
    assert len(got) == len(ref) == 4
    for a, b in zip(got, ref):
        assert float((a - b).abs().max() / b.abs().max()) < 2e-5
