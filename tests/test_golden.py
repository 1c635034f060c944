"""Pin the oracle to the outputs of the UNMODIFIED reference scripts (tests/golden/*.npz, produced
by tests/golden/make_golden.py from /root/reference on seeded synthetic checkpoints and inputs).

CPU only.  Integer work (frame indices, token ids, special-token offsets) is bit-exact; floating
point within 2e-5 relative (fp32 torch on both sides; HF model code vs the oracle restatement)."""
import os

import numpy as np
import pytest
import torch

from mertools_b200 import synthetic as S
from oracle import fusion as OF
from oracle import encoders as E
from oracle import pipeline as P

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / np.abs(b).max())


def _t(sd):
    return {k: torch.from_numpy(v) for k, v in sd.items()}


@pytest.mark.slow
def test_visual_oracle_matches_reference_script():
    g = np.load(os.path.join(G, "visual_golden.npz"))
    sd = _t(S.vit_state_dict(seed=0))
    clips = S.synth_frames(int(g["n_clips"]), 8, seed=int(g["seed"]))
    with torch.no_grad():
        utt = P.visual_clip_features(sd, clips[0], nframe=int(g["nframe"]))
        fra = P.visual_clip_features(sd, clips[1], nframe=int(g["nframe"]), feature_level="FRAME")
    assert utt.shape == g["utt0"].shape == (768,) and utt.dtype == g["utt0"].dtype
    assert _rel(utt, g["utt0"]) < 2e-5
    assert fra.shape == g["fra1"].shape == (64, 768)
    assert _rel(fra, g["fra1"]) < 2e-5


@pytest.mark.slow
def test_visual_oracle_matches_reference_script_on_112px_faces():
    """Resize step (a2): 112x112 crops through the unmodified script vs the oracle's PIL-resize restatement."""
    g = np.load(os.path.join(G, "visual112_golden.npz"))
    sd = _t(S.vit_state_dict(seed=0))
    clips = S.synth_frames(int(g["n_clips"]), 8, size=int(g["size"]), seed=int(g["seed"]))
    with torch.no_grad():
        utt = P.visual_clip_features(sd, clips[0], nframe=int(g["nframe"]))
        fra = P.visual_clip_features(sd, clips[1], nframe=None, feature_level="FRAME")
    assert utt.shape == g["utt0"].shape == (768,)
    assert _rel(utt, g["utt0"]) < 2e-5
    assert _rel(fra, g["fra1"]) < 2e-5


def test_audio_oracle_matches_reference_script():
    g = np.load(os.path.join(G, "audio_golden.npz"))
    sd = _t(S.hubert_state_dict(seed=1))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        with torch.no_grad():
            utt = P.audio_clip_features(sd, w)
            fra = P.audio_clip_features(sd, w, feature_level="FRAME")
        assert utt.shape == (768,) and utt.dtype == g[f"utt{i}"].dtype
        assert _rel(utt, g[f"utt{i}"]) < 5e-5, f"clip {i} ({n} samples)"
        assert _rel(fra[::8], g[f"fra{i}"]) < 5e-5


def test_audio_oracle_matches_reference_script_hubert_large_family():
    g = np.load(os.path.join(G, "audio_large_golden.npz"))
    layers = int(g["layers"])
    sd = _t(S.hubert_state_dict(seed=int(g["seed"]), layers=layers, large=True))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        with torch.no_grad():
            utt = P.audio_clip_features(sd, w, layers=layers, heads=16)
            fra = P.audio_clip_features(sd, w, layers=layers, heads=16, feature_level="FRAME")
        assert utt.shape == (1024,) and utt.dtype == g[f"utt{i}"].dtype
        assert _rel(utt, g[f"utt{i}"]) < 5e-5, f"clip {i} ({n} samples)"
        assert _rel(fra[::16], g[f"fra{i}"]) < 5e-5


def test_text_oracle_and_token_ids_match_reference_script():
    transformers = pytest.importorskip("transformers")
    g = np.load(os.path.join(G, "text_golden.npz"))
    tok = transformers.BertTokenizer(os.path.join(G, "text_vocab.txt"))
    assert P.find_start_end_pos(tok) == (int(g["start"]), int(g["end"])) == (1, -1)
    sd = _t(S.bert_state_dict(int(g["vocab_size"]), seed=2))
    for i, s in enumerate(g["sentences"]):
        s = str(s)
        if s:
            ids = tok(s)["input_ids"]
            np.testing.assert_array_equal(np.array(ids), g[f"ids{i}"])  # bit-exact token ids
        else:
            ids = None
        with torch.no_grad():
            utt = P.text_clip_features(sd, ids, 1, -1)
            fra = P.text_clip_features(sd, ids, 1, -1, feature_level="FRAME")
        assert utt.shape == g[f"utt{i}"].shape and utt.dtype == g[f"utt{i}"].dtype
        assert fra.shape == g[f"fra{i}"].shape
        if s:
            assert _rel(utt, g[f"utt{i}"]) < 2e-5 and _rel(fra, g[f"fra{i}"]) < 2e-5
        else:
            assert not utt.any() and not fra.any()


def test_fusion_oracle_matches_reference_trainer():
    g = np.load(os.path.join(G, "fusion_golden.npz"))
    tr = OF.Trainer(S.fusion_state_dict(seed=3), lr=1e-3, l2=1e-5)
    a, t, v, emo, val = S.synth_fusion_features(32, seed=7)
    T = torch.from_numpy
    for step, ref in enumerate(g["losses"]):
        ce, mse, tot, eo, vo, grads = tr.step(T(a), T(t), T(v), T(emo), T(val).view(-1, 1))
        assert abs(tot - ref) <= 1e-5 * max(1.0, abs(ref)), f"step {step}: {tot} vs {ref}"
        if step == 0:
            assert _rel(eo.numpy(), g["emos0"]) < 1e-5 and _rel(vo.numpy(), g["vals0"]) < 1e-5
            assert _rel(grads["fc_att.weight"].numpy(), g["grad_fc_att_w"]) < 1e-4
            assert _rel(grads["audio_encoder.linear_1.bias"].numpy(), g["grad_audio_l1_b"]) < 1e-4
    assert _rel(tr.sd["fc_out_1.weight"].detach().numpy(), g["final_fc_out_1_w"]) < 1e-4
    assert _rel(tr.sd["audio_encoder.linear_1.weight"].detach().numpy()[0], g["final_audio_l1_w_row0"]) < 1e-4


def test_frame_level_fusion_oracle_matches_reference_classes():
    """feat_type = frm_align: oracle Trainer vs the reference's Attention(LSTMEncoder) + losses + Adam."""
    from oracle import fusion as OF
    g = np.load(os.path.join(G, "fusion_frm_golden.npz"))
    sd = S.fusion_state_dict(seed=int(g["seed"]), feat_type="frm_align")
    a, t, v, emo, val = S.synth_fusion_sequences(int(g["batch"]), lens=tuple(int(x) for x in g["lens"]),
                                                 seed=int(g["data_seed"]))
    tt = torch.from_numpy
    tr = OF.Trainer(sd, lr=1e-3, l2=1e-5)
    for step in range(len(g["losses"])):
        ce, mse, tot, eo, vo, grads = tr.step(tt(a), tt(t), tt(v), tt(emo), tt(val).view(-1, 1))
        assert abs(tot - g["losses"][step]) < 1e-5 * max(1.0, abs(g["losses"][step])), step
        if step == 0:
            assert _rel(eo.numpy(), g["emos0"]) < 1e-5 and _rel(vo.numpy(), g["vals0"]) < 1e-5
            assert _rel(grads["audio_encoder.rnn.weight_hh_l0"][0].numpy(), g["grad_audio_whh_row0"]) < 1e-4
            assert _rel(grads["text_encoder.rnn.bias_ih_l0"].numpy(), g["grad_text_bih"]) < 1e-4
            assert _rel(grads["video_encoder.rnn.weight_ih_l0"][5].numpy(), g["grad_video_wih_row5"]) < 1e-4


def test_logmel_oracle_matches_reference_vggish_input():
    """oracle/pipeline.py:waveform_to_examples vs the reference's vggish_input.waveform_to_examples."""
    g = np.load(os.path.join(G, "logmel_golden.npz"))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        ex = P.waveform_to_examples(w, float(g["hop_sec"]))
        assert tuple(ex.shape) == tuple(g[f"n{i}"])
        sub = ex[:: max(1, len(ex) // 4)]
        assert np.abs(sub - g[f"ex{i}"]).max() < 1e-5


def _topn_data(g):
    rng = np.random.default_rng(5000 + int(g["data_seed"]))
    dims, B = [int(d) for d in g["dims"]], int(g["batch"])
    feats = [rng.standard_normal((B, d), dtype=np.float32) for d in dims]
    return dims, feats, rng.integers(0, 6, B).astype(np.int64), rng.uniform(-3, 3, B).astype(np.float32)


def test_attention_topn_oracle_matches_reference_class():
    """oracle Trainer vs the reference's Attention_TOPN + losses + Adam (MER2026 toolkit), 15 steps."""
    from oracle import fusion as OF
    g = np.load(os.path.join(G, "fusion_topn_golden.npz"))
    dims, feats, emo, val = _topn_data(g)
    tr = OF.Trainer(S.fusion_topn_state_dict(dims, seed=int(g["seed"])), lr=1e-3, l2=1e-5)
    tt = torch.from_numpy
    for step in range(len(g["losses"])):
        ce, mse, tot, eo, vo, grads = tr.step([tt(f) for f in feats], None, None, tt(emo), tt(val).view(-1, 1))
        assert abs(tot - g["losses"][step]) < 1e-5 * max(1.0, abs(g["losses"][step])), step
        if step == 0:
            assert _rel(eo.numpy(), g["emos0"]) < 1e-5
            assert _rel(grads["fc_att.weight"].numpy(), g["grad_fc_att_w"]) < 1e-4
            assert _rel(grads["encoder3.linear_1.bias"].numpy(), g["grad_enc3_l1_b"]) < 1e-4


class _OracleBert:
    """Checker-side stand-in for BertEncoder.forward(want_tokens=True): last-four sums from the oracle."""

    def __init__(self, sd, layers):
        self.sd, self.layers = sd, layers

    def forward(self, id_lists, start=0, end=None, want_tokens=True):
        toks = []
        for ids in id_lists:
            with torch.no_grad():
                hs = E.bert_hidden_states(self.sd, torch.tensor([ids]), layers=self.layers)
            toks.append(torch.stack(hs)[[-4, -3, -2, -1]].sum(dim=0)[0])
        return None, torch.cat(toks)


def test_english_word_alignment_host_logic_matches_reference_function():
    """Sentence splitting + sub-word -> word merging + save rules of extract_bert_embedding_english
    (MER2023 extract_text_embedding_LZ.py:168-311) against outputs of the unmodified reference function; the
    encoder behind the host logic is the oracle here (the CUDA encoder has its own parity tests)."""
    transformers = pytest.importorskip("transformers")
    from mertools_b200.extract import text_english as TE
    g = np.load(os.path.join(G, "text_words_golden.npz"))
    tok = transformers.BertTokenizer(os.path.join(G, "text_words_vocab.txt"), do_lower_case=True)
    layers = int(g["layers"])
    sd = _t(S.bert_state_dict(len(tok), seed=int(g["seed"]), layers=layers))
    enc = _OracleBert(sd, layers)
    for name, sent in zip(g["names"], g["sentences"]):
        emb = TE.transcript_word_features(enc, tok, str(sent), lower=True)
        fra = TE.save_word_features(None, emb, "FRAME", 768)
        utt = TE.save_word_features(None, emb, "UTTERANCE", 768)
        assert fra.shape == g[f"fra_{name}"].shape and _rel(fra, g[f"fra_{name}"]) < 5e-5, name
        assert utt.shape == (768,) and _rel(utt, g[f"utt_{name}"]) < 5e-5, name
    assert TE.split_words_and_sentences("Wow!!! ok.", True) == [["wow"], ["ok"]]


def _ferplus_clips():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_ferplus", os.path.join(G, "make_golden_ferplus.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.golden_clips()


def test_ferplus_oracle_matches_reference_extractor_golden():
    """resnet50_ferplus_dag restatement + compose_transforms restatement against outputs of the unmodified
    reference functions (extract_ferplus_embedding.py: load_model, compose_transforms, FaceDataset, get_feature)."""
    from oracle import pipeline as P
    g = np.load(os.path.join(G, "ferplus_golden.npz"))
    for se, prefix in ((False, ""), (True, "se_")):   # resnet50_ferplus_dag, senet50_ferplus_dag
        sd = _t(S.ferplus_resnet50_state_dict(int(g["seed"]), se=se))
        for vid, frames in _ferplus_clips().items():
            x = P.ferplus_preprocess(frames)
            assert np.array_equal(x.numpy()[:, :, ::16, ::16], g[f"x_{vid}"]), vid   # Resize / crop / scale: bit-exact
            for level, key in (("FRAME", "fra"), ("UTTERANCE", "utt")):
                got, ref = P.ferplus_clip_features(sd, frames, level), g[f"{prefix}{key}_{vid}"]
                assert got.shape == ref.shape and _rel(got, ref) < 1e-5, (se, vid, level)


def test_manet_oracle_matches_reference_extractor_golden():
    """MA-Net restatement (ResNet trunk, four CBAM patch branches, multi-scale branch) + the script's transform
    against outputs of the unmodified reference model definition and dataset / save rules."""
    import importlib.util
    from oracle import pipeline as P
    spec = importlib.util.spec_from_file_location("make_golden_manet", os.path.join(G, "make_golden_manet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(G, "manet_golden.npz"))
    sd = _t(S.manet_state_dict(int(g["seed"])))
    for vid, frames in mod.golden_clips().items():
        for level, key in (("FRAME", "fra"), ("UTTERANCE", "utt")):
            got, ref = P.manet_clip_features(sd, frames, level), g[f"{key}_{vid}"]
            assert got.shape == ref.shape and _rel(got, ref) < 1e-5, (vid, level)


def test_resnet18_oracle_matches_the_msceleb_reference_classes_golden():
    """extract_msceleb_embedding.py defines its own ResNet-18 (same parameter names as torchvision's): outputs of
    those reference classes + transform + save rules (make_golden_msceleb.py) against the oracle pipeline that also
    serves the ImageNet extractor — bit for bit."""
    import importlib.util
    from oracle import pipeline as P
    spec = importlib.util.spec_from_file_location("make_golden_msceleb", os.path.join(G, "make_golden_msceleb.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(G, "msceleb_golden.npz"))
    sd = _t(S.resnet18_state_dict(int(g["seed"])))
    for vid, frames in mod.golden_clips().items():
        for level, key in (("FRAME", "fra"), ("UTTERANCE", "utt")):
            got, ref = P.imagenet_clip_features(sd, frames, level), g[f"{key}_{vid}"]
            assert got.shape == ref.shape and _rel(got, ref) < 1e-6, (vid, level)


def test_emonet_oracle_matches_reference_extractor_golden():
    """EmoNet restatement (pre-activation ConvBlocks, two hourglasses, heat-map mask, emotion tower) + the
    DataAugmentor / ToTensor restatement against outputs of the unmodified reference model, dataset and augmentor."""
    import importlib.util
    from oracle import pipeline as P
    spec = importlib.util.spec_from_file_location("make_golden_emonet", os.path.join(G, "make_golden_emonet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(G, "emonet_golden.npz"))
    sd = _t(S.emonet_state_dict(int(g["seed"])))
    for vid, frames in mod.golden_clips().items():
        x = (P.emonet_preprocess(frames[:1]).numpy()[0] * 255.0).round().astype(np.uint8)
        assert np.array_equal(x[:, ::8, ::8], g[f"x_{vid}"]), vid          # cv2.resize restatement: bit-exact
        for level, key in (("FRAME", "fra"), ("UTTERANCE", "utt")):
            got, ref = P.emonet_clip_features(sd, frames, level), g[f"{key}_{vid}"]
            assert got.shape == ref.shape and _rel(got, ref) < 2e-5, (vid, level)


def test_audio_oracle_matches_reference_script_data2vec_audio():
    """Data2VecAudioModel branch of the oracle (LayerNorm convs, chain of five k = 19 positional convs, post-LN)
    against the unmodified reference extract() on a data2vec-audio-base-960h-style checkpoint."""
    g = np.load(os.path.join(G, "audio_data2vec_golden.npz"))
    layers = int(g["layers"])
    sd = _t(S.hubert_state_dict(seed=int(g["seed"]), layers=layers, data2vec=True))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        with torch.no_grad():
            utt = P.audio_clip_features(sd, w, layers=layers)
            fra = P.audio_clip_features(sd, w, layers=layers, feature_level="FRAME")
        assert utt.shape == (768,) and _rel(utt, g[f"utt{i}"]) < 5e-5, f"clip {i} ({n} samples)"
        assert _rel(fra[::16], g[f"fra{i}"]) < 5e-5


def test_audio_oracle_matches_reference_script_whisper_branch():
    """Whisper branch (extract_audio_huggingface.py:83-110): log-mel restatement + encoder / decoder restatement against
    the unmodified reference extract() on a whisper-base-shaped checkpoint."""
    g = np.load(os.path.join(G, "audio_whisper_golden.npz"))
    layers = int(g["layers"])
    sd = _t(S.whisper_state_dict(seed=int(g["seed"]), enc_layers=layers, dec_layers=layers))
    for i, n in enumerate(g["lens"]):
        w = S.synth_waves(1, int(n), seed=int(g["seed0"]) + i)[0].astype(np.float64) / 32768.0
        for level, key in (("UTTERANCE", "utt"), ("FRAME", "fra")):
            got, ref = P.whisper_clip_features(sd, w, int(g["start"]), level), g[f"{key}{i}"]
            assert got.shape == ref.shape and _rel(got, ref) < 5e-5, (i, level)


def test_frame_index_selection_and_cv2_path_match_load_video_from_npy_golden():
    """``select_frame_indices`` (all four readtypes, seeded draws) + the cv2.resize restatement + BGR->RGB against outputs
    of the reference's own ``load_video_from_npy`` source (make_golden_video_npy.py)."""
    import importlib.util
    from mertools_b200.extract.visual import select_frame_indices
    from oracle import pipeline as P
    spec = importlib.util.spec_from_file_location("make_golden_video_npy", os.path.join(G, "make_golden_video_npy.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    g = np.load(os.path.join(G, "video_npy_golden.npz"))
    for ci, (readtype, n_frms, vlen, size) in enumerate(mod.CASES):
        frames = mod.golden_clip(vlen, size, 200 + ci)
        np.random.seed(1000 + ci)
        idx = select_frame_indices(vlen, n_frms, readtype)
        rgb = np.stack([P.cv2_resize_linear_u8(frames[i], 224, 224)[..., ::-1] for i in idx])      # [T, 224, 224, 3]
        x = rgb.transpose(3, 0, 1, 2).astype(np.float32)
        assert list(x.shape) == list(g[f"shape{ci}"]), (readtype, x.shape)
        assert np.array_equal(x[:, :, ::16, ::16].astype(np.uint8), g[f"probe{ci}"]), readtype
        assert float(x.sum(dtype=np.float64)) == float(g[f"sum{ci}"][0]), readtype


def test_vggish_oracle_matches_the_reference_graph_and_extractor_golden():
    """tests/golden/make_golden_vggish.py ran the UNMODIFIED vggish_slim.define_vggish_slim / load_vggish_slim_checkpoint
    and extract_vggish_embedding.extract (UTTERANCE: 0.5 s hop + mean, FRAME: 0.05 s hop; batches of 3 examples) over a
    torch-backed stand-in for the TensorFlow / tf_slim calls they make (tests/golden/tf_slim_shim.py): the graph
    structure, variable names and extractor logic are the reference's; the float arithmetic is torch's."""
    g = np.load(os.path.join(G, "vggish_golden.npz"))
    raw = S.vggish_state_dict(seed=8)
    assert sorted(g["variable_names"].tolist()) == sorted(k + ":0" for k in raw)  # the graph's variables = the checkpoint names
    sd = {k: torch.from_numpy(v) for k, v in raw.items()}
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())  # noqa: E731
    with torch.no_grad():
        assert rel(E.vggish_embeddings(sd, torch.from_numpy(g["patches"])).numpy(), g["patch_embeddings"]) < 1e-5
        for name, n, seed in zip(g["clip_names"], g["clip_samples"], g["clip_seeds"]):
            w = S.synth_waves(1, int(n), seed=int(seed))[0].astype(np.int16)
            for level, hop in (("UTTERANCE", 0.5), ("FRAME", 0.05)):
                ex = P.waveform_to_examples(w / 32768.0, hop)
                emb = E.vggish_embeddings(sd, torch.from_numpy(ex.astype(np.float32))).numpy()
                if level == "UTTERANCE":      # extract_vggish_embedding.py:54-58
                    emb = emb.squeeze()
                    emb = emb.mean(axis=0) if emb.ndim != 1 else emb
                ref = g[f"{name}_{level}"]
                assert emb.shape == ref.shape and rel(emb, ref) < 1e-5, (name, level)


def test_tf_slim_stand_in_keeps_the_conventions_the_vggish_pin_relies_on():
    """tests/golden/tf_slim_shim.py (the torch-backed stand-in that let the unmodified VGGish graph definition run):
    variable naming under variable_scope / slim.repeat, arg_scope precedence (inner scope over outer, explicit argument
    over both), SAME padding of stride-1 convolutions and of 2x2 / stride-2 pools on odd sizes, NHWC flatten order, and
    feed-dict evaluation of a named tensor."""
    import sys
    import torch.nn.functional as F
    sys.path.insert(0, G)
    import tf_slim_shim as shim
    saved = {k: sys.modules.get(k) for k in ("tensorflow", "tensorflow.compat", "tensorflow.compat.v1", "tf_slim")}
    try:
        tf, slim = shim.install(5, 7)
        with tf.Graph().as_default(), tf.Session() as sess:
            with slim.arg_scope([slim.conv2d, slim.fully_connected], activation_fn=tf.nn.relu), \
                 slim.arg_scope([slim.conv2d], kernel_size=[3, 3], stride=1, padding="SAME"), \
                 slim.arg_scope([slim.max_pool2d], kernel_size=[2, 2], stride=2, padding="SAME"), \
                 tf.variable_scope("net"):
                x = tf.placeholder(tf.float32, shape=(None, 5, 7), name="in")
                y = tf.reshape(x, [-1, 5, 7, 1])
                y = slim.repeat(y, 2, slim.conv2d, 4, scope="c")
                y = slim.conv2d(y, 3, activation_fn=None, scope="lin")       # explicit argument beats the scope default
                y = slim.max_pool2d(y, scope="p")                            # 5 x 7 -> 3 x 4 (SAME, odd sizes)
                y = slim.fully_connected(slim.flatten(y), 2, scope="fc")
                out = tf.identity(y, name="out")
            names = sorted(v.name for v in tf.global_variables())
            assert names == sorted(["net/c/c_1/weights:0", "net/c/c_1/biases:0", "net/c/c_2/weights:0", "net/c/c_2/biases:0",
                                    "net/lin/weights:0", "net/lin/biases:0", "net/fc/weights:0", "net/fc/biases:0"])
            rng = np.random.default_rng(0)
            vals = {}
            for v in tf.global_variables():
                vals[v.name] = torch.from_numpy(rng.standard_normal(v.shape).astype(np.float32))
                v.value = vals[v.name]
            assert tuple(vals["net/fc/weights:0"].shape) == (3 * 4 * 3, 2)       # SAME pool of 5 x 7, 3 channels, NHWC
            xin = rng.standard_normal((2, 5, 7)).astype(np.float32)
            [got] = sess.run([sess.graph.get_tensor_by_name("net/out:0")],
                             feed_dict={sess.graph.get_tensor_by_name("net/in:0"): xin})
        # the same network written directly in torch (NCHW), TF conventions applied by hand
        conv = lambda t, n: F.conv2d(t, vals[f"net/{n}/weights:0"].permute(3, 2, 0, 1), vals[f"net/{n}/biases:0"], padding=1)  # noqa: E731
        t = torch.from_numpy(xin)[:, None]
        t = torch.relu(conv(torch.relu(conv(t, "c/c_1")), "c/c_2"))
        t = conv(t, "lin")                                                        # no activation
        t = F.max_pool2d(F.pad(t, (0, 1, 0, 1), value=float("-inf")), 2, 2)       # SAME: the odd row / column pads at the end
        t = t.permute(0, 2, 3, 1).reshape(2, -1)                                  # NHWC flatten
        ref = torch.relu(t @ vals["net/fc/weights:0"] + vals["net/fc/biases:0"]).numpy()
        assert got.shape == ref.shape == (2, 2) and np.abs(got - ref).max() < 1e-6 * np.abs(ref).max()
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
