"""MER2023 English word-aligned text extraction on the CUDA BERT encoder vs outputs of the unmodified
reference function (tests/golden/text_words_golden.npz, made by make_golden_words.py).  Tolerance 1e-3
relative on the utterance mean (north_star), 2e-3 on single word rows."""
import os

import numpy as np
import pytest

from mertools_b200 import synthetic as S

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-3


def _rel(a, b):
    return float(np.abs(np.asarray(a, dtype=np.float64) - b).max() / np.abs(b).max())


def test_english_word_aligned_text_features_vs_reference_golden(cuda):
    """extract_bert_embedding_english (MER2023 extract_text_embedding_LZ.py:168-311): host word / sentence logic
    around the CUDA BERT encoder, against outputs of the unmodified reference function."""
    transformers = pytest.importorskip("transformers")
    from mertools_b200.encoders import BertEncoder
    from mertools_b200.extract import text_english as TE
    g = np.load(os.path.join(G, "text_words_golden.npz"))
    tok = transformers.BertTokenizer(os.path.join(G, "text_words_vocab.txt"), do_lower_case=True)
    enc = BertEncoder(S.bert_state_dict(len(tok), seed=int(g["seed"]), layers=int(g["layers"])), device=cuda)
    for name, sent in zip(g["names"], g["sentences"]):
        emb = TE.transcript_word_features(enc, tok, str(sent), lower=True)
        fra = TE.save_word_features(None, emb, "FRAME", 768)
        utt = TE.save_word_features(None, emb, "UTTERANCE", 768)
        assert fra.shape == g[f"fra_{name}"].shape and _rel(fra, g[f"fra_{name}"]) < 2 * TOL, name
        assert _rel(utt, g[f"utt_{name}"]) < TOL, name
