"""Golden fixture for the English word-aligned text path: the UNMODIFIED reference function
MER2023/feature_extraction/text/extract_text_embedding_LZ.py:extract_bert_embedding_english (sentence
splitting, sub-word -> word alignment, FRAME / UTTERANCE save rules) on a seeded BERT-base-style checkpoint
(4 layers) and a synthetic WordPiece vocabulary whose words split into several pieces.

Run once in the build container (needs /root/reference + transformers; NOT on the GPU box):
    python tests/golden/make_golden_words.py
Writes tests/golden/text_words_golden.npz and tests/golden/text_words_vocab.txt.  Stubs: patched `config`,
`Module.to` / `BatchEncoding.to` as identities (the function hard-codes cuda:N; CPU-only container).
"""
import importlib.util
import os
import shutil
import sys
import tempfile
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
REF = "/root/reference/MER2023/feature_extraction/text"
OUT = os.path.dirname(os.path.abspath(__file__))

from mertools_b200 import synthetic as S  # noqa: E402

LAYERS, SEED = 4, 12
SENTENCES = {
    "clip0": "I'm really happy today, the movie was unbelievable! Did you like it?",
    "clip1": "No.",
    "clip2": "well it was okay I guess but the ending felt rushed and nobody laughed",
    "clip3": "Wow!!! Absolutely   fantastic, 10 out of 10.",
}
WORDS = ["the", "movie", "was", "really", "happy", "today", "did", "you", "like", "it", "no", "well", "okay", "but",
         "and", "out", "of", "un", "##believ", "##able", "fant", "##astic", "end", "##ing", "laugh", "##ed", "rush",
         "nobody", "felt", "guess", "wow", "absolute", "##ly", "10"]


def vocab():
    chars = list("abcdefghijklmnopqrstuvwxyz0123456789") + ["'", ",", ".", "!", "?"]
    v = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"] + chars + ["##" + c for c in chars if c.isalnum()]
    for w in WORDS:
        if w not in v:
            v.append(w)
    return v


def main():
    import pandas as pd
    from transformers import BertConfig, BertModel, BertTokenizer
    work = tempfile.mkdtemp(prefix="mer_golden_words_")
    name = "bert-base-uncased"
    mdir = os.path.join(work, "tools", "transformers", name)
    os.makedirs(mdir)
    v = vocab()
    with open(os.path.join(OUT, "text_words_vocab.txt"), "w") as f:
        f.write("\n".join(v) + "\n")
    shutil.copy(os.path.join(OUT, "text_words_vocab.txt"), os.path.join(mdir, "vocab.txt"))
    BertTokenizer(os.path.join(mdir, "vocab.txt"), do_lower_case=True).save_pretrained(mdir)
    m = BertModel(BertConfig(vocab_size=len(v), num_hidden_layers=LAYERS))
    sd = {k: torch.from_numpy(x) for k, x in S.bert_state_dict(len(v), seed=SEED, layers=LAYERS).items()}
    m.load_state_dict(sd, strict=False)
    m.save_pretrained(mdir)
    csv = os.path.join(work, "trans.csv")
    pd.DataFrame({"name": list(SENTENCES), "sentence": list(SENTENCES.values())}).to_csv(csv, index=False)
    cfg = types.ModuleType("config")
    cfg.PATH_TO_PRETRAINED_MODELS = os.path.join(work, "tools")
    sys.modules["config"] = cfg
    sys.path.insert(0, REF)
    spec = importlib.util.spec_from_file_location("ref_text_lz", os.path.join(REF, "extract_text_embedding_LZ.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    # the function hard-codes cuda:{gpu}: moving the model / the inputs there becomes the identity
    from transformers import BatchEncoding
    orig_mto, orig_bto = torch.nn.Module.to, BatchEncoding.to
    torch.nn.Module.to = lambda self, *a, **k: self
    BatchEncoding.to = lambda self, *a, **k: self
    out = {}
    try:
        for level in ("FRAME", "UTTERANCE"):
            sdir = os.path.join(work, "feat")
            ref.extract_bert_embedding_english(name, csv, sdir, level, gpu=0)
            d = os.path.join(sdir, f"{name}-4-{level[:3]}")
            for clip in SENTENCES:
                out[f"{level[:3].lower()}_{clip}"] = np.load(os.path.join(d, f"{clip}.npy"))
    finally:
        torch.nn.Module.to, BatchEncoding.to = orig_mto, orig_bto
    np.savez(os.path.join(OUT, "text_words_golden.npz"), layers=LAYERS, seed=SEED, names=np.array(list(SENTENCES)),
             sentences=np.array(list(SENTENCES.values())), **out)
    print({k: v.shape for k, v in out.items()})
    shutil.rmtree(work, ignore_errors=True)


if __name__ == "__main__":
    main()
