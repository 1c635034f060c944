"""Lexical feature extraction — B200 mirror of
MERBench/feature_extraction/text/extract_text_huggingface.py (BERT / RoBERTa-base branch).

Keeps ``extract_embedding(model_name, trans_dir, save_dir, feature_level, gpu, punc_case, language,
model_dir)`` (:139), ``find_start_end_pos`` (:90-114) and the save-dir naming (:148-157).  Token ids
come from the HF tokenizer on the host exactly as in the reference (:222, bit-exact by
construction); the batch-1 model loop (:208-231) becomes one packed variable-length device pass over
many sentences (embedding gather + LayerNorm, 12 post-LN layers, last-four sum, strip specials,
mean).
"""
from __future__ import annotations

import os
import time

import numpy as np

from ..encoders import BertEncoder
from . import common


def find_start_end_pos(tokenizer):
    """How many special tokens wrap a sentence: probe with a 6-character Chinese sentence and look
    for the [start:end] slice that decodes back to it (reference :90-114).  BERT/RoBERTa -> (1,-1)."""
    probe = "今天天气真好"
    ids = tokenizer(probe, return_tensors="pt")["input_ids"][0]
    strip = lambda s: s.replace(" ", "")  # noqa: E731
    start = None
    for start in (0, 1, 2):
        dec = strip(tokenizer.decode(ids[start:]))
        if dec == probe:
            print(f"start: {start};  end: {None}")
            return start, None
        if dec.startswith(probe):
            break
    end = None
    for end in (-1, -2):
        if strip(tokenizer.decode(ids[start:end])) == probe:
            break
    assert strip(tokenizer.decode(ids[start:end])) == probe
    print(f"start: {start};  end: {end}")
    return start, end


class TextExtractor:
    def __init__(self, state_dict, tokenizer, device="cuda", ln_eps=1e-12, position_offset=0,
                 max_tokens_per_launch=16384):
        self.enc = BertEncoder(state_dict, device=device, ln_eps=ln_eps, position_offset=position_offset)
        self.tokenizer = tokenizer
        self.start, self.end = find_start_end_pos(tokenizer)
        self.max_tokens = max_tokens_per_launch

    def tokenize(self, sentence):
        return self.tokenizer(sentence, return_tensors="pt")["input_ids"][0].tolist()

    def extract_sentences(self, sentences, feature_level="UTTERANCE", save_files=None):
        """sentences: list of str (None / NaN / '' give the reference's zero vector, :236-249)."""
        import pandas as pd
        ids, where = [], []
        for i, s in enumerate(sentences):
            if s is not None and not pd.isna(s) and len(s) > 0:
                ids.append(self.tokenize(s))
                where.append(i)
        res = [[] for _ in sentences]  # [] -> the reference's zero vector (nothing to embed)
        b = 0
        while b < len(ids):
            e, tok = b, 0
            while e < len(ids) and (e == b or tok + len(ids[e]) <= self.max_tokens):
                tok += len(ids[e])
                e += 1
            utt, toks = self.enc.forward(ids[b:e], start=self.start, end=self.end,
                                         want_tokens=(feature_level == "FRAME"))
            utt = utt.cpu().numpy()
            toks = toks.cpu().numpy() if toks is not None else None
            o = 0
            for j in range(b, e):
                n = len(ids[j])
                lo, hi = (self.start or 0), n + (self.end or 0)
                if hi > lo:  # something is left after stripping the special tokens (:228-231)
                    res[where[j]] = toks[o + lo:o + hi] if feature_level == "FRAME" else utt[j - b]
                o += n
            b = e
        return [common.save_feature(save_files[i] if save_files is not None else None, r,
                                    feature_level, self.enc.hidden) for i, r in enumerate(res)]


def extract_embedding(model_name, trans_dir, save_dir, feature_level, gpu=-1, punc_case=None,
                      language="chinese", model_dir=None, config=None, sentences_per_launch=256):
    """Same signature, naming and outputs as the reference (:139-252)."""
    import pandas as pd
    from transformers import AutoConfig, AutoTokenizer
    if config is None:
        from .. import config as config  # noqa: PLW0127
    print("=" * 30 + f' Extracting "{model_name}" ' + "=" * 30)
    start_time = time.time()
    if punc_case is None and language == "chinese" and model_dir is None:
        save_dir = os.path.join(save_dir, f"{model_name}-{feature_level[:3]}")
    elif punc_case is not None:
        save_dir = os.path.join(save_dir, f"{model_name}-punc{punc_case}-{feature_level[:3]}")
    elif language == "english":
        save_dir = os.path.join(save_dir, f"{model_name}-langeng-{feature_level[:3]}")
    elif model_dir is not None:
        prefix_name = "-".join(model_dir.split("/")[-2:])
        save_dir = os.path.join(save_dir, f"{prefix_name}-{model_name}-{feature_level[:3]}")
    os.makedirs(save_dir, exist_ok=True)
    if model_dir is None:
        model_dir = os.path.join(config.PATH_TO_PRETRAINED_MODELS, f"transformers/{model_name}")
    assert gpu != -1, "mertools_b200 has no CPU path (reference: gpu=-1 means CPU)"
    from .. import shard
    gpu = shard.device_index(gpu)
    cfg = AutoConfig.from_pretrained(model_dir)
    assert cfg.model_type in ("bert", "roberta", "xlm-roberta"), \
        f"only BERT/RoBERTa-base encoders are on the B200 path, got {cfg.model_type}"
    tokenizer = AutoTokenizer.from_pretrained(model_dir, use_fast=False)
    roberta = cfg.model_type != "bert"
    ext = TextExtractor(common.load_hf_state_dict(model_dir), tokenizer, device=f"cuda:{gpu}",
                        ln_eps=cfg.layer_norm_eps,
                        position_offset=(cfg.pad_token_id + 1) if roberta else 0)
    df = pd.read_csv(trans_dir)
    col = "chinese" if language == "chinese" else "english"
    by_name = dict(zip(df["name"], df[col]))
    # one process per GPU under torchrun: this rank's share of the rows that do not have their .npy yet
    names, rank, world = shard.my_work(list(by_name), lambda n: os.path.join(save_dir, f"{n}.npy"))
    sents = [by_name[n] for n in names]
    if world > 1:
        print(f"rank {rank}/{world}: {len(names)} sentences on cuda:{gpu}")
    for s in range(0, len(names), sentences_per_launch):
        files = [os.path.join(save_dir, f"{n}.npy") for n in names[s:s + sentences_per_launch]]
        ext.extract_sentences(sents[s:s + sentences_per_launch], feature_level, save_files=files)
    print(f"Total {len(df)} files done! Time used ({model_name}): {time.time() - start_time:.1f}s.")


def build_parser():
    """Flags of extract_text_huggingface.py:270-281."""
    import argparse
    parser = argparse.ArgumentParser(description="Run.")
    parser.add_argument("--dataset", type=str, help="input dataset")
    parser.add_argument("--gpu", type=int, default=1, help="gpu id")
    parser.add_argument("--model_name", type=str, help="name of pretrained model")
    parser.add_argument("--feature_level", type=str, default="UTTERANCE", choices=["UTTERANCE", "FRAME"], help="output types")
    parser.add_argument("--punc_case", type=str, default=None, help="test punc impact to the performance")
    parser.add_argument("--language", type=str, default="chinese", help="used language")
    parser.add_argument("--model_dir", type=str, default=None, help="used user-defined model_dir")
    return parser


def main(args, config=None):
    """Script body (:283-299): transcription CSV and save directory from config.py, then extract_embedding."""
    if config is None:
        from .. import config as config  # noqa: PLW0127
    trans_dir = config.PATH_TO_TRANSCRIPTIONS[args.dataset]
    save_dir = config.PATH_TO_FEATURES[args.dataset]
    if args.punc_case is not None:
        assert args.punc_case in ["case1", "case2", "case3"]
        trans_dir = trans_dir[:-4] + f"-{args.punc_case}.csv"
        assert os.path.exists(trans_dir)
    extract_embedding(model_name=args.model_name, trans_dir=trans_dir, save_dir=save_dir, feature_level=args.feature_level,
                      gpu=args.gpu, punc_case=args.punc_case, language=args.language, model_dir=args.model_dir, config=config)


if __name__ == "__main__":
    main(build_parser().parse_args())
