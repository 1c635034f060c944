"""Frame-level feature shaping for feat_type = frm_align / frm_unalign (host side, numpy).

Contract (names and results) of MERBench/toolkit/utils/read_data.py:72-125 as Data_Feat applies it
(toolkit/data/feat_data.py:33-44): every clip is first compressed by ``feat_scale``, then either averaged to one row
('utt'), or -- optionally after audio / video are brought to the text length ('frm_align') -- padded to the longest
clip of its modality over the whole split.  Padding always goes IN FRONT because LSTMEncoder reads the final state.

Everything is one primitive, :func:`front_pool`: place the T rows at the END of a zeroed float64 buffer of
``dst_len * ceil(T / dst_len)`` rows and average each run of ``ceil(T / dst_len)`` consecutive rows.  The three
list-level rules only differ in how the target length of a clip is chosen, so they are instances of
:func:`_retarget` with a length policy.  Results are bit-identical to the reference's arrays, dtype included
(tests/test_host_logic.py against tests/golden/frame_shaping_golden.npz, generated from the unmodified module).
"""
from __future__ import annotations

import math

import numpy as np

_FEAT_TYPES = ("utt", "frm_align", "frm_unalign")


def front_pool(rows: np.ndarray, dst_len: int) -> np.ndarray:
    """[T, D] -> [dst_len, D].  T == dst_len: the input itself (dtype kept); otherwise float64: zero rows in front,
    then the mean over runs of ceil(T / dst_len) rows (run length 1 when T < dst_len, i.e. plain front padding)."""
    n_rows, dim = rows.shape
    if n_rows == dst_len:
        return rows
    run = max(1, -(-n_rows // dst_len))
    buf = np.zeros((dst_len * run, dim))
    buf[dst_len * run - n_rows:] = rows
    return buf if run == 1 else buf.reshape(dst_len, run, dim).mean(axis=1)


# the reference's name for the primitive (read_data.py:74)
func_mapping_feature = front_pool


def _retarget(modalities, target_len):
    """Apply front_pool clip by clip, in place; ``target_len(m, i)`` = destination length of clip i of modality m."""
    for m, clips in enumerate(modalities):
        for i, clip in enumerate(clips):
            clips[i] = front_pool(clip, target_len(m, i))
    return modalities


def align_to_utt(audios, texts, videos):
    """One row per clip: the mean over its frames (read_data.py:93-98)."""
    for clips in (audios, texts, videos):
        clips[:] = [clip.mean(axis=0) for clip in clips]
    return audios, texts, videos


def feature_scale_compress(audios, texts, videos, scale_factor=1):
    """Every clip to ceil(T / scale_factor) rows (read_data.py:101-106)."""
    mods = (audios, texts, videos)
    return _retarget(mods, lambda m, i: math.ceil(len(mods[m][i]) / scale_factor))


def align_to_text(audios, texts, videos):
    """Audio and video of clip i to the number of text rows of clip i (read_data.py:109-115)."""
    text_len = [len(clip) for clip in texts]
    return _retarget((audios, texts, videos), lambda m, i: text_len[i])


def pad_to_maxlen_pre_modality(audios, texts, videos):
    """Every clip to the longest clip of its own modality (read_data.py:118-126)."""
    longest = [max(map(len, clips)) for clips in (audios, texts, videos)]
    return _retarget((audios, texts, videos), lambda m, i: longest[m])


def shape_split(audios, texts, videos, feat_type, feat_scale):
    """Data_Feat.__init__ (feat_data.py:33-44) for one split: lists of (T_i, D) arrays -> lists ready for
    ``np.array(...)`` in the collater ([D] rows for 'utt', equal-length [T, D] otherwise)."""
    assert feat_scale >= 1 and feat_type in _FEAT_TYPES
    mods = feature_scale_compress(list(audios), list(texts), list(videos), feat_scale)
    if feat_type == "utt":
        return align_to_utt(*mods)
    if feat_type == "frm_align":
        mods = align_to_text(*mods)
    return pad_to_maxlen_pre_modality(*mods)
