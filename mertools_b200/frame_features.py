"""Frame-level feature shaping of the reference's data path (feat_type = frm_align / frm_unalign), host side.

Mirrors MERBench/toolkit/utils/read_data.py:72-125 and the order in which Data_Feat applies them
(toolkit/data/feat_data.py:33-44): pre-compress every clip by ``feat_scale`` (mean-pool groups of adjacent
frames, zero-padded IN FRONT), optionally align audio / video to the text length, then pad every modality to
its maximum length over the whole split -- again in front, because LSTMEncoder reads the final state.
Pure numpy, same arithmetic and dtypes (float64 as soon as a zero block is concatenated), so the arrays are
bit-identical to the reference's (tests/test_host_logic.py against tests/golden/frame_shaping_golden.npz).
"""
from __future__ import annotations

import math

import numpy as np


def func_mapping_feature(feature, dst_len):
    """(seqlen, featdim) -> (dst_len, featdim): identity, zero pre-padding, or mean-pooling of
    ceil(seqlen / dst_len) adjacent frames after zero pre-padding to a multiple (read_data.py:74-90)."""
    featlen, featdim = feature.shape
    if featlen == dst_len:
        return feature
    if featlen < dst_len:
        pad_feature = np.zeros((dst_len - featlen, featdim))
        return np.concatenate((pad_feature, feature), axis=0)
    if featlen // dst_len == featlen / dst_len:
        pad_len = 0
        pool_size = featlen // dst_len
    else:
        pad_len = dst_len - featlen % dst_len
        pool_size = featlen // dst_len + 1
    pad_feature = np.zeros((pad_len, featdim))
    feature = np.concatenate([pad_feature, feature]).reshape(dst_len, pool_size, featdim)
    return np.mean(feature, axis=1)


def align_to_utt(audios, texts, videos):
    """read_data.py:93-98."""
    for ii in range(len(audios)):
        audios[ii] = np.mean(audios[ii], axis=0)
        texts[ii] = np.mean(texts[ii], axis=0)
        videos[ii] = np.mean(videos[ii], axis=0)
    return audios, texts, videos


def feature_scale_compress(audios, texts, videos, scale_factor=1):
    """read_data.py:101-106."""
    for ii in range(len(audios)):
        audios[ii] = func_mapping_feature(audios[ii], math.ceil(len(audios[ii]) / scale_factor))
        texts[ii] = func_mapping_feature(texts[ii], math.ceil(len(texts[ii]) / scale_factor))
        videos[ii] = func_mapping_feature(videos[ii], math.ceil(len(videos[ii]) / scale_factor))
    return audios, texts, videos


def align_to_text(audios, texts, videos):
    """read_data.py:109-115."""
    for ii in range(len(audios)):
        dst_len = len(texts[ii])
        audios[ii] = func_mapping_feature(audios[ii], dst_len)
        texts[ii] = func_mapping_feature(texts[ii], dst_len)
        videos[ii] = func_mapping_feature(videos[ii], dst_len)
    return audios, texts, videos


def pad_to_maxlen_pre_modality(audios, texts, videos):
    """read_data.py:118-126."""
    audio_maxlen = max(len(feature) for feature in audios)
    text_maxlen = max(len(feature) for feature in texts)
    video_maxlen = max(len(feature) for feature in videos)
    for ii in range(len(audios)):
        audios[ii] = func_mapping_feature(audios[ii], audio_maxlen)
        texts[ii] = func_mapping_feature(texts[ii], text_maxlen)
        videos[ii] = func_mapping_feature(videos[ii], video_maxlen)
    return audios, texts, videos


def shape_split(audios, texts, videos, feat_type, feat_scale):
    """Data_Feat.__init__ (feat_data.py:33-44) for one split: lists of (T_i, D) arrays -> lists ready for
    ``np.array(...)`` in the collater ([D] rows for 'utt', equal-length [T, D] otherwise)."""
    assert feat_scale >= 1 and feat_type in ("utt", "frm_align", "frm_unalign")
    audios, texts, videos = feature_scale_compress(list(audios), list(texts), list(videos), feat_scale)
    if feat_type == "utt":
        return align_to_utt(audios, texts, videos)
    if feat_type == "frm_align":
        audios, texts, videos = align_to_text(audios, texts, videos)
    return pad_to_maxlen_pre_modality(audios, texts, videos)
