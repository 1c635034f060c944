"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the per-clip logic of the three reference extractors, around the encoder
restatements of oracle/encoders.py.  Every function cites the reference lines it follows
(paths relative to MERBench/feature_extraction/ in /root/reference).  Pinned against the outputs
of the UNMODIFIED reference scripts by tests/test_oracle.py + tests/golden/ (the reference has
no tests of its own for this path: SURVEY.md §4).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import encoders as E


# ------------------------------------------------------------------------------------------------
# visual  (visual/extract_vision_huggingface.py)
# ------------------------------------------------------------------------------------------------
def resample_frames_uniform_indices(vlen, nframe=16):
    """Frame indices of ``resample_frames_uniform`` (:44-56): uniform float stride, truncate to
    int, pad with the last index, cut to nframe.  Same rule in toolkit/utils/functions.py:85-104
    (load_video_from_npy, n_frms=8, 'uniform')."""
    n_upd = min(nframe, vlen)
    indices = np.arange(0, vlen, vlen / n_upd).astype(int).tolist()
    while len(indices) < nframe:
        indices.append(indices[-1])
    return indices[:nframe]


def split_into_batch(inputs, bsize=32):
    """:58-63."""
    return [inputs[i * bsize:(i + 1) * bsize] for i in range(math.ceil(len(inputs) / bsize))]


def _pil_filter(name):
    """(filter function, support) of Pillow's resampling filters (Resample.c: bilinear_filter, bicubic_filter
    with a = -0.5)."""
    if name == "bilinear":
        return (lambda x: 1.0 - x if x < 1.0 else 0.0), 1.0

    def bicubic(x, a=-0.5):
        if x < 1.0:
            return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        if x < 2.0:
            return (((x - 5) * x + 8) * x - 4) * a
        return 0.0
    return bicubic, 2.0


def pil_bilinear_coeffs(in_size, out_size, filter="bilinear"):
    """8-bit coefficient table of Pillow's ``Image.resize(..., BILINEAR | BICUBIC)`` along one axis
    (Pillow src/libImaging/Resample.c: precompute_coeffs + normalize_coeffs_8bpc; Pillow is an un-vendored
    dependency of HF ``ViTImageProcessor.resize``, reached from extract_vision_huggingface.py:137-138;
    the container's Pillow 12.2 and transformers 5.5 agree with this table for up-scaling, pinned in
    tests/test_oracle.py).  Returns (xmin[out], count[out], kk[out, ksize] int32), PRECISION_BITS = 22."""
    fn, fsupport = _pil_filter(filter)
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = fsupport * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        lo = int(center - support + 0.5)
        lo = max(lo, 0)
        hi = int(center + support + 0.5)
        hi = min(hi, in_size)
        n = hi - lo
        w = np.zeros(n, np.float64)
        for x in range(n):
            w[x] = fn(abs((x + lo - center + 0.5) * ss))
        tot = 0.0                      # Pillow accumulates the weights in index order
        for x in range(n):
            tot += w[x]
        if tot != 0.0:
            w = w / tot
        xmin[xx], cnt[xx] = lo, n
        # normalize_coeffs_8bpc: round half away from zero (bicubic weights can be negative)
        kk[xx, :n] = np.trunc(w * (1 << 22) + np.where(w < 0, -0.5, 0.5)).astype(np.int32)
    return xmin, cnt, kk


def _pil_resample_axis(img, out_size, axis, filter="bilinear"):
    """One Pillow resampling pass over uint8 ``img`` [..., H, W, C] along ``axis`` (-3 rows, -2 columns):
    out = clip8((2^21 + sum_k in[xmin + k] * kk[k]) >> 22)."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, cnt, kk = pil_bilinear_coeffs(in_size, out_size, filter)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.uint8)
    for xx in range(out_size):
        acc = np.full(src.shape[1:], 1 << 21, np.int64)
        for k in range(cnt[xx]):
            acc += src[xmin[xx] + k] * int(kk[xx, k])
        out[xx] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def pil_resize_bilinear_u8(frames, out_h=224, out_w=224, filter="bilinear"):
    """``PIL.Image.resize((out_w, out_h), BILINEAR | BICUBIC)`` on uint8 [..., H, W, C]: horizontal pass,
    then vertical pass, uint8 in between (ImagingResample)."""
    f = np.asarray(frames)
    assert f.dtype == np.uint8
    return _pil_resample_axis(_pil_resample_axis(f, out_w, -2, filter), out_h, -3, filter)


def vit_preprocess(frames_bgr):
    """``func_opencv_to_image`` (:29-31, BGR->RGB) + HF ``ViTImageProcessor`` as called at
    :137-138: resize to 224x224 (PIL bilinear on uint8; the identity for 224x224 frames), rescale by
    1/255, then (x - 0.5) / 0.5, fp32, NCHW."""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    if f.shape[1:3] != (224, 224):
        f = pil_resize_bilinear_u8(f, 224, 224)
    rgb = f[..., ::-1].astype(np.float32)
    x = rgb * np.float32(1.0 / 255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def imagenet_preprocess(frames_bgr):
    """FaceDataset.__getitem__ (dataset.py:40-47: BGR->RGB PIL image) + the transform of
    extract_imagenet_embedding.py:52-54: Resize((224, 224)) (PIL bilinear), ToTensor, Normalize."""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    if f.shape[1:3] != (224, 224):
        f = pil_resize_bilinear_u8(f, 224, 224)
    rgb = f[..., ::-1].astype(np.float32) / np.float32(255.0)
    x = (rgb - np.asarray(IMAGENET_MEAN, np.float32)) / np.asarray(IMAGENET_STD, np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def imagenet_clip_features(sd, frames_bgr, feature_level="UTTERANCE"):
    """One video through extract_imagenet_embedding.py:57-94: batches of 32 frames, [N, 512] embeddings,
    FRAME -> [T, 512] (zeros((1, 512)) when empty), UTTERANCE -> mean over frames."""
    frames = np.asarray(frames_bgr)
    if len(frames) == 0:
        return np.zeros((1, 512)) if feature_level == "FRAME" else np.zeros((512,))
    x = imagenet_preprocess(frames)
    emb = torch.cat([E.resnet18_features(sd, b) for b in split_into_batch(x, 32)], dim=0).float().numpy()
    emb = np.array(emb).squeeze()
    if feature_level == "FRAME":
        return emb[np.newaxis, :] if emb.ndim == 1 else emb
    return np.mean(emb, axis=0) if emb.ndim == 2 else emb


def whisper_mel_filters(n_freq=201, n_mel=80, fmin=0.0, fmax=8000.0, sr=16000):
    """The [201, 80] filter bank WhisperFeatureExtractor builds (HF audio_utils.mel_filter_bank with
    norm="slaney", mel_scale="slaney"): triangles on the Slaney mel scale (linear below 1 kHz, logarithmic above),
    each scaled by 2 / (its band width in Hz).  Equal to ``WhisperFeatureExtractor().mel_filters`` (tests)."""
    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        m = 3.0 * f / 200.0
        lg = f >= 1000.0
        m[lg] = 15.0 + np.log(f[lg] / 1000.0) * (27.0 / np.log(6.4))
        return m

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f = 200.0 * m / 3.0
        lg = m >= 15.0
        f[lg] = 1000.0 * np.exp((np.log(6.4) / 27.0) * (m[lg] - 15.0))
        return f
    fft_freqs = np.linspace(0, sr // 2, n_freq)
    f_pts = mel_to_hz(np.linspace(hz_to_mel(np.array([fmin]))[0], hz_to_mel(np.array([fmax]))[0], n_mel + 2))
    fdiff = np.diff(f_pts)
    slopes = f_pts[None, :] - fft_freqs[:, None]
    fb = np.maximum(0, np.minimum(-slopes[:, :-2] / fdiff[:-1], slopes[:, 2:] / fdiff[1:]))
    return fb * (2.0 / (f_pts[2:n_mel + 2] - f_pts[:n_mel]))[None, :]


def whisper_log_mel(wave):
    """``WhisperFeatureExtractor()(wave, sampling_rate=16000).input_features[0]`` (the reference's Whisper branch,
    extract_audio_huggingface.py:85): zero-pad / cut to 30 s, reflect-padded STFT (periodic Hann 400, hop 160),
    power spectrum, mel filter bank, log10 (floor 1e-10), drop the last frame, clamp to max - 8, (x + 4) / 4.
    Returns float32 [80, 3000]."""
    x = np.zeros(480000, np.float64)
    n = min(len(wave), 480000)
    x[:n] = np.asarray(wave, np.float64)[:n]
    p = np.pad(x, (200, 200), mode="reflect")
    nfr = 1 + (len(p) - 400) // 160
    idx = np.arange(400)[None, :] + 160 * np.arange(nfr)[:, None]
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(400) / 400)
    spec = np.abs(np.fft.rfft(p[idx] * win, n=400, axis=1)) ** 2
    lm = np.log10(np.maximum(1e-10, spec @ whisper_mel_filters())).T[:, :-1]
    lm = np.maximum(lm, lm.max() - 8.0)
    return ((lm + 4.0) / 4.0).astype(np.float32)


def whisper_clip_features(sd, wave, start_token, feature_level="UTTERANCE", heads=8):
    """One file through the Whisper branch of extract_audio_huggingface.py:83-110: log-mel features, the model with
    ``decoder_input_ids = [[start, start]]``, ``last_hidden_state[0]`` = [2, D]; UTTERANCE -> mean over the two rows."""
    f = torch.from_numpy(whisper_log_mel(wave))[None]
    with torch.no_grad():
        feat = E.whisper_last_hidden_state(sd, f, torch.tensor([[start_token, start_token]]), heads=heads)[0].numpy()
    feat = np.array(feat).squeeze()
    if feature_level == "UTTERANCE" and len(feat.shape) != 1:
        feat = np.mean(feat, axis=0)
    return feat


def cv2_resize_linear_u8(img, oh, ow):
    """``cv2.resize(img, (ow, oh))`` (INTER_LINEAR, the default) for uint8 [H, W, C] images, restated from OpenCV's
    fixed-point resize (imgproc/src/resize.cpp: 11-bit coefficients, HResizeLinear then the 8-bit VResizeLinear
    ``((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2 >> 2``; the x fraction is reset at the borders, the y
    rows are only clamped; an exact 2x downscale takes the 2 x 2 area average).  Bit-exact against cv2 4.13 on the
    sizes tests/test_oracle.py tries.  Used by the EmoNet extractor's DataAugmentor (emonet/data_augmentation.py:77)."""
    img = np.asarray(img)
    ih, iw = img.shape[:2]
    a = img.astype(np.int32)
    if ih == 2 * oh and iw == 2 * ow:
        return ((a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2).astype(np.uint8)

    def frac(isz, osz):
        f = ((np.arange(osz) + 0.5) * (isz / osz) - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int32)
        return s, (f - s).astype(np.float32)

    def q(f):
        return (np.rint((np.float32(1.0) - f) * np.float32(2048)).astype(np.int32),
                np.rint(f * np.float32(2048)).astype(np.int32))
    sx, fx = frac(iw, ow)
    fx[sx < 0] = 0
    sx[sx < 0] = 0
    hi = sx >= iw - 1
    fx[hi] = 0
    sx[hi] = iw - 1
    ax0, ax1 = q(fx)
    sy, fy = frac(ih, oh)
    ay0, ay1 = q(fy)
    h = a[:, sx] * ax0[None, :, None] + a[:, np.minimum(sx + 1, iw - 1)] * ax1[None, :, None]
    r0, r1 = h[np.clip(sy, 0, ih - 1)], h[np.clip(sy + 1, 0, ih - 1)]
    out = (((ay0[:, None, None] * (r0 >> 4)) >> 16) + ((ay1[:, None, None] * (r1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def emonet_preprocess(frames_bgr):
    """FaceDatasetForEmoNet.__getitem__ (dataset.py:72-86: BGR -> RGB) + DataAugmentor(256, 256) without a bounding
    box (emonet/data_augmentation.py:68-87: cv2.resize to 256 x 256; the two warpAffine calls are identities at
    rotation 0 / scale 1 / translation 0) + ToTensor.  Returns [N, 3, 256, 256] in [0, 1]."""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    rgb = np.stack([cv2_resize_linear_u8(x[..., ::-1], 256, 256) for x in f])
    return torch.from_numpy(np.ascontiguousarray((rgb.astype(np.float32) / np.float32(255.0)).transpose(0, 3, 1, 2)))


def emonet_clip_features(sd, frames_bgr, feature_level="UTTERANCE"):
    """One video through extract_emonet_embedding.py:62-94: batches of 32 frames -> [N, 256]; FRAME -> [T, 256]
    (zeros((1, D)) when empty), UTTERANCE -> mean over frames."""
    frames = np.asarray(frames_bgr)
    if len(frames) == 0:
        return np.zeros((1, 256)) if feature_level == "FRAME" else np.zeros((256,))
    x = emonet_preprocess(frames)
    emb = torch.cat([E.emonet_embedding(sd, b) for b in split_into_batch(x, 32)], dim=0).float().numpy()
    emb = np.array(emb).squeeze()
    if feature_level == "FRAME":
        return emb[np.newaxis, :] if emb.ndim == 1 else emb
    return np.mean(emb, axis=0) if emb.ndim == 2 else emb


def dinov2_preprocess(frames_bgr, mean=IMAGENET_MEAN, std=IMAGENET_STD, resize=256, crop=224):
    """``AutoImageProcessor`` of dinov2-large / -giant = BitImageProcessor (preprocessor_config.json: shortest edge 256,
    BICUBIC, centre crop 224, rescale 1 / 255, ImageNet mean / std) on ``func_opencv_to_image`` frames (BGR -> RGB PIL,
    extract_vision_huggingface.py:37-41,137-138).  Returns [n, 3, 224, 224]."""
    f = np.asarray(frames_bgr)
    h, w = f.shape[1:3]
    nh, nw = (resize, int(resize * w / h)) if h <= w else (int(resize * h / w), resize)
    if (nh, nw) != (h, w):
        f = pil_resize_bilinear_u8(f, nh, nw, filter="bicubic")
    top, left = (nh - crop) // 2, (nw - crop) // 2
    f = f[:, top:top + crop, left:left + crop]
    x = (f[..., ::-1].astype(np.float32) / np.float32(255.0) - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def dinov2_clip_features(sd, frames_bgr, feature_level="UTTERANCE", heads=16, nframe=64, bsize=32):
    """One video through the DINOv2 branch (:135-145, :175-189): ``resample_frames_uniform(frames, 64)``, batches of 32,
    ``hidden_states[-1].sum(dim=1)`` per frame -> FRAME [64, D]; UTTERANCE -> mean over the frames."""
    f = np.asarray(frames_bgr)
    f = f[resample_frames_uniform_indices(len(f), nframe)]
    x = dinov2_preprocess(f)
    with torch.no_grad():
        emb = torch.cat([E.dinov2_hidden_states(sd, x[s:s + bsize], heads=heads)[-1].sum(dim=1)
                         for s in range(0, len(x), bsize)]).numpy().squeeze()
    return np.mean(emb, axis=0) if feature_level == "UTTERANCE" and emb.ndim == 2 else emb


def videomae_preprocess(frames_bgr, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """extract_vision_huggingface.py:148-152: ``resample_frames_uniform(frames)`` (16 frames), BGR -> RGB
    (func_opencv_to_numpy), VideoMAEImageProcessor: shortest edge -> 224 (bilinear, Pillow arithmetic as for the other
    processors), centre crop 224, / 255, normalise.  Returns [1, 16, 3, 224, 224]."""
    f = np.asarray(frames_bgr)
    f = f[resample_frames_uniform_indices(len(f), 16)]
    h, w = f.shape[1:3]
    nh, nw = (224, int(224 * w / h)) if h <= w else (int(224 * h / w), 224)
    if (nh, nw) != (h, w):
        f = pil_resize_bilinear_u8(f, nh, nw)
    top, left = (nh - 224) // 2, (nw - 224) // 2
    f = f[:, top:top + 224, left:left + 224]
    x = (f[..., ::-1].astype(np.float32) / np.float32(255.0) - np.asarray(mean, np.float32)) / np.asarray(std, np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))[None]


def videomae_clip_features(sd, frames_bgr, feature_level="UTTERANCE", heads=12, mean=IMAGENET_MEAN, std=IMAGENET_STD):
    """One video through the VideoMAE branch (:147-159, :175-189): last_hidden_state [1568, D] -> view(8, 196, D)
    .mean(1) -> FRAME [8, D]; UTTERANCE -> mean over the 8 tubelets."""
    with torch.no_grad():
        hs = E.videomae_last_hidden_state(sd, videomae_preprocess(frames_bgr, mean, std), heads=heads)
    emb = hs.view(8, 196, -1).mean(dim=1).numpy().squeeze()
    return np.mean(emb, axis=0) if feature_level == "UTTERANCE" and emb.ndim == 2 else emb


def manet_preprocess(frames_bgr):
    """FaceDataset.__getitem__ (dataset.py:40-47) + the transform of extract_manet_embedding.py:60-61:
    Resize((224, 224)) (PIL bilinear) and ToTensor only (no normalisation).  Returns [N, 3, 224, 224] in [0, 1]."""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    if f.shape[1:3] != (224, 224):
        f = pil_resize_bilinear_u8(f, 224, 224)
    rgb = f[..., ::-1].astype(np.float32) / np.float32(255.0)
    return torch.from_numpy(np.ascontiguousarray(rgb.transpose(0, 3, 1, 2)))


def manet_clip_features(sd, frames_bgr, feature_level="UTTERANCE"):
    """One video through extract_manet_embedding.py:69-103: batches of 32 frames -> [N, 1024]; FRAME -> [T, 1024]
    (zeros((1, D)) when empty), UTTERANCE -> mean over frames."""
    frames = np.asarray(frames_bgr)
    if len(frames) == 0:
        return np.zeros((1, 1024)) if feature_level == "FRAME" else np.zeros((1024,))
    x = manet_preprocess(frames)
    emb = torch.cat([E.manet_embedding(sd, b) for b in split_into_batch(x, 32)], dim=0).float().numpy()
    emb = np.array(emb).squeeze()
    if feature_level == "FRAME":
        return emb[np.newaxis, :] if emb.ndim == 1 else emb
    return np.mean(emb, axis=0) if emb.ndim == 2 else emb


def ferplus_preprocess(frames_bgr):
    """FaceDataset.__getitem__ (dataset.py:40-47) + compose_transforms (extract_ferplus_embedding.py:62-74) for the
    FER+ models (meta std == [1, 1, 1]): Resize(256) (shorter side -> 256, PIL bilinear), CenterCrop(224), ToTensor,
    x * 255, Normalize(mean, 1).  frames: uint8 [N, H, W, 3] BGR.  Returns [N, 3, 224, 224]."""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    h, w = f.shape[1:3]
    if min(h, w) != 256:  # torchvision Resize(int): the long side is int(256 * long / short)
        nh, nw = (256, int(256 * w / h)) if h <= w else (int(256 * h / w), 256)
        f = pil_resize_bilinear_u8(f, nh, nw)
    h, w = f.shape[1:3]
    top, left = int(round((h - 224) / 2.0)), int(round((w - 224) / 2.0))
    f = f[:, top:top + 224, left:left + 224]
    rgb = (f[..., ::-1].astype(np.float32) / np.float32(255.0)) * np.float32(255.0)
    x = rgb - np.asarray(E.FERPLUS_MEAN, np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def ferplus_clip_features(sd, frames_bgr, feature_level="UTTERANCE"):
    """One video through extract_ferplus_embedding.py:150-194 (resnet50_ferplus_dag, layer conv5_3_3x3_relu): batches
    of 32 frames -> [N, 512]; FRAME -> [T, 512] (zeros((1, D)) when empty), UTTERANCE -> mean over frames."""
    frames = np.asarray(frames_bgr)
    if len(frames) == 0:
        return np.zeros((1, 512)) if feature_level == "FRAME" else np.zeros((512,))
    x = ferplus_preprocess(frames)
    emb = torch.cat([E.ferplus_resnet50_features(sd, b) for b in split_into_batch(x, 32)], dim=0).float().numpy()
    emb = np.array(emb).squeeze()
    if feature_level == "FRAME":
        return emb[np.newaxis, :] if emb.ndim == 1 else emb
    return np.mean(emb, axis=0) if emb.ndim == 2 else emb


CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def clip_preprocess(frames_bgr, size=224):
    """``func_opencv_to_image`` (BGR->RGB) + HF ``CLIPImageProcessor`` as the CLIP branch calls it
    (extract_vision_huggingface.py:115-116): resize the SHORTER edge to 224 (PIL bicubic on uint8, the
    longer edge scaled by the same ratio and truncated), center crop 224x224, x/255, per-channel
    (x - mean) / std, fp32 NCHW.  (transformers 5.5 resizes through torchvision and differs from Pillow by
    at most one uint8 level on some pixels; the pinned 4.28 processor is Pillow -- this follows Pillow.)"""
    f = np.asarray(frames_bgr)
    assert f.dtype == np.uint8 and f.ndim == 4 and f.shape[-1] == 3, f.shape
    h, w = f.shape[1:3]
    short, long = (w, h) if w <= h else (h, w)
    new_short, new_long = size, int(size * long / short)
    nh, nw = (new_long, new_short) if w <= h else (new_short, new_long)
    if (nh, nw) != (h, w):
        f = pil_resize_bilinear_u8(f, nh, nw, filter="bicubic")
    top, left = (nh - size) // 2, (nw - size) // 2
    f = f[:, top:top + size, left:left + size]
    rgb = f[..., ::-1].astype(np.float32) * np.float32(1.0 / 255.0)
    x = (rgb - np.asarray(CLIP_MEAN, np.float32)) / np.asarray(CLIP_STD, np.float32)
    return torch.from_numpy(np.ascontiguousarray(x.transpose(0, 3, 1, 2)))


def clip_visual_features(sd, frames_bgr, layers=12, heads=12, feature_level="UTTERANCE", dtype=torch.float32):
    """One clip through the CLIP branch (:114-122) and the save logic (:171-189): every frame (no
    resampling), batches of 32, ``get_image_features``; UTTERANCE -> mean over frames."""
    inputs = clip_preprocess(np.asarray(frames_bgr))
    embs = [E.clip_image_features(sd, b, layers=layers, heads=heads, dtype=dtype)[0] for b in split_into_batch(inputs, 32)]
    emb = np.array(torch.cat(embs, dim=0).float().squeeze().numpy()).squeeze()
    if feature_level == "FRAME":
        return emb[np.newaxis, :] if emb.ndim == 1 else emb
    return np.mean(emb, axis=0) if emb.ndim == 2 else emb


def visual_clip_features(sd, frames_bgr, nframe=None, layers=12, feature_level="UTTERANCE",
                         dtype=torch.float32):
    """One clip through the DINOv2/data2vec branch (:135-145 / :125-133) and the save logic
    (:171-189).  ``nframe=64`` reproduces the DINOv2 branch of the script (resample to 64 frames);
    ``nframe=None`` the data2vec branch (all frames); ``nframe=8`` the 8-frame variant of
    toolkit/utils/functions.py:79.  Returns the array the script would ``np.save``."""
    frames = np.asarray(frames_bgr)
    if nframe is not None:
        frames = frames[resample_frames_uniform_indices(len(frames), nframe)]
    inputs = vit_preprocess(frames)
    embs = []
    for batch in split_into_batch(inputs, 32):
        if "encoder.layer.0.lambda_1" in sd:   # Data2VecVisionModel (data2vec-vision-base-ft1k): the BEiT graph
            hs = E.data2vec_vision_hidden_states(sd, batch, dtype=dtype)
        else:
            hs = E.vit_hidden_states(sd, batch, layers=layers, dtype=dtype)
        embs.append(torch.stack(hs)[-1].sum(dim=1))  # :143-144
    emb = torch.cat(embs, dim=0).float().squeeze().numpy()  # :171
    emb = np.array(emb).squeeze()
    if feature_level == "FRAME":
        if len(emb) == 0:
            emb = np.zeros((1, 768))
        elif emb.ndim == 1:
            emb = emb[np.newaxis, :]
        return emb
    if len(emb) == 0:
        return np.zeros((768,))
    if emb.ndim == 2:
        emb = np.mean(emb, axis=0)
    return emb


# ------------------------------------------------------------------------------------------------
# audio  (audio/extract_audio_huggingface.py)
# ------------------------------------------------------------------------------------------------
def wav2vec2_normalize(samples):
    """HF ``Wav2Vec2FeatureExtractor(do_normalize=True)`` as called at :94: cast the float64
    samples of ``sf.read`` to float32 FIRST, then (x - mean) / sqrt(var + 1e-7) with numpy's
    float32 reductions (HF feature_extraction_wav2vec2.py:78-97,205-236)."""
    x = np.asarray(samples)
    if x.dtype == np.float64:
        x = x.astype(np.float32)
    x = np.asarray(x, dtype=np.float32)
    return (x - x.mean()) / np.sqrt(x.var() + 1e-7)


def audio_split_into_batch(input_values, maxlen=16000 * 10):
    """:40-50 — waveforms longer than 10 s are zero-padded to a multiple of 10 s AFTER
    normalisation and reshaped into independent rows."""
    if len(input_values[0]) <= maxlen:
        return input_values
    bs, wavlen = input_values.shape
    assert bs == 1
    tgt = math.ceil(wavlen / maxlen) * maxlen
    out = torch.zeros((1, tgt))
    out[:, :wavlen] = input_values
    return out.view(-1, maxlen)


def audio_clip_features(sd, samples, layers=12, feature_level="UTTERANCE", dtype=torch.float32, heads=12):
    """One wav through ``extract`` (:72-110): normalise, chunk, HuBERT, sum of the last four
    hidden states (:98), flatten (B*T, D) (:100), UTTERANCE -> mean over axis 0 (:105-108)."""
    iv = torch.from_numpy(wav2vec2_normalize(samples))[None]
    iv = audio_split_into_batch(iv)
    hs = E.hubert_hidden_states(sd, iv, layers=layers, heads=heads, dtype=dtype)
    feat = torch.stack(hs)[[-4, -3, -2, -1]].sum(dim=0)
    feat = feat.reshape(-1, feat.shape[-1]).float().squeeze().numpy()
    if feature_level == "UTTERANCE":
        feat = np.array(feat).squeeze()
        if feat.ndim != 1:
            feat = np.mean(feat, axis=0)
    return feat


# ------------------------------------------------------------------------------------------------
# text  (text/extract_text_huggingface.py)
# ------------------------------------------------------------------------------------------------
def find_start_end_pos(tokenizer):
    """:90-114 — probe the tokenizer with '今天天气真好' to find how many special tokens wrap a
    sentence.  BERT/RoBERTa-style tokenizers give (1, -1)."""
    sentence = "今天天气真好"
    input_ids = tokenizer(sentence, return_tensors="pt")["input_ids"][0]
    start, end = None, None
    for start in range(0, 3, 1):
        outputs = tokenizer.decode(input_ids[start:]).replace(" ", "")
        if outputs == sentence:
            return start, None
        if outputs.startswith(sentence):
            break
    for end in range(-1, -3, -1):
        outputs = tokenizer.decode(input_ids[start:end]).replace(" ", "")
        if outputs == sentence:
            break
    assert tokenizer.decode(input_ids[start:end]).replace(" ", "") == sentence
    return start, end


def text_clip_features(sd, input_ids, start, end, layers=12, feature_level="UTTERANCE",
                       position_offset=0, eps=1e-12, dtype=torch.float32):
    """One sentence through :222-249: ids -> BERT -> sum of last four hidden states -> strip
    [start:end] -> UTTERANCE mean.  ``input_ids`` empty / None reproduces the empty-sentence branch
    (zeros, float64, :236-249)."""
    embeddings = []
    if input_ids is not None and len(input_ids) > 0:
        hs = E.bert_hidden_states(sd, input_ids, layers=layers, eps=eps,
                                  position_offset=position_offset, dtype=dtype)
        out = torch.stack(hs)[[-4, -3, -2, -1]].sum(dim=0).float().numpy()
        embeddings = out[0, start:end]
    emb = np.array(embeddings).squeeze()
    if feature_level == "FRAME":
        if len(emb) == 0:
            return np.zeros((1, 768))
        if emb.ndim == 1:
            emb = emb[np.newaxis, :]
        return emb
    if len(emb) == 0:
        return np.zeros((768,))
    if emb.ndim == 2:
        emb = np.mean(emb, axis=0)
    return emb


# ------------------------------------------------------------------------------------------------
# log-mel front-end  (audio/vggish/mel_features.py, vggish_input.py, vggish_params.py)
# ------------------------------------------------------------------------------------------------
def log_mel_spectrogram(data, sample_rate=16000, log_offset=0.01, window_secs=0.025, hop_secs=0.010,
                        num_mel_bins=64, lower_edge_hertz=125.0, upper_edge_hertz=7500.0):
    """mel_features.log_mel_spectrogram (:166-223) with the VGGish constants (vggish_params.py:22-34):
    frames (no padding, :21-45) x periodic Hann (:48-69) -> |rfft| (:72-93) -> HTK mel matrix (:96-164)
    -> log(mel + offset).  float64 numpy like the reference."""
    data = np.asarray(data, dtype=np.float64)
    wl = int(round(sample_rate * window_secs))
    hl = int(round(sample_rate * hop_secs))
    nfft = 2 ** int(np.ceil(np.log(wl) / np.log(2.0)))
    nfrm = 1 + int(np.floor((data.shape[0] - wl) / hl))
    frames = np.stack([data[i * hl:i * hl + wl] for i in range(nfrm)])
    window = 0.5 - 0.5 * np.cos(2 * np.pi / wl * np.arange(wl))
    spec = np.abs(np.fft.rfft(frames * window, nfft))
    nbins = spec.shape[1]

    def h2m(hz):
        return 1127.0 * np.log(1.0 + hz / 700.0)
    bins_mel = h2m(np.linspace(0.0, sample_rate / 2.0, nbins))
    edges = np.linspace(h2m(lower_edge_hertz), h2m(upper_edge_hertz), num_mel_bins + 2)
    mel = np.empty((nbins, num_mel_bins))
    for i in range(num_mel_bins):
        lo, ce, up = edges[i:i + 3]
        mel[:, i] = np.maximum(0.0, np.minimum((bins_mel - lo) / (ce - lo), (up - bins_mel) / (up - ce)))
    mel[0, :] = 0.0
    return np.log(np.dot(spec, mel) + log_offset)


def waveform_to_examples(data, hop_sec, num_frames=96):
    """vggish_input.waveform_to_examples (:37-82) for 16 kHz mono input: [num_examples, 96, 64]."""
    lm = log_mel_spectrogram(data)
    hop = int(round(hop_sec * 100.0))
    n = 1 + int(np.floor((lm.shape[0] - num_frames) / hop))
    return np.stack([lm[i * hop:i * hop + num_frames] for i in range(n)]) if n > 0 else np.zeros((0, num_frames, lm.shape[1]))
