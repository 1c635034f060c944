"""Device-side encoder objects: the "model" seam of the reference extractors.

Each class takes an HF-named ``state_dict`` (what ``AutoModel.from_pretrained(...).state_dict()``
holds in the reference, extract_*_huggingface.py) and exposes the fused forward + readout that the
reference spells as ``model(x, output_hidden_states=True).hidden_states`` followed by
``torch.stack(hs)[...]`` arithmetic.  All compute happens inside libmer_b200.so.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L
from . import weights as W


class MerVitModel(C.Structure):
    _fields_ = [("n_layers", C.c_int), ("ln_eps", C.c_float), ("patch_w", C.c_void_p),
                ("patch_b", C.c_void_p), ("cls_pos0", C.c_void_p), ("pos_rest", C.c_void_p),
                ("layers", C.POINTER(W.MerLayerWeights))]


class _Workspace:
    """Grow-only device scratch buffer (torch owns the memory)."""

    def __init__(self, device):
        self.device = device
        self.buf = None

    def get(self, nbytes):
        if self.buf is None or self.buf.numel() < nbytes:
            self.buf = None
            self.buf = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self.buf


class VitEncoder:
    """ViT-B/16 frame encoder (HF ``ViTModel``) + ``hidden_states[-1].sum(dim=1)`` readout.

    Reference: MERBench/feature_extraction/visual/extract_vision_huggingface.py:135-145."""

    def __init__(self, state_dict, device="cuda", ln_eps=1e-12):
        L.check(L.lib().mer_check_device())
        sd = W._np(state_dict)
        self.device = torch.device(device)
        self.pk = W.Packed(self.device)
        self.n_layers = W.count_layers(sd, VIT_PROBE)
        pw = sd["embeddings.patch_embeddings.projection.weight"]
        assert pw.shape == (768, 3, 16, 16), f"ViT-B/16 only, got patch weight {pw.shape}"
        pos = sd["embeddings.position_embeddings"][0]
        assert pos.shape == (197, 768), f"224x224 / patch 16 only, got pos {pos.shape}"
        self.patch_w = self.pk.keep(pw.reshape(768, 768), tf32=True)
        self.patch_b = self.pk.keep(sd["embeddings.patch_embeddings.projection.bias"])
        self.cls_pos0 = self.pk.keep(sd["embeddings.cls_token"].reshape(768) + pos[0])
        self.pos_rest = self.pk.keep(pos[1:])
        self.layers = W.pack_layers(sd, W.VIT_NAMES, self.n_layers, self.pk)
        self.model = MerVitModel(self.n_layers, ln_eps, self.patch_w.data_ptr(),
                                 self.patch_b.data_ptr(), self.cls_pos0.data_ptr(),
                                 self.pos_rest.data_ptr(), self.layers)
        self.ws = _Workspace(self.device)
        self._fwd = L.declare("mer_vit_forward", [C.POINTER(MerVitModel), C.c_void_p, C.c_int,
                                                  C.c_void_p, C.c_longlong, C.c_void_p,
                                                  C.c_void_p, C.c_void_p])
        L.lib().mer_vit_workspace_bytes.restype = C.c_longlong
        L.lib().mer_vit_workspace_bytes.argtypes = [C.c_int]

    def frame_features(self, frames_bgr_u8: torch.Tensor, return_hidden=False):
        """frames: uint8 CUDA tensor [N,224,224,3] (BGR).  Returns [N,768] fp32 (CUDA)."""
        assert frames_bgr_u8.dtype == torch.uint8 and frames_bgr_u8.is_cuda
        assert tuple(frames_bgr_u8.shape[1:]) == (224, 224, 3), \
            f"mer_vit_forward takes 224x224x3 frames, got {tuple(frames_bgr_u8.shape)}"
        frames = frames_bgr_u8.contiguous()
        n = frames.shape[0]
        need = L.lib().mer_vit_workspace_bytes(n)
        ws = self.ws.get(need)
        out = torch.empty(n, 768, dtype=torch.float32, device=self.device)
        hidden = None
        if return_hidden:
            hidden = torch.empty(self.n_layers + 1, n * 197, 768, dtype=torch.float32,
                                 device=self.device)
        L.check(self._fwd(C.byref(self.model), L.ptr(frames), n, L.ptr(ws), ws.numel(),
                          L.ptr(out), L.ptr(hidden), L.stream_ptr()))
        if return_hidden:
            return out, hidden.view(self.n_layers + 1, n, 197, 768)
        return out


VIT_PROBE = "encoder.layer.{i}.layernorm_before.weight"
