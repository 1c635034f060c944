"""HF ``state_dict`` -> device operands of libmer_b200.so.

The packing is pure layout work done once at load time: concatenate Q|K|V, flatten the patch
conv, fold the HuBERT weight-norm, permute conv kernels to [out, tap, in] for the time-major
implicit GEMM, and round every tensor-core operand to TF32 (round-to-nearest, on the device).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class MerLayerWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in (
        "ln1_g", "ln1_b", "w_qkv", "b_qkv", "w_o", "b_o", "ln2_g", "ln2_b",
        "w_fc1", "b_fc1", "w_fc2", "b_fc2")]


def _dev(x, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    return x.detach().to(device=device, dtype=torch.float32).contiguous()


class Packed:
    """Owns the device tensors (keeps them alive) and the ctypes layer array."""

    def __init__(self, device):
        self.device = device
        self.tensors = []

    def keep(self, x, tf32=False, split=False, f16=False):
        """tf32: round-to-nearest in place (TF32 GEMM operand); split: bf16 hi|lo rows (BF16X3);
        f16: IEEE fp16, round-to-nearest, saturating (F16 GEMM operand)."""
        t = _dev(x, self.device)
        if tf32:
            L.round_tf32_(t)
        if split:
            t = L.split_bf16(t)
        if f16:
            t = t.clamp(-65504.0, 65504.0).to(torch.float16)
        self.tensors.append(t)
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.tensors)


def pack_layers(sd, names, n_layers, pk: Packed, split=False, f16=False):
    """names: dict role -> key template with ``{i}``.  Default: tf32-rounded fp32 GEMM weights
    (MER_GEMM_TF32 stack); split=True: bf16 hi|lo rows (MER_GEMM_BF16X3 stack); f16=True: fp16
    (MER_GEMM_F16 stack)."""
    kw = dict(f16=True) if f16 else (dict(split=True) if split else dict(tf32=True))
    arr = (MerLayerWeights * n_layers)()
    for i in range(n_layers):
        g = lambda role: sd[names[role].format(i=i)]  # noqa: E731
        wq = np.concatenate([np.asarray(g("q_w")), np.asarray(g("k_w")), np.asarray(g("v_w"))], 0)
        bq = np.concatenate([np.asarray(g("q_b")), np.asarray(g("k_b")), np.asarray(g("v_b"))], 0)
        ent = dict(
            ln1_g=pk.keep(g("ln1_g")), ln1_b=pk.keep(g("ln1_b")),
            w_qkv=pk.keep(wq, **kw), b_qkv=pk.keep(bq),
            w_o=pk.keep(g("o_w"), **kw), b_o=pk.keep(g("o_b")),
            ln2_g=pk.keep(g("ln2_g")), ln2_b=pk.keep(g("ln2_b")),
            w_fc1=pk.keep(g("fc1_w"), **kw), b_fc1=pk.keep(g("fc1_b")),
            w_fc2=pk.keep(g("fc2_w"), **kw), b_fc2=pk.keep(g("fc2_b")),
        )
        for k, t in ent.items():
            setattr(arr[i], k, t.data_ptr())
    return arr


def _np(sd):
    return {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
            for k, v in sd.items()}


VIT_NAMES = dict(
    ln1_g="encoder.layer.{i}.layernorm_before.weight", ln1_b="encoder.layer.{i}.layernorm_before.bias",
    q_w="encoder.layer.{i}.attention.attention.query.weight", q_b="encoder.layer.{i}.attention.attention.query.bias",
    k_w="encoder.layer.{i}.attention.attention.key.weight", k_b="encoder.layer.{i}.attention.attention.key.bias",
    v_w="encoder.layer.{i}.attention.attention.value.weight", v_b="encoder.layer.{i}.attention.attention.value.bias",
    o_w="encoder.layer.{i}.attention.output.dense.weight", o_b="encoder.layer.{i}.attention.output.dense.bias",
    ln2_g="encoder.layer.{i}.layernorm_after.weight", ln2_b="encoder.layer.{i}.layernorm_after.bias",
    fc1_w="encoder.layer.{i}.intermediate.dense.weight", fc1_b="encoder.layer.{i}.intermediate.dense.bias",
    fc2_w="encoder.layer.{i}.output.dense.weight", fc2_b="encoder.layer.{i}.output.dense.bias",
)

CLIP_NAMES = dict(
    ln1_g="vision_model.encoder.layers.{i}.layer_norm1.weight", ln1_b="vision_model.encoder.layers.{i}.layer_norm1.bias",
    q_w="vision_model.encoder.layers.{i}.self_attn.q_proj.weight", q_b="vision_model.encoder.layers.{i}.self_attn.q_proj.bias",
    k_w="vision_model.encoder.layers.{i}.self_attn.k_proj.weight", k_b="vision_model.encoder.layers.{i}.self_attn.k_proj.bias",
    v_w="vision_model.encoder.layers.{i}.self_attn.v_proj.weight", v_b="vision_model.encoder.layers.{i}.self_attn.v_proj.bias",
    o_w="vision_model.encoder.layers.{i}.self_attn.out_proj.weight", o_b="vision_model.encoder.layers.{i}.self_attn.out_proj.bias",
    ln2_g="vision_model.encoder.layers.{i}.layer_norm2.weight", ln2_b="vision_model.encoder.layers.{i}.layer_norm2.bias",
    fc1_w="vision_model.encoder.layers.{i}.mlp.fc1.weight", fc1_b="vision_model.encoder.layers.{i}.mlp.fc1.bias",
    fc2_w="vision_model.encoder.layers.{i}.mlp.fc2.weight", fc2_b="vision_model.encoder.layers.{i}.mlp.fc2.bias",
)

HUBERT_NAMES = dict(
    q_w="encoder.layers.{i}.attention.q_proj.weight", q_b="encoder.layers.{i}.attention.q_proj.bias",
    k_w="encoder.layers.{i}.attention.k_proj.weight", k_b="encoder.layers.{i}.attention.k_proj.bias",
    v_w="encoder.layers.{i}.attention.v_proj.weight", v_b="encoder.layers.{i}.attention.v_proj.bias",
    o_w="encoder.layers.{i}.attention.out_proj.weight", o_b="encoder.layers.{i}.attention.out_proj.bias",
    ln1_g="encoder.layers.{i}.layer_norm.weight", ln1_b="encoder.layers.{i}.layer_norm.bias",
    fc1_w="encoder.layers.{i}.feed_forward.intermediate_dense.weight", fc1_b="encoder.layers.{i}.feed_forward.intermediate_dense.bias",
    fc2_w="encoder.layers.{i}.feed_forward.output_dense.weight", fc2_b="encoder.layers.{i}.feed_forward.output_dense.bias",
    ln2_g="encoder.layers.{i}.final_layer_norm.weight", ln2_b="encoder.layers.{i}.final_layer_norm.bias",
)

BERT_NAMES = dict(
    q_w="encoder.layer.{i}.attention.self.query.weight", q_b="encoder.layer.{i}.attention.self.query.bias",
    k_w="encoder.layer.{i}.attention.self.key.weight", k_b="encoder.layer.{i}.attention.self.key.bias",
    v_w="encoder.layer.{i}.attention.self.value.weight", v_b="encoder.layer.{i}.attention.self.value.bias",
    o_w="encoder.layer.{i}.attention.output.dense.weight", o_b="encoder.layer.{i}.attention.output.dense.bias",
    ln1_g="encoder.layer.{i}.attention.output.LayerNorm.weight", ln1_b="encoder.layer.{i}.attention.output.LayerNorm.bias",
    fc1_w="encoder.layer.{i}.intermediate.dense.weight", fc1_b="encoder.layer.{i}.intermediate.dense.bias",
    fc2_w="encoder.layer.{i}.output.dense.weight", fc2_b="encoder.layer.{i}.output.dense.bias",
    ln2_g="encoder.layer.{i}.output.LayerNorm.weight", ln2_b="encoder.layer.{i}.output.LayerNorm.bias",
)


def count_layers(sd, template):
    n = 0
    while template.format(i=n) in sd:
        n += 1
    return n
