"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement, in plain torch fp32/fp64 tensor ops, of the three encoder forwards the
reference extractors reach through ``transformers.AutoModel`` (an un-vendored third-party
dependency of the reference: ``transformers==4.28.0`` pinned in MERBench/environment.yml:44; this
image ships 5.5.0).  Each function returns the ``hidden_states`` tuple the reference scripts
consume (``output_hidden_states=True``), i.e. 1 + num_layers tensors ``[B, T, 768]``.

Pinned by tests/test_oracle.py against (a) the HF classes themselves, layer by layer, on the
synthetic checkpoints of mertools_b200/synthetic.py and (b) the golden fixtures produced by the
UNMODIFIED reference scripts (tests/golden/make_golden.py).  The reference holds no test or
golden vector of its own for this path, so those two are the pin.

HF = site-packages/transformers (5.5.0).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


def _t(sd, name, dtype):
    v = sd[name]
    if not isinstance(v, torch.Tensor):
        v = torch.from_numpy(v)
    return v.to(dtype)


def _linear(x, sd, prefix, dtype):
    return F.linear(x, _t(sd, prefix + ".weight", dtype), _t(sd, prefix + ".bias", dtype))


def _ln(x, sd, prefix, eps, dtype):
    return F.layer_norm(x, (x.shape[-1],), _t(sd, prefix + ".weight", dtype),
                        _t(sd, prefix + ".bias", dtype), eps)


def _mha(q, k, v, heads, bias=None):
    """softmax(q k^T / sqrt(d) (+ bias [B, heads, T, T])) v, no mask (HF eager_attention_forward,
    modeling_vit.py:171-196; HubertAttention modeling_hubert.py:262-345; BertSelfAttention)."""
    B, T, D = q.shape
    hd = D // heads
    q = q.view(B, T, heads, hd).transpose(1, 2)
    k = k.view(B, T, heads, hd).transpose(1, 2)
    v = v.view(B, T, heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    p = torch.softmax(s if bias is None else s + bias, dim=-1)
    return (p @ v).transpose(1, 2).reshape(B, T, D)


def wavlm_relative_buckets(T, num_buckets=320, max_distance=800):
    """HF WavLMAttention._relative_positions_bucket on ``j - i`` for i, j < T (modeling_wavlm.py:243-271; the float32
    log arithmetic of the original is kept: bucket boundaries depend on it).  int64 [T, T]."""
    rel = torch.arange(T)[None, :] - torch.arange(T)[:, None]
    nb = num_buckets // 2
    buckets = (rel > 0).to(torch.long) * nb
    rel = rel.abs()
    max_exact = nb // 2
    large = torch.log(rel.float() / max_exact) / math.log(max_distance / max_exact) * (nb - max_exact)
    large = torch.min((max_exact + large).to(torch.long), torch.full_like(rel, nb - 1))
    return buckets + torch.where(rel < max_exact, rel, large)


def wavlm_gated_position_bias(sd, prefix, y, pos_bias, heads, dtype=torch.float32):
    """WavLMAttention.forward steps 1-4 (modeling_wavlm.py:165-180): per (clip, head, query) gate from the head's 64
    input features through ``gru_rel_pos_linear`` (64 -> 8 = 2 x 4, summed over the 4), sigmoid,
    ``gate_a * (gate_b * const - 1) + 2``, times the shared position bias [heads, T, T]."""
    B, T, D = y.shape
    g = y.view(B, T, heads, D // heads).permute(0, 2, 1, 3)
    proj = _linear(g, sd, prefix + "gru_rel_pos_linear", dtype).view(B, heads, T, 2, 4).sum(-1)
    ga, gb = torch.sigmoid(proj).chunk(2, dim=-1)
    gate = ga * (gb * _t(sd, prefix + "gru_rel_pos_const", dtype) - 1.0) + 2.0          # [B, heads, T, 1]
    return gate * pos_bias[None]


# ------------------------------------------------------------------------------------------------
# ViT  (HF models/vit/modeling_vit.py)
# ------------------------------------------------------------------------------------------------
def vit_hidden_states(sd, pixel_values, layers=12, heads=12, eps=1e-12, dtype=torch.float32):
    """``ViTModel(pixel_values, output_hidden_states=True).hidden_states``.

    pixel_values [N,3,224,224].  Embeddings: patch conv k=s=16 (:151,166) -> prepend CLS ->
    + position embeddings (:117-124).  Layers are PRE-LN (:328-346).  hidden_states[-1] is the
    last layer's output BEFORE ViTModel.layernorm (:455), which is what the reference reads
    (extract_vision_huggingface.py:143-144)."""
    x = pixel_values.to(dtype)
    w = _t(sd, "embeddings.patch_embeddings.projection.weight", dtype)
    b = _t(sd, "embeddings.patch_embeddings.projection.bias", dtype)
    x = F.conv2d(x, w, b, stride=w.shape[-1]).flatten(2).transpose(1, 2)  # [N,196,768]
    cls = _t(sd, "embeddings.cls_token", dtype).expand(x.shape[0], -1, -1)
    x = torch.cat([cls, x], dim=1) + _t(sd, "embeddings.position_embeddings", dtype)
    hs = [x]
    for i in range(layers):
        p = f"encoder.layer.{i}."
        h = _ln(x, sd, p + "layernorm_before", eps, dtype)
        q = _linear(h, sd, p + "attention.attention.query", dtype)
        k = _linear(h, sd, p + "attention.attention.key", dtype)
        v = _linear(h, sd, p + "attention.attention.value", dtype)
        a = _linear(_mha(q, k, v, heads), sd, p + "attention.output.dense", dtype)
        x = x + a
        h = _ln(x, sd, p + "layernorm_after", eps, dtype)
        h = F.gelu(_linear(h, sd, p + "intermediate.dense", dtype))
        x = x + _linear(h, sd, p + "output.dense", dtype)
        hs.append(x)
    return tuple(hs)


def beit_relative_position_index(window=14):
    """HF Data2VecVisionRelativePositionBias.generate_relative_position_index (modeling_data2vec_vision.py:534-556):
    index into the (2w-1)^2 + 3 table for every (query, key) pair of the 1 + w*w tokens; the three extra rows are
    cls->token, token->cls and cls->cls.  int64 [1 + w*w, 1 + w*w]."""
    n = (2 * window - 1) ** 2 + 3
    c = torch.stack(torch.meshgrid(torch.arange(window), torch.arange(window), indexing="ij")).flatten(1)
    rel = (c[:, :, None] - c[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += window - 1
    rel[:, :, 1] += window - 1
    rel[:, :, 0] *= 2 * window - 1
    idx = torch.zeros((window * window + 1,) * 2, dtype=rel.dtype)
    idx[1:, 1:] = rel.sum(-1)
    idx[0, 0:] = n - 3
    idx[0:, 0] = n - 2
    idx[0, 0] = n - 1
    return idx


def data2vec_vision_hidden_states(sd, pixel_values, heads=12, eps=1e-12, dtype=torch.float32):
    """``Data2VecVisionModel(pixel_values, output_hidden_states=True).hidden_states`` (extract_vision_huggingface.py:
    124-133 reads ``[-1].sum(dim=1)``): patch conv with bias + class token, NO absolute positions; pre-LN BEiT layers
    whose attention adds a per-layer relative position bias to the scaled scores (key projection has no bias) and
    whose two branches are scaled by ``lambda_1`` / ``lambda_2``."""
    x = pixel_values.to(dtype)
    w = _t(sd, "embeddings.patch_embeddings.projection.weight", dtype)
    window = x.shape[2] // w.shape[-1]
    x = F.conv2d(x, w, _t(sd, "embeddings.patch_embeddings.projection.bias", dtype), stride=w.shape[-1])
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([_t(sd, "embeddings.cls_token", dtype).expand(x.shape[0], -1, -1), x], dim=1)
    idx = beit_relative_position_index(window)
    hs = [x]
    i = 0
    while f"encoder.layer.{i}.output.dense.weight" in sd:
        p = f"encoder.layer.{i}."
        a = p + "attention.attention."
        y = _ln(x, sd, p + "layernorm_before", eps, dtype)
        q = _linear(y, sd, a + "query", dtype)
        k = F.linear(y, _t(sd, a + "key.weight", dtype))
        v = _linear(y, sd, a + "value", dtype)
        tname = a + "relative_position_bias.relative_position_bias_table"
        if tname not in sd:   # use_shared_relative_position_bias, or none at all
            tname = "encoder.relative_position_bias.relative_position_bias_table"
        bias = _t(sd, tname, dtype)[idx].permute(2, 0, 1)[None] if tname in sd else None
        att = _linear(_mha(q, k, v, heads, bias), sd, p + "attention.output.dense", dtype)
        x = x + att * _t(sd, p + "lambda_1", dtype)
        h = F.gelu(_linear(_ln(x, sd, p + "layernorm_after", eps, dtype), sd, p + "intermediate.dense", dtype))
        x = x + _linear(h, sd, p + "output.dense", dtype) * _t(sd, p + "lambda_2", dtype)
        hs.append(x)
        i += 1
    return tuple(hs)


def dinov2_position_embeddings(sd, grid_h, grid_w, dtype=torch.float32):
    """HF Dinov2Embeddings.interpolate_pos_encoding (modeling_dinov2.py:57-95, transformers 5.x: bicubic resize of the
    [side, side] table to ``size=(grid_h, grid_w)``, align_corners=False; 4.x releases passed a scale factor computed
    from grid + 0.1 instead, which samples at slightly different coordinates)."""
    pos = _t(sd, "embeddings.position_embeddings", torch.float32)
    n = pos.shape[1] - 1
    side = int(n ** 0.5)
    if n == grid_h * grid_w and grid_h == grid_w:
        return pos.to(dtype)
    grid = pos[:, 1:].reshape(1, side, side, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid, size=(grid_h, grid_w), mode="bicubic", align_corners=False)
    return torch.cat([pos[:, :1], grid.permute(0, 2, 3, 1).reshape(1, grid_h * grid_w, -1)], dim=1).to(dtype)


def dinov2_hidden_states(sd, pixel_values, heads=16, eps=1e-6, dtype=torch.float32):
    """``Dinov2Model(pixel_values, output_hidden_states=True).hidden_states`` (extract_vision_huggingface.py:135-145
    reads ``[-1]``: the last layer's output BEFORE ``Dinov2Model.layernorm``).  Patch conv k = s = 14 with bias,
    class token, interpolated positions; PRE-LN layers with LayerScale on both branches (modeling_dinov2.py
    Dinov2Layer): x += lambda1 * attn(norm1(x)); x += lambda2 * mlp(norm2(x)), erf GELU."""
    x = pixel_values.to(dtype)
    w = _t(sd, "embeddings.patch_embeddings.projection.weight", dtype)
    gh, gw = x.shape[2] // w.shape[-1], x.shape[3] // w.shape[-1]
    x = F.conv2d(x, w, _t(sd, "embeddings.patch_embeddings.projection.bias", dtype), stride=w.shape[-1])
    x = x.flatten(2).transpose(1, 2)
    x = torch.cat([_t(sd, "embeddings.cls_token", dtype).expand(x.shape[0], -1, -1), x], dim=1)
    x = x + dinov2_position_embeddings(sd, gh, gw, dtype)
    hs = [x]
    i = 0
    while f"encoder.layer.{i}.norm2.weight" in sd:
        p = f"encoder.layer.{i}."
        h = _ln(x, sd, p + "norm1", eps, dtype)
        q = _linear(h, sd, p + "attention.attention.query", dtype)
        k = _linear(h, sd, p + "attention.attention.key", dtype)
        v = _linear(h, sd, p + "attention.attention.value", dtype)
        a = _linear(_mha(q, k, v, heads), sd, p + "attention.output.dense", dtype)
        x = x + a * _t(sd, p + "layer_scale1.lambda1", dtype)
        h = _ln(x, sd, p + "norm2", eps, dtype)
        if p + "mlp.weights_in.weight" in sd:   # Dinov2SwiGLUFFN (dinov2-giant): silu(x1) * x2 of one fused projection
            x1, x2 = _linear(h, sd, p + "mlp.weights_in", dtype).chunk(2, dim=-1)
            h = _linear(F.silu(x1) * x2, sd, p + "mlp.weights_out", dtype)
        else:
            h = _linear(F.gelu(_linear(h, sd, p + "mlp.fc1", dtype)), sd, p + "mlp.fc2", dtype)
        x = x + h * _t(sd, p + "layer_scale2.lambda1", dtype)
        hs.append(x)
        i += 1
    return tuple(hs)


# ------------------------------------------------------------------------------------------------
# HuBERT  (HF models/hubert/modeling_hubert.py)
# ------------------------------------------------------------------------------------------------
def clip_image_features(sd, pixel_values, layers=12, heads=12, eps=1e-5, dtype=torch.float32):
    """``CLIPModel.get_image_features(pixel_values)`` (extract_vision_huggingface.py:114-122; HF
    modeling_clip.py: CLIPVisionEmbeddings, pre_layrnorm, pre-LN CLIPEncoderLayer with quick_gelu,
    post_layernorm on the class token, visual_projection).  sd: CLIPModel state_dict names
    (``vision_model.*``, ``visual_projection.weight``).  Returns (image_embeds [N, proj], hidden states)."""
    v = "vision_model."
    w = _t(sd, v + "embeddings.patch_embedding.weight", dtype)
    x = F.conv2d(pixel_values.to(dtype), w, None, stride=w.shape[-1]).flatten(2).transpose(1, 2)
    cls = _t(sd, v + "embeddings.class_embedding", dtype).expand(x.shape[0], 1, -1)
    x = torch.cat([cls, x], dim=1) + _t(sd, v + "embeddings.position_embedding.weight", dtype)[None]
    x = _ln(x, sd, v + "pre_layrnorm", eps, dtype)
    hs = [x]
    for i in range(layers):
        p = f"{v}encoder.layers.{i}."
        y = _ln(x, sd, p + "layer_norm1", eps, dtype)
        q = _linear(y, sd, p + "self_attn.q_proj", dtype)
        k = _linear(y, sd, p + "self_attn.k_proj", dtype)
        vv = _linear(y, sd, p + "self_attn.v_proj", dtype)
        x = x + _linear(_mha(q, k, vv, heads), sd, p + "self_attn.out_proj", dtype)
        h = _linear(_ln(x, sd, p + "layer_norm2", eps, dtype), sd, p + "mlp.fc1", dtype)
        h = h * torch.sigmoid(1.702 * h)  # quick_gelu
        x = x + _linear(h, sd, p + "mlp.fc2", dtype)
        hs.append(x)
    pooled = _ln(x[:, 0], sd, v + "post_layernorm", eps, dtype)
    return F.linear(pooled, _t(sd, "visual_projection.weight", dtype)), tuple(hs)


def resnet18_features(sd, x, dtype=torch.float32, eps=1e-5):
    """``nn.Sequential(*list(torchvision.models.resnet18().children())[:-1])(x).squeeze()`` in eval mode
    (extract_imagenet_embedding.py:24,47-49; torchvision resnet.py BasicBlock): conv1 7x7/2 + bn + relu,
    maxpool 3x3/2, four stages of two BasicBlocks (1x1/2 downsample in the first block of stages 2-4),
    global average pool.  x: [N, 3, 224, 224] normalised.  Returns [N, 512]."""
    def cb(x, conv, bn, stride, pad):
        y = F.conv2d(x, _t(sd, conv + ".weight", dtype), None, stride=stride, padding=pad)
        return F.batch_norm(y, _t(sd, bn + ".running_mean", dtype), _t(sd, bn + ".running_var", dtype),
                            _t(sd, bn + ".weight", dtype), _t(sd, bn + ".bias", dtype), False, 0.0, eps)
    y = F.relu(cb(x.to(dtype), "conv1", "bn1", 2, 3))
    y = F.max_pool2d(y, 3, 2, 1)
    for li in range(1, 5):
        for b in range(2):
            p = f"layer{li}.{b}."
            stride = 2 if (li > 1 and b == 0) else 1
            idt = y
            o = F.relu(cb(y, p + "conv1", p + "bn1", stride, 1))
            o = cb(o, p + "conv2", p + "bn2", 1, 1)
            if p + "downsample.0.weight" in sd:
                idt = cb(y, p + "downsample.0", p + "downsample.1", stride, 0)
            y = F.relu(o + idt)
    return y.mean(dim=(2, 3))


VGGISH_CONVS = ("conv1", "conv2", "conv3/conv3_1", "conv3/conv3_2", "conv4/conv4_1", "conv4/conv4_2")
VGGISH_FCS = ("fc1/fc1_1", "fc1/fc1_2", "fc2")
VGGISH_POOL_AFTER = ("conv1", "conv2", "conv3/conv3_2", "conv4/conv4_2")


def vggish_embeddings(sd, examples, dtype=torch.float32):
    """``define_vggish_slim`` (vggish_slim.py:37-100) restated: sd holds the TF variables under their checkpoint
    names (``vggish/<scope>/weights`` HWIO or [in, out], ``vggish/<scope>/biases``); every conv is 3x3 / stride 1
    / 'SAME' + ReLU, every pool 2x2 / stride 2 (even sizes: 'SAME' == 'VALID'), ``slim.flatten`` runs over the
    NHWC map, all three fully connected layers end in ReLU (the arg_scope default, :58-62), the result is the
    ``vggish/embedding`` tensor the reference fetches (extract_vggish_embedding.py:35-49).
    examples: [n, 96, 64] log-mel patches.  Returns [n, 128].
    PARITY UNPINNED for this function: TensorFlow / tf_slim are not installed here, so the reference graph
    cannot be executed; the restatement follows the definition file only."""
    x = examples.to(dtype)[:, None]
    for name in VGGISH_CONVS:
        w = _t(sd, f"vggish/{name}/weights", dtype).permute(3, 2, 0, 1)  # HWIO -> OIHW
        x = F.relu(F.conv2d(x, w, _t(sd, f"vggish/{name}/biases", dtype), padding=1))
        if name in VGGISH_POOL_AFTER:
            x = F.max_pool2d(x, 2)
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)
    for name in VGGISH_FCS:
        x = F.relu(x @ _t(sd, f"vggish/{name}/weights", dtype) + _t(sd, f"vggish/{name}/biases", dtype))
    return x


FERPLUS_BLOCKS = (3, 4, 6, 3)
FERPLUS_MEAN = (131.0912, 103.8827, 91.4953)   # resnet50_ferplus_dag.py:11-13 (RGB, on the 0..255 scale; std 1)


def ferplus_resnet50_features(sd, x, dtype=torch.float32, eps=1e-5):
    """What extract_ferplus_embedding.py keeps of ``resnet50_ferplus_dag`` for its default
    ``--layer_name conv5_3_3x3_relu`` (:81-115): the forward (resnet50_ferplus_dag.py:178-355) up to the ReLU
    after conv5_3_3x3 + BN, then AvgPool2d(7) and a (no-op) ReLU -> [N, 512].  Caffe-style bottlenecks: 1x1 reduce
    (stride 2 in conv3_1 / conv4_1 / conv5_1) - 3x3 - 1x1 increase, a projection shortcut in the first block of each
    stage, one BatchNorm after every conv, MaxPool2d(3, 2, padding 0, ceil_mode=True) after the stem.
    A state_dict with ``<block>_1x1_down`` / ``_1x1_up`` entries is ``senet50_ferplus_dag`` (senet50_ferplus_dag.py:
    255-470): every block's increase output y is rescaled per channel by sigmoid(up(relu(down(mean_hw(y))))) before
    the shortcut is added.
    x: [N, 3, 224, 224] = RGB pixels on the 0..255 scale minus FERPLUS_MEAN."""
    def cb(x, name, stride=1, pad=0):
        y = F.conv2d(x, _t(sd, name + ".weight", dtype), None, stride=stride, padding=pad)
        return F.batch_norm(y, _t(sd, name + "_bn.running_mean", dtype), _t(sd, name + "_bn.running_var", dtype),
                            _t(sd, name + "_bn.weight", dtype), _t(sd, name + "_bn.bias", dtype), False, 0.0, eps)
    y = F.relu(cb(x.to(dtype), "conv1_7x7_s2", 2, 3))
    y = F.max_pool2d(y, 3, 2, 0, ceil_mode=True)
    for si, nblk in enumerate(FERPLUS_BLOCKS):
        for b in range(1, nblk + 1):
            p = f"conv{si + 2}_{b}_"
            stride = 2 if (si > 0 and b == 1) else 1
            o = F.relu(cb(y, p + "1x1_reduce", stride))
            o = F.relu(cb(o, p + "3x3", 1, 1))
            if si == 3 and b == nblk:
                return F.relu(o.mean(dim=(2, 3)))
            o = cb(o, p + "1x1_increase")
            if p + "1x1_down.weight" in sd:
                z = o.mean(dim=(2, 3), keepdim=True)
                z = F.relu(F.conv2d(z, _t(sd, p + "1x1_down.weight", dtype), _t(sd, p + "1x1_down.bias", dtype)))
                o = torch.sigmoid(F.conv2d(z, _t(sd, p + "1x1_up.weight", dtype), _t(sd, p + "1x1_up.bias", dtype))) * o
            idt = cb(y, p + "1x1_proj", stride) if b == 1 else y
            y = F.relu(idt + o)
    raise AssertionError("unreachable")


def manet_embedding(sd, x, dtype=torch.float32, eps=1e-5):
    """``manet(...)(x, return_embedding=True)`` of the reference's MA-Net (feature_extraction/visual/manet/model/
    manet.py:222-270, blocks :16-153; CBAM: manet/model/attention.py:27-84) in eval mode -> [N, 1024] =
    cat(local-branch embedding, multi-scale-branch embedding).  x: [N, 3, 224, 224] in [0, 1] (ToTensor only,
    extract_manet_embedding.py:60-61)."""
    def cb(x, conv, bn, stride=1, pad=0):
        y = F.conv2d(x, _t(sd, conv + ".weight", dtype), None, stride=stride, padding=pad)
        return F.batch_norm(y, _t(sd, bn + ".running_mean", dtype), _t(sd, bn + ".running_var", dtype),
                            _t(sd, bn + ".weight", dtype), _t(sd, bn + ".bias", dtype), False, 0.0, eps)

    def shortcut(p, x, stride):
        return cb(x, p + "downsample.0", p + "downsample.1", stride) if p + "downsample.0.weight" in sd else x

    def basic(p, x, stride):                                  # BasicBlock :16-43
        o = F.relu(cb(x, p + "conv1", p + "bn1", stride, 1))
        o = cb(o, p + "conv2", p + "bn2", 1, 1)
        return F.relu(o + shortcut(p, x, stride))

    def cbam(p, y):                                           # attention.py:27-84
        def mlp(v):
            h = F.relu(F.linear(v, _t(sd, p + "ChannelGate.mlp.1.weight", dtype), _t(sd, p + "ChannelGate.mlp.1.bias", dtype)))
            return F.linear(h, _t(sd, p + "ChannelGate.mlp.3.weight", dtype), _t(sd, p + "ChannelGate.mlp.3.bias", dtype))
        att = mlp(y.mean(dim=(2, 3))) + mlp(y.amax(dim=(2, 3)))
        y = y * torch.sigmoid(att)[:, :, None, None]
        comp = torch.cat((y.amax(dim=1, keepdim=True), y.mean(dim=1, keepdim=True)), dim=1)
        s = cb(comp, p + "SpatialGate.spatial.conv", p + "SpatialGate.spatial.bn", 1, 3)
        return y * torch.sigmoid(s)

    def attention(p, x, stride):                              # AttentionBlock :121-153
        o = F.relu(cb(x, p + "conv1", p + "bn1", stride, 1))
        o = cbam(p + "cbam.", cb(o, p + "conv2", p + "bn2", 1, 1))
        return F.relu(o + shortcut(p, x, stride))

    def mulscale(p, x, stride):                               # MulScaleBlock :46-118
        t = F.relu(cb(x, p + "conv1", p + "bn1", stride, 1))
        sw = t.shape[1] // 4
        sp = torch.split(t, sw, 1)
        total = 0
        for chain in (1, 2):
            outs, prev = [], None
            for i in range(4):
                inp = sp[i] if prev is None else F.relu(prev) + sp[i]
                prev = cb(inp, p + f"conv{chain}_2_{i + 1}", p + f"bn{chain}_2_{i + 1}", 1, 1)
                outs.append(prev)
            total = total + torch.cat(outs, dim=1)
        return F.relu(total + shortcut(p, x, stride))

    y = F.relu(cb(x.to(dtype), "conv1", "bn1", 2, 3))
    y = F.max_pool2d(y, 3, 2, 1)
    for b in range(2):
        y = basic(f"layer1.{b}.", y, 1)
    for b in range(2):
        y = basic(f"layer2.{b}.", y, 2 if b == 0 else 1)
    local = []
    for pi, (y0, x0) in enumerate(((0, 0), (0, 14), (14, 0), (14, 14)), start=1):
        o = y[:, :, y0:y0 + 14, x0:x0 + 14]
        for b in range(2):
            o = attention(f"layer3_1_p{pi}.{b}.", o, 2 if b == 0 else 1)
        for b in range(2):
            o = attention(f"layer4_1_p{pi}.{b}.", o, 1)
        local.append(o)
    b1 = torch.cat([torch.cat(local[:2], dim=3), torch.cat(local[2:], dim=3)], dim=2).mean(dim=(2, 3))
    o = y
    for b in range(2):
        o = mulscale(f"layer3_2.{b}.", o, 2 if b == 0 else 1)
    for b in range(2):
        o = mulscale(f"layer4_2.{b}.", o, 2 if b == 0 else 1)
    return torch.cat([b1, o.mean(dim=(2, 3))], dim=1)


def emonet_embedding(sd, x, dtype=torch.float32, eps=1e-5):
    """``EmoNet()(x, return_embedding=True)`` of the reference (feature_extraction/visual/emonet/models/emonet.py:
    173-222, blocks :20-112; attention=True, temporal_smoothing=False) in eval mode: the 256-d vector after the
    emotion tower's 4 x 4 average pool.  Every norm is a BatchNorm2d (the file rebinds nn.InstanceNorm2d).
    x: [N, 3, 256, 256] in [0, 1]."""
    def bn(x, name):
        return F.batch_norm(x, _t(sd, name + ".running_mean", dtype), _t(sd, name + ".running_var", dtype),
                            _t(sd, name + ".weight", dtype), _t(sd, name + ".bias", dtype), False, 0.0, eps)

    def conv(x, name, stride=1, pad=0):
        b = _t(sd, name + ".bias", dtype) if name + ".bias" in sd else None
        return F.conv2d(x, _t(sd, name + ".weight", dtype), b, stride=stride, padding=pad)

    def block(p, x):                                          # ConvBlock :20-64 (pre-activation, concatenated outputs)
        o1 = conv(F.relu(bn(x, p + "bn1")), p + "conv1", 1, 1)
        o2 = conv(F.relu(bn(o1, p + "bn2")), p + "conv2", 1, 1)
        o3 = conv(F.relu(bn(o2, p + "bn3")), p + "conv3", 1, 1)
        res = conv(F.relu(bn(x, p + "downsample.0")), p + "downsample.2") if p + "downsample.2.weight" in sd else x
        return torch.cat((o1, o2, o3), 1) + res

    def hourglass(p, level, inp):                             # HourGlass._forward :87-109
        up1 = block(p + f"b1_{level}.", inp)
        low = block(p + f"b2_{level}.", F.max_pool2d(inp, 2, stride=2))
        low = hourglass(p, level - 1, low) if level > 1 else block(p + f"b2_plus_{level}.", low)
        low = block(p + f"b3_{level}.", low)
        return up1 + F.interpolate(low, scale_factor=2, mode="nearest")

    x = F.relu(bn(conv(x.to(dtype), "conv1", 2, 3), "bn1"))
    x = F.max_pool2d(block("conv2.", x), 2, stride=2)
    x = block("conv4.", block("conv3.", x))
    previous, feats, heat = x, [], None
    for i in range(2):
        ll = block(f"top_m_{i}.", hourglass(f"m{i}.", 4, previous))
        ll = F.relu(bn(conv(ll, f"conv_last{i}"), f"bn_end{i}"))
        heat = conv(ll, f"l{i}")
        if i < 1:
            ll = conv(ll, f"bl{i}")
            previous = previous + ll + conv(heat, f"al{i}")
        feats.append(ll)
    hg = torch.cat(feats, dim=1) * heat.sum(dim=1, keepdim=True)
    y = conv(torch.cat((x, hg), dim=1), "conv1x1_input_emo_2")
    for i in range(4):
        y = F.max_pool2d(block(f"emo_net_2.{2 * i}.", y), 2, 2)
    return F.avg_pool2d(y, 4).flatten(1)


def whisper_last_hidden_state(sd, input_features, decoder_input_ids, heads=8, dtype=torch.float32, eps=1e-5):
    """``WhisperModel(input_features, decoder_input_ids=ids).last_hidden_state`` (the reference's Whisper branch,
    extract_audio_huggingface.py:83-91) in eval mode -> [B, len(ids), d_model].  Encoder (HF modeling_whisper.py
    WhisperEncoder): GELU(conv1 k3 p1), GELU(conv2 k3 s2 p1), + embed_positions, pre-LN layers (k_proj has no bias),
    layer_norm.  Decoder: embed_tokens + embed_positions, pre-LN layers of causal self-attention, cross-attention over
    the encoder output and an FFN, layer_norm.  input_features: [B, 80, 3000]."""
    def mha(xq, xkv, p, causal):
        B, Tq, D = xq.shape
        hd = D // heads
        q = _linear(xq, sd, p + "q_proj", dtype).view(B, Tq, heads, hd).transpose(1, 2)
        k = F.linear(xkv, _t(sd, p + "k_proj.weight", dtype)).view(B, -1, heads, hd).transpose(1, 2)
        v = _linear(xkv, sd, p + "v_proj", dtype).view(B, -1, heads, hd).transpose(1, 2)
        s = q @ k.transpose(-1, -2) / hd ** 0.5
        if causal:
            s = s + torch.full((Tq, k.shape[2]), float("-inf")).triu(1)
        o = (torch.softmax(s, dim=-1) @ v).transpose(1, 2).reshape(B, Tq, D)
        return _linear(o, sd, p + "out_proj", dtype)

    def ffn(x, p):
        return _linear(F.gelu(_linear(x, sd, p + "fc1", dtype)), sd, p + "fc2", dtype)
    x = input_features.to(dtype)
    x = F.gelu(F.conv1d(x, _t(sd, "encoder.conv1.weight", dtype), _t(sd, "encoder.conv1.bias", dtype), padding=1))
    x = F.gelu(F.conv1d(x, _t(sd, "encoder.conv2.weight", dtype), _t(sd, "encoder.conv2.bias", dtype), stride=2, padding=1))
    x = x.transpose(1, 2) + _t(sd, "encoder.embed_positions.weight", dtype)
    i = 0
    while f"encoder.layers.{i}.fc1.weight" in sd:
        p = f"encoder.layers.{i}."
        y = _ln(x, sd, p + "self_attn_layer_norm", eps, dtype)
        x = x + mha(y, y, p + "self_attn.", False)
        x = x + ffn(_ln(x, sd, p + "final_layer_norm", eps, dtype), p)
        i += 1
    enc = _ln(x, sd, "encoder.layer_norm", eps, dtype)
    ids = torch.as_tensor(decoder_input_ids, dtype=torch.long)
    y = _t(sd, "decoder.embed_tokens.weight", dtype)[ids] + _t(sd, "decoder.embed_positions.weight", dtype)[:ids.shape[1]]
    i = 0
    while f"decoder.layers.{i}.fc1.weight" in sd:
        p = f"decoder.layers.{i}."
        z = _ln(y, sd, p + "self_attn_layer_norm", eps, dtype)
        y = y + mha(z, z, p + "self_attn.", True)
        y = y + mha(_ln(y, sd, p + "encoder_attn_layer_norm", eps, dtype), enc, p + "encoder_attn.", False)
        y = y + ffn(_ln(y, sd, p + "final_layer_norm", eps, dtype), p)
        i += 1
    return _ln(y, sd, "decoder.layer_norm", eps, dtype)


def videomae_sinusoid_table(n_position, d):
    """``get_sinusoid_encoding_table`` of HF modeling_videomae.py: angle[p, j] = p / 10000^(2 (j // 2) / d), sine on
    the even columns, cosine on the odd ones (float64 math, float32 result)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d)
    tab = pos / np.power(10000.0, 2 * (j // 2) / d)[None, :]
    tab[:, 0::2] = np.sin(tab[:, 0::2])
    tab[:, 1::2] = np.cos(tab[:, 1::2])
    return torch.from_numpy(tab.astype(np.float32))


def videomae_last_hidden_state(sd, pixel_values, heads=12, eps=1e-12, dtype=torch.float32):
    """``VideoMAEModel(pixel_values).last_hidden_state`` with ``use_mean_pooling=True`` (no final LayerNorm: the
    released videomae-base / -large encoders; extract_vision_huggingface.py:147-159) in eval mode.
    pixel_values: [B, 16, 3, 224, 224] -> [B, 1568, hidden]; token = (tubelet, patch row, patch column)."""
    x = pixel_values.to(dtype).permute(0, 2, 1, 3, 4)                       # [B, C, T, H, W]
    w = _t(sd, "embeddings.patch_embeddings.projection.weight", dtype)
    x = F.conv3d(x, w, _t(sd, "embeddings.patch_embeddings.projection.bias", dtype), stride=w.shape[2:])
    x = x.flatten(2).transpose(1, 2)
    x = x + videomae_sinusoid_table(x.shape[1], x.shape[2]).to(dtype)
    i = 0
    while f"encoder.layer.{i}.output.dense.weight" in sd:
        p = f"encoder.layer.{i}."
        y = _ln(x, sd, p + "layernorm_before", eps, dtype)
        q = F.linear(y, _t(sd, p + "attention.attention.query.weight", dtype), _t(sd, p + "attention.attention.q_bias", dtype))
        k = F.linear(y, _t(sd, p + "attention.attention.key.weight", dtype))
        v = F.linear(y, _t(sd, p + "attention.attention.value.weight", dtype), _t(sd, p + "attention.attention.v_bias", dtype))
        x = x + _linear(_mha(q, k, v, heads), sd, p + "attention.output.dense", dtype)
        h = F.gelu(_linear(_ln(x, sd, p + "layernorm_after", eps, dtype), sd, p + "intermediate.dense", dtype))
        x = x + _linear(h, sd, p + "output.dense", dtype)
        i += 1
    # use_mean_pooling=False checkpoints (the self-supervised videomae-base / -large the reference lists) end with
    # VideoMAEModel.layernorm; the fine-tuned ones (use_mean_pooling=True) have no such parameter
    return _ln(x, sd, "layernorm", eps, dtype) if "layernorm.weight" in sd else x


def hubert_pos_conv_weight(sd, dtype=torch.float32):
    """Effective weight of the weight-normed positional conv (:45-92): W = g * v / ||v||, the
    norm taken over dims (0,1) per kernel tap (weight_norm dim=2).  Older checkpoints name the
    parameters weight_g / weight_v."""
    pre = "encoder.pos_conv_embed.conv."
    if pre + "parametrizations.weight.original0" in sd:
        g = _t(sd, pre + "parametrizations.weight.original0", torch.float64)
        v = _t(sd, pre + "parametrizations.weight.original1", torch.float64)
    elif pre + "weight_g" in sd:
        g = _t(sd, pre + "weight_g", torch.float64)
        v = _t(sd, pre + "weight_v", torch.float64)
    else:
        return _t(sd, pre + "weight", dtype)
    norm = v.pow(2).sum(dim=(0, 1), keepdim=True).sqrt()
    return (g * v / norm).to(dtype)


def hubert_hidden_states(sd, input_values, layers=12, heads=12, eps=1e-5,
                         conv_stride=(5, 2, 2, 2, 2, 2, 2), pos_groups=16, dtype=torch.float32,
                         stable_layer_norm=None):
    """``HubertModel(input_values, output_hidden_states=True).hidden_states`` in eval mode.

    input_values [B, L] (already zero-mean/unit-variance).  Feature encoder (:178-213): conv0 +
    GroupNorm(512 groups) + GELU (:154-175), conv1..6 + GELU (:106-124), no biases -- or, for the
    ``feat_extract_norm="layer"`` family (hubert-large; recognised by a LayerNorm on conv layer 1),
    every conv (+ bias) followed by LayerNorm over channels and GELU (HubertLayerNormConvLayer).  Feature
    projection (:216-231): LayerNorm(512) -> Linear.  Positional conv (:45-103): grouped conv
    k=128 pad=64, drop last frame, GELU; add.  Then either encoder.layer_norm (:440-442) and POST-LN
    layers (:372-405), or -- ``do_stable_layer_norm`` (hubert-large; default: same as the layer-norm
    feature extractor) -- PRE-LN layers (HubertEncoderLayerStableLayerNorm) with hidden states taken
    before each layer and encoder.layer_norm applied after the last one (HubertEncoderStableLayerNorm)."""
    x = input_values.to(dtype)[:, None, :]
    n_conv = len(conv_stride)
    layer_norm_convs = "feature_extractor.conv_layers.1.layer_norm.weight" in sd
    data2vec = "encoder.pos_conv_embed.layers.0.conv.weight" in sd   # Data2VecAudioModel: see below
    if stable_layer_norm is None:
        stable_layer_norm = layer_norm_convs and not data2vec
    for i in range(n_conv):
        w = _t(sd, f"feature_extractor.conv_layers.{i}.conv.weight", dtype)
        bname = f"feature_extractor.conv_layers.{i}.conv.bias"
        x = F.conv1d(x, w, _t(sd, bname, dtype) if bname in sd else None, stride=conv_stride[i])
        if layer_norm_convs:
            x = _ln(x.transpose(1, 2), sd, f"feature_extractor.conv_layers.{i}.layer_norm", 1e-5, dtype).transpose(1, 2)
        elif i == 0:
            x = F.group_norm(x, w.shape[0],
                             _t(sd, "feature_extractor.conv_layers.0.layer_norm.weight", dtype),
                             _t(sd, "feature_extractor.conv_layers.0.layer_norm.bias", dtype), 1e-5)
        x = F.gelu(x)
    x = x.transpose(1, 2)  # [B,T,512]
    x = _ln(x, sd, "feature_projection.layer_norm", eps, dtype)
    x = _linear(x, sd, "feature_projection.projection", dtype)
    if data2vec:
        # Data2VecAudioPositionalConvEmbedding (HF models/data2vec/modeling_data2vec_audio.py): a chain of grouped
        # convs (k = 19, pad 9) each followed by LayerNorm(elementwise_affine=False) and GELU; the chain's output
        # is added to its input once
        pos, l = x.transpose(1, 2), 0
        while f"encoder.pos_conv_embed.layers.{l}.conv.weight" in sd:
            w = _t(sd, f"encoder.pos_conv_embed.layers.{l}.conv.weight", dtype)
            pos = F.conv1d(pos, w, _t(sd, f"encoder.pos_conv_embed.layers.{l}.conv.bias", dtype),
                           padding=w.shape[-1] // 2, groups=pos_groups)
            if w.shape[-1] % 2 == 0:
                pos = pos[:, :, :-1]
            pos = F.gelu(F.layer_norm(pos.transpose(1, 2), (pos.shape[1],), None, None, 1e-5)).transpose(1, 2)
            l += 1
        x = x + pos.transpose(1, 2)
    else:
        wpos = hubert_pos_conv_weight(sd, dtype)
        pos = F.conv1d(x.transpose(1, 2), wpos, _t(sd, "encoder.pos_conv_embed.conv.bias", dtype),
                       padding=wpos.shape[-1] // 2, groups=pos_groups)
        if wpos.shape[-1] % 2 == 0:
            pos = pos[:, :, :-1]
        x = x + F.gelu(pos).transpose(1, 2)
    hs = []
    if not stable_layer_norm:
        x = _ln(x, sd, "encoder.layer_norm", eps, dtype)
        hs.append(x)
    wavlm, pos_bias = "encoder.layers.0.attention.gru_rel_pos_linear.weight" in sd, None
    for i in range(layers):
        p = f"encoder.layers.{i}."
        if stable_layer_norm:
            hs.append(x)
            y = _ln(x, sd, p + "layer_norm", eps, dtype)
        else:
            y = x
        q = _linear(y, sd, p + "attention.q_proj", dtype)
        k = _linear(y, sd, p + "attention.k_proj", dtype)
        v = _linear(y, sd, p + "attention.v_proj", dtype)
        bias = None
        if wavlm:   # WavLMModel: layer 0's bucket embedding gives the position bias of every layer, gated per layer
            if pos_bias is None:
                emb = _t(sd, "encoder.layers.0.attention.rel_attn_embed.weight", dtype)           # [buckets, heads]
                pos_bias = emb[wavlm_relative_buckets(y.shape[1], emb.shape[0])].permute(2, 0, 1)  # [heads, T, T]
            bias = wavlm_gated_position_bias(sd, p + "attention.", y, pos_bias, heads, dtype)
        a = _linear(_mha(q, k, v, heads, bias), sd, p + "attention.out_proj", dtype)
        if stable_layer_norm:
            x = x + a
            h = F.gelu(_linear(_ln(x, sd, p + "final_layer_norm", eps, dtype), sd,
                               p + "feed_forward.intermediate_dense", dtype))
            x = x + _linear(h, sd, p + "feed_forward.output_dense", dtype)
        else:
            x = _ln(x + a, sd, p + "layer_norm", eps, dtype)
            h = F.gelu(_linear(x, sd, p + "feed_forward.intermediate_dense", dtype))
            h = _linear(h, sd, p + "feed_forward.output_dense", dtype)
            x = _ln(x + h, sd, p + "final_layer_norm", eps, dtype)
            hs.append(x)
    if stable_layer_norm:
        hs.append(_ln(x, sd, "encoder.layer_norm", eps, dtype))
    return tuple(hs)


def hubert_num_frames(n_samples, conv_kernel=(10, 3, 3, 3, 3, 2, 2), conv_stride=(5, 2, 2, 2, 2, 2, 2)):
    t = n_samples
    for k, s in zip(conv_kernel, conv_stride):
        t = (t - k) // s + 1
    return t


# ------------------------------------------------------------------------------------------------
# BERT / RoBERTa  (HF models/bert/modeling_bert.py, models/roberta/modeling_roberta.py)
# ------------------------------------------------------------------------------------------------
def bert_hidden_states(sd, input_ids, layers=12, heads=12, eps=1e-12, position_offset=0,
                       dtype=torch.float32):
    """``BertModel(input_ids, output_hidden_states=True).hidden_states`` (token_type_ids = 0, no
    padding: the reference tokenises one sentence at a time, extract_text_huggingface.py:222).
    RoBERTa: same graph with position ids starting at pad_token_id + 1 = 2
    (modeling_roberta.py:157-159) -> ``position_offset=2``."""
    ids = torch.as_tensor(input_ids, dtype=torch.long)
    if ids.dim() == 1:
        ids = ids[None]
    T = ids.shape[1]
    pos = torch.arange(T) + position_offset
    x = (_t(sd, "embeddings.word_embeddings.weight", dtype)[ids]
         + _t(sd, "embeddings.token_type_embeddings.weight", dtype)[0]
         + _t(sd, "embeddings.position_embeddings.weight", dtype)[pos])
    x = _ln(x, sd, "embeddings.LayerNorm", eps, dtype)
    hs = [x]
    for i in range(layers):
        p = f"encoder.layer.{i}."
        q = _linear(x, sd, p + "attention.self.query", dtype)
        k = _linear(x, sd, p + "attention.self.key", dtype)
        v = _linear(x, sd, p + "attention.self.value", dtype)
        a = _linear(_mha(q, k, v, heads), sd, p + "attention.output.dense", dtype)
        x = _ln(x + a, sd, p + "attention.output.LayerNorm", eps, dtype)
        h = F.gelu(_linear(x, sd, p + "intermediate.dense", dtype))
        h = _linear(h, sd, p + "output.dense", dtype)
        x = _ln(x + h, sd, p + "output.LayerNorm", eps, dtype)
        hs.append(x)
    return tuple(hs)
