"""Attention-fusion training on the device — mirror of MERBench ``toolkit/models`` (get_models,
Attention), ``toolkit/utils/loss.py`` and the step of ``main-release.py:train_or_eval_model``.

``FusionNet`` owns flat fp32 parameter / gradient / Adam-moment buffers (reference state_dict order)
and drives libmer_b200.so: eval forward, or one fused training step = forward + CELoss + MSELoss +
backward (+ one NCCL all-reduce of the flat gradient under data parallelism) + Adam, captured in a
CUDA graph.  ``get_models(args)`` / ``train_or_eval_model(...)`` keep the reference's names and
argument meaning.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

ENC = ("audio_encoder", "text_encoder", "video_encoder", "attention_mlp")


class MerFusionDims(C.Structure):
    _fields_ = [("audio_dim", C.c_int), ("text_dim", C.c_int), ("video_dim", C.c_int),
                ("hidden", C.c_int), ("out1", C.c_int), ("out2", C.c_int)]


LSTM_PARAMS = ("rnn.weight_ih_l0", "rnn.weight_hh_l0", "rnn.bias_ih_l0", "rnn.bias_hh_l0",
               "linear_1.weight", "linear_1.bias")


def param_names(feat_type="utt"):
    names = []
    for e in ENC:
        if feat_type != "utt" and e != "attention_mlp":  # LSTMEncoder (encoder.py:45-72)
            names += [f"{e}.{n}" for n in LSTM_PARAMS]
            continue
        for l in ("linear_1", "linear_2", "linear_3"):
            names += [f"{e}.{l}.weight", f"{e}.{l}.bias"]
    for l in ("fc_att", "fc_out_1", "fc_out_2"):
        names += [f"{l}.weight", f"{l}.bias"]
    return names


def param_shapes(audio_dim, text_dim, video_dim, hidden, out1, out2, feat_type="utt"):
    ins = dict(audio_encoder=audio_dim, text_encoder=text_dim, video_encoder=video_dim,
               attention_mlp=3 * hidden)
    shapes = {}
    for e in ENC:
        if feat_type != "utt" and e != "attention_mlp":
            shapes[f"{e}.rnn.weight_ih_l0"] = (4 * hidden, ins[e])
            shapes[f"{e}.rnn.weight_hh_l0"] = (4 * hidden, hidden)
            shapes[f"{e}.rnn.bias_ih_l0"] = (4 * hidden,)
            shapes[f"{e}.rnn.bias_hh_l0"] = (4 * hidden,)
            shapes[f"{e}.linear_1.weight"] = (hidden, hidden)
            shapes[f"{e}.linear_1.bias"] = (hidden,)
            continue
        shapes[f"{e}.linear_1.weight"] = (hidden, ins[e])
        shapes[f"{e}.linear_1.bias"] = (hidden,)
        for l in ("linear_2", "linear_3"):
            shapes[f"{e}.{l}.weight"] = (hidden, hidden)
            shapes[f"{e}.{l}.bias"] = (hidden,)
    for l, o in (("fc_att", 3), ("fc_out_1", out1), ("fc_out_2", out2)):
        shapes[f"{l}.weight"] = (o, hidden)
        shapes[f"{l}.bias"] = (o,)
    return shapes


class FusionNet:
    """Device-resident Attention fusion model.  feat_type 'utt': MLP encoders on [B, D] features;
    'frm_align' / 'frm_unalign': LSTM encoders on [B, T, D] sequences (attention.py:25-33)."""

    def __init__(self, audio_dim=768, text_dim=768, video_dim=768, hidden_dim=128, output_dim1=6,
                 output_dim2=1, dropout=0.0, grad_clip=-1.0, device="cuda", max_batch=4096, seed=0,
                 feat_type="utt"):
        L.check(L.lib().mer_check_device())
        assert feat_type in ("utt", "frm_align", "frm_unalign"), feat_type
        self.feat_type = feat_type
        self.frm = feat_type != "utt"
        self.device = torch.device(device)
        self.dims = MerFusionDims(audio_dim, text_dim, video_dim, hidden_dim, output_dim1, output_dim2)
        self.dropout, self.grad_clip, self.seed = float(dropout), float(grad_clip), int(seed)
        lib = L.lib()
        lib.mer_fusion_param_count.restype = C.c_longlong
        lib.mer_fusion_param_count.argtypes = [C.POINTER(MerFusionDims)]
        lib.mer_fusion_workspace_bytes.restype = C.c_longlong
        lib.mer_fusion_workspace_bytes.argtypes = [C.POINTER(MerFusionDims), C.c_int]
        vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong
        self._fwd = L.declare("mer_fusion_forward", [C.POINTER(MerFusionDims), vp, vp, vp, vp, i32, vp,
                                                     i64, vp, vp, vp, vp])
        self._fb = L.declare("mer_fusion_fwd_bwd", [C.POINTER(MerFusionDims), vp, vp, vp, vp, vp, vp, vp,
                                                    i32, f32, f32, C.c_ulonglong, vp, vp, vp, i64, vp,
                                                    vp, vp, vp, vp])
        self._adam = L.declare("mer_fusion_adam", [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32,
                                                   f32, vp, vp])
        if self.frm:
            lib.mer_fusion_frm_param_count.restype = C.c_longlong
            lib.mer_fusion_frm_param_count.argtypes = [C.POINTER(MerFusionDims)]
            lib.mer_fusion_frm_workspace_bytes.restype = C.c_longlong
            lib.mer_fusion_frm_workspace_bytes.argtypes = [C.POINTER(MerFusionDims)] + [i32] * 4
            self._fwd_frm = L.declare("mer_fusion_frm_forward", [C.POINTER(MerFusionDims), vp, vp, vp, vp, i32,
                                                                 i32, i32, i32, vp, i64, vp, vp, vp, vp])
            self._fb_frm = L.declare("mer_fusion_frm_fwd_bwd", [C.POINTER(MerFusionDims), vp, vp, vp, vp, vp,
                                                                i32, i32, i32, vp, vp, i32, f32, f32,
                                                                C.c_ulonglong, vp, vp, vp, i64, vp, vp, vp, vp,
                                                                vp])
        self.n_params = int((lib.mer_fusion_frm_param_count if self.frm else lib.mer_fusion_param_count)(
            C.byref(self.dims)))
        self.shapes = param_shapes(audio_dim, text_dim, video_dim, hidden_dim, output_dim1, output_dim2,
                                   feat_type)
        assert sum(int(np.prod(s)) for s in self.shapes.values()) == self.n_params
        z = lambda: torch.zeros(self.n_params, dtype=torch.float32, device=self.device)  # noqa: E731
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.loss = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.max_batch = max_batch
        self.ws = torch.empty(int(lib.mer_fusion_workspace_bytes(C.byref(self.dims), max_batch)) if not self.frm
                              else 0, dtype=torch.uint8, device=self.device)
        self.training = True
        self.graph_launches = 0  # kernels launched through CUDA-graph replays (not seen by the library counter)
        lib.mer_launch_count.restype = C.c_longlong
        self._graphs = {}
        self._static = None

    # ---- nn.Module-flavoured surface used by the reference loop --------------------------------
    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def cuda(self):
        return self

    def parameters(self):
        return [self.params]

    def named_views(self, flat=None):
        flat = self.params if flat is None else flat
        out, o = {}, 0
        for n in param_names(self.feat_type):
            k = int(np.prod(self.shapes[n]))
            out[n] = flat[o:o + k].view(self.shapes[n])
            o += k
        return out

    def state_dict(self):
        return {k: v.clone() for k, v in self.named_views().items()}

    def load_state_dict(self, sd):
        """Accepts the reference's names with or without the ``model.`` wrapper prefix
        (toolkit/models/__init__.py wraps the net as ``.model``)."""
        views = self.named_views()
        for n, dst in views.items():
            src = sd[n] if n in sd else sd["model." + n]
            if isinstance(src, np.ndarray):
                src = torch.from_numpy(src)
            dst.copy_(src.to(self.device, torch.float32))
        self._invalidate()
        return self

    def _invalidate(self):
        self._graphs.clear()

    # ---- forward / step --------------------------------------------------------------------------
    def _bufs(self, B):
        d = self.dims
        f = lambda n: torch.empty(B, n, dtype=torch.float32, device=self.device)  # noqa: E731
        return f(d.hidden), f(d.out1), f(d.out2)

    def forward(self, batch):
        """batch: dict with 'audios','texts','videos' fp32 CUDA [B,D].  Eval-mode forward; returns
        (features, emos_out, vals_out, interloss) like Attention.forward (attention.py:36-57)."""
        a, t, v = (batch[k].contiguous() for k in ("audios", "texts", "videos"))
        B = a.shape[0]
        assert B <= self.max_batch
        feats, emos, vals = self._bufs(B)
        if self.frm:
            ws = self._frm_ws(B, a, t, v)
            L.check(self._fwd_frm(C.byref(self.dims), L.ptr(self.params), L.ptr(a), L.ptr(t), L.ptr(v),
                                  a.shape[1], t.shape[1], v.shape[1], B, L.ptr(ws), ws.numel(), L.ptr(feats),
                                  L.ptr(emos), L.ptr(vals), L.stream_ptr()))
            return feats, emos, vals, torch.zeros((), dtype=torch.int64, device=self.device)
        L.check(self._fwd(C.byref(self.dims), L.ptr(self.params), L.ptr(a), L.ptr(t), L.ptr(v), B,
                          L.ptr(self.ws), self.ws.numel(), L.ptr(feats), L.ptr(emos), L.ptr(vals),
                          L.stream_ptr()))
        return feats, emos, vals, torch.zeros((), dtype=torch.int64, device=self.device)

    __call__ = forward

    def _frm_ws(self, B, a, t, v):
        """Workspace of the frame-level variant: depends on the three padded sequence lengths."""
        assert a.dim() == t.dim() == v.dim() == 3, "frame-level features are [B, T, D] (read_data.py:118-125)"
        need = int(L.lib().mer_fusion_frm_workspace_bytes(C.byref(self.dims), B, a.shape[1], t.shape[1],
                                                          v.shape[1]))
        if self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
            self._invalidate()
        return self.ws

    def _launch_step(self, a, t, v, emo, val, feats, emos_out, vals_out, lr, betas, eps, wd,
                     world, ext_masks):
        B = a.shape[0]
        masks = None
        if ext_masks is not None:
            arr = (C.c_void_p * 4)(*[m.data_ptr() if m is not None else None for m in ext_masks])
            masks = C.cast(arr, C.c_void_p)
        if self.frm:
            ws = self._frm_ws(B, a, t, v)
            L.check(self._fb_frm(C.byref(self.dims), L.ptr(self.params), L.ptr(self.grads), L.ptr(a), L.ptr(t),
                                 L.ptr(v), a.shape[1], t.shape[1], v.shape[1], L.ptr(emo), L.ptr(val), B,
                                 1.0 / (B * world), self.dropout, self.seed, L.ptr(self.step_counter), masks,
                                 L.ptr(ws), ws.numel(), L.ptr(self.loss), L.ptr(feats), L.ptr(emos_out),
                                 L.ptr(vals_out), L.stream_ptr()))
        else:
            L.check(self._fb(C.byref(self.dims), L.ptr(self.params), L.ptr(self.grads), L.ptr(a), L.ptr(t),
                             L.ptr(v), L.ptr(emo), L.ptr(val), B, 1.0 / (B * world), self.dropout,
                             self.seed, L.ptr(self.step_counter), masks, L.ptr(self.ws), self.ws.numel(),
                             L.ptr(self.loss), L.ptr(feats), L.ptr(emos_out), L.ptr(vals_out), L.stream_ptr()))
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads)  # SUM: loss already carries 1/global_batch
        L.check(self._adam(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.exp_avg),
                           L.ptr(self.exp_avg_sq), self.n_params, lr, betas[0], betas[1], eps, wd, 1.0,
                           self.grad_clip if self.grad_clip != -1 else 0.0, L.ptr(self.step_counter),
                           L.stream_ptr()))

    def train_step(self, a, t, v, emo, val, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                   world_size=1, ext_masks=None, use_graph=True):
        """One optimisation step on a device batch.  Returns (loss[3] device tensor, emos_out,
        vals_out).  With use_graph the launch sequence (incl. the NCCL all-reduce) is captured once
        per batch size and replayed; inputs are copied into static buffers first."""
        B = a.shape[0]
        assert B <= self.max_batch and emo.dtype == torch.int64 and val.dtype == torch.float32
        # under data parallelism the step is launched eagerly: the NCCL all-reduce sits between the
        # backward and the Adam kernels and is not captured (30 launches per step either way)
        if not use_graph or ext_masks is not None or world_size > 1:
            feats, emos_out, vals_out = self._bufs(B)
            self._launch_step(a.contiguous(), t.contiguous(), v.contiguous(), emo.contiguous(),
                              val.contiguous(), feats, emos_out, vals_out, lr, betas, eps,
                              weight_decay, world_size, ext_masks)
            return self.loss, emos_out, vals_out
        if self.frm:
            self._frm_ws(B, a, t, v)  # sized (and graphs invalidated on growth) before any capture
        key = (B, tuple(a.shape[1:]), tuple(t.shape[1:]), tuple(v.shape[1:]), lr, betas, eps, weight_decay,
               world_size, self.dropout, self.grad_clip)
        if key not in self._graphs:
            st = dict(a=torch.empty_like(a), t=torch.empty_like(t), v=torch.empty_like(v),
                      emo=torch.empty_like(emo), val=torch.empty_like(val))
            st["feats"], st["emos"], st["vals"] = self._bufs(B)
            for k, src in (("a", a), ("t", t), ("v", v), ("emo", emo), ("val", val)):
                st[k].copy_(src)
            # warm-up outside capture on a side stream, with state restored afterwards
            saved = [x.clone() for x in (self.params, self.exp_avg, self.exp_avg_sq, self.step_counter)]
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._launch_step(st["a"], st["t"], st["v"], st["emo"], st["val"], st["feats"],
                                  st["emos"], st["vals"], lr, betas, eps, weight_decay, world_size, None)
            torch.cuda.current_stream().wait_stream(s)
            for dst, src in zip((self.params, self.exp_avg, self.exp_avg_sq, self.step_counter), saved):
                dst.copy_(src)
            g = torch.cuda.CUDAGraph()
            n0 = L.lib().mer_launch_count()
            with torch.cuda.graph(g):
                self._launch_step(st["a"], st["t"], st["v"], st["emo"], st["val"], st["feats"],
                                  st["emos"], st["vals"], lr, betas, eps, weight_decay, world_size, None)
            self._graphs[key] = (g, st, int(L.lib().mer_launch_count() - n0))
        g, st, n_kernels = self._graphs[key]
        for k, src in (("a", a), ("t", t), ("v", v), ("emo", emo), ("val", val)):
            st[k].copy_(src, non_blocking=True)
        g.replay()
        self.graph_launches += n_kernels
        return self.loss, st["emos"], st["vals"]


class MerFusionTopnDims(C.Structure):
    _fields_ = [("n_feats", C.c_int), ("feat_dims", C.c_int * 18), ("hidden", C.c_int), ("out1", C.c_int),
                ("out2", C.c_int)]


class TopnFusionNet:
    """Attention_TOPN (MER2026_Track1/toolkit/models/attention_topn.py): the utterance-level fusion net over
    N <= 18 features ``batch['feat0'] .. batch['feat{N-1}']`` ([B, feat_dims[i]] each)."""

    def __init__(self, feat_dims, hidden_dim=128, output_dim1=6, output_dim2=1, dropout=0.0, grad_clip=-1.0,
                 device="cuda", seed=0):
        L.check(L.lib().mer_check_device())
        assert 1 <= len(feat_dims) <= 18
        self.device = torch.device(device)
        self.feat_dims = [int(d) for d in feat_dims]
        self.dims = MerFusionTopnDims(len(feat_dims), (C.c_int * 18)(*self.feat_dims), hidden_dim, output_dim1,
                                      output_dim2)
        self.dropout, self.grad_clip, self.seed = float(dropout), float(grad_clip), int(seed)
        lib = L.lib()
        lib.mer_fusion_topn_param_count.restype = C.c_longlong
        lib.mer_fusion_topn_param_count.argtypes = [C.POINTER(MerFusionTopnDims)]
        lib.mer_fusion_topn_workspace_bytes.restype = C.c_longlong
        lib.mer_fusion_topn_workspace_bytes.argtypes = [C.POINTER(MerFusionTopnDims), C.c_int]
        vp, i32, f32, i64 = C.c_void_p, C.c_int, C.c_float, C.c_longlong
        self._step = L.declare("mer_fusion_topn_step", [C.POINTER(MerFusionTopnDims), vp, vp, vp, vp, vp, i32, f32, f32,
                                                        C.c_ulonglong, vp, vp, vp, i64, vp, vp, vp, vp, vp])
        self._adam = L.declare("mer_fusion_adam", [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, f32, f32, vp, vp])
        self.n_params = int(lib.mer_fusion_topn_param_count(C.byref(self.dims)))
        self.names, self.shapes = [], {}
        H = hidden_dim
        for i, d in enumerate(self.feat_dims + [H * len(self.feat_dims)]):
            e = f"encoder{i}" if i < len(self.feat_dims) else "attention_mlp"
            for l, k in (("linear_1", d), ("linear_2", H), ("linear_3", H)):
                self.names += [f"{e}.{l}.weight", f"{e}.{l}.bias"]
                self.shapes[f"{e}.{l}.weight"], self.shapes[f"{e}.{l}.bias"] = (H, k), (H,)
        for l, o in (("fc_att", len(self.feat_dims)), ("fc_out_1", output_dim1), ("fc_out_2", output_dim2)):
            self.names += [f"{l}.weight", f"{l}.bias"]
            self.shapes[f"{l}.weight"], self.shapes[f"{l}.bias"] = (o, H), (o,)
        assert sum(int(np.prod(s)) for s in self.shapes.values()) == self.n_params
        z = lambda: torch.zeros(self.n_params, dtype=torch.float32, device=self.device)  # noqa: E731
        self.params, self.grads, self.exp_avg, self.exp_avg_sq = z(), z(), z(), z()
        self.step_counter = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.loss = torch.zeros(3, dtype=torch.float32, device=self.device)
        self.ws = torch.empty(0, dtype=torch.uint8, device=self.device)
        self.training = True

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    def named_views(self, flat=None):
        flat = self.params if flat is None else flat
        out, o = {}, 0
        for n in self.names:
            k = int(np.prod(self.shapes[n]))
            out[n] = flat[o:o + k].view(self.shapes[n])
            o += k
        return out

    def state_dict(self):
        return {k: v.clone() for k, v in self.named_views().items()}

    def load_state_dict(self, sd):
        for n, dst in self.named_views().items():
            src = sd[n] if n in sd else sd["model." + n]
            if isinstance(src, np.ndarray):
                src = torch.from_numpy(src)
            dst.copy_(src.to(self.device, torch.float32))
        return self

    def _run(self, feats, emo, val, ext_masks, world):
        B = feats[0].shape[0]
        need = int(L.lib().mer_fusion_topn_workspace_bytes(C.byref(self.dims), B))
        if self.ws.numel() < need:
            self.ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        feats = [f.contiguous() for f in feats]
        fp = (C.c_void_p * len(feats))(*[f.data_ptr() for f in feats])
        mp = None
        if ext_masks is not None:
            mp = (C.c_void_p * len(ext_masks))(*[m.data_ptr() if m is not None else None for m in ext_masks])
        d = self.dims
        out = [torch.empty(B, n, dtype=torch.float32, device=self.device) for n in (d.hidden, d.out1, d.out2)]
        L.check(self._step(C.byref(d), L.ptr(self.params), L.ptr(self.grads), C.cast(fp, C.c_void_p), L.ptr(emo),
                           L.ptr(val), B, 1.0 / (B * world), self.dropout if emo is not None else 0.0, self.seed,
                           L.ptr(self.step_counter), C.cast(mp, C.c_void_p) if mp is not None else None,
                           L.ptr(self.ws), self.ws.numel(), L.ptr(self.loss), L.ptr(out[0]), L.ptr(out[1]),
                           L.ptr(out[2]), L.stream_ptr()))
        return out

    def forward(self, batch):
        """batch: {'feat0': [B, d0], ...}; eval-mode forward -> (features, emos_out, vals_out, interloss)."""
        feats = [batch[f"feat{i}"] for i in range(len(self.feat_dims))]
        f, e, v = self._run(feats, None, None, None, 1)
        return f, e, v, torch.zeros((), dtype=torch.int64, device=self.device)

    __call__ = forward

    def train_step(self, feats, emo, val, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, world_size=1,
                   ext_masks=None):
        """One optimisation step (forward, CE + MSE, backward, optional all-reduce, Adam)."""
        assert emo.dtype == torch.int64 and val.dtype == torch.float32
        _, eo, vo = self._run(feats, emo.contiguous(), val.contiguous(), ext_masks, world_size)
        if world_size > 1:
            import torch.distributed as dist
            dist.all_reduce(self.grads)
        L.check(self._adam(L.ptr(self.params), L.ptr(self.grads), L.ptr(self.exp_avg), L.ptr(self.exp_avg_sq),
                           self.n_params, lr, betas[0], betas[1], eps, weight_decay, 1.0,
                           self.grad_clip if self.grad_clip != -1 else 0.0, L.ptr(self.step_counter), L.stream_ptr()))
        return self.loss, eo, vo


class _Wrapper:
    """``get_models`` wraps the chosen net as ``.model`` (toolkit/models/__init__.py:18-46)."""

    def __init__(self, net):
        self.model = net

    def __call__(self, batch):
        return self.model(batch)

    def __getattr__(self, k):
        return getattr(self.model, k)

    def train(self, mode=True):
        self.model.train(mode)
        return self

    def eval(self):
        self.model.eval()
        return self


def get_models(args):
    """args: .model ('attention'), .feat_type ('utt' | 'frm_align' | 'frm_unalign'),
    .audio_dim/.text_dim/.video_dim, .output_dim1/.output_dim2, .dropout, .hidden_dim, .grad_clip
    (models/__init__.py:18-46)."""
    if args.model == "attention_topn":  # MER2026 toolkit: args.audio_dim holds the list of feature widths
        return _Wrapper(TopnFusionNet(args.audio_dim, args.hidden_dim, args.output_dim1, args.output_dim2,
                                      dropout=args.dropout, grad_clip=args.grad_clip,
                                      device=getattr(args, "device", "cuda")))
    assert args.model == "attention", "only the Attention / Attention_TOPN fusion nets are on the B200 path"
    net = FusionNet(args.audio_dim, args.text_dim, args.video_dim, args.hidden_dim, args.output_dim1,
                    args.output_dim2, dropout=args.dropout, grad_clip=args.grad_clip,
                    device=getattr(args, "device", "cuda"), feat_type=args.feat_type)
    return _Wrapper(net)


class Adam:
    """Hyper-parameter holder standing where ``optim.Adam(model.parameters(), lr, weight_decay)``
    stands in main-release.py:205; the update itself runs inside FusionNet.train_step."""

    def __init__(self, params=None, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay

    def zero_grad(self):
        pass


def train_or_eval_model(args, model, reg_loss, cls_loss, dataloader, epoch, optimizer=None, train=False,
                        calculate_results=None, world_size=1):
    """Mirror of main-release.py:17-87.  ``reg_loss`` / ``cls_loss`` are accepted for signature
    compatibility; in training the two losses are computed inside the fused step (same formulas,
    toolkit/utils/loss.py).  Returns the same ``save_results`` dict."""
    vidnames, val_preds, val_labels, emo_probs, emo_labels, losses = [], [], [], [], [], []
    assert not train or optimizer is not None
    net = model.model if hasattr(model, "model") else model
    net.train(train)
    for data in dataloader:
        batch, emos, vals, bnames = data
        vidnames += bnames
        batch = {k: v.cuda(non_blocking=True) for k, v in batch.items()}
        emos, vals = emos.cuda(non_blocking=True), vals.cuda(non_blocking=True)
        if train:
            loss3, emos_out, vals_out = net.train_step(
                batch["audios"], batch["texts"], batch["videos"], emos.long(), vals.float().view(-1, 1),
                lr=optimizer.lr, betas=optimizer.betas, eps=optimizer.eps,
                weight_decay=optimizer.weight_decay, world_size=world_size)
            loss = loss3[2]
        else:
            _, emos_out, vals_out, _ = net(batch)
            loss = cls_loss(emos_out, emos) + reg_loss(vals_out, vals)
        emo_probs.append(emos_out.data.cpu().numpy())
        emo_labels.append(emos.data.cpu().numpy())
        val_preds.append(vals_out.data.cpu().numpy())
        val_labels.append(vals.data.cpu().numpy())
        losses.append(loss.data.cpu().numpy())
    emo_probs, emo_labels = np.concatenate(emo_probs), np.concatenate(emo_labels)
    val_preds, val_labels = np.concatenate(val_preds), np.concatenate(val_labels)
    results = {}
    if calculate_results is not None:
        results, _ = calculate_results(emo_probs, emo_labels, val_preds, val_labels)
    return dict(names=vidnames, loss=np.mean(losses), emo_probs=emo_probs, val_preds=val_preds,
                **results)


class CELoss(torch.nn.Module):
    """toolkit/utils/loss.py:5-15 (host-side torch; used for eval-loss reporting)."""

    def forward(self, pred, target):
        pred = torch.nn.functional.log_softmax(pred, 1)
        return torch.nn.functional.nll_loss(pred, target.long(), reduction="sum") / len(pred)


class MSELoss(torch.nn.Module):
    """toolkit/utils/loss.py:18-28."""

    def forward(self, pred, target):
        pred, target = pred.view(-1, 1), target.view(-1, 1)
        return torch.nn.functional.mse_loss(pred, target, reduction="sum") / len(pred)
