"""A torch-backed stand-in for the slice of ``tensorflow.compat.v1`` + ``tf_slim`` that the reference's VGGish code uses
(audio/vggish/vggish_slim.py:33-130, audio/extract_vggish_embedding.py:27-49), so that the UNMODIFIED reference files
can be imported and run in a container without TensorFlow: a lazy op graph (placeholders, variables, named tensors),
``Graph`` / ``Session.run(feed_dict)``, ``variable_scope`` naming, ``slim.arg_scope`` defaults, ``slim.repeat`` scoping
(``conv3/conv3_1``), SAME-padded NHWC ``conv2d`` / ``max_pool2d``, ``fully_connected``, ``flatten`` and a ``Saver`` whose
"checkpoint" is an .npz of variables under their TF names.

TEST INFRASTRUCTURE (used by tests/golden/make_golden_vggish.py only).  What it pins: the reference's graph definition
(layer order, widths, scopes / variable names, where the pools sit, the NHWC flatten, the ReLU on the embedding) and
its extractor loop (hop sizes, batching, save rules) -- executed from the reference's own source.  What it cannot pin:
TensorFlow's float arithmetic (fp32 convolutions here are torch's)."""
import contextlib
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ graph
class Node:
    def __init__(self, graph, name, fn=None, inputs=()):
        self.graph, self.fn, self.inputs = graph, fn, tuple(inputs)
        self.name = graph.unique(name) + ":0"
        graph.tensors[self.name] = self

    def eval(self, feed, cache):
        if self in feed:
            return torch.as_tensor(np.asarray(feed[self]), dtype=torch.float32)
        if self not in cache:
            assert self.fn is not None, f"placeholder {self.name} was not fed"
            cache[self] = self.fn(*[i.eval(feed, cache) for i in self.inputs])
        return cache[self]

    __hash__ = object.__hash__


class Variable(Node):
    def __init__(self, graph, name, shape):
        super().__init__(graph, name)
        self.shape, self.value = tuple(shape), None
        graph.variables.append(self)

    def eval(self, feed, cache):
        assert self.value is not None, f"variable {self.name} was never restored"
        return self.value


class Graph:
    def __init__(self):
        self.tensors, self.variables, self.scopes, self.names = {}, [], [], set()

    def scoped(self, name):
        return "/".join(self.scopes + [name])

    def unique(self, name):
        full, base, n = self.scoped(name), self.scoped(name), 0
        while full in self.names:
            n += 1
            full = f"{base}_{n}"
        self.names.add(full)
        return full

    @contextlib.contextmanager
    def as_default(self):
        _GRAPHS.append(self)
        try:
            yield self
        finally:
            _GRAPHS.pop()

    def get_tensor_by_name(self, name):
        return self.tensors[name]


_GRAPHS = [Graph()]


def _g():
    return _GRAPHS[-1]


class Session:
    def __init__(self):
        self.graph = _g()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        cache = {}
        with torch.no_grad():
            out = [t.eval(feed_dict or {}, cache).numpy() for t in fetches]
        return out


class Saver:
    """Saver(var_list).restore(sess, path): the checkpoint is an .npz keyed by variable name (without ':0')."""

    def __init__(self, var_list, name=None, write_version=None):
        self.vars = list(var_list)

    def restore(self, session, path):
        with np.load(path) as ck:
            for v in self.vars:
                a = ck[v.name[:-2]]
                assert tuple(a.shape) == v.shape, (v.name, a.shape, v.shape)
                v.value = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


@contextlib.contextmanager
def variable_scope(name, default_name=None, values=None):
    g = _g()
    g.scopes.append(name)
    try:
        yield
    finally:
        g.scopes.pop()


# ------------------------------------------------------------------------------------------------ slim
_ARG_SCOPES = []


@contextlib.contextmanager
def arg_scope(fns, **kw):
    _ARG_SCOPES.append(([getattr(f, "_key", f) for f in fns], kw))
    try:
        yield
    finally:
        _ARG_SCOPES.pop()


def _scoped(fn):
    def wrapper(*args, **kw):
        merged = {}
        for keys, defaults in _ARG_SCOPES:   # outer scopes first, inner ones override, explicit arguments win
            if wrapper in keys:
                merged.update(defaults)
        merged.update(kw)
        return fn(*args, **merged)
    wrapper._key = wrapper
    wrapper.__name__ = fn.__name__
    return wrapper


def _same_pad(size, k, s):
    total = max((-(-size // s) - 1) * s + k - size, 0)
    return total // 2, total - total // 2


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


@_scoped
def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=None, weights_initializer=None,
           biases_initializer=None, trainable=True, scope=None):
    g = _g()
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)
    cin = inputs.channels
    with variable_scope(scope):
        w = Variable(g, "weights", (kh, kw, cin, num_outputs))   # HWIO, as TF stores them
        b = Variable(g, "biases", (num_outputs,))

        def run(x, wv, bv):  # x: NHWC
            assert padding == "SAME"
            ph, pw = _same_pad(x.shape[1], kh, sh), _same_pad(x.shape[2], kw, sw)
            xp = F.pad(x.permute(0, 3, 1, 2), (pw[0], pw[1], ph[0], ph[1]))
            return F.conv2d(xp, wv.permute(3, 2, 0, 1), bv, stride=(sh, sw)).permute(0, 2, 3, 1)
        out = Node(g, "Conv2D", run, (inputs, w, b))
        out.channels = num_outputs
        if activation_fn is not None:
            out = activation_fn(out)
    return out


@_scoped
def max_pool2d(inputs, kernel_size, stride=2, padding="VALID", scope=None):
    kh, kw = _pair(kernel_size)
    sh, sw = _pair(stride)

    def run(x):
        ph = _same_pad(x.shape[1], kh, sh) if padding == "SAME" else (0, 0)
        pw = _same_pad(x.shape[2], kw, sw) if padding == "SAME" else (0, 0)
        xp = F.pad(x.permute(0, 3, 1, 2), (pw[0], pw[1], ph[0], ph[1]), value=float("-inf"))
        return F.max_pool2d(xp, (kh, kw), (sh, sw)).permute(0, 2, 3, 1)
    with variable_scope(scope):
        out = Node(_g(), "MaxPool", run, (inputs,))
    out.channels = inputs.channels
    return out


@_scoped
def fully_connected(inputs, num_outputs, activation_fn=None, weights_initializer=None, biases_initializer=None,
                    trainable=True, scope=None):
    g = _g()
    with variable_scope(scope):
        w = Variable(g, "weights", (inputs.channels, num_outputs))
        b = Variable(g, "biases", (num_outputs,))
        out = Node(g, "BiasAdd", lambda x, wv, bv: x @ wv + bv, (inputs, w, b))
        out.channels = num_outputs
        if activation_fn is not None:
            out = activation_fn(out)
    return out


def flatten(inputs, scope=None):
    out = Node(_g(), "flatten", lambda x: x.reshape(x.shape[0], -1), (inputs,))
    out.channels = inputs.flat_features
    return out


def repeat(inputs, repetitions, layer, *args, **kwargs):
    """tf_slim.repeat: variable_scope(scope) around `repetitions` calls with scope = f'{scope}_{i + 1}'."""
    scope = kwargs.pop("scope", None) or "Repeat"
    out = inputs
    with variable_scope(scope):
        for i in range(repetitions):
            out = layer(out, *args, scope=f"{scope}_{i + 1}", **kwargs)
    return out


# ------------------------------------------------------------------------------------------------ tf
def placeholder(dtype, shape=None, name=None):
    n = Node(_g(), name or "Placeholder")
    n.static_shape = tuple(shape)
    return n


def reshape(x, shape):
    shape = list(shape)
    out = Node(_g(), "Reshape", lambda v: v.reshape(shape), (x,))
    out.channels = shape[-1]
    out.spatial = tuple(shape[1:-1])
    return out


def identity(x, name=None):
    out = Node(_g(), name or "Identity", lambda v: v, (x,))
    out.channels = x.channels
    return out


def _relu(x):
    out = Node(_g(), "Relu", torch.relu, (x,))
    out.channels = x.channels
    return out


def install(num_frames, num_bands):
    """Put the stand-ins into sys.modules as `tensorflow`, `tensorflow.compat`, `tensorflow.compat.v1`, `tf_slim`.
    The flatten width needs the static spatial shape: VGGish's four SAME 2x2 pools halve (num_frames, num_bands)."""
    tf = types.ModuleType("tensorflow.compat.v1")
    tf.disable_v2_behavior = lambda: None
    tf.float32 = "float32"
    tf.Graph, tf.Session = Graph, Session
    tf.placeholder, tf.reshape, tf.identity, tf.variable_scope = placeholder, reshape, identity, variable_scope
    tf.truncated_normal_initializer = lambda stddev=1.0: ("truncated_normal", stddev)
    tf.zeros_initializer = lambda: ("zeros",)
    tf.global_variables = lambda: list(_g().variables)
    tf.nn = types.SimpleNamespace(relu=_relu)
    tf.train = types.SimpleNamespace(Saver=Saver)
    slim = types.ModuleType("tf_slim")
    slim.arg_scope, slim.conv2d, slim.max_pool2d, slim.fully_connected = arg_scope, conv2d, max_pool2d, fully_connected
    slim.repeat = repeat

    def flatten_static(inputs, scope=None):   # static feature count of the NHWC flatten (what TF infers from shapes)
        pools = 0
        node = inputs
        while node.inputs:                    # walk back to the placeholder, counting the pools on the way
            if node.name.rsplit("/", 1)[-1].startswith("MaxPool"):
                pools += 1
            node = node.inputs[0]
        inputs.flat_features = -(-num_frames // 2 ** pools) * -(-num_bands // 2 ** pools) * inputs.channels
        return flatten(inputs, scope)
    slim.flatten = flatten_static
    root = types.ModuleType("tensorflow")
    compat = types.ModuleType("tensorflow.compat")
    root.compat, compat.v1 = compat, tf
    sys.modules.update({"tensorflow": root, "tensorflow.compat": compat, "tensorflow.compat.v1": tf, "tf_slim": slim})
    return tf, slim
