"""An eager, float64 stand-in for the slice of the TensorFlow-1.4 API that the reference's model files call, so that
THEIR graph-construction code (variable scopes and names, layer wiring, residual sums, noise shape, loss formulas, tower
slicing, average_gradients, per-tensor clipping, optimizer / EMA grouping) can be executed in this container, where
TensorFlow 1.4 cannot be installed.  TEST INFRASTRUCTURE ONLY: used by tests/golden/make_reference_graph_golden.py (run
here, where /root/reference exists) to produce tests/golden/ref_graph_*.npz; nothing in the product imports it.

What this is and is not.  Everything the reference WRITES is executed from its own source files: models/lstm.py,
models/res_lstm_l.py, models/res_lstm_base.py, models/discriminator_lstm.py, models/discriminator_dnn.py, models/dnn.py,
models/rced.py, models/gan_rnn_placeholder.py, models/gan.py, models/dnn_trainer.py, models/dnn_trainer_single_gpu.py,
models/BNLSTMCell.py, utils/ops.py, utils/bnorm.py, io_funcs/tfrecords_io.py (splice_feats) and the training loop
`train_one_iteration` of scripts/train_gan_rnn_placeholder.py (its sess.run calls go to `Session`, which re-executes the
reference's graph-building code on the current variables for every run).  What TensorFlow itself provides -- the
op kernels and the library layers -- is restated here on torch float64 tensors (derivatives by torch autograd):
  * elementwise / reduction / shape ops: one line of torch each;
  * tf.contrib.layers.fully_connected (TF r1.4 contrib/layers/python/layers/layers.py): variable scope
    "fully_connected" made unique per enclosing scope, variables "weights" [in, out] and "biases" [out], matmul over the
    last axis, bias or normalizer, then activation_fn (default relu); weights_regularizer -> REGULARIZATION_LOSSES;
  * tf.contrib.rnn.LSTMCell (TF r1.4 python/ops/rnn_cell_impl.py, LSTMCell.call): variables "kernel" [(I + P), 4C],
    "bias" [4C] (zeros), "w_f_diag" / "w_i_diag" / "w_o_diag" [C], "projection/kernel" [C, P]; gate order i, j, f, o.
    The generator script checks this restatement against the reference's OWN statement of the same equations
    (models/BNLSTMCell.py:176-213, its three batch_norm calls replaced by the identity);
  * tf.contrib.layers.conv2d (scope "Conv", NHWC, SAME), tf.nn.conv2d / conv1d with TensorFlow's SAME rule for strides,
    tf.nn.conv2d_transpose taken literally as the gradient of conv2d (autograd), tf.pad SYMMETRIC;
  * MultiRNNCell ("multi_rnn_cell/cell_%d"), tf.nn.dynamic_rnn (scope "rnn"; past sequence_length: zero output, state
    copied through -- python/ops/rnn.py _rnn_step), GradientDescentOptimizer, AdamOptimizer (python/training/adam.py:
    lr_t = lr sqrt(1 - b2^t) / (1 - b1^t), var -= lr_t m / (sqrt(v) + eps)), ExponentialMovingAverage, clip_by_norm.
Static shapes: the reference's placeholders leave the time axis unknown ([B, None, D]); TensorFlow's shape inference
carries that None through every rank-3 tensor of these graphs, and utils/ops.py:19-30 (gaussian_noise_layer) depends on
it.  get_shape() therefore reports None for axis 1 of rank-3 tensors while STATE.unknown_time is set.
"""
import contextlib
import re
import sys
import types

import numpy as np
import torch

F64 = torch.float64


class _State(object):
    def __init__(self):
        self.reset()

    def reset(self):
        self.vars = {}                 # full name -> TT (creation order)
        self.init = None               # name -> ndarray: initial values; a variable missing here is an error when set
        self.init_used = set()
        self.scope = [VariableScope("", False)]
        self.scope_counts = {}
        self.collections = {}
        self.feeds = {}                # placeholder name -> ndarray
        self.noise = []                # queue of arrays handed out by tf.random_normal (unit variance)
        self.unknown_time = True
        self.rng = np.random.default_rng(0)
        self.anon = 0
        self.grad_log = []             # (optimizer, [(grad, var)]) per compute_gradients call
        self.apply_log = []            # (optimizer, [(grad, var)]) per apply_gradients call
        self.emas = []                 # ExponentialMovingAverage objects, creation order
        self.summaries = []            # (name, tensor) per tf.summary.histogram / scalar call
        self.retrace = False           # Session.run re-executes the reference's graph-building code on the CURRENT variables
        self.anon_list, self.anon_i = [], 0
        self.persist = {}              # (kind, creation index within one build) -> optimizer slots / EMA shadows
        self.created = {}              # kind -> objects created in the current build
        self.noise_fn = None           # fallback for tf.random_normal when the queue is empty
        self.allow_bn = False          # batch_norm stand-in (wiring only) enabled
        self.ctrl = []                 # stack of control_dependencies op lists
        self.ran_updates = []          # names of UPDATE_OPS executed, in order

    def begin_retrace(self, feeds):
        """One sess.run(fetches, feed_dict): the graph code runs again, on the same variable objects (current values), the
        same optimizer slots and EMA shadows; nothing may be created."""
        self.retrace = True
        self.feeds = dict(feeds)
        self.scope = [VariableScope("", False)]
        self.scope_counts, self.created, self.anon_i = {}, {}, 0
        self.grad_log, self.apply_log, self.emas, self.summaries = [], [], [], []

    def slots(self, kind, make):
        i = self.created.get(kind, 0)
        self.created[kind] = i + 1
        return self.persist.setdefault((kind, i), make())


class VariableScope(object):
    def __init__(self, name, reuse):
        self.name, self.reuse = name, reuse

    def reuse_variables(self):
        self.reuse = True


class Dimension(object):
    def __init__(self, v):
        self.value = v

    def __int__(self):
        return int(self.value)

    __index__ = __int__

    def __eq__(self, o):
        return self.value == (o.value if isinstance(o, Dimension) else o)

    def __hash__(self):
        return hash(self.value)

    def __mul__(self, o):
        return Dimension(self.value * int(o))

    __rmul__ = __mul__

    def __repr__(self):
        return "?" if self.value is None else str(self.value)


class TensorShape(object):
    def __init__(self, dims):
        self.dims = list(dims)

    def as_list(self):
        return list(self.dims)

    def with_rank(self, r):
        assert len(self.dims) == r
        return self

    def __getitem__(self, i):
        if isinstance(i, slice):
            return TensorShape(self.dims[i])
        return Dimension(self.dims[i])

    def __len__(self):
        return len(self.dims)

    def __iter__(self):
        return iter(Dimension(d) for d in self.dims)

    def __repr__(self):
        return "(%s)" % ", ".join("?" if d is None else str(d) for d in self.dims)


def _raw(x):
    if isinstance(x, TT):
        return x.v
    if torch.is_tensor(x):
        return x.to(F64)
    if isinstance(x, (list, tuple)) and any(isinstance(e, TT) for e in x):
        return torch.stack([_raw(e) for e in x])
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


class TT(object):
    """A tensor or a variable: identity equality / hashing like tf.Tensor and tf.Variable."""

    def __init__(self, v, name=None, trainable=False):
        self.v = v if torch.is_tensor(v) else _raw(v)
        self.name, self.trainable = name, trainable
        self.dtype = "float32"

    def _static(self):
        d = list(self.v.shape)
        if len(d) == 3 and STATE.unknown_time:
            d[1] = None
        return d

    def get_shape(self):
        return TensorShape(self._static())

    @property
    def shape(self):
        return TensorShape(self._static())

    def assign(self, value):
        return assign(self, value)

    def __getitem__(self, idx):
        return TT(self.v[idx])

    def __add__(self, o): return TT(self.v + _raw(o))
    __radd__ = __add__
    def __sub__(self, o): return TT(self.v - _raw(o))
    def __rsub__(self, o): return TT(_raw(o) - self.v)
    def __mul__(self, o): return TT(self.v * _raw(o))
    __rmul__ = __mul__
    def __truediv__(self, o): return TT(self.v / _raw(o))
    def __rtruediv__(self, o): return TT(_raw(o) / self.v)
    def __neg__(self): return TT(-self.v)
    def __pow__(self, o): return TT(self.v ** _raw(o))

    def numpy(self):
        return self.v.detach().numpy().copy()


class Op(object):
    """A deferred side effect (apply_gradients, EMA update, tf.group): run by calling it, like sess.run(op)."""

    def __init__(self, fns):
        self.fns = list(fns)

    def __call__(self):
        for f in self.fns:
            f()


# ---------------------------------------------------------------------------------------- scopes and variables
def _join(a, b):
    return b if not a else a + "/" + b


@contextlib.contextmanager
def variable_scope(name_or_scope=None, default_name=None, values=None, reuse=None):
    cur = STATE.scope[-1]
    if isinstance(name_or_scope, VariableScope):          # re-enter an existing scope (tf.variable_scope(tf.get_variable_scope()))
        new = VariableScope(name_or_scope.name, name_or_scope.reuse if reuse is None else reuse)
    else:
        if name_or_scope is None:                         # default_name, made unique inside the enclosing scope
            base = _join(cur.name, default_name)
            n = STATE.scope_counts.get(base, 0)
            full = base if n == 0 else "%s_%d" % (base, n)
            STATE.scope_counts[base] = n + 1
        else:
            full = _join(cur.name, name_or_scope)
            STATE.scope_counts[full] = STATE.scope_counts.get(full, 0) + 1
        new = VariableScope(full, cur.reuse if reuse is None else (reuse or cur.reuse))
    STATE.scope.append(new)
    try:
        yield new
    finally:
        STATE.scope.pop()
        # python/ops/variable_scope.py close_variable_subscopes: leaving a scope resets the unique-name counters below it,
        # which is what makes a second pass through "g_model" (reuse) arrive at fully_connected, fully_connected_1, ... again
        if not isinstance(name_or_scope, VariableScope):
            for k in list(STATE.scope_counts):
                if k.startswith(new.name + "/"):
                    STATE.scope_counts[k] = 0


def get_variable_scope():
    return STATE.scope[-1]


@contextlib.contextmanager
def _noop(*a, **k):
    yield


name_scope = device = _noop


@contextlib.contextmanager
def control_dependencies(ops):
    """What is created inside depends on `ops`.  The one use in the reference's graph code is around compute_gradients
    (gan_rnn_placeholder.py:163-175, gan.py:139-143, dnn_trainer*.py): the stand-in records the ops on the optimizer whose
    compute_gradients runs inside, and that optimizer's apply op runs them first."""
    STATE.ctrl.append([o for o in (ops or []) if isinstance(o, Op)])
    try:
        yield
    finally:
        STATE.ctrl.pop()


def _get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, regularizer=None, **_):
    sc = STATE.scope[-1]
    full = _join(sc.name, name)
    if STATE.retrace:
        if full not in STATE.vars:
            raise ValueError("re-executed graph code asks for a new variable %s" % full)
        if regularizer is not None:
            _add_reg(STATE.vars[full], regularizer)
        return STATE.vars[full]
    if full in STATE.vars:
        if not sc.reuse:
            raise ValueError("Variable %s already exists, disallowed (reuse not set)" % full)
        return STATE.vars[full]
    if sc.reuse:
        raise ValueError("Variable %s does not exist (scope is reusing)" % full)
    shape = [int(s) for s in (shape if isinstance(shape, (list, tuple, TensorShape)) else [shape])]
    if STATE.init is not None:
        if full not in STATE.init:
            raise KeyError("the reference graph creates %s %s, which the supplied parameter set does not name" % (full, shape))
        val = np.asarray(STATE.init[full], np.float64)
        if list(val.shape) != shape:
            raise ValueError("%s: reference graph shape %s, supplied %s" % (full, shape, list(val.shape)))
        STATE.init_used.add(full)
    else:
        val = (initializer or xavier_initializer())(shape)
    t = torch.tensor(val, dtype=F64, requires_grad=bool(trainable))
    v = TT(t, name=full + ":0", trainable=bool(trainable))
    STATE.vars[full] = v
    if regularizer is not None:
        _add_reg(v, regularizer)
    return v


def _add_reg(v, regularizer):
    """one REGULARIZATION_LOSSES entry per variable (added where TensorFlow adds it: at creation), re-evaluated on the
    variable's current value when the graph code is re-executed"""
    items = STATE.collections.setdefault(GraphKeys.REGULARIZATION_LOSSES, [])
    t = regularizer(v)
    t.var = v
    for i, old in enumerate(items):
        if getattr(old, "var", None) is v:
            items[i] = t
            return
    items.append(t)


def get_variable(*a, **k):
    return _get_variable(*a, **k)


def Variable(value, trainable=True, name=None, **_):
    if STATE.retrace:
        v = STATE.anon_list[STATE.anon_i]
        STATE.anon_i += 1
        return v
    n = "Variable" if STATE.anon == 0 else "Variable_%d" % STATE.anon
    STATE.anon += 1
    full = _join(STATE.scope[-1].name, name or n)
    v = TT(torch.tensor(np.asarray(value, np.float64), dtype=F64, requires_grad=bool(trainable)), name=full + ":0",
           trainable=bool(trainable))
    STATE.vars.setdefault("__anon__/" + full, v)
    STATE.anon_list.append(v)
    return v


def trainable_variables():
    return [v for k, v in STATE.vars.items() if v.trainable and not k.startswith("__anon__/")]


def get_collection(key, scope=None):
    items = STATE.collections.get(key, [])
    if scope is None:
        return list(items)
    return [i for i in items if re.match(scope, getattr(i, "name", "") or "")]


def placeholder(dtype, shape=None, name=None):
    val = np.asarray(STATE.feeds[name], np.float64)
    want = list(shape)
    assert len(want) == val.ndim and all(w is None or int(w) == s for w, s in zip(want, val.shape)), (name, want, val.shape)
    return TT(torch.tensor(val, dtype=F64), name=name + ":0")


def assign(ref, value):
    with torch.no_grad():
        ref.v.copy_(_raw(value))
    return ref


# ------------------------------------------------------------------------------------------------ initializers
def xavier_initializer(uniform=True, **_):
    def f(shape, **__):
        fan_in = shape[0] if len(shape) < 2 else int(np.prod(shape[:-1]))
        fan_out = shape[-1]
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return STATE.rng.uniform(-lim, lim, size=shape)
    return f


def zeros_initializer(**_):
    return lambda shape, **__: np.zeros(shape)


def ones_initializer(**_):
    return lambda shape, **__: np.ones(shape)


def constant_initializer(value=0.0, **_):
    return lambda shape, **__: np.full(shape, float(value))


def truncated_normal_initializer(mean=0.0, stddev=1.0, **_):
    return lambda shape, **__: mean + stddev * np.clip(STATE.rng.standard_normal(shape), -2, 2)


def random_normal_initializer(mean=0.0, stddev=1.0, **_):
    return lambda shape, **__: mean + stddev * STATE.rng.standard_normal(shape)


def l2_regularizer(scale, scope=None):
    def f(w):
        t = TT(scale * 0.5 * (w.v ** 2).sum())
        t.name = (w.name or "").rsplit("/", 1)[0] + "/kernel/Regularizer/l2_regularizer:0"
        return t
    return f


# --------------------------------------------------------------------------------------------------------- ops
def _t(f):
    return lambda x, *a, **k: TT(f(_raw(x)))


tanh, sigmoid = _t(torch.tanh), _t(torch.sigmoid)
square, sqrt, ones_like, zeros_like = _t(torch.square), _t(torch.sqrt), _t(torch.ones_like), _t(torch.zeros_like)


def relu(x, name=None): return TT(torch.relu(_raw(x)))
def maximum(x, y, name=None): return TT(torch.maximum(_raw(x), _raw(y)))
def add(x, y, name=None): return TT(_raw(x) + _raw(y))
def matmul(a, b, **_): return TT(_raw(a) @ _raw(b))
def squared_difference(x, y, name=None): return TT((_raw(x) - _raw(y)) ** 2)
def clip_by_value(x, lo, hi, name=None): return TT(torch.clamp(_raw(x), float(lo), float(hi)))
def constant(v, **_): return TT(_raw(v))
def expand_dims(x, axis, name=None): return TT(_raw(x).unsqueeze(axis))
def squeeze(x, axis=None, **_): return TT(_raw(x).squeeze() if axis is None else _raw(x).squeeze(axis))
def reshape(x, shape, name=None): return TT(_raw(x).reshape([int(s) for s in shape]))
def concat(values, axis, name=None): return TT(torch.cat([_raw(v) for v in values], dim=axis))
def bias_add(x, b, **_): return TT(_raw(x) + _raw(b))
def l2_loss(x, name=None): return TT(0.5 * (_raw(x) ** 2).sum())


def split(value, num_or_size_splits, axis=0, **_):
    return [TT(c) for c in torch.chunk(_raw(value), int(num_or_size_splits), dim=axis)]


def slice_(x, begin, size, name=None):
    r = _raw(x)
    return TT(r[tuple(slice(b, None if s == -1 else b + s) for b, s in zip(begin, size))])


def shape(x, **_):
    return [int(d) for d in _raw(x).shape]


def pad(tensor, paddings, mode="CONSTANT", **_):
    """tf.pad; SYMMETRIC mirrors INCLUDING the edge element (numpy 'symmetric')."""
    r = _raw(tensor).detach().numpy()
    return TT(torch.tensor(np.pad(r, [tuple(int(q) for q in pr) for pr in paddings], mode={"CONSTANT": "constant", "SYMMETRIC": "symmetric",
                                                                                     "REFLECT": "reflect"}[mode])))


def _reduce(f):
    def g(x, axis=None, keep_dims=False, name=None, **_):
        r = _raw(x)
        return TT(f(r) if axis is None else f(r, dim=axis, keepdim=keep_dims))
    return g


reduce_mean, reduce_sum = _reduce(torch.mean), _reduce(torch.sum)


def moments(x, axes, **_):
    r = _raw(x)
    m = r.mean(dim=axes)
    return TT(m), TT(((r - m) ** 2).mean(dim=axes))


def batch_normalization(x, mean, variance, offset, scale, variance_epsilon, name=None):
    return TT((_raw(x) - _raw(mean)) * torch.rsqrt(_raw(variance) + variance_epsilon) * _raw(scale) + _raw(offset))


def random_normal(shape, mean=0.0, stddev=1.0, dtype=None, seed=None, name=None):
    """Unit-variance draws come from STATE.noise (queued by the caller, one array per call, shape checked): the
    generator script decides the numbers, the reference decides the SHAPE."""
    unit = np.asarray(STATE.noise.pop(0) if STATE.noise else STATE.noise_fn([int(q) for q in shape]), np.float64)
    assert list(unit.shape) == [int(s) for s in shape], ("tf.random_normal asked for", list(shape), "queued", unit.shape)
    return TT(mean + _raw(stddev) * torch.tensor(unit, dtype=F64))


def clip_by_norm(t, clip_norm, axes=None, name=None):
    """python/ops/clip_ops.py: t * clip_norm / max(||t||_2, clip_norm)."""
    r = _raw(t)
    n = torch.sqrt((r * r).sum())
    return TT(r * float(clip_norm) / torch.maximum(n, torch.tensor(float(clip_norm), dtype=F64)))


def mean_squared_error(labels, predictions, weights=1.0, **_):
    """tf.losses.mean_squared_error, default reduction SUM_BY_NONZERO_WEIGHTS with unit weights: the plain mean."""
    return TT(((_raw(predictions) - _raw(labels)) ** 2).mean())


def dropout(x, keep_prob, **_):
    if float(keep_prob) == 1.0:
        return x
    raise NotImplementedError("tf.nn.dropout with keep_prob < 1 is outside the fixtures")


# ------------------------------------------------------------------------------------------------ contrib.layers
def fully_connected(inputs, num_outputs, activation_fn=relu, normalizer_fn=None, normalizer_params=None,
                    weights_initializer=None, weights_regularizer=None, biases_initializer=zeros_initializer(),
                    biases_regularizer=None, reuse=None, variables_collections=None, outputs_collections=None,
                    trainable=True, scope=None):
    with variable_scope(scope, "fully_connected", [inputs], reuse=reuse):
        x = _raw(inputs)
        w = _get_variable("weights", [x.shape[-1], int(num_outputs)], initializer=weights_initializer or xavier_initializer(),
                          regularizer=weights_regularizer, trainable=trainable)
        out = TT(x @ w.v)
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            b = _get_variable("biases", [int(num_outputs)], initializer=biases_initializer, trainable=trainable)
            out = TT(out.v + b.v)
        if activation_fn is not None:
            out = activation_fn(out)
        return out


def conv2d(inputs, num_outputs, kernel_size, stride=1, padding="SAME", activation_fn=relu, normalizer_fn=None,
           normalizer_params=None, weights_initializer=None, weights_regularizer=None, biases_initializer=zeros_initializer(),
           reuse=None, trainable=True, scope=None, **_):
    """tf.contrib.layers.conv2d (contrib/layers/python/layers/layers.py `convolution`): scope "Conv" made unique, variables
    "weights" [kh, kw, C_in, C_out] and "biases" [C_out], NHWC, stride 1, SAME (odd kernels: k // 2 zeros per side)."""
    assert stride == 1 and padding == "SAME"
    kh, kw = (int(k) for k in kernel_size)
    assert kh % 2 == 1 and kw % 2 == 1
    with variable_scope(scope, "Conv", [inputs], reuse=reuse):
        x = _raw(inputs)                                                   # N, H, W, C
        w = _get_variable("weights", [kh, kw, x.shape[-1], int(num_outputs)], initializer=weights_initializer or xavier_initializer(),
                          regularizer=weights_regularizer, trainable=trainable)
        out = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w.v.permute(3, 2, 0, 1), padding=(kh // 2, kw // 2))
        out = TT(out.permute(0, 2, 3, 1))
        if normalizer_fn is not None:
            out = normalizer_fn(out, **(normalizer_params or {}))
        elif biases_initializer is not None:
            b = _get_variable("biases", [int(num_outputs)], initializer=biases_initializer, trainable=trainable)
            out = TT(out.v + b.v)
        if activation_fn is not None:
            out = activation_fn(out)
        return out


def _same(L, k, s):
    """TensorFlow SAME padding along one axis (core/framework/common_shape_fns.cc GetWindowedOutputSize): output length
    ceil(L / s), total padding max((out - 1) s + k - L, 0), the odd element after."""
    out = -(-L // s)
    total = max((out - 1) * s + k - L, 0)
    return out, total // 2, total - total // 2


def _conv2d_raw(x, w, strides):
    sh, sw = int(strides[1]), int(strides[2])
    _, pt, pb = _same(x.shape[1], w.shape[0], sh)
    _, pl, pr = _same(x.shape[2], w.shape[1], sw)
    xp = torch.nn.functional.pad(x.permute(0, 3, 1, 2), (pl, pr, pt, pb))
    return torch.nn.functional.conv2d(xp.contiguous(), w.permute(3, 2, 0, 1).contiguous(), stride=(sh, sw)).permute(0, 2, 3, 1)


def nn_conv2d(input, filter, strides, padding, **_):
    """tf.nn.conv2d, NHWC, filter [kh, kw, C_in, C_out]."""
    assert padding == "SAME" and strides[0] == 1 and strides[3] == 1
    return TT(_conv2d_raw(_raw(input), _raw(filter), strides))


def nn_conv1d(value, filters, stride, padding, **_):
    """tf.nn.conv1d: conv2d on [B, 1, L, C] with filter [1, k, C_in, C_out] (python/ops/nn_ops.py conv1d)."""
    assert padding == "SAME"
    return TT(_conv2d_raw(_raw(value).unsqueeze(1), _raw(filters).unsqueeze(0), [1, 1, int(stride), 1]).squeeze(1))


def nn_conv2d_transpose(value, filter, output_shape, strides, padding="SAME", **_):
    """tf.nn.conv2d_transpose is, by definition, conv2d_backprop_input: the gradient of conv2d(z, filter) wrt z (z of
    output_shape, filter [kh, kw, C_out, C_in]) contracted with `value` -- taken here literally, by autograd."""
    assert padding == "SAME"
    z = torch.zeros([int(s) for s in output_shape], dtype=F64, requires_grad=True)
    y = _conv2d_raw(z, _raw(filter), strides)
    v = _raw(value)
    assert list(y.shape) == list(v.shape), (list(y.shape), list(v.shape))
    return TT(torch.autograd.grad(y, z, grad_outputs=v, create_graph=True)[0])


def batch_norm(inputs, is_training=True, scale=False, renorm=False, decay=0.999, epsilon=0.001, scope=None, reuse=None, **_):
    """WIRING ONLY.  contrib batch_norm's arithmetic (renorm corrections, zero-debiased renorm averages) is TensorFlow library
    code and is NOT restated here: this stand-in normalises with the plain batch moments and exists so that the reference's
    graph code can be executed with batch_norm = True and the UPDATE_OPS it collects -- which copies of which network, run by
    which optimizer -- can be recorded.  Per call (= per graph copy) it registers one update op in GraphKeys.UPDATE_OPS,
    named after the variable scope like TensorFlow's (`<scope>/BatchNorm/AssignMovingAvg`), plus a serial number."""
    if not STATE.allow_bn:
        raise NotImplementedError("contrib batch_norm arithmetic is outside the numeric fixtures (tf_standin.batch_norm docstring)")
    with variable_scope(scope, "BatchNorm", [inputs], reuse=reuse) as sc:
        x = _raw(inputs)
        n = x.shape[-1]
        beta = _get_variable("beta", [n], initializer=zeros_initializer())
        gamma = _get_variable("gamma", [n], initializer=ones_initializer()) if scale else None
        mm = _get_variable("moving_mean", [n], initializer=zeros_initializer(), trainable=False)
        mv = _get_variable("moving_variance", [n], initializer=ones_initializer(), trainable=False)
        red = list(range(x.dim() - 1))
        if is_training:
            mean, var = x.mean(dim=red), x.var(dim=red, unbiased=False)
            def update(mean=mean.detach(), var=var.detach()):
                with torch.no_grad():
                    mm.v -= (1.0 - decay) * (mm.v - mean)
                    mv.v -= (1.0 - decay) * (mv.v - var)
            op = Op([update])
            op.name = "%s/AssignMovingAvg#%d" % (sc.name, len(STATE.collections.setdefault(GraphKeys.UPDATE_OPS, [])))
            STATE.collections[GraphKeys.UPDATE_OPS].append(op)
        else:
            mean, var = mm.v, mv.v
        y = (x - mean) * torch.rsqrt(var + epsilon)
        if gamma is not None:
            y = y * gamma.v
        return TT(y + beta.v)


def flatten(x, **_):
    r = _raw(x)
    return TT(r.reshape(r.shape[0], -1))


# --------------------------------------------------------------------------------------------------- contrib.rnn
class LSTMStateTuple(tuple):
    def __new__(cls, c, h):
        return tuple.__new__(cls, (c, h))

    c = property(lambda s: s[0])
    h = property(lambda s: s[1])


class RNNCell(object):
    _scope_name = "rnn_cell"

    def __init__(self, _reuse=None, **_):
        self._reuse = _reuse

    def zero_state(self, batch_size, dtype):
        def z(n):
            return TT(torch.zeros(int(batch_size), int(n), dtype=F64))
        ss = self.state_size
        if isinstance(ss, LSTMStateTuple):
            return LSTMStateTuple(z(ss.c), z(ss.h))
        if isinstance(ss, tuple):
            return tuple(LSTMStateTuple(z(s.c), z(s.h)) for s in ss)
        return z(ss)

    def __call__(self, inputs, state, scope=None):
        # layers/base.py Layer.__call__: the cell's variables live in a scope named after the layer; the first call creates
        # them, every later call (the next time step, traced once in TensorFlow) reuses them
        first = not getattr(self, "_built", False)
        with variable_scope(self._scope_name, reuse=(self._reuse or None) if first else True):
            out = self.call(inputs, state)
        self._built = True
        return out


class LSTMCell(RNNCell):
    _scope_name = "lstm_cell"

    def __init__(self, num_units, use_peepholes=False, cell_clip=None, initializer=None, num_proj=None, proj_clip=None,
                 forget_bias=1.0, state_is_tuple=True, activation=None, reuse=None, **_):
        super(LSTMCell, self).__init__(_reuse=reuse)
        assert state_is_tuple and cell_clip is None and proj_clip is None
        self.C, self.P, self.peep, self.fb = int(num_units), num_proj, use_peepholes, float(forget_bias)
        self.init, self.act = initializer, activation or tanh

    @property
    def state_size(self):
        return LSTMStateTuple(self.C, self.P or self.C)

    @property
    def output_size(self):
        return self.P or self.C

    def call(self, inputs, state):
        c_prev, m_prev = state
        x = _raw(inputs)
        P = self.P or self.C
        kernel = _get_variable("kernel", [x.shape[1] + P, 4 * self.C], initializer=self.init)
        bias = _get_variable("bias", [4 * self.C], initializer=zeros_initializer())
        z = torch.cat([x, _raw(m_prev)], 1) @ kernel.v + bias.v
        i, j, f, o = torch.chunk(z, 4, dim=1)
        cp = _raw(c_prev)
        if self.peep:
            w_f = _get_variable("w_f_diag", [self.C], initializer=self.init)
            w_i = _get_variable("w_i_diag", [self.C], initializer=self.init)
            w_o = _get_variable("w_o_diag", [self.C], initializer=self.init)
            c = torch.sigmoid(f + self.fb + w_f.v * cp) * cp + torch.sigmoid(i + w_i.v * cp) * _raw(self.act(TT(j)))
            m = torch.sigmoid(o + w_o.v * c) * _raw(self.act(TT(c)))
        else:
            c = torch.sigmoid(f + self.fb) * cp + torch.sigmoid(i) * _raw(self.act(TT(j)))
            m = torch.sigmoid(o) * _raw(self.act(TT(c)))
        if self.P is not None:
            with variable_scope("projection"):
                wp = _get_variable("kernel", [self.C, self.P], initializer=self.init)
            m = m @ wp.v
        return TT(m), LSTMStateTuple(TT(c), TT(m))


class MultiRNNCell(RNNCell):
    _scope_name = "multi_rnn_cell"

    def __init__(self, cells, state_is_tuple=True):
        super(MultiRNNCell, self).__init__()
        self.cells = list(cells)

    @property
    def state_size(self):
        return tuple(c.state_size for c in self.cells)

    def call(self, inputs, state):
        cur, new = inputs, []
        for i, cell in enumerate(self.cells):
            with variable_scope("cell_%d" % i):
                cur, s = cell(cur, state[i])
            new.append(s)
        return cur, tuple(new)


class DropoutWrapper(RNNCell):
    def __init__(self, cell, output_keep_prob=1.0, **_):
        raise NotImplementedError("DropoutWrapper (keep_prob < 1) is outside the fixtures")


def dynamic_rnn(cell, inputs, sequence_length=None, initial_state=None, dtype=None, time_major=False, scope=None, **_):
    """python/ops/rnn.py: scope "rnn"; batch-major inputs; at t >= sequence_length[b] the emitted output is zero and the
    state is copied through (_rnn_step / _copy_some_through)."""
    assert not time_major
    x = _raw(inputs)
    B, T = x.shape[0], x.shape[1]
    ln = None if sequence_length is None else _raw(sequence_length).to(torch.int64)        # math_ops.to_int32
    state = initial_state if initial_state is not None else cell.zero_state(B, dtype)
    outs = []

    def keep(new, old, live):
        if isinstance(new, LSTMStateTuple):
            return LSTMStateTuple(keep(new.c, old.c, live), keep(new.h, old.h, live))
        if isinstance(new, tuple):
            return tuple(keep(n, o, live) for n, o in zip(new, old))
        return TT(torch.where(live, new.v, old.v))
    with variable_scope(scope or "rnn"):
        for t in range(T):
            out, new_state = cell(TT(x[:, t]), state)
            if ln is None:
                state = new_state
                outs.append(out.v)
            else:
                live = (t < ln).reshape(B, 1)
                state = keep(new_state, state, live)
                outs.append(torch.where(live, out.v, torch.zeros_like(out.v)))
    return TT(torch.stack(outs, 1)), state


# --------------------------------------------------------------------------------------------------------- train
class _Optimizer(object):
    def __init__(self, learning_rate):
        self.lr = learning_rate
        self.deps = []                 # ops its compute_gradients calls were made to depend on (control_dependencies)

    def _run_deps(self):
        for o in self.deps:
            STATE.ran_updates.append(getattr(o, "name", "?"))
            o()

    def compute_gradients(self, loss, var_list=None):
        vs_ = list(var_list)
        g = torch.autograd.grad(_raw(loss), [v.v for v in vs_], retain_graph=True, allow_unused=True)
        gv = [(None if gi is None else TT(gi.detach().clone()), v) for gi, v in zip(g, vs_)]
        STATE.grad_log.append((self, gv))
        for ops in STATE.ctrl:
            for o in ops:
                if o not in self.deps:
                    self.deps.append(o)
        return gv

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = list(grads_and_vars)
        STATE.apply_log.append((self, gv))
        return Op([self._run_deps, lambda: self._apply(gv)])


    def minimize(self, loss, global_step=None, var_list=None, **_):
        return self.apply_gradients(self.compute_gradients(loss, var_list=var_list))


class GradientDescentOptimizer(_Optimizer):
    def _apply(self, gv):
        lr = float(_raw(self.lr))
        with torch.no_grad():
            for g, v in gv:
                v.v -= lr * g.v


class AdamOptimizer(_Optimizer):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **_):
        super(AdamOptimizer, self).__init__(learning_rate)
        self.b1, self.b2, self.eps = beta1, beta2, epsilon
        self.st = STATE.slots("adam", lambda: {"t": 0, "m": {}, "v": {}})      # beta powers and slots outlive a re-execution

    def _apply(self, gv):
        self.st["t"] += 1
        t = self.st["t"]
        lr_t = float(_raw(self.lr)) * np.sqrt(1.0 - self.b2 ** t) / (1.0 - self.b1 ** t)
        with torch.no_grad():
            for g, v in gv:
                m = self.st["m"].setdefault(v, torch.zeros_like(v.v))
                s = self.st["v"].setdefault(v, torch.zeros_like(v.v))
                m.mul_(self.b1).add_((1.0 - self.b1) * g.v)
                s.mul_(self.b2).add_((1.0 - self.b2) * g.v * g.v)
                v.v -= lr_t * m / (torch.sqrt(s) + self.eps)


class RMSPropOptimizer(_Optimizer):
    def _apply(self, gv):
        raise NotImplementedError


class ExponentialMovingAverage(object):
    def __init__(self, decay, num_updates=None, **_):
        self.decay = float(decay)
        self.shadow = STATE.slots("ema", dict)
        STATE.emas.append(self)

    def apply(self, var_list=None):
        vs_ = list(var_list)
        for v in vs_:
            self.shadow.setdefault(v, v.v.detach().clone())           # shadow variables start at the variable's value

        def run():
            with torch.no_grad():
                for v in vs_:
                    self.shadow[v] -= (1.0 - self.decay) * (self.shadow[v] - v.v)
        return Op([run])

    def average(self, v):
        return TT(self.shadow[v])


def group(*ops, **_):
    return Op(ops)


class Session(object):
    """sess.run for the eager stand-in.  Without a feed_dict: reads / eager assigns.  With one: `rebuild(feeds)` re-executes
    the reference's graph-building code on the current variables (STATE.begin_retrace) and returns the new model object; the
    fetches -- attributes of the ORIGINAL model object, as the reference's scripts pass them -- are looked up by attribute
    name on it.  Fetched tensors are evaluated on the pre-update state, then the fetched ops run (one run = one forward)."""
    graph = None

    def __init__(self, rebuild=None):
        self.rebuild, self.names, self.log = rebuild, {}, []

    def bind(self, model):
        self.names = {id(v): k for k, v in vars(model).items() if v is not None}

    @staticmethod
    def _value(o):
        if isinstance(o, (list, tuple)):
            return [Session._value(e) for e in o]
        if isinstance(o, TT):
            return o.numpy()
        return o

    def run(self, fetches, feed_dict=None):
        if fetches is None:
            return None
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        if feed_dict is None:
            out = [self._value(f) for f in fl]
        else:
            new = self.rebuild({k.name[:-2]: np.asarray(v, np.float64) for k, v in feed_dict.items()})
            names = [self.names[id(f)] for f in fl]
            self.log.append(names)
            objs = [getattr(new, n) for n in names]
            out = [None if isinstance(o, Op) else self._value(o) for o in objs]
            for o in objs:
                if isinstance(o, Op):
                    o()
        return out[0] if single else out


class _Dummy(object):
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, n):
        return _Dummy()

    def __call__(self, *a, **k):
        return None


class GraphKeys(object):
    UPDATE_OPS, REGULARIZATION_LOSSES, TRAINABLE_VARIABLES = "update_ops", "regularization_losses", "trainable_variables"


STATE = None


def install():
    """Registers the stand-in as `tensorflow` (+ contrib.layers / contrib.rnn) in sys.modules and returns (tf, STATE)."""
    global STATE
    STATE = _State()
    tf = types.ModuleType("tensorflow")
    me = sys.modules[__name__]
    for n in ("variable_scope get_variable_scope name_scope device control_dependencies get_variable Variable "
              "trainable_variables get_collection placeholder assign zeros_initializer ones_initializer constant_initializer "
              "truncated_normal_initializer random_normal_initializer tanh sigmoid square sqrt ones_like zeros_like maximum add matmul squared_difference "
              "clip_by_value shape pad constant expand_dims squeeze reshape concat split reduce_mean reduce_sum random_normal clip_by_norm "
              "group GraphKeys").split():
        setattr(tf, n, getattr(me, n))
    tf.slice = slice_
    tf.float32 = "float32"
    tf.int32 = "int32"
    tf.nn = types.SimpleNamespace(relu=relu, dynamic_rnn=dynamic_rnn, dropout=dropout, moments=moments, l2_loss=l2_loss,
                                  conv2d=nn_conv2d, conv1d=nn_conv1d, conv2d_transpose=nn_conv2d_transpose,
                                  batch_normalization=batch_normalization, bias_add=bias_add, tanh=tanh, sigmoid=sigmoid)
    tf.train = types.SimpleNamespace(GradientDescentOptimizer=GradientDescentOptimizer, AdamOptimizer=AdamOptimizer,
                                     RMSPropOptimizer=RMSPropOptimizer, ExponentialMovingAverage=ExponentialMovingAverage,
                                     Saver=_Dummy, get_checkpoint_state=lambda *a, **k: None)
    def _summ(name, x, *a, **k):
        STATE.summaries.append((name, x))
    tf.summary = types.SimpleNamespace(scalar=_summ, histogram=_summ,
                                       tensor_summary=lambda *a, **k: None, audio=lambda *a, **k: None,
                                       merge=lambda *a, **k: None, FileWriter=_Dummy)
    tf.losses = types.SimpleNamespace(mean_squared_error=mean_squared_error)
    tf.logging = types.SimpleNamespace(WARN=30, log_first_n=lambda *a, **k: None, info=lambda *a, **k: None)
    tf.Session = Session
    tf.errors = types.SimpleNamespace(OutOfRangeError=type("OutOfRangeError", (Exception,), {}))
    layers = types.ModuleType("tensorflow.contrib.layers")
    for n in "fully_connected conv2d batch_norm xavier_initializer l2_regularizer flatten".split():
        setattr(layers, n, getattr(me, n))
    rnn = types.ModuleType("tensorflow.contrib.rnn")
    for n in "LSTMCell MultiRNNCell DropoutWrapper RNNCell LSTMStateTuple".split():
        setattr(rnn, n, getattr(me, n))
    contrib = types.ModuleType("tensorflow.contrib")
    slim = types.ModuleType("tensorflow.contrib.slim")          # utils/misc.py imports it; nothing on these paths calls it
    contrib.layers, contrib.rnn, contrib.slim = layers, rnn, slim
    tf.contrib = contrib
    import queue
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers,
                        "tensorflow.contrib.rnn": rnn, "tensorflow.contrib.slim": slim,
                        "Queue": queue})                          # the reference's scripts are Python 2: `import Queue`
    return tf, STATE
