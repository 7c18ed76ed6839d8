"""CPU oracle for the RSRGAN GAN-training hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy restatement of the reference math (wangkenpu/rsrgan) for
the path `models/gan_rnn_placeholder.py` drives.  It is the *checker*: only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / reference
leg may import it.  Nothing under `rsrgan_b200/` imports it and the product has
no CPU fallback.

PARITY PINNED TO THE REFERENCE'S SOURCE, NOT TO TENSORFLOW'S KERNELS.  TensorFlow 1.4.0 (unvendored, not
installable here: no python2, no `tensorflow`, no network) cannot be run and the reference ships no golden vectors
for this path.  What CAN be run is everything the reference itself wrote:
`tests/golden/make_reference_graph_golden.py` imports models/gan_rnn_placeholder.py, lstm.py, res_lstm_l.py,
res_lstm_base.py, discriminator_lstm.py, discriminator_dnn.py, dnn.py, rced.py, gan.py, dnn_trainer*.py, BNLSTMCell.py,
utils/ops.py and utils/bnorm.py from /root/reference and executes them over an eager float64 stand-in for the
TensorFlow calls they make (tests/golden/tf_standin.py); the fixtures it wrote (tests/golden/ref_graph_*.npz: variable
names and shapes, forward values, losses, per-tower gradients, averaged + clipped gradients, SGD / Adam / EMA results)
are what tests/test_reference_graph.py holds this oracle to, at 1e-9.  That pins the wiring, the formulas and the LSTM
step (against models/BNLSTMCell.py:176-213) to reference code; TensorFlow's own library layers and op kernels (contrib
fully_connected / conv2d / LSTMCell / dynamic_rnn / batch_norm(renorm), Adam) remain restatements, cross-checked by
(i) an independent torch-autograd float64 implementation of the same forward (`oracle/torch_ref.py`; must agree to
1e-9), (ii) finite-difference gradient checks, (iii) torch.nn.LSTM(proj_size) with peepholes zeroed and gates
re-ordered, torch conv / conv_transpose / batch_norm.  The Kaldi ark reader is pinned bit for bit to the reference's own
`io_funcs/kaldi_io.py`, which imports here (tests/golden/make_kaldi_golden.py).

Every function cites the reference file:line it follows (paths relative to the
reference root).  Default dtype is float64; pass float32 arrays to get a
float32 run of the same statement.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np

# --------------------------------------------------------------------------
# elementary pieces
# --------------------------------------------------------------------------


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


def leakyrelu(x, alpha=0.3):
    """utils/ops.py:120-121  tf.maximum(x, alpha*x)."""
    return np.maximum(x, alpha * x)


ACT_NONE, ACT_RELU, ACT_LRELU, ACT_CLIP = 0, 1, 2, 3


def act_fwd(u, act):
    if act == ACT_NONE:
        return u
    if act == ACT_RELU:
        return np.maximum(u, 0.0)
    if act == ACT_LRELU:
        return leakyrelu(u)
    if act == ACT_CLIP:  # discriminator_dnn.py:93 clip_by_value(y, -0.5, 1.5)
        return np.clip(u, -0.5, 1.5)
    raise ValueError(act)


def act_bwd(u, dy, act):
    if act == ACT_NONE:
        return dy
    if act == ACT_RELU:
        return dy * (u > 0)
    if act == ACT_LRELU:
        # maximum(u, .3u): slope 1 for u>0, .3 for u<0 (u==0: TF splits; measure-zero)
        return dy * np.where(u > 0, 1.0, 0.3)
    if act == ACT_CLIP:  # TF clip_by_value gradient: pass inside [lo, hi]
        return dy * ((u >= -0.5) & (u <= 1.5))
    raise ValueError(act)


def linear_fwd(x, W, b, act=ACT_NONE):
    """tf.contrib.layers.fully_connected over the last axis
    (lstm.py:82-87,121-124; discriminator_dnn.py:61-93; discriminator_lstm.py:100-104)."""
    u = x @ W + b
    return act_fwd(u, act), (x, W, u, act)


def linear_bwd(dy, cache):
    x, W, u, act = cache
    du = act_bwd(u, dy, act)
    x2 = x.reshape(-1, x.shape[-1])
    du2 = du.reshape(-1, du.shape[-1])
    dW = x2.T @ du2
    db = du2.sum(0)
    dx = du @ W.T
    return dx, dW, db


# --- contrib batch_norm(renorm=True, scale=True) and dropout behind a fully_connected ---------------------
# models/dnn.py:56-62,79-94, models/discriminator_dnn.py:36-46,61-83, models/lstm.py:61-67,82-87.
# TF-1.4 contrib batch_norm with renorm=True resolves to tf.layers.BatchNormalization(momentum=0.999,
# epsilon=1e-3, center=True, scale=True, renorm=True, renorm_clipping=None, renorm_momentum=0.99), non-fused;
# fully_connected drops its bias when a normalizer_fn is given.  (TF upstream behaviour, not visible in the
# reference tree -- an assumption like the others listed in the module docstring.)

BN_EPS, BN_DECAY, BN_RENORM_DECAY = 1e-3, 0.999, 0.99
BN_STATE_KEYS = ("moving_mean", "moving_variance", "renorm_mean", "renorm_stddev", "renorm_mean_weight",
                 "renorm_stddev_weight")


def bn_init_state(n, dtype=np.float64):
    """moving_mean 0, moving_variance 1; TF 1.4 zero-initialises all four renorm variables ("we initialize
    renorm_stddev to 0, and maintain the (0-initialized) renorm_stddev_weight"), so the first batch sees r = 1, d = 0."""
    return OrderedDict(moving_mean=np.zeros(n, dtype), moving_variance=np.ones(n, dtype),
                       renorm_mean=np.zeros(n, dtype), renorm_stddev=np.zeros(n, dtype),
                       renorm_mean_weight=np.zeros((), dtype), renorm_stddev_weight=np.zeros((), dtype))


def bn_renorm_train_fwd(z, gamma, beta, st, update=True, eps=BN_EPS, decay=BN_DECAY, renorm_decay=BN_RENORM_DECAY):
    """training=True branch of BatchNormalization.call + _renorm_correction_and_moments.  Moments over every
    axis but the last (nn.moments: biased variance).  r, d use the PRE-update renorm averages and are constants
    for the gradient (stop_gradient).  `st` is updated in place when update=True (the UPDATE_OPS)."""
    z2 = z.reshape(-1, z.shape[-1])
    mean = z2.mean(0)
    var = ((z2 - mean) ** 2).mean(0)
    stddev = np.sqrt(var + eps)
    mixed_mean = st["renorm_mean"] + (1.0 - st["renorm_mean_weight"]) * mean
    mixed_std = st["renorm_stddev"] + (1.0 - st["renorm_stddev_weight"]) * stddev
    r = stddev / mixed_std
    d = (mean - mixed_mean) / mixed_std
    xhat = (z - mean) / stddev
    y = (xhat * r + d) * gamma + beta
    if update:
        k = 1.0 - renorm_decay
        # assign_moving_average(var, value, decay, zero_debias=False): var -= (var - value) * (1 - decay)
        st["renorm_mean"] = st["renorm_mean"] - (st["renorm_mean"] - mean) * k
        st["renorm_mean_weight"] = st["renorm_mean_weight"] - (st["renorm_mean_weight"] - 1.0) * k
        st["renorm_stddev"] = st["renorm_stddev"] - (st["renorm_stddev"] - stddev) * k
        st["renorm_stddev_weight"] = st["renorm_stddev_weight"] - (st["renorm_stddev_weight"] - 1.0) * k
        new_mean = st["renorm_mean"] / st["renorm_mean_weight"]
        new_std = st["renorm_stddev"] / st["renorm_stddev_weight"]
        new_var = new_std ** 2 - eps
        st["moving_mean"] = st["moving_mean"] - (st["moving_mean"] - new_mean) * (1.0 - decay)
        st["moving_variance"] = st["moving_variance"] - (st["moving_variance"] - new_var) * (1.0 - decay)
    return y, (xhat, r, d, gamma, stddev)


def bn_renorm_train_bwd(dy, cache):
    xhat, r, d, gamma, stddev = cache
    dy2, xh2 = dy.reshape(-1, dy.shape[-1]), xhat.reshape(-1, xhat.shape[-1])
    s1, s2 = dy2.sum(0), (dy2 * xh2).sum(0)
    n = dy2.shape[0]
    dgamma = r * s2 + d * s1
    dbeta = s1
    dz = (gamma * r / stddev) * (dy - s1 / n - xhat * (s2 / n))
    return dz, dgamma, dbeta


def bn_eval_fwd(z, gamma, beta, st, eps=BN_EPS):
    """training=False: nn.batch_normalization(z, moving_mean, moving_variance, beta, gamma, eps)."""
    return (z - st["moving_mean"]) / np.sqrt(st["moving_variance"] + eps) * gamma + beta


_U64 = np.uint64


def _splitmix64(x):
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = x ^ (x >> _U64(30))
        x = x * _U64(0xbf58476d1ce4e5b9)
        x = x ^ (x >> _U64(27))
        x = x * _U64(0x94d049bb133111eb)
        x = x ^ (x >> _U64(31))
    return x


def dropout_mask(seed, tick, salt, rows, n, keep_prob):
    """The counter-based mask of rsr_affine_act_drop (include/rsrgan_b200.h): tf.nn.dropout keeps an element with
    probability keep_prob and divides kept values by keep_prob (models/dnn.py:116-121); TF's own random stream
    cannot be reproduced, so the mask generator is part of the C ABI and restated here bit-exactly.
    One splitmix64 hash per (even, odd) column pair: element (r, c) with flat index i = r n + c uses
    h = splitmix64(key ^ (i >> 1)), bits 63..40 for even c and bits 39..16 for odd c (n is even)."""
    assert n % 2 == 0
    with np.errstate(over="ignore"):
        key = _splitmix64(_U64(seed) + _U64(0x9E3779B97F4A7C15) * (_U64(tick) * _U64(65536) + _U64(salt)))
        pair = np.arange(rows * n // 2, dtype=np.uint64)
        hsh = _splitmix64(key ^ pair)
        bits = np.empty((rows * n // 2, 2), np.uint32)
        bits[:, 0] = (hsh >> _U64(40)).astype(np.uint32)
        bits[:, 1] = ((hsh >> _U64(16)) & _U64(0xffffff)).astype(np.uint32)
    # the C ABI receives keep_prob as a float32: the threshold is floor(float32(keep_prob) * 2^24)
    return bits.reshape(rows, n) < np.uint32(int(float(np.float32(keep_prob)) * 16777216.0))


def gauss_noise(seed, tick, salt, n, stddev):
    """rsr_gauss_noise (include/rsrgan_b200.h): the discriminator's input noise (utils/ops.py:19-30) drawn from the same
    counter-based stream as the dropout mask -- Box-Muller of the two 24-bit fields of splitmix64(key ^ i)."""
    with np.errstate(over="ignore"):
        key = _splitmix64(_U64(seed) + _U64(0x9E3779B97F4A7C15) * (_U64(tick) * _U64(65536) + _U64(salt)))
        hsh = _splitmix64(key ^ np.arange(n, dtype=np.uint64))
    u1 = ((hsh >> _U64(40)).astype(np.float64) + 1.0) / 16777216.0
    u2 = ((hsh >> _U64(16)) & _U64(0xffffff)).astype(np.float64) / 16777216.0
    return (stddev * np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)).astype(np.float32)


def fc_block_fwd(p, name, h, act, opts, salt):
    """One fully_connected call site with its optional normalizer and dropout.  opts (all optional):
    bn_state {name/BatchNorm/<key>}, train (True), update (False), keep_prob (1.0), rng (seed, tick)."""
    opts = opts or {}
    train = opts.get("train", True)
    W = p[name + "/weights"]
    bn = (name + "/BatchNorm/gamma") in p
    bc = None
    if bn:
        z = h @ W
        gamma, beta = p[name + "/BatchNorm/gamma"], p[name + "/BatchNorm/beta"]
        st = OrderedDict((k, opts["bn_state"][name + "/BatchNorm/" + k]) for k in BN_STATE_KEYS)
        if train:
            y, bc = bn_renorm_train_fwd(z, gamma, beta, st, update=opts.get("update", False))
            for k in BN_STATE_KEYS:
                opts["bn_state"][name + "/BatchNorm/" + k] = st[k]
        else:
            y = bn_eval_fwd(z, gamma, beta, st)
    else:
        y = h @ W + p[name + "/biases"]
    a = act_fwd(y, act)
    keep = opts.get("keep_prob", 1.0) if train else 1.0
    mask = None
    if keep < 1.0:
        seed, tick = opts["rng"]
        rows, n = int(np.prod(a.shape[:-1])), a.shape[-1]
        # the library's rows are the flattened leading axes IN THE ORDER GIVEN (feed time-major data to match its
        # t*B+b rows) and its row pitch is the width padded to a multiple of 8
        mask = dropout_mask(seed, tick, salt, rows, -(-n // 8) * 8, keep)[:, :n].reshape(a.shape)
        a = np.where(mask, a / keep, 0.0)
    return a, (h, W, y, act, bc, mask, keep)


def seq_dropout_fwd(a, opts, salt):
    """tf.contrib.rnn.DropoutWrapper(cell, output_keep_prob=keep_prob) (models/lstm.py:99-102, models/res_lstm_l.py:
    96-99): the layer OUTPUT handed to the next layer is dropped, the recurrent state is not.  a: (B, T, P) batch-major;
    the mask is drawn in the library's layout -- time-major rows t*B+b, row pitch P padded to a multiple of 8."""
    opts = opts or {}
    keep = opts.get("keep_prob", 1.0) if opts.get("train", True) else 1.0
    if keep >= 1.0:
        return a, None
    B, T, P = a.shape
    seed, tick = opts["rng"]
    Pp = -(-P // 8) * 8
    m = dropout_mask(seed, tick, salt, T * B, Pp, keep).reshape(T, B, Pp)[:, :, :P].transpose(1, 0, 2)
    return np.where(m, a / keep, 0.0), (m, keep)


def seq_dropout_bwd(da, cache):
    if cache is None:
        return da
    m, keep = cache
    return np.where(m, da / keep, 0.0)


LSTM_SALT = 16       # dropout stream of LSTM layer l of a generator: salt0 + LSTM_SALT + l


def fc_block_bwd(da, cache, name, g):
    h, W, y, act, bc, mask, keep = cache
    if mask is not None:
        da = np.where(mask, da / keep, 0.0)
    dy = act_bwd(y, da, act)
    if bc is not None:
        dz, g[name + "/BatchNorm/gamma"], g[name + "/BatchNorm/beta"] = bn_renorm_train_bwd(dy, bc)
    else:
        dz = dy
        g[name + "/biases"] = dz.reshape(-1, dz.shape[-1]).sum(0)
    g[name + "/weights"] = h.reshape(-1, h.shape[-1]).T @ dz.reshape(-1, dz.shape[-1])
    return dz @ W.T


def conv1d_same_fwd(x, W, b, act=ACT_RELU):
    """tf.contrib.layers.conv2d(inputs, C_out, [splice, w], padding=SAME, relu) with splice = 1
    (models/rced.py:90-101): inputs NHWC (N, 1, L, C_in) -> here (N, L, C_in); W is the TF filter
    (1, w, C_in, C_out); stride 1; w odd, so SAME pads w//2 zeros on both sides.
        u[n, p, :] = b + sum_k x[n, p - w//2 + k, :] @ W[0, k]"""
    assert W.shape[0] == 1 and W.shape[1] % 2 == 1, "splice = 1, odd filter width"
    N, L, _ = x.shape
    w = W.shape[1]
    xp = np.zeros((N, L + w - 1, x.shape[2]), x.dtype)
    xp[:, w // 2:w // 2 + L] = x
    u = np.zeros((N, L, W.shape[3]), np.result_type(x, W)) + b
    for k in range(w):
        u += xp[:, k:k + L] @ W[0, k]
    return act_fwd(u, act), (xp, W, u, act)


def conv1d_same_bwd(dy, cache):
    xp, W, u, act = cache
    N, L, _ = u.shape
    w = W.shape[1]
    du = act_bwd(u, dy, act)
    dW = np.zeros_like(W)
    dxp = np.zeros_like(xp)
    for k in range(w):
        dW[0, k] = np.einsum("nlc,nld->cd", xp[:, k:k + L], du)
        dxp[:, k:k + L] += du @ W[0, k].T
    return dxp[:, w // 2:w // 2 + L], dW, du.sum((0, 1))


# --------------------------------------------------------------------------
# the 1-D convolution family of utils/ops.py (strided downconv, transposed deconv) and virtual batch norm
# (utils/bnorm.py): consumers models/discriminator.py:38-90 (SEGAN-style waveform discriminator)
# --------------------------------------------------------------------------
def same_pad(L, k, stride):
    """TensorFlow SAME padding along one axis: (output length, pad before, pad after)."""
    out = -(-L // stride)
    total = max((out - 1) * stride + k - L, 0)
    return out, total // 2, total - total // 2


def downconv_fwd(x, W, b=None, pool=2):
    """utils/ops.py:78-98 `downconv`: tf.nn.conv2d(x[:, :, None, :], W[k, 1, C_in, C_out], strides=[1, pool, 1, 1],
    padding='SAME') (+ bias_add), reshaped back to (B, ceil(L / pool), C_out).  x (B, L, C_in); W (k, C_in, C_out)
    (the singleton filter axis dropped)."""
    B, L, _ = x.shape
    k = W.shape[0]
    out, pl, pr = same_pad(L, k, pool)
    xp = np.pad(x, ((0, 0), (pl, pr), (0, 0)))
    cols = np.stack([xp[:, j:j + pool * out:pool] for j in range(k)], 2)           # (B, out, k, C_in)
    y = np.einsum("bokc,kcd->bod", cols, W)
    if b is not None:
        y = y + b
    return y, (cols, W, x.shape, pool, pl, b is not None)


def downconv_bwd(dy, cache):
    cols, W, xshape, pool, pl, has_b = cache
    B, L, C = xshape
    k, out = W.shape[0], dy.shape[1]
    dW = np.einsum("bokc,bod->kcd", cols, dy)
    dcols = np.einsum("bod,kcd->bokc", dy, W)
    dxp = np.zeros((B, max((out - 1) * pool + k, L + pl), C), dy.dtype)
    for j in range(k):
        dxp[:, j:j + pool * out:pool] += dcols[:, :, j]
    return dxp[:, pl:pl + L], dW, (dy.sum((0, 1)) if has_b else None)


def deconv_fwd(x, W, b=None, dilation=2):
    """utils/ops.py:277-310 `deconv`: tf.nn.conv2d_transpose(x[:, :, None, :], W[k, 1, C_out, C_in], output_shape =
    (B, dilation * L, 1, C_out), strides=[1, dilation, 1, 1]) (padding defaults to SAME) (+ bias): the gradient of the
    SAME strided convolution C_out -> C_in with that filter.  x (B, L, C_in); W (k, C_out, C_in) -> (B, dilation*L, C_out)."""
    B, L, _ = x.shape
    k = W.shape[0]
    Lo = dilation * L
    _, pl, _ = same_pad(Lo, k, dilation)
    yp = np.zeros((B, (L - 1) * dilation + k, W.shape[1]), x.dtype)
    contrib = np.einsum("boi,kci->bokc", x, W)                                     # (B, L, k, C_out)
    for j in range(k):
        yp[:, j:j + dilation * L:dilation] += contrib[:, :, j]
    y = yp[:, pl:pl + Lo]
    if y.shape[1] < Lo:                                                            # k < dilation never happens here
        y = np.pad(y, ((0, 0), (0, Lo - y.shape[1]), (0, 0)))
    if b is not None:
        y = y + b
    return y, (x, W, dilation, pl, b is not None)


def deconv_bwd(dy, cache):
    x, W, dilation, pl, has_b = cache
    B, L, _ = x.shape
    k = W.shape[0]
    dyp = np.zeros((B, (L - 1) * dilation + k + pl, dy.shape[2]), dy.dtype)
    dyp[:, pl:pl + dy.shape[1]] = dy
    cols = np.stack([dyp[:, j:j + dilation * L:dilation] for j in range(k)], 2)    # (B, L, k, C_out)
    dx = np.einsum("bokc,kci->boi", cols, W)
    dW = np.einsum("bokc,boi->kci", cols, x)
    return dx, dW, (dy.sum((0, 1)) if has_b else None)


def vbn_reference(x_ref, eps=1e-5):
    """utils/bnorm.py:17-37: statistics of the reference batch, (mean, mean of squares) per channel over (batch, time)."""
    return x_ref.mean((0, 1)), (x_ref ** 2).mean((0, 1)), x_ref.shape[0]


def vbn_fwd(x, gamma, beta, ref=None, eps=1e-5):
    """utils/bnorm.py:39-69.  ref None: the reference pass itself (statistics of x, :31-35); else ref = (mean, mean_sq,
    batch_size) of the reference batch and the live statistics are blended with weight 1 / (batch_size + 1) (:42-49)."""
    m_b, q_b = x.mean((0, 1)), (x ** 2).mean((0, 1))
    if ref is None:
        a, m, q = 1.0, m_b, q_b
    else:
        a = 1.0 / (ref[2] + 1.0)
        m, q = a * m_b + (1.0 - a) * ref[0], a * q_b + (1.0 - a) * ref[1]
    std = np.sqrt(eps + q - m ** 2)
    xhat = (x - m) / std
    return xhat * gamma + beta, (xhat, std, gamma, a)


def vbn_bwd(dy, cache):
    """d/dx through the batch statistics (weight a of the live batch), d/dgamma, d/dbeta."""
    xhat, std, gamma, a = cache
    n = dy.shape[0] * dy.shape[1]
    s1, s2 = dy.sum((0, 1)), (dy * xhat).sum((0, 1))
    dx = (gamma / std) * (dy - a * s1 / n - a * xhat * (s2 / n))
    return dx, s2, s1


def conv2d_same_fwd(x, W, b, act=ACT_RELU):
    """tf.contrib.layers.conv2d(inputs, C_out, [splice, w], padding=SAME, relu) for ANY splice (models/rced.py:90-101):
    x NHWC (N, H, L, C_in), W (kh, kw, C_in, C_out), stride 1, odd kh and kw (SAME pads kh//2 / kw//2 zeros per side).
        u[n, h, p, :] = b + sum_{i, k} x[n, h - kh//2 + i, p - kw//2 + k, :] @ W[i, k]
    The product path builds only splice = 1 (conv1d_same_fwd); this is the statement the 2-D case will be held to."""
    kh, kw = W.shape[:2]
    assert kh % 2 == 1 and kw % 2 == 1
    N, H, L, _ = x.shape
    xp = np.zeros((N, H + kh - 1, L + kw - 1, x.shape[3]), x.dtype)
    xp[:, kh // 2:kh // 2 + H, kw // 2:kw // 2 + L] = x
    u = np.zeros((N, H, L, W.shape[3]), np.result_type(x, W)) + b
    for i in range(kh):
        for k in range(kw):
            u += xp[:, i:i + H, k:k + L] @ W[i, k]
    return act_fwd(u, act), (xp, W, u, act)


def conv2d_same_bwd(dy, cache):
    xp, W, u, act = cache
    N, H, L, _ = u.shape
    kh, kw = W.shape[:2]
    du = act_bwd(u, dy, act)
    dW, dxp = np.zeros_like(W), np.zeros_like(xp)
    for i in range(kh):
        for k in range(kw):
            dW[i, k] = np.einsum("nhlc,nhld->cd", xp[:, i:i + H, k:k + L], du)
            dxp[:, i:i + H, k:k + L] += du @ W[i, k].T
    return dxp[:, kh // 2:kh // 2 + H, kw // 2:kw // 2 + L], dW, du.sum((0, 1, 2))


def toeplitz_taps(W, H):
    """The [kh, kw] SAME convolution over H stacked lines as a 1-D convolution over positions whose channels are
    (line, channel) pairs: W2[0, k, h_in * C_in + ci, h_out * C_out + co] = W[h_in - h_out + kh//2, k, ci, co] (zero where
    that row index falls outside the filter).  DESIGN.md section 9 item 4: how the overlapped-view GEMM will run the
    2-D RCED; the compact filter stays the parameter, the expansion is a derived operand."""
    kh, kw, ci, co = W.shape
    W2 = np.zeros((1, kw, H * ci, H * co), W.dtype)
    for h_out in range(H):
        for h_in in range(H):
            i = h_in - h_out + kh // 2
            if 0 <= i < kh:
                W2[0, :, h_in * ci:(h_in + 1) * ci, h_out * co:(h_out + 1) * co] = W[i]
    return W2


def toeplitz_fold_grad(dW2, kh, ci, co):
    """Gradient of the compact filter from the gradient of its Toeplitz expansion: the sum over the tied copies."""
    H = dW2.shape[2] // ci
    dW = np.zeros((kh, dW2.shape[1], ci, co), dW2.dtype)
    for h_out in range(H):
        for h_in in range(H):
            i = h_in - h_out + kh // 2
            if 0 <= i < kh:
                dW[i] += dW2[0, :, h_in * ci:(h_in + 1) * ci, h_out * co:(h_out + 1) * co]
    return dW


# --------------------------------------------------------------------------
# LSTMP cell with peepholes == tf.contrib.rnn.LSTMCell(use_peepholes=True,
# num_proj=P, forget_bias=1.0) under tf.nn.dynamic_rnn(sequence_length=...)
# call sites lstm.py:89-112, res_lstm_l.py:86-138, discriminator_lstm.py:70-91;
# in-repo statement of the same gate math: models/BNLSTMCell.py:160-216
# (gate order i, j, f, o at :176-179; peepholes/forget_bias at :191-192,203;
# bias-free projection at :207-213).
# --------------------------------------------------------------------------


def lstmp_fwd(x, lengths, K, b, w_i, w_f, w_o, W_p, forget_bias=1.0):
    """x (B,T,I) batch-major; K ((I+P),4C) with rows [inputs ; m_prev];
    returns out (B,T,P) (zero past lengths[b]) and a cache for lstmp_bwd."""
    B, T, I = x.shape
    C = w_i.shape[0]
    P = W_p.shape[1]
    dt = x.dtype
    c = np.zeros((B, C), dt)
    m = np.zeros((B, P), dt)
    out = np.zeros((B, T, P), dt)
    steps = []
    lengths = np.asarray(lengths).astype(np.int64)
    for t in range(T):
        act = (t < lengths)[:, None]
        xin = np.concatenate([x[:, t], m], axis=1)
        z = xin @ K + b
        zi, zj, zf, zo = z[:, :C], z[:, C:2 * C], z[:, 2 * C:3 * C], z[:, 3 * C:]
        ig = sigmoid(zi + w_i * c)
        fg = sigmoid(zf + forget_bias + w_f * c)
        jg = np.tanh(zj)
        c_new = fg * c + ig * jg
        og = sigmoid(zo + w_o * c_new)
        tc = np.tanh(c_new)
        mt = og * tc
        m_new = mt @ W_p
        steps.append((xin, c, ig, fg, jg, og, tc, c_new, mt, act))
        out[:, t] = np.where(act, m_new, 0.0)
        c = np.where(act, c_new, c)
        m = np.where(act, m_new, m)
    cache = (x.shape, K, w_i, w_f, w_o, W_p, steps)
    return out, cache


def lstmp_bwd(dout, cache):
    (B, T, I), K, w_i, w_f, w_o, W_p, steps = cache
    C = w_i.shape[0]
    P = W_p.shape[1]
    dt = dout.dtype
    dx = np.zeros((B, T, I), dt)
    dK = np.zeros_like(K)
    db = np.zeros(4 * C, dt)
    dw_i = np.zeros(C, dt)
    dw_f = np.zeros(C, dt)
    dw_o = np.zeros(C, dt)
    dW_p = np.zeros_like(W_p)
    dm_next = np.zeros((B, P), dt)
    dc_next = np.zeros((B, C), dt)
    for t in range(T - 1, -1, -1):
        xin, c_prev, ig, fg, jg, og, tc, c_new, mt, act = steps[t]
        a = act.astype(dt)
        dm = (dout[:, t] + dm_next) * a          # active rows only
        dmt = dm @ W_p.T
        dW_p += mt.T @ dm
        do_pre = dmt * tc * og * (1 - og)
        dc = (dc_next * a) + dmt * og * (1 - tc * tc) + do_pre * w_o
        dw_o += (do_pre * c_new).sum(0)
        df_pre = dc * c_prev * fg * (1 - fg)
        di_pre = dc * jg * ig * (1 - ig)
        dzj = dc * ig * (1 - jg * jg)
        dc_prev = dc * fg + df_pre * w_f + di_pre * w_i
        dw_f += (df_pre * c_prev).sum(0)
        dw_i += (di_pre * c_prev).sum(0)
        dz = np.concatenate([di_pre, dzj, df_pre, do_pre], axis=1)
        db += dz.sum(0)
        dK += xin.T @ dz
        dxin = dz @ K.T
        dx[:, t] = dxin[:, :I]
        # frozen rows pass their state gradient straight through
        dm_next = dxin[:, I:] + dm_next * (1 - a)
        dc_next = dc_prev + dc_next * (1 - a)
    return dx, dict(kernel=dK, bias=db, w_i_diag=dw_i, w_f_diag=dw_f,
                    w_o_diag=dw_o, proj=dW_p)


# --------------------------------------------------------------------------
# parameter containers (names follow SURVEY.md App. B == TF-1.4 variable names)
# --------------------------------------------------------------------------


def _cell_names(prefix):
    return [prefix + s for s in ("kernel", "bias", "w_f_diag", "w_i_diag",
                                 "w_o_diag", "projection/kernel")]


def xavier(rng, shape, dtype=np.float64):
    """tf.contrib.layers.xavier_initializer(): U(+-sqrt(6/(fan_in+fan_out)))."""
    if len(shape) == 1:
        fan_in = fan_out = shape[0]
    elif len(shape) == 4:      # conv filter (h, w, C_in, C_out): receptive field x channels
        fan_in, fan_out = shape[0] * shape[1] * shape[2], shape[0] * shape[1] * shape[3]
    else:
        fan_in, fan_out = shape[0], shape[1]
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return rng.uniform(-lim, lim, size=shape).astype(dtype)


def _init_cell(p, rng, prefix, I, C, P, dtype):
    p[prefix + "kernel"] = xavier(rng, (I + P, 4 * C), dtype)
    p[prefix + "bias"] = np.zeros(4 * C, dtype)
    p[prefix + "w_f_diag"] = xavier(rng, (C,), dtype)
    p[prefix + "w_i_diag"] = xavier(rng, (C,), dtype)
    p[prefix + "w_o_diag"] = xavier(rng, (C,), dtype)
    p[prefix + "projection/kernel"] = xavier(rng, (C, P), dtype)


def init_g_lstm(rng, in_dim=257, out_dim=40, cell=760, proj=280, layers=3,
                dtype=np.float64):
    """models/lstm.py:43-45,82-124 (ref-native sizes are the defaults)."""
    p = OrderedDict()
    p["g_model/fully_connected/weights"] = xavier(rng, (in_dim, proj), dtype)
    p["g_model/fully_connected/biases"] = np.zeros(proj, dtype)
    for l in range(layers):
        _init_cell(p, rng, "g_model/rnn/multi_rnn_cell/cell_%d/lstm_cell/" % l,
                   proj, cell, proj, dtype)
    p["g_model/fully_connected_1/weights"] = xavier(rng, (proj, out_dim), dtype)
    p["g_model/fully_connected_1/biases"] = np.zeros(out_dim, dtype)
    return p


def init_g_res_lstm_l(rng, in_dim=257, out_dim=40, cell=760, layers=4,
                      dtype=np.float64):
    """models/res_lstm_l.py:43-45,101-138,187-194 (proj == in_dim == 257)."""
    p = OrderedDict()
    for l in range(1, layers + 1):
        _init_cell(p, rng, "g_model/lstm_cell_%d/rnn/lstm_cell/" % l,
                   in_dim, cell, in_dim, dtype)
    p["g_model/forward_out/fully_connected/weights"] = xavier(rng, (in_dim, out_dim), dtype)
    p["g_model/forward_out/fully_connected/biases"] = np.zeros(out_dim, dtype)
    return p


def _bn_vars(p, name, n, dtype):
    """beta zeros, gamma ones (contrib batch_norm defaults with scale=True); replaces the layer's bias."""
    p[name + "/BatchNorm/beta"] = np.zeros(n, dtype)
    p[name + "/BatchNorm/gamma"] = np.ones(n, dtype)


def init_bn_state(p, dtype=np.float64):
    """Non-trainable batch_norm variables of every normalised layer in p, keyed by their TF names."""
    st = OrderedDict()
    for k in p:
        if k.endswith("/BatchNorm/gamma"):
            for kk, v in bn_init_state(p[k].shape[0], dtype).items():
                st[k[:-len("gamma")] + kk] = v
    return st


def init_g_dnn(rng, in_dim=257, out_dim=40, units=1024, hidden=3, dtype=np.float64, batch_norm=False):
    """models/dnn.py:34-35,79-110: in -> units x (1 + hidden) ReLU -> out, xavier weights, zero biases
    (hidden layers: BatchNorm beta / gamma instead of biases when batch_norm, :56-62)."""
    p = OrderedDict()
    dims = [in_dim] + [units] * (hidden + 1) + [out_dim]
    for l in range(hidden + 2):
        name = "g_model/fully_connected" + ("" if l == 0 else "_%d" % l)
        p[name + "/weights"] = xavier(rng, (dims[l], dims[l + 1]), dtype)
        if batch_norm and l < hidden + 1:
            _bn_vars(p, name, dims[l + 1], dtype)
        else:
            p[name + "/biases"] = np.zeros(dims[l + 1], dtype)
    return p


RCED_FILTERS = (12, 16, 20, 24, 32, 24, 20, 16, 12)      # models/rced.py:92
RCED_WIDTHS = (13, 11, 9, 7, 7, 7, 9, 11, 13)            # models/rced.py:93


def init_g_rced(rng, in_dim=257, out_dim=40, filters=RCED_FILTERS, widths=RCED_WIDTHS, dtype=np.float64, splice=1,
                batch_norm=False):
    """models/rced.py:90-114: nine conv2d [splice, w] (xavier, zero bias; contrib default scopes Conv, Conv_1, ...), then
    FC (splice * in_dim * filters[-1]) -> out_dim with bias 0.1 (:108-113).  batch_norm (:63-71,97): the convolutions
    have BatchNorm beta / gamma instead of biases, the output layer keeps its bias."""
    p = OrderedDict()
    cin = 1
    for l, (c, w) in enumerate(zip(filters, widths)):
        name = "g_model/Conv" + ("" if l == 0 else "_%d" % l)
        p[name + "/weights"] = xavier(rng, (splice, w, cin, c), dtype)
        if batch_norm:
            _bn_vars(p, name, c, dtype)
        else:
            p[name + "/biases"] = np.zeros(c, dtype)
        cin = c
    p["g_model/fully_connected/weights"] = xavier(rng, (splice * in_dim * cin, out_dim), dtype)
    p["g_model/fully_connected/biases"] = np.full(out_dim, 0.1, dtype)
    return p


def init_d_lstm(rng, in_dim=40, cell=256, proj=40, layers=2, dtype=np.float64):
    """models/discriminator_lstm.py:26-28,70-104."""
    p = OrderedDict()
    for l in range(layers):
        _init_cell(p, rng, "d_model/rnn/multi_rnn_cell/cell_%d/lstm_cell/" % l,
                   in_dim if l == 0 else proj, cell, proj, dtype)
    p["d_model/fully_connected/weights"] = xavier(rng, (proj, 1), dtype)
    p["d_model/fully_connected/biases"] = np.zeros(1, dtype)
    return p


def init_d_dnn(rng, in_dim=40, units=1024, hidden=3, dtype=np.float64, batch_norm=False):
    """models/discriminator_dnn.py:23-27,61-93: truncN(0, sqrt(2/units)) hidden
    weights (truncation at 2 sigma), xavier output layer, zero biases."""
    p = OrderedDict()
    std = math.sqrt(2.0 / units)

    def truncn(shape):
        v = rng.standard_normal(size=shape)
        bad = np.abs(v) > 2
        while bad.any():
            v[bad] = rng.standard_normal(size=int(bad.sum()))
            bad = np.abs(v) > 2
        return (v * std).astype(dtype)

    dims = [in_dim] + [units] * (hidden + 1)
    for l in range(hidden + 1):
        name = "d_model/fully_connected" + ("" if l == 0 else "_%d" % l)
        p[name + "/weights"] = truncn((dims[l], dims[l + 1]))
        if batch_norm:
            _bn_vars(p, name, dims[l + 1], dtype)
        else:
            p[name + "/biases"] = np.zeros(dims[l + 1], dtype)
    name = "d_model/fully_connected_%d" % (hidden + 1)
    p[name + "/weights"] = xavier(rng, (units, 1), dtype)
    p[name + "/biases"] = np.zeros(1, dtype)
    return p


# --------------------------------------------------------------------------
# networks: forward returns (y, cache); backward returns (dx, grads-dict)
# --------------------------------------------------------------------------


def _cell_params(p, prefix):
    return (p[prefix + "kernel"], p[prefix + "bias"], p[prefix + "w_i_diag"],
            p[prefix + "w_f_diag"], p[prefix + "w_o_diag"], p[prefix + "projection/kernel"])


def _cell_grads(g, prefix, cg):
    g[prefix + "kernel"] = cg["kernel"]
    g[prefix + "bias"] = cg["bias"]
    g[prefix + "w_f_diag"] = cg["w_f_diag"]
    g[prefix + "w_i_diag"] = cg["w_i_diag"]
    g[prefix + "w_o_diag"] = cg["w_o_diag"]
    g[prefix + "projection/kernel"] = cg["proj"]


def _cell_prefixes(p, scope):
    pre = sorted({k[:-len("kernel")] for k in p
                  if k.startswith(scope) and k.endswith("lstm_cell/kernel")})
    return pre


def g_lstm_fwd(p, x, lengths, opts=None, salt0=0):
    """models/lstm.py:82-124: FC(leakyrelu .3) [batch_norm(renorm) when built with it, :61-67,84-85] -> stacked LSTMP
    [DropoutWrapper on every layer's output when training with keep_prob < 1, :99-102] -> FC(linear).
    MultiRNNCell per-timestep stacking == layer-after-layer over the sequence."""
    caches = []
    h, c0 = fc_block_fwd(p, "g_model/fully_connected", x, ACT_LRELU, dict(opts or {}, keep_prob=1.0), salt0)
    caches.append(c0)
    for l, pre in enumerate(_cell_prefixes(p, "g_model/rnn/")):
        h, cc = lstmp_fwd(h, lengths, *_cell_params(p, pre))
        h, dc = seq_dropout_fwd(h, opts, salt0 + LSTM_SALT + l)
        caches.append((cc, dc))
    y, c1 = linear_fwd(h, p["g_model/fully_connected_1/weights"],
                       p["g_model/fully_connected_1/biases"], ACT_NONE)
    caches.append(c1)
    return y, caches


def g_lstm_bwd(p, dy, caches):
    g = OrderedDict()
    dh, dW, db = linear_bwd(dy, caches[-1])
    g["g_model/fully_connected_1/weights"] = dW
    g["g_model/fully_connected_1/biases"] = db
    pres = _cell_prefixes(p, "g_model/rnn/")
    for l in range(len(pres) - 1, -1, -1):
        cc, dc = caches[1 + l]
        dh, cg = lstmp_bwd(seq_dropout_bwd(dh, dc), cc)
        _cell_grads(g, pres[l], cg)
    dx = fc_block_bwd(dh, caches[0], "g_model/fully_connected", g)
    return dx, g


def g_res_lstm_l_fwd(p, x, lengths, residual=True, opts=None, salt0=0):
    """models/res_lstm_l.py:101-138,187-194: x_{l+1} = LSTMP_l(x_l) + x_l,
    y = FC(out_L + x_L).  residual=False is models/res_lstm_base.py:111-131,190.  Each of the four cells is wrapped
    in DropoutWrapper(output_keep_prob) when training with keep_prob < 1 (:96-99): the dropped output enters the sum."""
    caches = []
    xin = x
    pres = _cell_prefixes(p, "g_model/lstm_cell_")
    for l, pre in enumerate(pres):
        o, cc = lstmp_fwd(xin, lengths, *_cell_params(p, pre))
        o, dc = seq_dropout_fwd(o, opts, salt0 + LSTM_SALT + l)
        caches.append((cc, dc))
        xin = o + xin if residual else o
    y, c1 = linear_fwd(xin, p["g_model/forward_out/fully_connected/weights"],
                       p["g_model/forward_out/fully_connected/biases"], ACT_NONE)
    caches.append(c1)
    return y, caches


def g_res_lstm_l_bwd(p, dy, caches, residual=True):
    g = OrderedDict()
    dxin, dW, db = linear_bwd(dy, caches[-1])
    g["g_model/forward_out/fully_connected/weights"] = dW
    g["g_model/forward_out/fully_connected/biases"] = db
    pres = _cell_prefixes(p, "g_model/lstm_cell_")
    for l in range(len(pres) - 1, -1, -1):
        cc, dc = caches[l]
        dprev, cg = lstmp_bwd(seq_dropout_bwd(dxin, dc), cc)
        _cell_grads(g, pres[l], cg)
        dxin = dprev + dxin if residual else dprev
    return dxin, g


def d_lstm_fwd(p, x, lengths, noise=None):
    """models/discriminator_lstm.py:60-104.  `noise` is the (B,1,D) draw of
    utils/ops.py:19-30 (already scaled by disc_noise_std); None == std 0."""
    caches = []
    h = x if noise is None else x + noise
    for pre in _cell_prefixes(p, "d_model/rnn/"):
        h, cc = lstmp_fwd(h, lengths, *_cell_params(p, pre))
        caches.append(cc)
    y, c1 = linear_fwd(h, p["d_model/fully_connected/weights"],
                       p["d_model/fully_connected/biases"], ACT_NONE)
    caches.append(c1)
    return y, caches


def d_lstm_bwd(p, dy, caches):
    g = OrderedDict()
    dh, dW, db = linear_bwd(dy, caches[-1])
    g["d_model/fully_connected/weights"] = dW
    g["d_model/fully_connected/biases"] = db
    pres = _cell_prefixes(p, "d_model/rnn/")
    for l in range(len(pres) - 1, -1, -1):
        dh, cg = lstmp_bwd(dh, caches[l])
        _cell_grads(g, pres[l], cg)
    return dh, g


def _dnn_names(p):
    names = sorted({k.rsplit("/", 1)[0] for k in p if k.startswith("d_model/fully_connected") and k.endswith("/weights")},
                   key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)
    return names


def d_dnn_fwd(p, x, lengths=None, noise=None, opts=None, salt0=256):
    """models/discriminator_dnn.py:61-93 applied per frame (fully_connected
    broadcasts over leading dims; SURVEY App. C-15 adapter: lengths ignored).  opts: see fc_block_fwd."""
    caches = []
    h = x
    names = _dnn_names(p)
    for i, n in enumerate(names[:-1]):
        h, c = fc_block_fwd(p, n, h, ACT_RELU, opts, salt0 + i)
        caches.append(c)
    y, c = linear_fwd(h, p[names[-1] + "/weights"], p[names[-1] + "/biases"], ACT_CLIP)
    caches.append(c)
    return y, caches


def d_dnn_bwd(p, dy, caches):
    g = OrderedDict()
    names = _dnn_names(p)
    dh, dW, db = linear_bwd(dy, caches[-1])
    g[names[-1] + "/weights"], g[names[-1] + "/biases"] = dW, db
    for n, c in zip(reversed(names[:-1]), reversed(caches[:-1])):
        dh = fc_block_bwd(dh, c, n, g)
    return dh, g


def _fc_names(p, scope):
    names = sorted({k.rsplit("/", 1)[0] for k in p if k.startswith(scope + "/fully_connected") and k.endswith("/weights")},
                   key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)
    return names


def g_dnn_fwd(p, x, lengths=None, opts=None, salt0=0):
    """models/dnn.py:79-110 applied per frame (no sequence dependence; lengths unused).  opts: see fc_block_fwd."""
    caches = []
    h = x
    names = _fc_names(p, "g_model")
    for i, n in enumerate(names[:-1]):
        h, c = fc_block_fwd(p, n, h, ACT_RELU, opts, salt0 + i)
        caches.append(c)
    y, c = linear_fwd(h, p[names[-1] + "/weights"], p[names[-1] + "/biases"], ACT_NONE)
    caches.append(c)
    return y, caches


def g_dnn_bwd(p, dy, caches):
    g = OrderedDict()
    names = _fc_names(p, "g_model")
    dh, dW, db = linear_bwd(dy, caches[-1])
    g[names[-1] + "/weights"], g[names[-1] + "/biases"] = dW, db
    for n, c in zip(reversed(names[:-1]), reversed(caches[:-1])):
        dh = fc_block_bwd(dh, c, n, g)
    return dh, g


def _conv_names(p):
    return sorted({k.rsplit("/", 1)[0] for k in p if k.startswith("g_model/Conv") and k.endswith("/weights")},
                  key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)


def g_rced_fwd(p, x, lengths=None, opts=None, salt0=0):
    """models/rced.py:46-57,90-114: every frame (leading dims flattened) is `splice` stacked lines of `input_dim` bins
    with one channel, NHWC (N, splice, input_dim, 1) (:46-57; splice = the filter height of the first convolution, 1 in
    BASELINE configs[3], 11 in run_dnn.sh:129-140); nine ReLU conv2d [splice, w]; the NHWC tensor is flattened
    line-major / position / channel-minor (tf.reshape, :106) into the linear output layer.
    With BatchNorm variables in p (:63-71,97): conv2d without bias -> batch_norm(renorm) over (N, H, W) per channel ->
    relu; opts as in fc_block_fwd (bn_state, train, update)."""
    opts = opts or {}
    train = opts.get("train", True)
    names = _conv_names(p)
    H = p[names[0] + "/weights"].shape[0]
    lead, L = x.shape[:-1], x.shape[-1] // H
    h = x.reshape(-1, H, L, 1)
    caches = []
    for n in names:
        bn = (n + "/BatchNorm/gamma") in p
        W = p[n + "/weights"]
        b = np.zeros(W.shape[-1], W.dtype) if bn else p[n + "/biases"]
        act = ACT_NONE if bn else ACT_RELU
        if H == 1:
            h1, c = conv1d_same_fwd(h[:, 0], W, b, act)
            h = h1[:, None]
        else:
            h, c = conv2d_same_fwd(h, W, b, act)
        bc = None
        if bn:
            gamma, beta = p[n + "/BatchNorm/gamma"], p[n + "/BatchNorm/beta"]
            st = OrderedDict((k, opts["bn_state"][n + "/BatchNorm/" + k]) for k in BN_STATE_KEYS)
            if train:
                y, bc = bn_renorm_train_fwd(h, gamma, beta, st, update=opts.get("update", False))
                for k in BN_STATE_KEYS:
                    opts["bn_state"][n + "/BatchNorm/" + k] = st[k]
            else:
                y = bn_eval_fwd(h, gamma, beta, st)
            h = act_fwd(y, ACT_RELU)
            bc = (bc, y)
        caches.append((c, bc))
    flat = h.reshape(h.shape[0], -1)
    y, c = linear_fwd(flat, p["g_model/fully_connected/weights"], p["g_model/fully_connected/biases"], ACT_NONE)
    caches.append((c, h.shape))
    return y.reshape(*lead, -1), caches


def g_rced_bwd(p, dy, caches):
    g = OrderedDict()
    c, hshape = caches[-1]
    H = hshape[1]
    dh, dW, db = linear_bwd(dy.reshape(-1, dy.shape[-1]), c)
    g["g_model/fully_connected/weights"] = dW
    g["g_model/fully_connected/biases"] = db
    dh = dh.reshape(hshape)
    for n, (cc, bc) in zip(reversed(_conv_names(p)), reversed(caches[:-1])):
        if bc is not None:
            bn_cache, y = bc
            dh, g[n + "/BatchNorm/gamma"], g[n + "/BatchNorm/beta"] = bn_renorm_train_bwd(act_bwd(y, dh, ACT_RELU), bn_cache)
        if H == 1:
            d1, dW, db = conv1d_same_bwd(dh[:, 0], cc)
            dh = d1[:, None]
        else:
            dh, dW, db = conv2d_same_bwd(dh, cc)
        g[n + "/weights"] = dW
        if bc is None:
            g[n + "/biases"] = db
    return dh[..., 0].reshape(dy.shape[:-1] + (-1,)), g


GENERATORS = {
    "dnn": (g_dnn_fwd, g_dnn_bwd),
    "rced": (g_rced_fwd, g_rced_bwd),
    "lstm": (g_lstm_fwd, g_lstm_bwd),
    "res_lstm_l": (g_res_lstm_l_fwd, g_res_lstm_l_bwd),
    "res_lstm_base": (lambda p, x, l, **kw: g_res_lstm_l_fwd(p, x, l, False, **kw),
                      lambda p, dy, c: g_res_lstm_l_bwd(p, dy, c, False)),
}
DISCRIMINATORS = {
    "lstm": (d_lstm_fwd, d_lstm_bwd),
    "dnn": (d_dnn_fwd, d_dnn_bwd),
}

# --------------------------------------------------------------------------
# losses  (models/gan_rnn_placeholder.py:244-260; same formulas models/gan.py:200-208)
# --------------------------------------------------------------------------


def lsgan_mse_losses(d_rl_logits, d_fk_logits, g, y, d_real=1.0, d_fake=0.0,
                     mse_lambda=10.0, output_dim=40):
    """All means are over every element including padded frames (App. C-1)."""
    d_rl = np.mean((d_rl_logits - d_real) ** 2)
    d_fk = np.mean((d_fk_logits - d_fake) ** 2)
    g_adv = np.mean((d_fk_logits - d_real) ** 2)
    g_mse = 0.5 * np.mean((g - y) ** 2) * output_dim
    return dict(d_rl_loss=d_rl, d_fk_loss=d_fk, d_loss=d_rl + d_fk,
                g_adv_loss=g_adv, g_mse_loss=g_mse, g_l2_loss=0.0,
                g_loss=g_adv + mse_lambda * g_mse)


def l2_loss_g(p, l2_scale):
    """gan_rnn_placeholder.py:253-258: l2_scale * sum(0.5*||v||^2) over G vars
    whose name does not contain "bias"."""
    return l2_scale * sum(0.5 * float((v * v).sum()) for k, v in p.items() if "bias" not in k)


# --------------------------------------------------------------------------
# update rules  (gan_rnn_placeholder.py:144-150,177-189)
# --------------------------------------------------------------------------


def average_gradients(tower_grads):
    """utils/ops.py:343-376: per-variable mean over towers."""
    out = OrderedDict()
    for k in tower_grads[0]:
        out[k] = np.mean(np.stack([tg[k] for tg in tower_grads], 0), 0)
    return out


def clip_by_norm(g, max_norm=15.0):
    """tf.clip_by_norm per tensor: g * max_norm / max(||g||, max_norm)."""
    n = math.sqrt(float((g.astype(np.float64) ** 2).sum()))
    return g * (max_norm / max(n, max_norm))


def sgd_update(p, g, lr):
    return OrderedDict((k, p[k] - lr * g[k]) for k in p)


def adam_update_tf(p, g, m, v, t, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer: lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    theta -= lr_t * m / (sqrt(v) + eps)   (eps on the un-corrected sqrt(v))."""
    t = t + 1
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    pn, mn, vn = OrderedDict(), OrderedDict(), OrderedDict()
    for k in p:
        mn[k] = beta1 * m[k] + (1 - beta1) * g[k]
        vn[k] = beta2 * v[k] + (1 - beta2) * g[k] * g[k]
        pn[k] = p[k] - lr_t * mn[k] / (np.sqrt(vn[k]) + eps)
    return pn, mn, vn, t


def ema_update(shadow, p, decay=0.9999):
    """tf.train.ExponentialMovingAverage(0.9999).apply without num_updates,
    evaluated on the post-update weights (the tf.group has no ordering; we fix
    'after')."""
    return OrderedDict((k, shadow[k] - (1 - decay) * (shadow[k] - p[k])) for k in p)


def exponential_decay(iteration, num_jobs, num_iters, init_lr, multiply_jobs=True):
    """utils/ops.py:378-391."""
    final = 0.0001 * init_lr
    if iteration + 1 >= num_iters:
        cur = final
    else:
        cur = init_lr * math.exp(iteration * math.log(final / init_lr) / num_iters)
    return num_jobs * cur if multiply_jobs else cur


# --------------------------------------------------------------------------
# CMVN  (io_funcs/convert_cmvn_to_numpy.py:29-47; make_tfrecords.py:84-87;
#        scripts/train_gan_rnn_placeholder.py:286-287)
# --------------------------------------------------------------------------


def cmvn_from_stats(stats):
    """Kaldi global stats (2, D+1): row0 = sums | count, row1 = sumsq | 0."""
    n = stats[0][-1]
    s = stats[:, :-1]
    mean = s[0] / n
    std = np.sqrt(s[1] / n - mean ** 2)
    return mean, std


def cmvn_apply(x, mean, std):
    """(x-mean)/std in float64, stored as float32."""
    return ((x.astype(np.float64) - mean) / std).astype(np.float32)


def cmvn_invert(y, mean, std):
    """y*std+mean (decode), written to ark as float32 (kaldi_io.py:269)."""
    return (y * std + mean).astype(np.float32)


# --------------------------------------------------------------------------
# one D update / one G update  (gan_rnn_placeholder.py:169-189,
# train_gan_rnn_placeholder.py:72-101)
# --------------------------------------------------------------------------


class GanState(object):
    """Weights + optimizer state of one GAN (all towers share it)."""

    def __init__(self, g_params, d_params, g_type="lstm", d_type="lstm"):
        self.g, self.d = g_params, d_params
        self.g_type, self.d_type = g_type, d_type
        z = lambda p: OrderedDict((k, np.zeros_like(v)) for k, v in p.items())
        self.adam_m, self.adam_v, self.adam_t = z(g_params), z(g_params), 0
        self.d_adam_m, self.d_adam_v, self.d_adam_t = z(d_params), z(d_params), 0   # models/gan.py:125 (Adam for D)
        self.g_ema = OrderedDict((k, v.copy()) for k, v in g_params.items())
        self.d_ema = OrderedDict((k, v.copy()) for k, v in d_params.items())


def tower_losses_and_grads(st, x, y, lengths, which, noise_rl=None, noise_fk=None,
                           mse_lambda=10.0, d_real=1.0, d_fake=0.0, l2_scale=0.0, g_opts=None, d_opts=None,
                           d_cat=None, l2_weights_only=False, update_ops=None):
    """One tower of build_model_single_gpu (gan_rnn_placeholder.py:191-298) plus
    compute_gradients wrt d_vars (which='d') or g_vars (which='g').
    update_ops: which batch_norm UPDATE_OPS the optimizer op depends on (they mutate opts['bn_state'] in place):
      'own' -- gan_rnn_placeholder.py:163-175: d_opt runs the d_model updates (both discriminator passes), g_opt the
               g_model ones;
      'all' -- models/gan.py:139-143: both optimizers depend on the whole collection;
      None  -- whatever the caller put into the opts dicts ('update' key, default off)."""
    gf, gb = GENERATORS[st.g_type]
    df, db_ = DISCRIMINATORS[st.d_type]
    if update_ops is not None:
        if g_opts is not None:
            g_opts = dict(g_opts, update=update_ops == "all" or which == "g")     # bn_state stays the caller's object
        if d_opts is not None:
            d_opts = dict(d_opts, update=update_ops == "all" or which == "d")
    # g_opts / d_opts: batch_norm state, dropout stream (fc_block_fwd); salts: G layers 0.., D(labels) 256.., D(G(x)) 512..
    g_out, gc = gf(st.g, x, lengths) if g_opts is None else gf(st.g, x, lengths, opts=g_opts, salt0=0)
    # d_cat = (c0, c1): the frame-level GAN of models/gan.py:159-174 feeds D tf.concat([inputs[..., c0:c1], .], -1)
    d_in = (lambda v: v) if d_cat is None else (lambda v: np.concatenate([x[..., d_cat[0]:d_cat[1]], v], -1))
    if d_opts is None:
        lr_, crl = df(st.d, d_in(y), lengths, noise_rl)
        lf_, cfk = df(st.d, d_in(g_out), lengths, noise_fk)
    else:
        lr_, crl = df(st.d, d_in(y), lengths, noise_rl, opts=d_opts, salt0=256)
        lf_, cfk = df(st.d, d_in(g_out), lengths, noise_fk, opts=d_opts, salt0=512)
    losses = lsgan_mse_losses(lr_, lf_, g_out, y, d_real, d_fake, mse_lambda, y.shape[-1])
    reg = (lambda k: k.endswith("weights")) if l2_weights_only else (lambda k: "bias" not in k)
    if l2_scale > 0.0:
        losses["g_l2_loss"] = l2_scale * sum(0.5 * float((v * v).sum()) for k, v in st.g.items() if reg(k))
        losses["g_loss"] += losses["g_l2_loss"]
    n_logit = lr_.size
    if which == "d":
        _, g_rl = db_(st.d, 2.0 * (lr_ - d_real) / n_logit, crl)
        _, g_fk = db_(st.d, 2.0 * (lf_ - d_fake) / n_logit, cfk)
        grads = OrderedDict((k, g_rl[k] + g_fk[k]) for k in st.d)
    else:
        dg_adv, _ = db_(st.d, 2.0 * (lf_ - d_real) / n_logit, cfk)
        if d_cat is not None:
            dg_adv = dg_adv[..., d_cat[1] - d_cat[0]:]          # the generator's block of the concatenated input
        dg = dg_adv + mse_lambda * 0.5 * y.shape[-1] * 2.0 * (g_out - y) / g_out.size
        _, gg = gb(st.g, dg, gc)
        grads = OrderedDict((k, gg[k]) for k in st.g)
        if l2_scale > 0.0:
            for k in grads:
                if reg(k):
                    grads[k] = grads[k] + l2_scale * st.g[k]
    return losses, grads, g_out


def d_step(st, towers, lr_d, max_norm=15.0, ema_decay=0.9999, adam=False, **kw):
    """towers: list of dicts(x, y, lengths, noise_rl, noise_fk). SGD on theta_D (gan_rnn_placeholder.py:144), or Adam
    (adam=True: models/gan.py:125)."""
    res = [tower_losses_and_grads(st, t["x"], t["y"], t["lengths"], "d",
                                  t.get("noise_rl"), t.get("noise_fk"), **kw) for t in towers]
    avg = average_gradients([r[1] for r in res])
    clipped = OrderedDict((k, clip_by_norm(v, max_norm)) for k, v in avg.items())
    if adam:
        st.d, st.d_adam_m, st.d_adam_v, st.d_adam_t = adam_update_tf(
            st.d, clipped, st.d_adam_m, st.d_adam_v, st.d_adam_t, lr_d)
    else:
        st.d = sgd_update(st.d, clipped, lr_d)
    st.d_ema = ema_update(st.d_ema, st.d, ema_decay)
    return [r[0] for r in res], clipped


def g_step(st, towers, lr_g, max_norm=15.0, ema_decay=0.9999, **kw):
    res = [tower_losses_and_grads(st, t["x"], t["y"], t["lengths"], "g",
                                  t.get("noise_rl"), t.get("noise_fk"), **kw) for t in towers]
    avg = average_gradients([r[1] for r in res])
    clipped = OrderedDict((k, clip_by_norm(v, max_norm)) for k, v in avg.items())
    st.g, st.adam_m, st.adam_v, st.adam_t = adam_update_tf(
        st.g, clipped, st.adam_m, st.adam_v, st.adam_t, lr_g)
    st.g_ema = ema_update(st.g_ema, st.g, ema_decay)
    return [r[0] for r in res], clipped


# --------------------------------------------------------------------------
# MSE-only generator training: models/dnn_trainer_single_gpu.py:93-133
# (BASELINE.json configs[0]; also the RCED trainer's loss)
# --------------------------------------------------------------------------


def mse_losses_and_grads(g_params, g_type, x, y, l2_scale=0.0, g_opts=None):
    """g_mse = 0.5 * output_dim * mean((G(x)-y)^2) (:109-110); g_l2 = sum over WEIGHTS (contrib
    l2_regularizer is attached to weights only, dnn.py:64-67,85-86) of l2_scale * 0.5 ||W||^2 (:111-115);
    gradients of g_mse + g_l2 wrt every g_ variable (:102-104 minimize, no clipping)."""
    gf, gb = GENERATORS[g_type]
    g_out, gc = gf(g_params, x, None) if g_opts is None else gf(g_params, x, None, opts=g_opts, salt0=0)
    out_dim = y.shape[-1]
    losses = dict(g_mse_loss=0.5 * out_dim * float(np.mean((g_out - y) ** 2)), g_l2_loss=0.0)
    _, grads = gb(g_params, out_dim * (g_out - y) / g_out.size, gc)
    grads = OrderedDict((k, grads[k]) for k in g_params)
    if l2_scale > 0.0:
        losses["g_l2_loss"] = l2_scale * sum(0.5 * float((v * v).sum()) for k, v in g_params.items()
                                             if k.endswith("weights"))
        for k in grads:
            if k.endswith("weights"):
                grads[k] = grads[k] + l2_scale * g_params[k]
    losses["g_loss"] = losses["g_mse_loss"] + losses["g_l2_loss"]
    return losses, grads, g_out


class MseState(object):
    def __init__(self, g_params, g_type="dnn"):
        self.g, self.g_type = g_params, g_type
        z = lambda: OrderedDict((k, np.zeros_like(v)) for k, v in g_params.items())
        self.adam_m, self.adam_v, self.adam_t = z(), z(), 0


def mse_step(st, x, y, lr, l2_scale=0.0, g_opts=None):
    """One DNNTrainer update: Adam(lr) on g_mse + g_l2, no clipping, no EMA.  With batch_norm the UPDATE_OPS run
    with the step (dnn_trainer_single_gpu.py:101-104): pass g_opts with update=True."""
    losses, grads, g_out = mse_losses_and_grads(st.g, st.g_type, x, y, l2_scale, g_opts)
    st.g, st.adam_m, st.adam_v, st.adam_t = adam_update_tf(st.g, grads, st.adam_m, st.adam_v, st.adam_t, lr)
    return losses, grads
