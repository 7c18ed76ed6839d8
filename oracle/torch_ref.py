"""Second, independent CPU statement of the same path: torch float64 forward +
autograd backward.  TEST INFRASTRUCTURE ONLY (see oracle/rsr_oracle.py header).

Used (a) to pin the hand-derived numpy backward of rsr_oracle.py (must agree to
1e-9), and (b) as the multi-threaded CPU baseline `bench.py --impl reference`
times ("CPU restatement of the reference, TF1 unavailable").

Parameter dicts use the same TF-1.4 variable names as rsr_oracle.py.
"""
from __future__ import annotations

from collections import OrderedDict

import torch


def to_torch(p, dtype=torch.float64, requires_grad=False):
    return OrderedDict((k, torch.tensor(v, dtype=dtype, requires_grad=requires_grad))
                       for k, v in p.items())


def lstmp(x, lengths, K, b, w_i, w_f, w_o, W_p, forget_bias=1.0):
    """tf.contrib.rnn.LSTMCell(peepholes, num_proj) under dynamic_rnn; gate order
    i,j,f,o (models/BNLSTMCell.py:176-213)."""
    B, T, _ = x.shape
    C = w_i.shape[0]
    P = W_p.shape[1]
    c = x.new_zeros(B, C)
    m = x.new_zeros(B, P)
    outs = []
    for t in range(T):
        act = (t < lengths).unsqueeze(1)
        z = torch.cat([x[:, t], m], 1) @ K + b
        zi, zj, zf, zo = z.split(C, dim=1)
        cn = torch.sigmoid(zf + forget_bias + w_f * c) * c + torch.sigmoid(zi + w_i * c) * torch.tanh(zj)
        mn = (torch.sigmoid(zo + w_o * cn) * torch.tanh(cn)) @ W_p
        outs.append(torch.where(act, mn, torch.zeros_like(mn)))
        c = torch.where(act, cn, c)
        m = torch.where(act, mn, m)
    return torch.stack(outs, 1)


def _cells(p, scope):
    return sorted({k[:-len("kernel")] for k in p
                   if k.startswith(scope) and k.endswith("lstm_cell/kernel")})


def _cell(p, pre, x, lengths):
    return lstmp(x, lengths, p[pre + "kernel"], p[pre + "bias"], p[pre + "w_i_diag"],
                 p[pre + "w_f_diag"], p[pre + "w_o_diag"], p[pre + "projection/kernel"])


def lrelu(x):
    return torch.maximum(x, 0.3 * x)


def _seq_drop(a, opts, salt):
    """DropoutWrapper(output_keep_prob): mask drawn in the library layout (time-major rows, padded pitch)."""
    from . import rsr_oracle as O
    opts = opts or {}
    keep = opts.get("keep_prob", 1.0) if opts.get("train", True) else 1.0
    if keep >= 1.0:
        return a
    B, T, P = a.shape
    Pp = -(-P // 8) * 8
    seed, tick = opts["rng"]
    m = O.dropout_mask(seed, tick, salt, T * B, Pp, keep).reshape(T, B, Pp)[:, :, :P].transpose(1, 0, 2)
    return torch.where(torch.as_tensor(m.copy()), a / keep, torch.zeros_like(a))


def g_lstm(p, x, lengths, opts=None, salt0=0):
    h = _fc_block(p, "g_model/fully_connected", x, lrelu, dict(opts or {}, keep_prob=1.0), salt0)
    for l, pre in enumerate(_cells(p, "g_model/rnn/")):
        h = _seq_drop(_cell(p, pre, h, lengths), opts, salt0 + 16 + l)
    return h @ p["g_model/fully_connected_1/weights"] + p["g_model/fully_connected_1/biases"]


def g_res_lstm_l(p, x, lengths, residual=True, opts=None, salt0=0):
    xin = x
    for l, pre in enumerate(_cells(p, "g_model/lstm_cell_")):
        o = _seq_drop(_cell(p, pre, xin, lengths), opts, salt0 + 16 + l)
        xin = o + xin if residual else o
    return xin @ p["g_model/forward_out/fully_connected/weights"] + \
        p["g_model/forward_out/fully_connected/biases"]


def d_lstm(p, x, lengths, noise=None):
    h = x if noise is None else x + noise
    for pre in _cells(p, "d_model/rnn/"):
        h = _cell(p, pre, h, lengths)
    return h @ p["d_model/fully_connected/weights"] + p["d_model/fully_connected/biases"]


def bn_renorm_train(z, gamma, beta, st, eps=1e-3):
    """tf.layers.BatchNormalization(renorm=True).call, training branch, written independently of the numpy oracle:
    torch moments + nn.batch_normalization form x * inv + (offset - mean * inv); r and d detached (stop_gradient).
    `st` holds the PRE-update renorm variables (torch tensors or floats); the state update is not differentiated."""
    red = tuple(range(z.dim() - 1))
    mean = z.mean(red)
    var = z.var(red, unbiased=False)
    stddev = torch.sqrt(var + eps)
    denom = st["renorm_stddev"] + (1.0 - st["renorm_stddev_weight"]) * stddev
    r = (stddev / denom).detach()
    d = ((mean - (st["renorm_mean"] + (1.0 - st["renorm_mean_weight"]) * mean)) / denom).detach()
    scale, offset = r * gamma, d * gamma + beta
    inv = torch.rsqrt(var + eps) * scale
    return z * inv + (offset - mean * inv)


def _fc_block(p, n, h, act, opts, salt):
    from . import rsr_oracle as O
    opts = opts or {}
    train = opts.get("train", True)
    if (n + "/BatchNorm/gamma") in p:
        z = h @ p[n + "/weights"]
        st = {k: torch.as_tensor(opts["bn_state"][n + "/BatchNorm/" + k], dtype=z.dtype) for k in O.BN_STATE_KEYS}
        if train:
            y = bn_renorm_train(z, p[n + "/BatchNorm/gamma"], p[n + "/BatchNorm/beta"], st)
        else:
            y = torch.nn.functional.batch_norm(z.reshape(-1, z.shape[-1]), st["moving_mean"], st["moving_variance"],
                                               p[n + "/BatchNorm/gamma"], p[n + "/BatchNorm/beta"], False, 0.0,
                                               1e-3).reshape(z.shape)
    else:
        y = h @ p[n + "/weights"] + p[n + "/biases"]
    a = act(y)
    keep = opts.get("keep_prob", 1.0) if train else 1.0
    if keep < 1.0:
        seed, tick = opts["rng"]
        rows = int(a.numel() // a.shape[-1])
        m = torch.as_tensor(O.dropout_mask(seed, tick, salt, rows, a.shape[-1], keep).reshape(tuple(a.shape)))
        a = torch.where(m, a / keep, torch.zeros_like(a))
    return a


def d_dnn(p, x, lengths=None, noise=None, opts=None, salt0=256):
    names = sorted({k.rsplit("/", 1)[0] for k in p if k.startswith("d_model/fully_connected") and k.endswith("/weights")},
                   key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)
    h = x
    for i, n in enumerate(names[:-1]):
        h = _fc_block(p, n, h, torch.relu, opts, salt0 + i)
    y = h @ p[names[-1] + "/weights"] + p[names[-1] + "/biases"]
    return torch.clamp(y, -0.5, 1.5)


def _fc_names(p, scope):
    return sorted({k.rsplit("/", 1)[0] for k in p if k.startswith(scope + "/fully_connected") and k.endswith("/weights")},
                  key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)


def g_dnn(p, x, lengths=None, opts=None, salt0=0):
    """models/dnn.py:79-110."""
    names = _fc_names(p, "g_model")
    h = x
    for i, n in enumerate(names[:-1]):
        h = _fc_block(p, n, h, torch.relu, opts, salt0 + i)
    return h @ p[names[-1] + "/weights"] + p[names[-1] + "/biases"]


def g_rced(p, x, lengths=None):
    """models/rced.py:90-114 with splice = 1, through torch.nn.functional.conv2d on the NHWC->NCHW tensor."""
    import torch.nn.functional as F
    lead, L = x.shape[:-1], x.shape[-1]
    h = x.reshape(-1, 1, 1, L)                                  # N, C=1, H=splice=1, W=257
    names = sorted({k.rsplit("/", 1)[0] for k in p if k.startswith("g_model/Conv")},
                   key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)
    for n in names:
        W = p[n + "/weights"].permute(3, 2, 0, 1)               # TF HWIO -> torch OIHW
        h = torch.relu(F.conv2d(h, W, p[n + "/biases"], padding=(0, W.shape[3] // 2)))
    flat = h.permute(0, 2, 3, 1).reshape(h.shape[0], -1)        # NHWC flatten (tf.reshape, :106)
    y = flat @ p["g_model/fully_connected/weights"] + p["g_model/fully_connected/biases"]
    return y.reshape(*lead, -1)


GEN = {"lstm": g_lstm, "res_lstm_l": g_res_lstm_l, "dnn": g_dnn, "rced": g_rced,
       "res_lstm_base": lambda p, x, l, **kw: g_res_lstm_l(p, x, l, False, **kw)}
DIS = {"lstm": d_lstm, "dnn": d_dnn}


def losses(gp, dp, g_type, d_type, x, y, lengths, noise_rl=None, noise_fk=None,
           mse_lambda=10.0, d_real=1.0, d_fake=0.0, g_opts=None, d_opts=None, d_cat=None):
    """models/gan_rnn_placeholder.py:196-260; d_cat = (c0, c1): models/gan.py:159-174 (D sees concat([x[c0:c1], .]))."""
    g = GEN[g_type](gp, x, lengths) if g_opts is None else GEN[g_type](gp, x, lengths, opts=g_opts, salt0=0)
    d_in = (lambda v: v) if d_cat is None else (lambda v: torch.cat([x[..., d_cat[0]:d_cat[1]], v], -1))
    if d_opts is None:
        rl = DIS[d_type](dp, d_in(y), lengths, noise_rl)
        fk = DIS[d_type](dp, d_in(g), lengths, noise_fk)
    else:
        rl = DIS[d_type](dp, d_in(y), lengths, noise_rl, opts=d_opts, salt0=256)
        fk = DIS[d_type](dp, d_in(g), lengths, noise_fk, opts=d_opts, salt0=512)
    d_rl = ((rl - d_real) ** 2).mean()
    d_fk = ((fk - d_fake) ** 2).mean()
    g_adv = ((fk - d_real) ** 2).mean()
    g_mse = 0.5 * ((g - y) ** 2).mean() * y.shape[-1]
    return dict(d_rl_loss=d_rl, d_fk_loss=d_fk, d_loss=d_rl + d_fk, g_adv_loss=g_adv,
                g_mse_loss=g_mse, g_loss=g_adv + mse_lambda * g_mse), g


def grads(gp, dp, g_type, d_type, x, y, lengths, which, **kw):
    ls, g = losses(gp, dp, g_type, d_type, x, y, lengths, **kw)
    params = dp if which == "d" else gp
    loss = ls["d_loss"] if which == "d" else ls["g_loss"]
    gs = torch.autograd.grad(loss, list(params.values()))
    return ls, OrderedDict(zip(params.keys(), gs)), g
