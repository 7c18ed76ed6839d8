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


def g_lstm(p, x, lengths):
    h = lrelu(x @ p["g_model/fully_connected/weights"] + p["g_model/fully_connected/biases"])
    for pre in _cells(p, "g_model/rnn/"):
        h = _cell(p, pre, h, lengths)
    return h @ p["g_model/fully_connected_1/weights"] + p["g_model/fully_connected_1/biases"]


def g_res_lstm_l(p, x, lengths, residual=True):
    xin = x
    for pre in _cells(p, "g_model/lstm_cell_"):
        o = _cell(p, pre, xin, lengths)
        xin = o + xin if residual else o
    return xin @ p["g_model/forward_out/fully_connected/weights"] + \
        p["g_model/forward_out/fully_connected/biases"]


def d_lstm(p, x, lengths, noise=None):
    h = x if noise is None else x + noise
    for pre in _cells(p, "d_model/rnn/"):
        h = _cell(p, pre, h, lengths)
    return h @ p["d_model/fully_connected/weights"] + p["d_model/fully_connected/biases"]


def d_dnn(p, x, lengths=None, noise=None):
    names = sorted({k.rsplit("/", 1)[0] for k in p if k.startswith("d_model/fully_connected")},
                   key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)
    h = x
    for n in names[:-1]:
        h = torch.relu(h @ p[n + "/weights"] + p[n + "/biases"])
    y = h @ p[names[-1] + "/weights"] + p[names[-1] + "/biases"]
    return torch.clamp(y, -0.5, 1.5)


def _fc_names(p, scope):
    return sorted({k.rsplit("/", 1)[0] for k in p if k.startswith(scope + "/fully_connected")},
                  key=lambda s: int(s.split("_")[-1]) if s[-1].isdigit() else 0)


def g_dnn(p, x, lengths=None):
    """models/dnn.py:79-110."""
    names = _fc_names(p, "g_model")
    h = x
    for n in names[:-1]:
        h = torch.relu(h @ p[n + "/weights"] + p[n + "/biases"])
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
       "res_lstm_base": lambda p, x, l: g_res_lstm_l(p, x, l, False)}
DIS = {"lstm": d_lstm, "dnn": d_dnn}


def losses(gp, dp, g_type, d_type, x, y, lengths, noise_rl=None, noise_fk=None,
           mse_lambda=10.0, d_real=1.0, d_fake=0.0):
    """models/gan_rnn_placeholder.py:196-260."""
    g = GEN[g_type](gp, x, lengths)
    rl = DIS[d_type](dp, y, lengths, noise_rl)
    fk = DIS[d_type](dp, g, lengths, noise_fk)
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
