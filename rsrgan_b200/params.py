"""Flat parameter storage of one network (generator or discriminator).

Every TF-1.4 variable of the reference (names in SURVEY.md App. B) is one *segment* of a
flat fp32 device buffer, stored in the layout the kernels consume (zero-padded, LSTM gate
columns packed -- see packing.py) and padded to a multiple of 1024 elements so that the
fused update sweep (rsr_seg_sumsq / rsr_clip_*_ema) handles whole blocks of one tensor.
Five parallel buffers share the segment table: theta, grad, ema (tf.train.ExponentialMovingAverage
shadows, models/gan_rnn_placeholder.py:149-150), Adam m / v (models/gan_rnn_placeholder.py:147)
and theta16, the 16-bit operand copy the tensor cores read, which the update kernel refreshes.

Padding invariant: padded elements are exactly zero in theta and receive exactly zero
gradients (padded activations are zero), so they stay zero under SGD and Adam.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from . import packing

BLOCK = 1024


class Seg(object):
    __slots__ = ("name", "kind", "tf_shape", "dev_shape", "off", "size", "meta", "index")

    def __init__(self, name, kind, tf_shape, dev_shape, meta=None):
        self.name, self.kind, self.tf_shape, self.dev_shape = name, kind, tuple(tf_shape), tuple(dev_shape)
        self.meta = meta or {}
        self.off = self.size = self.index = 0


def fc_w(name, n_in, n_out):
    return Seg(name, "fc_w", (n_in, n_out), (packing.round_up(n_in, 8), packing.round_up(n_out, 8)))


def fc_w_cat(name, n_a, n_b, n_out):
    """First-layer weight of a fully_connected fed tf.concat([a, b], -1) (models/gan.py:159-174: a = centre-frame LPS,
    b = MFCC).  TF rows are [a ; b]; the device rows are [b ; a ; zero pad] so that the b block -- the part of the data
    gradient that continues into the generator -- starts at column 0 of the activation buffer (16-byte aligned for TMA)."""
    return Seg(name, "fc_w_cat", (n_a + n_b, n_out), (packing.round_up(n_a + n_b, 8), packing.round_up(n_out, 8)),
               dict(n_a=n_a, n_b=n_b))


def fc_b(name, n_out):
    return Seg(name, "vec", (n_out,), (packing.round_up(n_out, 8),))


def conv_w(name, width, c_in, c_out):
    """tf.contrib.layers.conv2d filter (1, w, C_in, C_out) (models/rced.py:94-101, splice = 1) stored as the GEMM
    B operand [w * Cin_p, Cout_p] of the overlapped-view convolution (channels padded to multiples of 8)."""
    cip, cop = packing.round_up(c_in, 8), packing.round_up(c_out, 8)
    return Seg(name, "conv_w", (1, width, c_in, c_out), (width * cip, cop), dict(W=width, Cin_p=cip, Cout_p=cop))


def conv_w2d(name, kh, width, c_in, c_out):
    """tf.contrib.layers.conv2d filter (kh, w, C_in, C_out) of the [splice, w] convolutions (models/rced.py:94-101,
    splice > 1), stored COMPACT exactly as TensorFlow holds it; the GEMM operand -- its block-Toeplitz expansion over
    the stacked lines -- is derived after every update (rsr_conv_toeplitz_expand)."""
    return Seg(name, "conv_w2d", (kh, width, c_in, c_out), (kh * width * c_in, c_out), dict(kh=kh, W=width, Ci=c_in, Co=c_out))


def fc_w_lines(name, lines, positions, chans, n_out):
    """FC over a flattened NHWC frame of `lines` stacked lines (models/rced.py:106-113): TF rows = (h * L + pos) * chans
    + ch; device rows = pos * Cp + h * chans + ch, Cp = lines * chans padded to a multiple of 8 (zero rows)."""
    cp = packing.round_up(lines * chans, 8)
    return Seg(name, "fc_w_lines", (lines * positions * chans, n_out), (positions * cp, packing.round_up(n_out, 8)),
               dict(H=lines, L=positions, C=chans, Cp=cp))


def fc_w_frames(name, positions, chans, n_out):
    """FC over a flattened channels-last frame (models/rced.py:106-113): TF rows = pos * chans + ch; device rows =
    pos * Cp + ch with zero rows for the padded channels."""
    cp = packing.round_up(chans, 8)
    return Seg(name, "fc_w_frames", (positions * chans, n_out), (positions * cp, packing.round_up(n_out, 8)),
               dict(L=positions, C=chans, Cp=cp))


def lstm_cell(prefix, I, C, P):
    """The six variables of one tf.contrib.rnn.LSTMCell(use_peepholes, num_proj) in TF creation order."""
    Ip, Pp, Cp = packing.round_up(I, 8), packing.round_up(P, 8), packing.cell_pad(C)
    meta = dict(I=I, C=C, P=P, Ip=Ip, Pp=Pp, Cp=Cp)
    return [
        Seg(prefix + "kernel", "lstm_kernel", (I + P, 4 * C), (Ip + Pp, 4 * Cp), meta),
        Seg(prefix + "bias", "lstm_bias", (4 * C,), (4 * Cp,), meta),
        Seg(prefix + "w_f_diag", "peep", (C,), (Cp,), meta),
        Seg(prefix + "w_i_diag", "peep", (C,), (Cp,), meta),
        Seg(prefix + "w_o_diag", "peep", (C,), (Cp,), meta),
        Seg(prefix + "projection/kernel", "proj", (C, P), (Cp, Pp), meta),
    ]


def to_dev_layout(seg, a):
    """numpy array in TF layout -> numpy array in device layout (zero padded / gate packed)."""
    a = np.asarray(a)
    assert a.shape == seg.tf_shape, (seg.name, a.shape, seg.tf_shape)
    out = np.zeros(seg.dev_shape, dtype=np.float32)
    if seg.kind == "fc_w":
        out[:a.shape[0], :a.shape[1]] = a
    elif seg.kind == "fc_w_cat":
        n_a, n_b = seg.meta["n_a"], seg.meta["n_b"]
        out[:n_b, :a.shape[1]] = a[n_a:]
        out[n_b:n_b + n_a, :a.shape[1]] = a[:n_a]
    elif seg.kind in ("vec", "peep"):
        out[:a.shape[0]] = a
    elif seg.kind == "lstm_bias":
        out[:] = packing.pack_cols(a, seg.meta["C"])
    elif seg.kind == "proj":
        out[:a.shape[0], :a.shape[1]] = a
    elif seg.kind == "conv_w":
        m = seg.meta
        out.reshape(m["W"], m["Cin_p"], m["Cout_p"])[:, :a.shape[2], :a.shape[3]] = a[0]
    elif seg.kind == "fc_w_frames":
        m = seg.meta
        out.reshape(m["L"], m["Cp"], -1)[:, :m["C"], :a.shape[1]] = a.reshape(m["L"], m["C"], -1)
    elif seg.kind == "conv_w2d":
        out[:] = a.reshape(seg.dev_shape)
    elif seg.kind == "fc_w_lines":
        m = seg.meta
        t = a.reshape(m["H"], m["L"], m["C"], -1).transpose(1, 0, 2, 3).reshape(m["L"], m["H"] * m["C"], -1)
        out.reshape(m["L"], m["Cp"], -1)[:, :m["H"] * m["C"], :a.shape[1]] = t
    elif seg.kind == "lstm_kernel":
        m = seg.meta
        p = packing.pack_cols(a, m["C"])
        out[:m["I"]] = p[:m["I"]]
        out[m["Ip"]:m["Ip"] + m["P"]] = p[m["I"]:]
    else:
        raise ValueError(seg.kind)
    return out


def from_dev_layout(seg, d):
    d = np.asarray(d).reshape(seg.dev_shape)
    if seg.kind == "fc_w":
        return d[:seg.tf_shape[0], :seg.tf_shape[1]].copy()
    if seg.kind == "fc_w_cat":
        n_a, n_b = seg.meta["n_a"], seg.meta["n_b"]
        return np.concatenate([d[n_b:n_b + n_a, :seg.tf_shape[1]], d[:n_b, :seg.tf_shape[1]]], 0)
    if seg.kind in ("vec", "peep"):
        return d[:seg.tf_shape[0]].copy()
    if seg.kind == "lstm_bias":
        return packing.unpack_cols(d, seg.meta["C"])
    if seg.kind == "proj":
        return d[:seg.tf_shape[0], :seg.tf_shape[1]].copy()
    if seg.kind == "conv_w":
        m = seg.meta
        return d.reshape(m["W"], m["Cin_p"], m["Cout_p"])[None, :, :seg.tf_shape[2], :seg.tf_shape[3]].copy()
    if seg.kind == "fc_w_frames":
        m = seg.meta
        return d.reshape(m["L"], m["Cp"], -1)[:, :m["C"], :seg.tf_shape[1]].reshape(seg.tf_shape).copy()
    if seg.kind == "conv_w2d":
        return d.reshape(seg.tf_shape).copy()
    if seg.kind == "fc_w_lines":
        m = seg.meta
        t = d.reshape(m["L"], m["Cp"], -1)[:, :m["H"] * m["C"], :seg.tf_shape[1]]
        return t.reshape(m["L"], m["H"], m["C"], -1).transpose(1, 0, 2, 3).reshape(seg.tf_shape).copy()
    if seg.kind == "lstm_kernel":
        m = seg.meta
        u = packing.unpack_cols(d, m["C"])
        return np.concatenate([u[:m["I"]], u[m["Ip"]:m["Ip"] + m["P"]]], 0)
    raise ValueError(seg.kind)


class ParamStore(object):
    def __init__(self, handle, segs, adam):
        self.h = handle
        self.segs = OrderedDict()
        off = 0
        for i, s in enumerate(segs):
            n = int(np.prod(s.dev_shape))
            s.off, s.size, s.index = off, packing.round_up(n, BLOCK), i
            off += s.size
            assert s.name not in self.segs, s.name
            self.segs[s.name] = s
        self.n = off
        dev = handle.device
        z = lambda: torch.zeros(self.n, dtype=torch.float32, device=dev)
        self.theta, self.grad, self.ema = z(), z(), z()
        self.m, self.v = (z(), z()) if adam else (None, None)
        self.theta16 = torch.zeros(self.n, dtype=handle.h16, device=dev)
        seg_id = np.concatenate([np.full(s.size // BLOCK, s.index, np.int32) for s in self.segs.values()])
        self.seg_id = torch.tensor(seg_id, device=dev)
        # L2 regulariser covers variables whose name does not contain "bias" (gan_rnn_placeholder.py:254)
        self.seg_l2 = torch.tensor(np.array([0 if "bias" in s.name else 1 for s in self.segs.values()], np.int32),
                                   device=dev)
        self.sumsq = torch.zeros(len(self.segs), dtype=torch.float32, device=dev)
        # device-resident hyper-parameters, see rsr_clip_adam_ema in include/rsrgan_b200.h
        self.hyper = torch.tensor([0.0, 0.9, 0.999, 1e-8, 0.9, 0.999, 0.0, 0.0], dtype=torch.float32, device=dev)
        self.adam = adam

    # -- views ---------------------------------------------------------------------------
    def view(self, name, buf="theta"):
        s = self.segs[name]
        n = int(np.prod(s.dev_shape))
        return getattr(self, buf)[s.off:s.off + n].view(*s.dev_shape)

    def n_params(self):
        return sum(int(np.prod(s.tf_shape)) for s in self.segs.values())

    # -- host <-> device -----------------------------------------------------------------
    def load_tf(self, params, buf="theta"):
        """params: dict TF-variable-name -> numpy array in TF layout."""
        flat = np.zeros(self.n, np.float32)
        for s in self.segs.values():
            d = to_dev_layout(s, params[s.name])
            flat[s.off:s.off + d.size] = d.reshape(-1)
        t = torch.from_numpy(flat).to(self.h.device)
        getattr(self, buf).copy_(t)
        if buf == "theta":
            self.refresh16()

    def export_tf(self, buf="theta", dtype=np.float32):
        flat = getattr(self, buf).detach().float().cpu().numpy()
        out = OrderedDict()
        for s in self.segs.values():
            n = int(np.prod(s.dev_shape))
            out[s.name] = from_dev_layout(s, flat[s.off:s.off + n]).astype(dtype)
        return out

    def refresh16(self):
        self.h.cast16(self.theta, self.theta16)

    def set_lr(self, lr):
        self.hyper[0:1].fill_(float(lr))

    def state_dict(self):
        d = OrderedDict()
        for buf in ("theta", "ema", "m", "v"):
            if getattr(self, buf) is not None:
                d[buf] = self.export_tf(buf)
        d["hyper"] = self.hyper.cpu().numpy()
        return d

    def load_state_dict(self, d):
        for buf in ("theta", "ema", "m", "v"):
            if buf in d and getattr(self, buf) is not None:
                self.load_tf(d[buf], buf)
        self.hyper.copy_(torch.tensor(np.asarray(d["hyper"], np.float32)))
