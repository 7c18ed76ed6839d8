"""The strided 1-D convolution family of the reference's utils/ops.py and its virtual batch norm (utils/bnorm.py), as
sequences of C-ABI kernel calls -- the operators of the SEGAN-style waveform discriminator, models/discriminator.py:38-90:

    downconv  (utils/ops.py:78-98)    tf.nn.conv2d(x[:, :, None, :], W[k, 1, C_in, C_out], strides [1, 2, 1, 1], SAME) (+ b)
    conv1d    (utils/ops.py:138-156)  tf.nn.conv1d(x, W[k, C_in, C_out], stride 1, SAME) (+ b)     -> nets.Conv1dSame geometry
    deconv    (utils/ops.py:277-310)  tf.nn.conv2d_transpose(x, W[k, 1, C_out, C_in], strides [1, 2, 1, 1]) (+ b)
    VBN       (utils/bnorm.py:11-69)  reference / live virtual batch normalisation
    leakyrelu (utils/ops.py:120-121)  max(x, 0.3 x), fused behind the normalisation

Every convolution is an rsr_gemm over a strided, OVERLAPPED view of a channels-last 16-bit sequence buffer (no im2col):
include/rsrgan_b200.h, "Strided members of the family".  torch is the allocator and the view factory only; there is no CPU
path.  Layout (`Seq`): sequence b of a batch occupies rows [G + b*S, G + b*S + L) of a [G + B*S + G, Cp] buffer; rows
L..S-1 of every sequence and the G guard rows at both ends are zero -- the SAME padding, shared between neighbours.  A
stride-2 layer maps pitch S to S/2 and length L to L/2, so S - L must be at least 2 * 15 per stride-2 level below it.
"""
from __future__ import annotations

import torch

from . import packing
from .ops import ACT_LRELU, ACT_NONE

F32 = torch.float32


def same_pad(L, k, stride):
    """TensorFlow SAME padding along one axis: (output length, pad before, pad after)."""
    out = -(-L // stride)
    total = max((out - 1) * stride + k - L, 0)
    return out, total // 2, total - total // 2


class Seq(object):
    """Geometry of one level of a channels-last sequence stack."""
    GUARD = 16

    def __init__(self, B, L, S):
        assert S % 2 == 0 and L % 2 == 0 and S - L >= 16, (L, S)
        self.B, self.L, self.S = B, L, S

    @property
    def rows(self):
        return self.B * self.S

    def half(self):
        return Seq(self.B, self.L // 2, self.S // 2)

    def double(self):
        return Seq(self.B, self.L * 2, self.S * 2)

    def alloc(self, h, cp):
        """-> (whole buffer incl. guards, view of the B*S sequence rows)"""
        t = torch.zeros(self.rows + 2 * self.GUARD, cp, dtype=h.h16, device=h.device)
        return t, t[self.GUARD:self.GUARD + self.rows]

    def window(self, whole, cp, taps, first, step=1):
        """View A[m, k*cp + c] = whole[GUARD + step*m + first + k, c]  (m < rows / step): the GEMM A operand."""
        return whole.as_strided((self.rows // step, taps * cp), (step * cp, 1),
                                whole.storage_offset() + (self.GUARD + first) * cp)

    def phase(self, whole, cp, r):
        """View of the rows 2u + r of every sequence (u < S/2): [rows/2, cp] with row pitch 2*cp."""
        return whole.as_strided((self.rows // 2, cp), (2 * cp, 1), whole.storage_offset() + (self.GUARD + r) * cp)

    def stage(self, h, x, cp):
        """host / device (B, L, C) fp32 -> zero-padded 16-bit buffer of this level (test and input glue)."""
        whole, body = self.alloc(h, cp)
        v = body.view(self.B, self.S, cp)
        v[:, :self.L, :x.shape[2]] = torch.as_tensor(x, device=h.device).to(h.h16)
        return whole

    def unstage(self, body, c):
        return body.view(self.B, self.S, -1)[:, :self.L, :c].float()


class _StridedConv(object):
    """Shared machinery of downconv and deconv: the filter W[k, a, b] (16-bit, [k*Ap, Bp] row-major) and its two phase
    operands.  `contract` = the channel axis the STRIDED-WINDOW product contracts (a), `emit` = the one it produces (b)."""

    def __init__(self, h, kwidth, a, b):
        self.h, self.k = h, kwidth
        self.a, self.b = a, b
        self.ap, self.bp = packing.round_up(a, 8), packing.round_up(b, 8)
        dev = h.device
        self.w16 = torch.zeros(kwidth * self.ap, self.bp, dtype=h.h16, device=dev)          # [k, Ap, Bp]
        self.bias = torch.zeros(max(self.ap, self.bp), dtype=F32, device=dev)
        self.gw = torch.zeros(kwidth * self.ap, self.bp, dtype=F32, device=dev)             # weight gradient (accumulated)
        self.gb = torch.zeros(max(self.ap, self.bp), dtype=F32, device=dev)
        self.nj = [(kwidth - c + 1) // 2 for c in (0, 1)]                                   # taps of parity c
        self.wph = [torch.zeros(self.nj[c] * self.bp, self.ap, dtype=h.h16, device=dev) for c in (0, 1)]

    def load(self, W, b=None):
        """W (k, a, b) fp32 (the TF filter with its singleton axis dropped), b (channels,) or None."""
        W = torch.as_tensor(W, dtype=F32, device=self.h.device)
        w = torch.zeros(self.k, self.ap, self.bp, dtype=F32, device=self.h.device)
        w[:, :self.a, :self.b] = W
        self.w16.copy_(w.reshape(self.k * self.ap, self.bp).to(self.h.h16))
        self.bias.zero_()
        if b is not None:
            self.bias[:len(b)] = torch.as_tensor(b, dtype=F32, device=self.h.device)
        self.refresh()

    def refresh(self):
        """phase operands after every weight update"""
        for c in (0, 1):
            self.h.conv_w_phase(self.w16, self.k, self.ap, self.bp, 2, c, self.wph[c])

    # -- the two products -----------------------------------------------------------------------------------------
    def strided(self, seq_in, x_whole, pl, y_body, bias, act, out32=None):
        """Y[o] = act(sum_k X[2o + k - pl] W[k] + b): x at level seq_in (channels a) -> y at seq_in.half() (channels b)"""
        h = self.h
        A = seq_in.window(x_whole, self.ap, self.k, -pl, step=2)
        h.gemm(A, self.w16, seq_in.rows // 2, self.bp, self.k * self.ap, b_mn=True, bias=bias, act=act,
               out16=y_body, out32=out32)

    def strided_dw(self, seq_in, x_whole, pl, dy_body):
        """dW[k, a, b] += sum_o X[2o + k - pl, a] dY[o, b]"""
        A = seq_in.window(x_whole, self.ap, self.k, -pl, step=2)
        self.h.gemm(A, dy_body, self.k * self.ap, self.bp, seq_in.rows // 2, a_mn=True, b_mn=True, beta=1.0, out32=self.gw)

    def phased(self, seq_lo, lo_whole, pl, hi_whole, bias=None, act=ACT_NONE, dact_src_whole=None, dact=ACT_NONE):
        """HI[2u + r] = sum_j LO[u + e - j] W[2j + c]^T (+ b): lo at level seq_lo (channels b) -> hi at seq_lo.double()
        (channels a); one GEMM per output parity r, rows written with pitch 2*Ap."""
        h, hi = self.h, seq_lo.double()
        for r in (0, 1):
            c = (r + pl) & 1
            e = (r + pl - c) // 2
            nj = self.nj[c]
            A = seq_lo.window(lo_whole, self.bp, nj, e - nj + 1)
            h.gemm(A, self.wph[c], seq_lo.rows, self.ap, nj * self.bp, b_mn=True, bias=bias, act=act,
                   dact_src=None if dact_src_whole is None else hi.phase(dact_src_whole, self.ap, r), dact=dact,
                   out16=hi.phase(hi_whole, self.ap, r))

    def phased_dw(self, seq_lo, lo_body, pl, hi_whole):
        """dW[k, a, b] += sum_o HI[2o + k - pl, a] LO[o, b]   (the weight gradient when the forward is `phased`)"""
        hi = seq_lo.double()
        A = hi.window(hi_whole, self.ap, self.k, -pl, step=2)
        self.h.gemm(A, lo_body, self.k * self.ap, self.bp, seq_lo.rows, a_mn=True, b_mn=True, beta=1.0, out32=self.gw)


class DownConv1d(_StridedConv):
    """utils/ops.py:78-98 `downconv(x, output_dim, kwidth, pool=2)`: (B, L, C_in) -> (B, L/2, C_out)."""

    def __init__(self, h, c_in, c_out, kwidth=31):
        super(DownConv1d, self).__init__(h, kwidth, c_in, c_out)

    def fwd(self, seq, x_whole, act=ACT_NONE, want32=False, bias=True):
        """-> (y_whole 16-bit at seq.half(), z32 [rows/2, Cout_p] fp32 pre-activation or None)"""
        h, lo = self.h, seq.half()
        self._pl = same_pad(seq.L, self.k, 2)[1]
        y_whole, y = lo.alloc(h, self.bp)
        z32 = torch.zeros(lo.rows, self.bp, dtype=F32, device=h.device) if want32 else None
        self.strided(seq, x_whole, self._pl, y, self.bias if bias else None, act, out32=z32)
        h.conv_mask_rows(y, lo.B, lo.S, lo.L, self.bp)
        return y_whole, z32

    def bwd(self, seq, x_whole, dy_whole, want_dx=True, x_act=ACT_NONE):
        """dy_whole: gradient wrt the PRE-activation output at seq.half() (padding rows zero).  Accumulates gw / gb;
        returns the gradient wrt the input (times x_act'(x) when the producer of x was activated), padding rows zero."""
        h, lo = self.h, seq.half()
        dy = dy_whole[lo.GUARD:lo.GUARD + lo.rows]
        self.strided_dw(seq, x_whole, self._pl, dy)
        h.colsum16(dy, lo.rows, self.bp, self.gb[:self.bp], accumulate=True)
        if not want_dx:
            return None
        dx_whole, dx = seq.alloc(h, self.ap)
        self.phased(lo, dy_whole, self._pl, dx_whole, dact_src_whole=x_whole if x_act != ACT_NONE else None, dact=x_act)
        h.conv_mask_rows(dx, seq.B, seq.S, seq.L, self.ap)
        return dx_whole


class Deconv1d(_StridedConv):
    """utils/ops.py:277-310 `deconv(x, output_shape, kwidth, dilation=2)`: (B, L, C_in) -> (B, 2L, C_out); the filter is
    W[k, C_out, C_in] (conv2d_transpose's [height, width, output_channels, in_channels] with width 1 dropped)."""

    def __init__(self, h, c_in, c_out, kwidth=31):
        super(Deconv1d, self).__init__(h, kwidth, c_out, c_in)

    def fwd(self, seq, x_whole, act=ACT_NONE, bias=True):
        h, hi = self.h, seq.double()
        self._pl = same_pad(hi.L, self.k, 2)[1]
        y_whole, y = hi.alloc(h, self.ap)
        self.phased(seq, x_whole, self._pl, y_whole, bias=self.bias if bias else None, act=act)
        h.conv_mask_rows(y, hi.B, hi.S, hi.L, self.ap)
        return y_whole

    def bwd(self, seq, x_whole, dy_whole, want_dx=True, x_act=ACT_NONE):
        h, hi = self.h, seq.double()
        x = x_whole[seq.GUARD:seq.GUARD + seq.rows]
        dy = dy_whole[hi.GUARD:hi.GUARD + hi.rows]
        self.phased_dw(seq, x, self._pl, dy_whole)
        h.colsum16(dy, hi.rows, self.ap, self.gb[:self.ap], accumulate=True)
        if not want_dx:
            return None
        dx_whole, dx = seq.alloc(h, self.bp)
        A = hi.window(dy_whole, self.ap, self.k, -self._pl, step=2)
        h.gemm(A, self.w16, seq.rows, self.bp, self.k * self.ap, b_mn=True, out16=dx,
               dact_src=x if x_act != ACT_NONE else None, dact=x_act)
        h.conv_mask_rows(dx, seq.B, seq.S, seq.L, self.bp)
        return dx_whole


class Conv1d(object):
    """utils/ops.py:138-156 `conv1d(x, kwidth, num_kernels)`: tf.nn.conv1d(x, W[k, C_in, C_out], stride 1, SAME) (+ b) --
    the `logits_conv` of models/discriminator.py:80-82 (kwidth 31, one kernel)."""

    def __init__(self, h, c_in, c_out, kwidth=31):
        self.h, self.k = h, kwidth
        self.a, self.b = c_in, c_out
        self.ap, self.bp = packing.round_up(c_in, 8), packing.round_up(c_out, 8)
        dev = h.device
        self.w16 = torch.zeros(kwidth * self.ap, self.bp, dtype=h.h16, device=dev)
        self.wflip = torch.zeros(kwidth * self.bp, self.ap, dtype=h.h16, device=dev)
        self.bias = torch.zeros(self.bp, dtype=F32, device=dev)
        self.gw = torch.zeros(kwidth * self.ap, self.bp, dtype=F32, device=dev)
        self.gb = torch.zeros(self.bp, dtype=F32, device=dev)

    def load(self, W, b=None):
        w = torch.zeros(self.k, self.ap, self.bp, dtype=F32, device=self.h.device)
        w[:, :self.a, :self.b] = torch.as_tensor(W, dtype=F32, device=self.h.device)
        self.w16.copy_(w.reshape(self.k * self.ap, self.bp).to(self.h.h16))
        self.bias.zero_()
        if b is not None:
            self.bias[:self.b] = torch.as_tensor(b, dtype=F32, device=self.h.device)
        self.refresh()

    def refresh(self):
        self.h.conv_w_phase(self.w16, self.k, self.ap, self.bp, 1, 0, self.wflip)

    def fwd(self, seq, x_whole, act=ACT_NONE, want32=False):
        h = self.h
        _, self._pl, _ = same_pad(seq.L, self.k, 1)
        y_whole, y = seq.alloc(h, self.bp)
        z32 = torch.zeros(seq.rows, self.bp, dtype=F32, device=h.device) if want32 else None
        h.gemm(seq.window(x_whole, self.ap, self.k, -self._pl), self.w16, seq.rows, self.bp, self.k * self.ap, b_mn=True,
               bias=self.bias, act=act, out16=y, out32=z32)
        h.conv_mask_rows(y, seq.B, seq.S, seq.L, self.bp)
        return y_whole, z32

    def bwd(self, seq, x_whole, dy_whole, want_dx=True, x_act=ACT_NONE):
        h = self.h
        dy = dy_whole[seq.GUARD:seq.GUARD + seq.rows]
        h.gemm(seq.window(x_whole, self.ap, self.k, -self._pl), dy, self.k * self.ap, self.bp, seq.rows, a_mn=True,
               b_mn=True, beta=1.0, out32=self.gw)
        h.colsum16(dy, seq.rows, self.bp, self.gb, accumulate=True)
        if not want_dx:
            return None
        dx_whole, dx = seq.alloc(h, self.ap)
        # dx[i] = sum_k dy[i + pl - k] W[k]^T: a stride-1 window of dy against the flipped taps
        h.gemm(seq.window(dy_whole, self.bp, self.k, self._pl - self.k + 1), self.wflip, seq.rows, self.ap, self.k * self.bp,
               b_mn=True, out16=dx, dact_src=x_whole[seq.GUARD:seq.GUARD + seq.rows] if x_act != ACT_NONE else None,
               dact=x_act)
        h.conv_mask_rows(dx, seq.B, seq.S, seq.L, self.ap)
        return dx_whole


class VBN(object):
    """utils/bnorm.py:11-69 on the fp32 pre-activation of a convolution, leaky ReLU fused behind it
    (models/discriminator.py:55-64).  The first call is the reference pass (statistics of that batch, kept);
    later calls are live passes blended with weight 1 / (reference batch size + 1)."""

    def __init__(self, h, channels, eps=1e-5):
        self.h, self.c, self.cp, self.eps = h, channels, packing.round_up(channels, 8), eps
        dev = h.device
        self.gamma = torch.ones(self.cp, dtype=F32, device=dev)
        self.beta = torch.zeros(self.cp, dtype=F32, device=dev)
        self.ggamma, self.gbeta = torch.zeros(self.cp, dtype=F32, device=dev), torch.zeros(self.cp, dtype=F32, device=dev)
        self.ref, self.ref_batch = None, None
        self.coef = torch.zeros(8, self.cp, dtype=F32, device=dev)
        self.scratch = torch.zeros(768, self.cp, dtype=F32, device=dev)

    def fwd(self, seq, z32, act=ACT_LRELU):
        """z32 [seq.rows, Cp] fp32 with the padding rows of every sequence ZERO and excluded from the statistics ->
        16-bit activation buffer at seq's level.  The statistics run over the B*L valid rows only."""
        h = self.h
        # the stream kernels take a flat [rows, N] matrix with a row pitch: gather the B*L valid rows once (the padding rows
        # of the fp32 pre-activation are not zero -- the GEMM writes bias + partial windows there -- and must not count)
        zc = self._compact(seq, z32)
        n = seq.B * seq.L
        if self.ref is None:
            self.ref = torch.zeros(2, self.cp, dtype=F32, device=h.device)
            self.ref_batch = seq.B
            h.vbn_stats(zc, n, self.cp, self.gamma, self.beta, self.coef, self.scratch, eps=self.eps, stats_out=self.ref)
            self._w = 1.0
        else:
            self._w = 1.0 / (self.ref_batch + 1.0)
            h.vbn_stats(zc, n, self.cp, self.gamma, self.beta, self.coef, self.scratch, eps=self.eps,
                        batch_weight=self._w, ref_stats=self.ref)
        self._zc, self._act = zc, act
        y_whole, y = seq.alloc(h, self.cp)
        yc = torch.zeros(n, self.cp, dtype=h.h16, device=h.device)
        h.affine_act_drop(zc, n, self.cp, self.coef[0], self.coef[1], act, 1.0, None, 0, yc)
        self._scatter(seq, yc, y)
        return y_whole

    def bwd(self, seq, da_whole):
        """da_whole: gradient wrt the activated output -> dz_whole (16-bit, padding rows zero); accumulates ggamma / gbeta."""
        h = self.h
        n = seq.B * seq.L
        dac = self._compact16(seq, da_whole[seq.GUARD:seq.GUARD + seq.rows])
        dzc = torch.zeros(n, self.cp, dtype=h.h16, device=h.device)
        h.vbn_bwd(dac, self._zc, n, self.cp, self._act, self._w, self.coef, self.ggamma, self.gbeta, dzc, self.scratch)
        dz_whole, dz = seq.alloc(h, self.cp)
        self._scatter(seq, dzc, dz)
        return dz_whole

    # valid rows <-> padded layout: plain strided copies (cudaMemcpy2D through torch's allocator-side copy engine)
    def _compact(self, seq, z32):
        out = torch.empty(seq.B * seq.L, self.cp, dtype=F32, device=self.h.device)
        out.view(seq.B, seq.L, self.cp).copy_(z32.view(seq.B, seq.S, self.cp)[:, :seq.L])
        return out

    def _compact16(self, seq, body):
        out = torch.empty(seq.B * seq.L, self.cp, dtype=self.h.h16, device=self.h.device)
        out.view(seq.B, seq.L, self.cp).copy_(body.view(seq.B, seq.S, self.cp)[:, :seq.L])
        return out

    def _scatter(self, seq, compact, body):
        body.view(seq.B, seq.S, self.cp)[:, :seq.L].copy_(compact.view(seq.B, seq.L, self.cp))
