"""Network definitions of the GAN hot path, expressed as sequences of C-ABI kernel calls.

Mirrors the reference's L2 layer:
  * generators  `lstm` (models/lstm.py:41-129), `res_lstm_l` (models/res_lstm_l.py:41-199),
    `res_lstm_base` (models/res_lstm_base.py:111-131,190) and the frame-level `dnn`
    (models/dnn.py:32-114) and `rced` (models/rced.py:34-119, splice = 1);
  * discriminators `discriminator_lstm` (models/discriminator_lstm.py:24-110) and
    `discriminator_dnn` (models/discriminator_dnn.py:21-98, applied per frame).
The reference hard-codes the layer sizes inside those files; here they are constructor
arguments whose defaults are the reference values.

Every activation lives time-major ([T*B, ld] rows = t*B + b) in zero-padded 16-bit buffers
(fp32 copies only where a residual or a loss needs them).  torch is the allocator only; all
arithmetic is done by librsrgan_sm100.so (ops.Handle).  No function here has a CPU path.
"""
from __future__ import annotations

import contextlib
import os

import torch

from . import packing, params
from .ops import ACT_LRELU, ACT_NONE, ACT_RELU

F32 = torch.float32


class Workspace(object):
    """Named device buffers, zero-filled at allocation, grown on demand (rows only).

    `generation` counts REPLACEMENTS of an existing buffer (a longer batch needed more rows): the old tensor is
    freed, so anything that recorded its address -- the CUDA graphs GAN_RNN captures per batch shape -- is stale
    from then on.  GAN_RNN compares generations before every replay and re-captures (gan_rnn._schedule_graphed).
    Growth is geometric so that a stream of slowly lengthening batches settles after a few replacements."""

    def __init__(self, handle):
        self.h = handle
        self.bufs = {}
        self.generation = 0

    def get(self, key, rows, cols, dtype):
        t = self.bufs.get(key)
        if t is None or t.shape[0] < rows or t.shape[1] != cols or t.dtype != dtype:
            alloc = rows
            if t is not None:
                self.generation += 1
                if t.shape[1] == cols and t.dtype == dtype:
                    alloc = max(rows, t.shape[0] + t.shape[0] // 4)
            t = torch.zeros(alloc, cols, dtype=dtype, device=self.h.device)
            self.bufs[key] = t
        return t[:rows]

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values())


class FC(object):
    """tf.contrib.layers.fully_connected over the last axis (models/lstm.py:82-87,121-124;
    models/discriminator_dnn.py:61-93; models/discriminator_lstm.py:100-104; models/dnn.py:79-110)."""

    def __init__(self, net, scope, n_in, n_out, act, cat=None):
        self.net, self.scope, self.n_in, self.n_out, self.act = net, scope, n_in, n_out, act
        self.inp, self.outp = packing.round_up(n_in, 8), packing.round_up(n_out, 8)
        self.wname, self.bname = scope + "/weights", scope + "/biases"
        self.cat = cat          # (n_a, n_b): the input is tf.concat([a, b], -1), stored [b | a] on the device

    def _wseg(self):
        if self.cat is not None:
            return params.fc_w_cat(self.wname, self.cat[0], self.cat[1], self.n_out)
        return params.fc_w(self.wname, self.n_in, self.n_out)

    def segs(self):
        return [self._wseg(), params.fc_b(self.bname, self.n_out)]

    def refresh(self):
        pass

    def fwd(self, ctx, x16, rows, want16=True, want32=False):
        net, h = self.net, self.net.h
        y16 = net.ws.get((ctx, self.scope, "y16"), rows, self.outp, h.h16) if want16 else None
        y32 = net.ws.get((ctx, self.scope, "y32"), rows, self.outp, F32) if want32 else None
        if self.n_out == 1 and self.act == ACT_NONE and want32 and not want16 and self.inp >= 64:
            # one output unit: a dot product per row (HBM stream) instead of a padded tensor-core tile
            h.fc1_fwd(x16, rows, self.inp, net.P.view(self.wname, "theta16"), net.P.view(self.bname), y32)
            return y16, y32
        h.gemm(x16, net.P.view(self.wname, "theta16"), rows, self.outp, self.inp, b_mn=True,
               bias=net.P.view(self.bname), act=self.act, out16=y16, out32=y32)
        return y16, y32

    def head_ok(self):
        """this layer is a one-unit head the fused head kernel applies to (rsr_fc1_head)"""
        return self.n_out == 1 and self.act == ACT_NONE and self.inp >= 64

    def head(self, ctx, x16, rows, spec, want_dx, x_act):
        """Forward, LSGAN loss terms, d loss / d logit and (want_dx) this layer's data gradient in one kernel.
        spec: dict(which, clip, d_real, d_fake, grad_target, gscale, losses, dlogit16); x_act: the activation that
        produced x16 when its derivative belongs to this layer's data gradient (plain FC producer), else ACT_NONE.
        Returns the fp32 logits; the data gradient is left where bwd(dx_done=True) expects it."""
        net, h = self.net, self.net.h
        y32 = net.ws.get((ctx, self.scope, "y32"), rows, self.outp, F32)
        dx16 = net.ws.get((ctx, self.scope, "dx16"), rows, self.inp, h.h16) if want_dx else None
        h.fc1_head(x16, rows, self.inp, net.P.view(self.wname, "theta16"), net.P.view(self.bname), spec["which"],
                   spec["clip"], spec["d_real"], spec["d_fake"], spec["grad_target"], spec["gscale"], spec["losses"], y32,
                   dlogit16=spec["dlogit16"], dact=x_act, dx16=dx16)
        return y32

    def bwd(self, ctx, x16, dy16, rows, want_dw=True, want_dx=True, prev_y16=None, prev_act=ACT_NONE,
            resid32=None, want32=False, dw_side=True, dx_done=False):
        """dy16: gradient wrt this layer's PRE-activation.  Returns the gradient wrt the input,
        multiplied by prev_act'(prev_y16) when the producer of x16 was an activated FC.
        dw_side=False keeps the weight gradient on the calling stream (last layer of a backward pass: the side stream
        still holds the previous layer's weight-gradient GEMMs and nothing else is left for the main stream to do)."""
        net, h = self.net, self.net.h
        dx16 = dx32 = None
        if want_dx and dx_done:      # the fused head kernel has written it (FC.head)
            dx16 = net.ws.get((ctx, self.scope, "dx16"), rows, self.inp, h.h16)
        elif want_dx:    # the producer layer waits for this: main stream first
            dx16 = net.ws.get((ctx, self.scope, "dx16"), rows, self.inp, h.h16)
            dx32 = net.ws.get((ctx, self.scope, "dx32"), rows, self.inp, F32) if want32 else None
            if self.n_out == 1 and resid32 is None and not want32 and self.inp >= 64:
                h.fc1_bwd_dx(dy16, rows, self.inp, net.P.view(self.wname, "theta16"), dx16, dact_src=prev_y16,
                             dact=prev_act)      # outer product masked by the producer's relu' (HBM stream)
            else:
                h.gemm(dy16, net.P.view(self.wname, "theta16"), rows, self.inp, self.outp, resid=resid32,
                       dact_src=prev_y16, dact=prev_act, out16=dx16, out32=dx32)
        if want_dw:
            with (h.side_stream() if dw_side else contextlib.nullcontext()):
                h.gemm(x16, dy16, self.inp, self.outp, rows, a_mn=True, b_mn=True, beta=1.0,
                       out32=net.P.view(self.wname, "grad"))
                h.colsum16(dy16, rows, self.outp, net.P.view(self.bname, "grad"), accumulate=True)
        return dx16, dx32


CTX_SALT = {"g": 0, "rl": 256, "fk": 512}     # dropout stream per layer call: salt = CTX_SALT[ctx] + layer index
LSTM_SALT = 16                                # ... + LSTM_SALT + l for the DropoutWrapper of LSTM layer l


class FCBN(FC):
    """fully_connected with its optional normalizer and dropout: tf.contrib.layers.fully_connected(x, n_out,
    activation_fn, normalizer_fn=batch_norm, normalizer_params={is_training, scale=True, renorm=True}) then
    tf.nn.dropout(keep_prob) -- models/dnn.py:56-62,79-102, models/discriminator_dnn.py:36-46,61-83,
    models/lstm.py:61-67,82-87.  With the normalizer the layer has BatchNorm/beta and BatchNorm/gamma instead of a bias.
    The tensor-core GEMM writes the fp32 pre-activation; statistics, normalisation, activation, dropout and their
    gradients are the HBM-stream kernels of csrc/batchnorm.cu.  `bwd` takes the gradient wrt the layer OUTPUT."""

    STATE_KEYS = ("moving_mean", "moving_variance", "renorm_mean", "renorm_stddev", "renorm_mean_weight",
                  "renorm_stddev_weight")

    def __init__(self, net, scope, n_in, n_out, act, bn, index, drop=True, cat=None):
        super(FCBN, self).__init__(net, scope, n_in, n_out, act, cat=cat)
        self.bn, self.index, self.drop = bool(bn), index, drop
        self.betaname, self.gname = scope + "/BatchNorm/beta", scope + "/BatchNorm/gamma"
        self.state = None
        if self.bn:
            # moving_mean 0, moving_variance 1, the four renorm variables 0 (TF 1.4 zero-initialises them)
            self.state = torch.zeros(6, self.outp, dtype=F32, device=net.h.device)
            self.state[1].fill_(1.0)
        self._mode = {}

    def segs(self):
        if not self.bn:
            return super(FCBN, self).segs()
        return [self._wseg(), params.fc_b(self.betaname, self.n_out), params.fc_b(self.gname, self.n_out)]

    def _bufs(self, ctx, rows):
        ws = self.net.ws
        return (ws.get((ctx, self.scope, "z32"), rows, self.outp, F32),
                ws.get((ctx, self.scope, "bn_coef"), 8, self.outp, F32),
                ws.get((ctx, self.scope, "bn_scratch"), 768, self.outp, F32))

    def fwd(self, ctx, x16, rows, want16=True, want32=False):
        net, h = self.net, self.net.h
        z32, coef, scratch = self._bufs(ctx, rows)
        y16 = net.ws.get((ctx, self.scope, "y16"), rows, self.outp, h.h16)
        # training-mode batch_norm: the GEMM's epilogue leaves the (count, mean, M2) of every 128-row block and column in
        # `scratch` (rsr_gemm_args.stats), so the statistics need only the fixed-order merge of those partials -- the
        # pre-activation is not read a second time (models/dnn.py:56-62, models/discriminator_dnn.py:36-46)
        epi_stats = (self.bn and net.training and hasattr(h, "bn_train_finish") and self.outp % 32 == 0
                     and rows <= h.BN_STATS_ROWS_MAX and os.environ.get("RSR_NO_EPILOGUE_STATS") != "1")
        h.gemm(x16, net.P.view(self.wname, "theta16"), rows, self.outp, self.inp, b_mn=True, out32=z32,
               **(dict(stats=scratch) if epi_stats else {}))
        keep = net.keep_prob if net.training and self.drop else 1.0
        if self.bn:
            if epi_stats:
                h.bn_train_finish((rows + 127) // 128, rows, self.outp, net.P.view(self.gname), net.P.view(self.betaname),
                                  self.state, coef, scratch, update_state=net.bn_update)
            elif net.training:
                h.bn_train_stats(z32, rows, self.outp, net.P.view(self.gname), net.P.view(self.betaname), self.state,
                                 coef, scratch, update_state=net.bn_update)
            else:
                h.bn_eval_coef(self.outp, net.P.view(self.gname), net.P.view(self.betaname), self.state, coef)
            A, Bc = coef[0], coef[1]
        else:
            A, Bc = None, net.P.view(self.bname)
        self._mode[ctx] = keep
        h.affine_act_drop(z32, rows, self.outp, A, Bc, self.act, keep, net.rng, CTX_SALT[ctx] + self.index, y16)
        return y16, None

    def bwd(self, ctx, x16, da16, rows, want_dw=True, want_dx=True, resid32=None, want32=False, dw_side=True, **_):
        net, h = self.net, self.net.h
        z32, coef, scratch = self._bufs(ctx, rows)
        dz16 = net.ws.get((ctx, self.scope, "dz16"), rows, self.outp, h.h16)
        P = net.P
        if self.bn:
            h.bn_bwd(da16, z32, rows, self.outp, self.act, self._mode[ctx], net.rng, CTX_SALT[ctx] + self.index, True,
                     coef, None, P.view(self.gname, "grad") if want_dw else None,
                     P.view(self.betaname, "grad") if want_dw else None, dz16, scratch)
        else:
            h.bn_bwd(da16, z32, rows, self.outp, self.act, self._mode[ctx], net.rng, CTX_SALT[ctx] + self.index, False,
                     None, P.view(self.bname), None, P.view(self.bname, "grad") if want_dw else None, dz16, scratch)
        dx16 = dx32 = None
        if want_dx:
            dx16 = net.ws.get((ctx, self.scope, "dx16"), rows, self.inp, h.h16)
            dx32 = net.ws.get((ctx, self.scope, "dx32"), rows, self.inp, F32) if want32 else None
            h.gemm(dz16, P.view(self.wname, "theta16"), rows, self.inp, self.outp, resid=resid32, out16=dx16,
                   out32=dx32)
        if want_dw:
            with (h.side_stream() if dw_side else contextlib.nullcontext()):
                h.gemm(x16, dz16, self.inp, self.outp, rows, a_mn=True, b_mn=True, beta=1.0,
                       out32=P.view(self.wname, "grad"))
        return dx16, dx32

    # -- non-trainable variables <-> TF names ------------------------------------------------
    def export_state(self):
        out = {}
        if self.bn:
            st = self.state.detach().cpu().numpy()
            for i, k in enumerate(self.STATE_KEYS):
                out[self.scope + "/BatchNorm/" + k] = st[i, :self.n_out].copy() if i < 4 else st[i, 0].copy()
        return out

    def load_state(self, d):
        if self.bn:
            for i, k in enumerate(self.STATE_KEYS):
                v = torch.as_tensor(d[self.scope + "/BatchNorm/" + k], dtype=F32)
                if i < 4:
                    self.state[i].zero_()
                    self.state[i, :self.n_out] = v.to(self.state.device)
                else:
                    self.state[i].fill_(float(v))
            self.state[1, self.n_out:] = 1.0


class LSTMP(object):
    """tf.contrib.rnn.LSTMCell(C, use_peepholes=True, num_proj=P, forget_bias=1.0) under
    tf.nn.dynamic_rnn(sequence_length) -- models/lstm.py:89-112, models/res_lstm_l.py:86-138,
    models/discriminator_lstm.py:70-91.  Three kernels: hoisted input GEMM (X K_x + b), the
    persistent recurrence with Wc = W_proj K_h resident on chip, hoisted projection GEMM."""

    def __init__(self, net, prefix, I, C, P):
        self.net, self.prefix, self.I, self.C, self.P = net, prefix, I, C, P
        self.Ip, self.Pp, self.Cp = packing.round_up(I, 8), packing.round_up(P, 8), packing.cell_pad(C)
        h = net.h
        self.wc16 = torch.zeros(self.Cp, 4 * self.Cp, dtype=h.h16, device=h.device)    # Wc   (backward operand)
        self.wcT16 = torch.zeros(4 * self.Cp, self.Cp, dtype=h.h16, device=h.device)   # Wc^T (forward operand)
        self.scratch = torch.zeros(7 * self.Cp, dtype=F32, device=h.device)            # sink for unwanted db/dw
        self.Ik = packing.round_up(I, 16)
        self.kxT16 = torch.zeros(4 * self.Cp, self.Ik, dtype=h.h16, device=h.device)   # K_x^T (fused forward operand)
        self.fused = True            # cleared the first time the library says the fused variant does not apply
        self.wpT16 = torch.zeros(self.Pp, self.Cp, dtype=h.h16, device=h.device)       # W_p^T (layer-wavefront operand)
        self.wave = False            # set by the network on the first layer of a stacked pair (fwd_wave / bwd_wave) ...
        self.wave_next = None        # ... together with the layer stacked on it
        self.wave_declined = set()   # batch sizes the library declined the wavefront launch for: ("f" | "b", B)
        self.fT16 = None             # F = W_p K_x(next layer) [Cp, 4Cp]: operand of the backward wavefront
        if self.Cp > 512 and hasattr(h, "overlap"):
            # L2-exchange recurrence kernels (Cp > 512) spin on counters of co-resident CTAs: nothing else
            # may take SMs while they run, so the side stream is switched off for this model
            h.overlap = False

    def segs(self):
        return params.lstm_cell(self.prefix, self.I, self.C, self.P)

    def rec_flops(self, B, T):
        """ALGORITHMIC flops of the recurrent half per sequence: m_{t-1} K_h plus the
        projection, 2*B*(P*4C + C*P) per step -- the unfolded reference form, not the folded
        B x C x 4C product the kernel executes."""
        return 2.0 * B * T * (self.P * 4 * self.C + self.C * self.P)

    def _w(self, buf="theta16"):
        P = self.net.P
        K = P.view(self.prefix + "kernel", buf)
        return K[:self.Ip], K[self.Ip:], P.view(self.prefix + "projection/kernel", buf)

    def refresh(self):
        """Wc = W_proj K_h after every weight update (fp32 accumulate, rounded once to 16 bit)."""
        h = self.net.h
        _, Kh16, Wp16 = self._w()
        h.gemm(Wp16, Kh16, self.Cp, 4 * self.Cp, self.Pp, b_mn=True, out16=self.wc16)
        h.gemm(Kh16, Wp16, 4 * self.Cp, self.Cp, self.Pp, a_mn=True, out16=self.wcT16)
        if self.fused:
            Kx16 = self._w()[0]
            h.transpose16(Kx16, self.Ip, 4 * self.Cp, self.kxT16)
            if self.wave:
                h.transpose16(Wp16, self.Cp, self.Pp, self.wpT16)
                if self.wave_next is not None:
                    if self.fT16 is None:
                        self.fT16 = torch.zeros(self.Cp, 4 * self.Cp, dtype=h.h16, device=h.device)
                    h.gemm(Wp16, self.wave_next._w()[0], self.Cp, 4 * self.Cp, self.Pp, b_mn=True, out16=self.fT16)

    def fwd(self, ctx, x16, B, T, lengths, save=True, want32=False):
        """x16 [T*B, Ip] -> out_seq16 [(T+1)*B, Pp] (slot 0 = zero initial state; rows B.. are
        the outputs) and optionally out32 [T*B, Pp]."""
        net, h, P = self.net, self.net.h, self.net.P
        rows, Cp = T * B, self.Cp
        key = (ctx, self.prefix, B)
        Kx16, _, Wp16 = self._w()
        mt = net.ws.get(key + ("mt",), rows + B, Cp, h.h16)
        out = net.ws.get(key + ("out",), rows + B, self.Pp, h.h16)
        sv = net.ws.get(key + ("save",), rows, 5 * Cp, F32) if save else None
        o32 = net.ws.get(key + ("o32",), rows, self.Pp, F32) if want32 else None
        peep = (P.view(self.prefix + "w_i_diag"), P.view(self.prefix + "w_f_diag"), P.view(self.prefix + "w_o_diag"))
        done = False
        if self.fused:   # input GEMM + recurrent GEMM + gate epilogue in one kernel, no Zx round trip through HBM
            done = h.lstmp_fused_fwd(B, T, self.I, Cp, x16, self.kxT16, P.view(self.prefix + "bias"), self.wcT16,
                                     peep[0], peep[1], peep[2], lengths, mt, sv,
                                     work=self.rec_flops(B, T) + 2.0 * rows * self.I * 4 * self.C)
            self.fused = done
        if not done:
            zx = net.ws.get((ctx, "zx", Cp, B), rows, 4 * Cp, F32)        # shared by layers of equal width
            h.gemm(x16, Kx16, rows, 4 * Cp, self.Ip, b_mn=True, bias=P.view(self.prefix + "bias"), out32=zx)
            h.lstmp_rec_fwd(B, T, Cp, zx, self.wcT16, peep[0], peep[1], peep[2], lengths, mt, sv,
                            work=self.rec_flops(B, T))
        h.gemm(mt[B:], Wp16, rows, self.Pp, Cp, b_mn=True, out16=out[B:], out32=o32)
        return out, o32

    def _fused_operands(self):
        P = self.net.P
        return (self.kxT16, P.view(self.prefix + "bias"), self.wcT16, P.view(self.prefix + "w_i_diag"),
                P.view(self.prefix + "w_f_diag"), P.view(self.prefix + "w_o_diag"))

    def fwd_wave(self, nxt, ctx, x16, B, T, lengths, save=True, want32=False):
        """This layer and the one stacked on it (`nxt`) as ONE wavefront launch: layer 2 runs a few time steps behind
        layer 1 instead of after it (rsr_lstmp_wave_fwd).  Returns (out_seq16 of this layer, out_seq16 of nxt, out32 of
        nxt or None) -- or None when the launch does not apply (the caller then runs the two layers one by one)."""
        net, h = self.net, self.net.h
        if not (self.wave and self.fused and nxt.fused and self.Cp == nxt.Cp and nxt.I == self.P and nxt.Ip == self.Pp) \
                or ("f", B) in self.wave_declined:
            return None
        rows, Cp = T * B, self.Cp
        k1, k2 = (ctx, self.prefix, B), (ctx, nxt.prefix, B)
        mt1 = net.ws.get(k1 + ("mt",), rows + B, Cp, h.h16)
        out1 = net.ws.get(k1 + ("out",), rows + B, self.Pp, h.h16)
        sv1 = net.ws.get(k1 + ("save",), rows, 5 * Cp, F32) if save else None
        mt2 = net.ws.get(k2 + ("mt",), rows + B, Cp, h.h16)
        out2 = net.ws.get(k2 + ("out",), rows + B, nxt.Pp, h.h16)
        sv2 = net.ws.get(k2 + ("save",), rows, 5 * Cp, F32) if save else None
        o32 = net.ws.get(k2 + ("o32",), rows, nxt.Pp, F32) if want32 else None
        work = (self.rec_flops(B, T) + nxt.rec_flops(B, T) + 2.0 * rows * (self.I + nxt.I) * 4 * self.C)
        if not h.lstmp_wave_fwd(B, T, Cp, self.I, self.P, lengths, x16, self._fused_operands(), mt1, sv1, self.wpT16, out1,
                                nxt._fused_operands(), mt2, sv2, work=work):
            self.wave_declined.add(("f", B))
            return None
        h.gemm(mt2[B:], nxt._w()[2], rows, nxt.Pp, Cp, b_mn=True, out16=out2[B:], out32=o32)
        return out1, out2, o32

    def _peep(self):
        P = self.net.P
        return tuple(P.view(self.prefix + n) for n in ("w_i_diag", "w_f_diag", "w_o_diag"))

    def _rec_grads(self):
        P = self.net.P
        return tuple(P.view(self.prefix + n, "grad") for n in ("bias", "w_i_diag", "w_f_diag", "w_o_diag"))

    def bwd_wave(self, upper, ctx, x16, x16_upper, dout32_upper, B, T, lengths, prev_y16=None, prev_act=ACT_NONE,
                 want32=False, want_dw=True, want_dx=True, resid32=None):
        """Backward of this layer and the one stacked on it (`upper`, whose bwd_pre has run) as ONE wavefront launch
        (rsr_lstmp_wave_bwd), then this layer's data gradient on the calling stream and both layers' weight gradients
        on the side stream.  Returns (dx16, dx32) of this layer -- or None when the launch does not apply."""
        net, h = self.net, self.net.h
        if not (self.wave and self.wave_next is upper and self.fT16 is not None and self.Cp == upper.Cp
                and upper.Ip == self.Pp) or ("b", B) in self.wave_declined:
            return None
        rows, Cp = T * B, self.Cp
        k1, k2 = (ctx, self.prefix, B), (ctx, upper.prefix, B)
        sv1 = net.ws.get(k1 + ("save",), rows, 5 * Cp, F32)
        sv2 = net.ws.get(k2 + ("save",), rows, 5 * Cp, F32)
        dz1 = net.ws.get(k1 + ("dz",), rows + B, 4 * Cp, h.h16)
        dz2 = net.ws.get(k2 + ("dz",), rows + B, 4 * Cp, h.h16)
        dmt2 = net.ws.get((ctx, "dmt", Cp, B), rows, Cp, F32)
        part = net.ws.get((ctx, "wave_part", Cp, B), T * (B + 48), Cp, F32)   # zero at allocation and after every launch
        h.fill32(dz1[rows:].view(F32), 0.0)                # dz_{T} = 0 (no step after the last one; T varies per batch)
        h.fill32(dz2[rows:].view(F32), 0.0)

        def sink(l):
            s = l.scratch
            return (s[:4 * Cp], s[4 * Cp:5 * Cp], s[5 * Cp:6 * Cp], s[6 * Cp:])
        g1, g2 = (self._rec_grads(), upper._rec_grads()) if want_dw else (sink(self), sink(upper))
        if not h.lstmp_wave_bwd(B, T, Cp, lengths, dmt2, (upper.wc16,) + upper._peep(), sv2, dz2, g2,
                                self.fT16, part, (self.wc16,) + self._peep(), sv1, dz1, g1,
                                work=self.rec_flops(B, T) + upper.rec_flops(B, T),
                                # weight gradients that hide behind the per-layer launches on the side stream: only the
                                # 32-utterance launch pays (cfg-2, B = 128: 3.089 ms per schedule with the 48-utterance
                                # launch against 3.057 without)
                                max_nbp=32 if (want_dw and getattr(h, "overlap", False)) else 0):
            self.wave_declined.add(("b", B))
            return None
        dx16 = dx32 = None
        if want_dx:
            dx16 = net.ws.get(k1 + ("dx16",), rows, self.Ip, h.h16)
            dx32 = net.ws.get(k1 + ("dx32",), rows, self.Ip, F32) if want32 else None
            h.gemm(dz1, self._w()[0], rows, self.Ip, 4 * Cp, resid=resid32, dact_src=prev_y16, dact=prev_act,
                   out16=dx16, out32=dx32)
        if want_dw:
            upper.bwd_side(ctx, x16_upper, dout32_upper, B, T)
            with h.side_stream():      # gradient wrt this layer's output, for its projection's weight gradient only
                du16 = net.ws.get(k2 + ("dx16",), rows, upper.Ip, h.h16)
                du32 = net.ws.get(k2 + ("dx32",), rows, upper.Ip, F32)
                h.gemm(dz2, upper._w()[0], rows, upper.Ip, 4 * Cp, out16=du16, out32=du32)
            self.bwd_side(ctx, x16, du32, B, T)
        return dx16, dx32

    def bwd(self, ctx, x16, dout16, dout32, B, T, lengths, want_dw=True, want_dx=True, prev_y16=None,
            prev_act=ACT_NONE, resid32=None, want32=False):
        """dout16/dout32: gradient wrt the layer output [T*B, Pp] (dout32 needed iff want_dw)."""
        self.bwd_pre(ctx, dout16, B, T)
        dx16, dx32 = self.bwd_main(ctx, dout16, B, T, lengths, want_dw=want_dw, want_dx=want_dx, prev_y16=prev_y16,
                                   prev_act=prev_act, resid32=resid32, want32=want32)
        if want_dw:
            self.bwd_side(ctx, x16, dout32, B, T)
        return dx16, dx32

    # The backward pass in three pieces, so that a network can enqueue the NEXT layer's bwd_pre (a small GEMM on the
    # critical path) before THIS layer's weight-gradient GEMMs start competing for SMs on the side stream.
    def bwd_pre(self, ctx, dout16, B, T):
        """dmt = dOut W_proj^T ; the recurrence kernel adds dz_{t+1} Wc^T"""
        net, h = self.net, self.net.h
        rows, Cp = T * B, self.Cp
        dmt = net.ws.get((ctx, "dmt", Cp, B), rows, Cp, F32)
        h.gemm(dout16, self._w()[2], rows, Cp, self.Pp, out32=dmt)

    def bwd_main(self, ctx, dout16, B, T, lengths, want_dw=True, want_dx=True, prev_y16=None, prev_act=ACT_NONE,
                 resid32=None, want32=False, after_rec=None):
        """after_rec(mark): called just in front of the launch of the recurrence kernel with a mark of that point -- the
        place to enqueue side-stream work that should take the SMs the recurrence leaves free rather than the ones it is
        about to need."""
        net, h, P = self.net, self.net.h, self.net.P
        rows, Cp = T * B, self.Cp
        key = (ctx, self.prefix, B)
        Kx16, Kh16, Wp16 = self._w()
        sv = net.ws.get(key + ("save",), rows, 5 * Cp, F32)
        dmt = net.ws.get((ctx, "dmt", Cp, B), rows, Cp, F32)
        dz = net.ws.get(key + ("dz",), rows + B, 4 * Cp, h.h16)       # per layer: the side stream reads it after we return
        h.fill32(dz[rows:].view(F32), 0.0)                 # dz_{T} = 0 (no step after the last one; T varies per batch)
        if after_rec is not None:
            # In a captured graph only dependencies order kernels: side work that became ready a microsecond before this
            # recurrence (a persistent 148-CTA GEMM) held its cluster slots for the whole GEMM (~22 us).  So the side
            # stream is made to wait for THIS point and to start with a one-block spacer kernel: the recurrence's CTAs
            # are placed first, the GEMM takes the SMs it leaves free.
            after_rec(h.mark())
        if want_dw:
            gb = P.view(self.prefix + "bias", "grad")
            gi, gf, go = (P.view(self.prefix + n, "grad") for n in ("w_i_diag", "w_f_diag", "w_o_diag"))
        else:
            s = self.scratch
            gb, gi, gf, go = s[:4 * Cp], s[4 * Cp:5 * Cp], s[5 * Cp:6 * Cp], s[6 * Cp:]
        h.lstmp_rec_bwd(B, T, Cp, dmt, self.wc16, P.view(self.prefix + "w_i_diag"),
                        P.view(self.prefix + "w_f_diag"), P.view(self.prefix + "w_o_diag"), lengths, sv,
                        dz, gb, gi, gf, go, work=self.rec_flops(B, T))
        self.rec_mark = h.mark()       # dz is complete here: all the weight gradients of this layer need
        dx16 = dx32 = None
        if want_dx:    # the next (earlier) layer waits for this: main stream, before the weight gradients
            dx16 = net.ws.get(key + ("dx16",), rows, self.Ip, h.h16)
            dx32 = net.ws.get(key + ("dx32",), rows, self.Ip, F32) if want32 else None
            h.gemm(dz, Kx16, rows, self.Ip, 4 * Cp, resid=resid32, dact_src=prev_y16, dact=prev_act,
                   out16=dx16, out32=dx32)
        return dx16, dx32

    def bwd_side(self, ctx, x16, dout32, B, T, after=None):
        """weight gradients, on the side stream: they overlap the next layer's recurrence.  after: a Handle.mark() the
        side stream waits for instead of the calling stream's tail."""
        net, h, P = self.net, self.net.h, self.net.P
        rows, Cp = T * B, self.Cp
        key = (ctx, self.prefix, B)
        Kx16, Kh16, Wp16 = self._w()
        mt = net.ws.get(key + ("mt",), rows + B, Cp, h.h16)
        out = net.ws.get(key + ("out",), rows + B, self.Pp, h.h16)
        dz = net.ws.get(key + ("dz",), rows + B, 4 * Cp, h.h16)
        with h.side_stream(after=after):
            gK = P.view(self.prefix + "kernel", "grad")
            # dK = [x_t , m_{t-1}]^T dz_t  (two row blocks of the TF kernel)
            h.gemm(x16, dz, self.Ip, 4 * Cp, rows, a_mn=True, b_mn=True, beta=1.0, out32=gK[:self.Ip])
            h.gemm(out, dz, self.Pp, 4 * Cp, rows, a_mn=True, b_mn=True, beta=1.0, out32=gK[self.Ip:])
            # dW_proj = mt_t^T (dOut_t + dz_{t+1} K_h^T)
            dmtot = net.ws.get((ctx, "dmtot", self.Pp, B), rows, self.Pp, h.h16)
            h.gemm(dz[B:], Kh16, rows, self.Pp, 4 * Cp, resid=dout32, out16=dmtot)
            h.gemm(mt[B:], dmtot, Cp, self.Pp, rows, a_mn=True, b_mn=True, beta=1.0,
                   out32=P.view(self.prefix + "projection/kernel", "grad"))


class ConvBN(object):
    """normalizer_fn=batch_norm on a convolution of the frame layout (models/rced.py:63-71,94-97: {is_training,
    scale=True, renorm=True} on all nine conv2d, which then have BatchNorm/beta and BatchNorm/gamma instead of a bias).
    The GEMM writes the fp32 pre-activation; moments pool every (frame, line, position) of a channel
    (rsr_bn_train_stats_lines), the normalise kernel also zeroes the padding rows, and the backward kernel takes the
    gradient wrt the layer OUTPUT (include/rsrgan_b200.h, "batch_norm behind the convolutions")."""

    STATE_KEYS = FCBN.STATE_KEYS

    def _bn_init(self, bn, lines, chans):
        self.bn, self.bn_H, self.bn_C = bool(bn), lines, chans
        self.betaname, self.gname = self.scope + "/BatchNorm/beta", self.scope + "/BatchNorm/gamma"
        self.state = None
        if self.bn:
            self.state = torch.zeros(6, packing.round_up(chans, 8), dtype=F32, device=self.net.h.device)
            self.state[1].fill_(1.0)

    def _bn_bufs(self, ctx, frames):
        ws, rows = self.net.ws, frames * self.fl.S
        return (ws.get((ctx, self.scope, "z32", frames), rows, self.cop, F32),
                ws.get((ctx, self.scope, "bn_coef"), 8, self.cop, F32),
                ws.get((ctx, self.scope, "bn_scratch"), 768, self.cop, F32))

    def _bn_fwd(self, ctx, x_whole, taps16, frames):
        net, h, fl = self.net, self.net.h, self.fl
        y_whole, y = fl.buf(net, (ctx, self.scope, "y16", frames), frames, self.cop)
        z32, coef, scratch = self._bn_bufs(ctx, frames)
        h.gemm(fl.window(x_whole, frames, self.cip, self.W), taps16, frames * fl.S, self.cop, self.W * self.cip, b_mn=True,
               out32=z32)
        gamma, beta = net.P.view(self.gname), net.P.view(self.betaname)
        if net.training:
            h.bn_train_stats_lines(z32, frames, fl.S, fl.L, self.bn_H, self.bn_C, self.cop, gamma, beta, self.state, coef,
                                   scratch, update_state=net.bn_update)
        else:
            h.bn_eval_coef_lines(self.cop, self.bn_H, self.bn_C, gamma, beta, self.state, coef)
        h.affine_act_lines(z32, frames, fl.S, fl.L, self.cop, coef[0], coef[1], ACT_RELU, y)
        return y_whole

    def _bn_bwd(self, ctx, da_whole, frames):
        """da_whole: gradient wrt this layer's OUTPUT -> (whole, data rows) of the gradient wrt its pre-activation"""
        net, h, fl, P = self.net, self.net.h, self.fl, self.net.P
        z32, coef, scratch = self._bn_bufs(ctx, frames)
        dz_whole, dz = fl.buf(net, (ctx, self.scope, "dz16", frames), frames, self.cop)
        h.bn_bwd_lines(da_whole[fl.GUARD:fl.GUARD + frames * fl.S], z32, frames, fl.S, fl.L, self.bn_H, self.bn_C, self.cop,
                       ACT_RELU, coef, P.view(self.gname, "grad"), P.view(self.betaname, "grad"), dz, scratch)
        return dz_whole, dz

    def export_state(self):
        out = {}
        if self.bn:
            st = self.state.detach().cpu().numpy()
            for i, k in enumerate(self.STATE_KEYS):
                out[self.scope + "/BatchNorm/" + k] = st[i, :self.bn_C].copy() if i < 4 else st[i, 0].copy()
        return out

    def load_state(self, d):
        if self.bn:
            for i, k in enumerate(self.STATE_KEYS):
                v = torch.as_tensor(d[self.scope + "/BatchNorm/" + k], dtype=F32)
                if i < 4:
                    self.state[i].zero_()
                    self.state[i, :self.bn_C] = v.to(self.state.device)
                else:
                    self.state[i].fill_(float(v))
            self.state[1, self.bn_C:] = 1.0


class ConvFrames(object):
    """The frame layout shared by the layers of the convolutional generator (models/rced.py:46-57, splice = 1):
    frame r = rows [r*S, r*S+S) of a channels-last 16-bit buffer, positions 0..L-1 data, rows L..S-1 zero (the
    SAME padding shared with the next frame), GUARD zero rows before the first / after the last frame."""
    GUARD = 8

    def __init__(self, L, max_width):
        self.L = L
        self.S = packing.round_up(L + max_width // 2, 8)
        assert max_width // 2 <= self.GUARD

    def buf(self, net, key, frames, cp):
        """-> (whole buffer incl. guards, view of the frames*S data rows)"""
        t = net.ws.get(key, frames * self.S + 2 * self.GUARD, cp, net.h.h16)
        return t, t[self.GUARD:self.GUARD + frames * self.S]

    def window(self, whole, frames, cp, width):
        """Overlapped view A[m, k*cp + c] = act[m - width//2 + k, c]: the GEMM A operand of a SAME convolution
        (include/rsrgan_b200.h, "1-D convolution family").  No copy."""
        return whole.as_strided((frames * self.S, width * cp), (cp, 1),
                                whole.storage_offset() + (self.GUARD - width // 2) * cp)


class Conv1dSame(ConvBN):
    """tf.contrib.layers.conv2d(inputs, C_out, [1, w], padding=SAME, activation_fn=relu) -- models/rced.py:94-101 with
    splice = 1.  Forward, data gradient and weight gradient are rsr_gemm calls over overlapped views.
    Without a normalizer `bwd` takes the gradient wrt this layer's PRE-activation and hands the producer the same; with
    batch_norm (ConvBN) both are gradients wrt the layer OUTPUT."""

    def __init__(self, net, scope, fl, width, c_in, c_out, bn=False):
        self.net, self.scope, self.fl, self.W, self.c_in, self.c_out = net, scope, fl, width, c_in, c_out
        self.cip, self.cop = packing.round_up(c_in, 8), packing.round_up(c_out, 8)
        self.wname, self.bname = scope + "/weights", scope + "/biases"
        h = net.h
        self.wflip16 = torch.zeros(width * self.cop, self.cip, dtype=h.h16, device=h.device)
        self._bn_init(bn, 1, c_out)

    def _tail_segs(self):
        if self.bn:
            return [params.fc_b(self.betaname, self.c_out), params.fc_b(self.gname, self.c_out)]
        return [params.fc_b(self.bname, self.c_out)]

    def segs(self):
        return [params.conv_w(self.wname, self.W, self.c_in, self.c_out)] + self._tail_segs()

    def refresh(self):
        """taps of the transposed convolution (data gradient) after every weight update"""
        self.net.h.conv_w_flip(self.net.P.view(self.wname, "theta16"), self.W, self.cip, self.cop, self.wflip16)

    def fwd(self, ctx, x_whole, frames):
        net, h, fl = self.net, self.net.h, self.fl
        if self.bn:
            return self._bn_fwd(ctx, x_whole, net.P.view(self.wname, "theta16"), frames)
        y_whole, y = fl.buf(net, (ctx, self.scope, "y16", frames), frames, self.cop)
        rows = frames * fl.S
        h.gemm(fl.window(x_whole, frames, self.cip, self.W), net.P.view(self.wname, "theta16"), rows, self.cop,
               self.W * self.cip, b_mn=True, bias=net.P.view(self.bname), act=ACT_RELU, out16=y)
        h.conv_mask_rows(y, frames, fl.S, fl.L, self.cop)
        return y_whole

    def bwd(self, ctx, x_whole, dy_whole, frames, want_dx=True):
        """dy_whole: gradient wrt this layer's PRE-activation (padding rows zero).  Returns the gradient wrt the
        producer's pre-activation (times relu'(x), which also zeroes its padding rows)."""
        net, h, fl = self.net, self.net.h, self.fl
        rows = frames * fl.S
        G = fl.GUARD
        if self.bn:
            dy_whole, dy = self._bn_bwd(ctx, dy_whole, frames)
        else:
            dy = dy_whole[G:G + rows]
        dx_whole = None
        if want_dx:
            dx_whole, dx = fl.buf(net, (ctx, self.scope, "dx16", frames), frames, self.cip)
            h.gemm(fl.window(dy_whole, frames, self.cop, self.W), self.wflip16, rows, self.cip, self.W * self.cop,
                   b_mn=True, dact_src=None if self.bn else x_whole[G:G + rows],
                   dact=ACT_NONE if self.bn else ACT_RELU, out16=dx)
        with h.side_stream():
            h.gemm(fl.window(x_whole, frames, self.cip, self.W), dy, self.W * self.cip, self.cop, rows, a_mn=True,
                   b_mn=True, beta=1.0, out32=net.P.view(self.wname, "grad"))
            if not self.bn:
                h.colsum16(dy, rows, self.cop, net.P.view(self.bname, "grad"), accumulate=True)
        return dx_whole


class Conv2dLines(Conv1dSame):
    """tf.contrib.layers.conv2d(inputs, C_out, [splice, w], padding=SAME, relu) over frames of H = splice stacked lines
    (models/rced.py:46-57,94-101; run_dnn.sh:129-140 trains 40 bins x 11 lines).  The lines are CHANNELS of one position
    (channel = line * C + c), so the layer is Conv1dSame with block-Toeplitz taps (include/rsrgan_b200.h, "[splice, w]
    convolutions"): the flat parameter buffer holds TensorFlow's compact filter (kh, w, C_in, C_out) and per-channel bias,
    the GEMM operands (taps, flipped taps, tiled bias) are derived after every update, and the gradient of the expansion is
    folded back over the tied copies in a fixed order."""

    def __init__(self, net, scope, fl, kh, width, c_in, c_out, lines, bn=False):
        self.net, self.scope, self.fl, self.W, self.c_in, self.c_out = net, scope, fl, width, c_in, c_out
        self.kh, self.H = kh, lines
        self.cip, self.cop = packing.round_up(lines * c_in, 8), packing.round_up(lines * c_out, 8)
        self.wname, self.bname = scope + "/weights", scope + "/biases"
        h = net.h
        self.taps16 = torch.zeros(width * self.cip, self.cop, dtype=h.h16, device=h.device)
        self.wflip16 = torch.zeros(width * self.cop, self.cip, dtype=h.h16, device=h.device)
        self.dw2 = torch.zeros(width * self.cip, self.cop, dtype=F32, device=h.device)
        self.bias_t = torch.zeros(self.cop, dtype=F32, device=h.device)
        self.db_t = torch.zeros(self.cop, dtype=F32, device=h.device)
        self._bn_init(bn, lines, c_out)

    def segs(self):
        return [params.conv_w2d(self.wname, self.kh, self.W, self.c_in, self.c_out)] + self._tail_segs()

    def refresh(self):
        h, P = self.net.h, self.net.P
        h.conv_toeplitz_expand(P.view(self.wname, "theta16"), self.kh, self.W, self.c_in, self.c_out, self.H, self.cip,
                               self.cop, self.taps16)
        h.conv_w_flip(self.taps16, self.W, self.cip, self.cop, self.wflip16)
        if not self.bn:
            h.vec_tile(P.view(self.bname), self.c_out, self.H, self.cop, self.bias_t)

    def fwd(self, ctx, x_whole, frames):
        net, h, fl = self.net, self.net.h, self.fl
        if self.bn:
            return self._bn_fwd(ctx, x_whole, self.taps16, frames)
        y_whole, y = fl.buf(net, (ctx, self.scope, "y16", frames), frames, self.cop)
        rows = frames * fl.S
        h.gemm(fl.window(x_whole, frames, self.cip, self.W), self.taps16, rows, self.cop, self.W * self.cip, b_mn=True,
               bias=self.bias_t, act=ACT_RELU, out16=y)
        h.conv_mask_rows(y, frames, fl.S, fl.L, self.cop)
        return y_whole

    def bwd(self, ctx, x_whole, dy_whole, frames, want_dx=True):
        net, h, fl, P = self.net, self.net.h, self.fl, self.net.P
        rows = frames * fl.S
        G = fl.GUARD
        if self.bn:
            dy_whole, dy = self._bn_bwd(ctx, dy_whole, frames)
        else:
            dy = dy_whole[G:G + rows]
        dx_whole = None
        if want_dx:
            dx_whole, dx = fl.buf(net, (ctx, self.scope, "dx16", frames), frames, self.cip)
            h.gemm(fl.window(dy_whole, frames, self.cop, self.W), self.wflip16, rows, self.cip, self.W * self.cop,
                   b_mn=True, dact_src=None if self.bn else x_whole[G:G + rows],
                   dact=ACT_NONE if self.bn else ACT_RELU, out16=dx)
        with h.side_stream():
            h.fill32(self.dw2, 0.0)
            h.gemm(fl.window(x_whole, frames, self.cip, self.W), dy, self.W * self.cip, self.cop, rows, a_mn=True,
                   b_mn=True, beta=1.0, out32=self.dw2)
            h.conv_toeplitz_fold(self.dw2, self.kh, self.W, self.c_in, self.c_out, self.H, self.cip, self.cop,
                                 P.view(self.wname, "grad"))
            if not self.bn:
                h.fill32(self.db_t, 0.0)
                h.colsum16(dy, rows, self.cop, self.db_t, accumulate=True)
                h.vec_fold(self.db_t, self.c_out, self.H, P.view(self.bname, "grad"))
        return dx_whole


class FCFrames(object):
    """fully_connected over the flattened NHWC frame (models/rced.py:106-113): one GEMM whose A rows are whole
    frames of the channels-last buffer (row pitch S*Cp, K = L*Cp; padded channels meet zero weight rows).  lines > 1:
    the frame is `lines` stacked lines stored as channels (Conv2dLines); the weight rows are re-indexed accordingly."""

    def __init__(self, net, scope, fl, chans, n_out, lines=1, x_act=True):
        """x_act: apply relu'(x) to the returned gradient (the producer wants the gradient wrt its pre-activation);
        False when the producer is normalised and takes the gradient wrt its output (ConvBN)."""
        self.net, self.scope, self.fl, self.chans, self.n_out, self.H = net, scope, fl, chans, n_out, lines
        self.x_act = x_act
        self.cp, self.outp = packing.round_up(lines * chans, 8), packing.round_up(n_out, 8)
        self.wname, self.bname = scope + "/weights", scope + "/biases"

    def segs(self):
        w = (params.fc_w_frames(self.wname, self.fl.L, self.chans, self.n_out) if self.H == 1 else
             params.fc_w_lines(self.wname, self.H, self.fl.L, self.chans, self.n_out))
        return [w, params.fc_b(self.bname, self.n_out)]

    def refresh(self):
        pass

    def _frames(self, whole, frames):
        fl = self.fl
        return whole.as_strided((frames, fl.L * self.cp), (fl.S * self.cp, 1),
                                whole.storage_offset() + fl.GUARD * self.cp)

    def fwd(self, ctx, x_whole, frames):
        net, h = self.net, self.net.h
        y32 = net.ws.get((ctx, self.scope, "y32"), frames, self.outp, F32)
        h.gemm(self._frames(x_whole, frames), net.P.view(self.wname, "theta16"), frames, self.outp,
               self.fl.L * self.cp, b_mn=True, bias=net.P.view(self.bname), out32=y32)
        return y32

    def bwd(self, ctx, x_whole, dy16, frames):
        """-> gradient wrt the last convolution's pre-activation, in the frame layout"""
        net, h, fl = self.net, self.net.h, self.fl
        K = fl.L * self.cp
        dx_whole, _ = fl.buf(net, (ctx, self.scope, "dx16", frames), frames, self.cp)
        xf = self._frames(x_whole, frames)
        h.gemm(dy16, net.P.view(self.wname, "theta16"), frames, K, self.outp, dact_src=xf if self.x_act else None,
               dact=ACT_RELU if self.x_act else ACT_NONE, out16=self._frames(dx_whole, frames))
        with h.side_stream():
            h.gemm(xf, dy16, K, self.outp, frames, a_mn=True, b_mn=True, beta=1.0,
                   out32=net.P.view(self.wname, "grad"))
            h.colsum16(dy16, frames, self.outp, net.P.view(self.bname, "grad"), accumulate=True)
        return dx_whole


class Net(object):
    """Common part: parameter store, workspace, weight-derived operands."""

    def __init__(self, handle, layers_fn, adam):
        self.h = handle
        self.ws = Workspace(handle)
        # batch_norm / dropout mode (FCBN layers): `training` = is_training of the graph being run (False for the
        # cross-validation / inference models), `bn_update` = whether THIS network's UPDATE_OPS run with the step
        # about to be taken (models/dnn_trainer_single_gpu.py:101-104; models/gan_rnn_placeholder.py:163-175: d_opt runs
        # the d_model ones, g_opt the g_model ones -- set per step by GAN_RNN._mode), `rng` = device
        # {seed, tick} of the dropout stream (rsr_affine_act_drop)
        self.training, self.bn_update = True, False
        self.keep_prob = getattr(self, "keep_prob", 1.0)
        self.rng = torch.tensor([1234, 0], dtype=torch.int64, device=handle.device)
        self.layers = layers_fn(self)
        segs = []
        for l in self.layers:
            segs += l.segs()
        self.P = params.ParamStore(handle, segs, adam)

    def load_tf(self, p):
        self.P.load_tf(p)
        self.P.ema.copy_(self.P.theta)
        self.refresh()

    @property
    def fcbn(self):
        return any(isinstance(l, FCBN) for l in self.layers)

    @property
    def has_bn_state(self):
        """any layer with non-trainable batch_norm variables (FCBN, ConvBN) or a dropout stream to checkpoint"""
        return self.fcbn or any(getattr(l, "bn", False) for l in self.layers)

    def bn_state_tf(self):
        """Non-trainable batch_norm variables keyed by their TF names (saved with the checkpoint, like tf.train.Saver
        saves every global variable, models/gan_rnn_placeholder.py:26-34)."""
        out = {}
        for l in self.layers:
            if hasattr(l, "export_state"):
                out.update(l.export_state())
        return out

    def load_bn_state_tf(self, d):
        for l in self.layers:
            if hasattr(l, "load_state"):
                l.load_state(d)

    def tick(self):
        """Advance the dropout stream (once per update, after the backward pass regenerated its masks)."""
        if self.keep_prob < 1.0:
            self.h.rng_tick(self.rng)

    def refresh(self):
        """Weight-derived operands after an update.  The per-layer refreshes are tiny and independent: alternate
        them between the two streams and join."""
        h = self.h
        for i, l in enumerate(self.layers):
            if i & 1 and hasattr(h, "side_stream"):
                with h.side_stream():
                    l.refresh()
            else:
                l.refresh()
        if hasattr(h, "join"):
            h.join()


RCED_FILTERS = (12, 16, 20, 24, 32, 24, 20, 16, 12)      # models/rced.py:92
RCED_WIDTHS = (13, 11, 9, 7, 7, 7, 9, 11, 13)            # models/rced.py:93


class Generator(Net):
    def __init__(self, handle, g_type="lstm", in_dim=257, out_dim=40, cell=760, proj=280, layers=None,
                 units=1024, batch_norm=False, keep_prob=1.0, splice=1):
        """splice (rced only): the input frame is `splice` stacked lines of in_dim / splice bins (models/rced.py:46-57)."""
        self.g_type, self.in_dim, self.out_dim = g_type, in_dim, out_dim
        self.splice = int(splice) if g_type == "rced" else 1
        # dnn: tf.nn.dropout behind every hidden layer (FCBN); LSTM generators: DropoutWrapper(output_keep_prob) on every
        # LSTM layer's output (models/lstm.py:99-102, models/res_lstm_l.py:96-99) = _drop_fwd / _drop_bwd below
        # rced builds keep_prob but never applies a dropout op (models/rced.py:73-77,102-103): a no-op there
        self.keep_prob = 1.0 if g_type == "rced" else float(keep_prob)
        self._dropping = False
        # res_lstm_l / res_lstm_base build normalizer_params but never pass them on (models/res_lstm_l.py:58-67,81-82):
        # batch_norm is a no-op there, exactly as in the reference
        special = g_type == "dnn" and (batch_norm or self.keep_prob < 1.0)
        if g_type == "lstm":
            # models/lstm.py:43-45: cell 760, projection 280, 3 layers
            L = 3 if layers is None else layers

            def mk(net):
                ls = [FCBN(net, "g_model/fully_connected", in_dim, proj, ACT_LRELU, True, 0, drop=False) if batch_norm else
                      FC(net, "g_model/fully_connected", in_dim, proj, ACT_LRELU)]
                ls += [LSTMP(net, "g_model/rnn/multi_rnn_cell/cell_%d/lstm_cell/" % i, proj, cell, proj)
                       for i in range(L)]
                for i in range(0, L - 1, 2):      # stacked pairs run as one wavefront launch (LSTMP.fwd_wave / bwd_wave)
                    ls[1 + i].wave = ls[1 + i].Cp <= 512 and float(keep_prob) >= 1.0
                    ls[1 + i].wave_next = ls[2 + i]
                return ls + [FC(net, "g_model/fully_connected_1", proj, out_dim, ACT_NONE)]
        elif g_type in ("res_lstm_l", "res_lstm_base"):
            # models/res_lstm_l.py:101-138: four LSTMP(760 -> in_dim) layers (the `lstm_num_layer = 3` at :45 is unused)
            L = 4 if layers is None else layers

            def mk(net):
                ls = [LSTMP(net, "g_model/lstm_cell_%d/rnn/lstm_cell/" % (i + 1), in_dim, cell, in_dim)
                      for i in range(L)]
                return ls + [FC(net, "g_model/forward_out/fully_connected", in_dim, out_dim, ACT_NONE)]
        elif g_type == "dnn":
            # models/dnn.py:34-35,79-110: in -> 1024 x (1+3) ReLU -> out
            L = 3 if layers is None else layers

            def mk(net):
                dims = [in_dim] + [units] * (L + 1)
                name = lambda i: "g_model/fully_connected" + ("" if i == 0 else "_%d" % i)
                if special:
                    ls = [FCBN(net, name(i), dims[i], dims[i + 1], ACT_RELU, batch_norm, i) for i in range(L + 1)]
                else:
                    ls = [FC(net, name(i), dims[i], dims[i + 1], ACT_RELU) for i in range(L + 1)]
                return ls + [FC(net, "g_model/fully_connected_%d" % (L + 1), units, out_dim, ACT_NONE)]
        elif g_type == "rced":
            # models/rced.py:92-101: nine [1, w] ReLU convolutions over the spectrum bins of each frame, then
            # FC (splice * bins * 12) -> out_dim.  splice = 1: 1-D convolutions (Conv1dSame); splice > 1: the [splice, w]
            # 2-D convolutions over the stacked lines (Conv2dLines).
            filt, wid = RCED_FILTERS, RCED_WIDTHS
            H = self.splice
            assert in_dim % H == 0, "rced input width must be splice * bins"
            self.frames = ConvFrames(in_dim // H, max(wid))

            def mk(net):
                ch = (1,) + filt
                name = lambda i: "g_model/Conv" + ("" if i == 0 else "_%d" % i)
                bn = bool(batch_norm)       # normalizer_fn=batch_norm on the nine conv2d, not on the output layer (:94-113)
                if H == 1:
                    ls = [Conv1dSame(net, name(i), self.frames, wid[i], ch[i], ch[i + 1], bn=bn) for i in range(len(filt))]
                else:
                    ls = [Conv2dLines(net, name(i), self.frames, H, wid[i], ch[i], ch[i + 1], H, bn=bn)
                          for i in range(len(filt))]
                return ls + [FCFrames(net, "g_model/fully_connected", self.frames, filt[-1], out_dim, lines=H,
                                      x_act=not bn)]
        else:
            raise ValueError("Unrecognized G type {}".format(g_type))   # models/gan_rnn_placeholder.py:131-132
        super(Generator, self).__init__(handle, mk, adam=True)
        self.residual = g_type == "res_lstm_l"

    # DropoutWrapper(output_keep_prob): the output handed to the next layer is dropped, the recurrent state is not
    def _drop_fwd(self, li, o32, rows, want32=False):
        h, ws, Pp = self.h, self.ws, o32.shape[1]
        xd16 = ws.get(("g", "drop16", li), rows, Pp, h.h16)
        xd32 = ws.get(("g", "drop32", li), rows, Pp, F32) if want32 else None
        zeros = ws.get(("g", "zeros", Pp), 1, Pp, F32)[0]
        h.affine_act_drop(o32, rows, Pp, None, zeros, ACT_NONE, self.keep_prob, self.rng,
                          CTX_SALT["g"] + LSTM_SALT + li, xd16, out32=xd32)
        return xd16, xd32

    def _drop_bwd(self, li, d16, o32, rows):
        h, ws, Pp = self.h, self.ws, o32.shape[1]
        dd16 = ws.get(("g", "ddrop16", li), rows, Pp, h.h16)
        dd32 = ws.get(("g", "ddrop32", li), rows, Pp, F32)
        zeros = ws.get(("g", "zeros", Pp), 1, Pp, F32)[0]
        h.bn_bwd(d16, o32, rows, Pp, ACT_NONE, self.keep_prob, self.rng, CTX_SALT["g"] + LSTM_SALT + li, False, None,
                 zeros, None, None, dd16, ws.get(("g", "drop_scratch"), 1, 8, F32), dz32=dd32)
        return dd16, dd32

    def fwd(self, x, B, T, lengths, train=True, x_time_major=False, reuse_staged=False, wave=True):
        """x fp32 (B, T, in_dim) batch-major on the device -> y32 [T*B, out_pad] time-major fp32.
        reuse_staged: the 16-bit time-major copy of this same x from the previous call is still valid (the second
        generator forward of a batch schedule sees the same minibatch).
        wave: stacked LSTMP pairs may run as one layer-wavefront launch (7 sixteen-CTA clusters instead of 4; measured
        faster even while D(labels) runs beside it on the side stream: 3.157 vs 3.175 ms per cfg-2 schedule)."""
        h, ws, rows = self.h, self.ws, T * B
        if self.g_type == "rced":
            self._B, self._T, self._len = B, T, lengths
            fl, Ls = self.frames, self.layers
            a, _ = fl.buf(self, ("g", "x16", rows), rows, Ls[0].cip)
            if not reuse_staged:
                if self.splice == 1:
                    h.conv_stage_frames(x, B, T, self.in_dim, fl.S, Ls[0].cip, a[fl.GUARD:], time_major_in=x_time_major)
                else:
                    h.conv_stage_lines(x, B, T, self.splice, fl.L, fl.S, Ls[0].cip, a[fl.GUARD:],
                                       time_major_in=x_time_major)
            self._acts = [a]
            for l in Ls[:-1]:
                a = l.fwd("g", a, rows)
                self._acts.append(a)
            return Ls[-1].fwd("g", a, rows)
        ip = packing.round_up(self.in_dim, 8)
        res = self.g_type in ("res_lstm_l", "res_lstm_base")
        x16 = ws.get(("g", "x16", B), rows, ip, h.h16)
        x32 = ws.get(("g", "x32", B), rows, ip, F32) if self.residual else None
        if not reuse_staged:
            h.stage_input(x, B, T, self.in_dim, out16=x16, out32=x32, time_major_in=x_time_major)
        self._B, self._T, self._len = B, T, lengths
        if self.g_type == "dnn":
            a = x16
            self._acts = [x16]
            for l in self.layers[:-1]:
                a, _ = l.fwd("g", a, rows)
                self._acts.append(a)
            _, y32 = self.layers[-1].fwd("g", a, rows, want16=False, want32=True)
            return y32
        drop = self._dropping = self.training and self.keep_prob < 1.0
        self._o32 = []
        if not res:
            h0, _ = self.layers[0].fwd("g", x16, rows)
            self._acts = [x16, h0]
            a = h0
            rec, li = self.layers[1:-1], 0
            while li < len(rec):
                l = rec[li]
                # two stacked layers as one wavefront launch (no DropoutWrapper between them)
                w = (l.fwd_wave(rec[li + 1], "g", a, B, T, lengths, save=train)
                     if (wave and not drop and li + 1 < len(rec)) else None)
                if w is not None:
                    self._o32 += [None, None]
                    self._acts += [w[0][B:], w[1][B:]]
                    a = w[1][B:]
                    li += 2
                    continue
                seq, o32 = l.fwd("g", a, B, T, lengths, save=train, want32=drop)
                a = self._drop_fwd(li, o32, rows)[0] if drop else seq[B:]
                self._o32.append(o32)
                self._acts.append(a)
                li += 1
            _, y32 = self.layers[-1].fwd("g", a, rows, want16=False, want32=True)
            return y32
        a16, a32 = x16, x32
        self._acts = [x16]
        for i, l in enumerate(self.layers[:-1]):
            seq, o32 = l.fwd("g", a16, B, T, lengths, save=train, want32=self.residual or drop)
            self._o32.append(o32)
            od16 = seq[B:]
            if drop:
                od16, o32 = self._drop_fwd(i, o32, rows, want32=self.residual)
            if self.residual:                      # x_{l+1} = out_l + x_l   (models/res_lstm_l.py:116,127,138,187)
                n16 = ws.get(("g", "xr16", i, B), rows, ip, h.h16)
                n32 = ws.get(("g", "xr32", i, B), rows, ip, F32)
                h.add_cast(o32, a32, rows * ip, out32=n32, out16=n16)
                a16, a32 = n16, n32
            else:
                a16 = od16
            self._acts.append(a16)
        _, y32 = self.layers[-1].fwd("g", a16, rows, want16=False, want32=True)
        return y32

    def bwd(self, dy16):
        """dy16 [T*B, out_pad]: (scaled) gradient wrt the generator output.  Accumulates into P.grad."""
        B, T, lengths, rows = self._B, self._T, self._len, self._T * self._B
        acts, Ls = self._acts, self.layers
        if self.g_type == "rced":
            d = Ls[-1].bwd("g", acts[-1], dy16, rows)
            for i in range(len(Ls) - 2, -1, -1):
                d = Ls[i].bwd("g", acts[i], d, rows, want_dx=i > 0)
            return
        if self.g_type == "dnn":
            d = dy16
            plain = not self.fcbn          # FCBN layers take the gradient wrt their OUTPUT and apply act' themselves
            for i in range(len(Ls) - 1, -1, -1):
                d, _ = Ls[i].bwd("g", acts[i], d, rows, want_dx=i > 0, prev_y16=acts[i] if i > 0 and plain else None,
                                 prev_act=ACT_RELU if plain else ACT_NONE)
            return
        if self.g_type == "lstm" and self._dropping:
            d16, d32 = Ls[-1].bwd("g", acts[-1], dy16, rows, want32=True)     # wrt the DROPPED output of the last LSTM
            for i in range(len(Ls) - 2, 0, -1):
                first = i == 1
                mask = first and not self.fcbn
                d16, d32 = self._drop_bwd(i - 1, d16, self._o32[i - 1], rows)
                d16, d32 = Ls[i].bwd("g", acts[i], d16, d32, B, T, lengths, prev_y16=acts[1] if mask else None,
                                     prev_act=ACT_LRELU if mask else ACT_NONE, want32=not first)
            Ls[0].bwd("g", acts[0], d16, rows, want_dx=False, dw_side=False)
            return
        if self.g_type == "lstm":
            d16, d32 = Ls[-1].bwd("g", acts[-1], dy16, rows, want32=True)
            Ls[-2].bwd_pre("g", d16, B, T)
            pending = []

            def flush(ev=None):
                if pending and ev is not None:
                    with self.h.side_stream(after=ev):
                        self.h.fill32(self.ws.get(("g", "spacer"), 1, 256, F32), 0.0)
                while pending:
                    l, a, d, ev0 = pending.pop(0)
                    l.bwd_side("g", a, d, B, T, after=False if ev is not None else ev0)
            i = len(Ls) - 2
            while i >= 1:
                dout32 = d32
                if i >= 2 and Ls[i - 1].wave:     # this layer and the one below it as one wavefront launch
                    mask = i - 1 == 1 and not self.fcbn
                    flush()
                    w = Ls[i - 1].bwd_wave(Ls[i], "g", acts[i - 1], acts[i], dout32, B, T, lengths,
                                           prev_y16=acts[1] if mask else None, prev_act=ACT_LRELU if mask else ACT_NONE,
                                           want32=i - 1 > 1)
                    if w is not None:
                        d16, d32 = w
                        i -= 2
                        if i >= 1:
                            Ls[i].bwd_pre("g", d16, B, T)
                        continue
                first = i == 1
                mask = first and not self.fcbn      # an FCBN first layer applies its own activation gradient
                # the weight gradients of the layer above are enqueued BEHIND this layer's recurrence kernel (the side
                # stream itself only waits for the point where their operands were complete): launched in front of it,
                # their persistent 148-CTA GEMM held the recurrence's cluster slots for ~20 us
                d16, d32 = Ls[i].bwd_main("g", d16, B, T, lengths, prev_y16=acts[1] if mask else None,
                                          prev_act=ACT_LRELU if mask else ACT_NONE, want32=not first, after_rec=flush)
                if not first:
                    Ls[i - 1].bwd_pre("g", d16, B, T)     # critical path first, then this layer's weight gradients
                # (first LSTMP layer: nothing is launched behind it that its weight gradients should yield to -- they start
                #  as soon as its recurrence is done, beside its data-gradient GEMM)
                pending.append((Ls[i], acts[i], dout32, Ls[i].rec_mark if first else self.h.mark()))
                i -= 1
            flush()
            Ls[0].bwd("g", acts[0], d16, rows, want_dx=False, dw_side=False)
            return
        d16, d32 = Ls[-1].bwd("g", acts[-1], dy16, rows, want32=True)
        for i in range(len(Ls) - 2, -1, -1):
            # in_{l+1} = drop(out_l) + in_l: the skip path carries d32 as it is, the layer sees the masked gradient
            o16, o32 = self._drop_bwd(i, d16, self._o32[i], rows) if self._dropping else (d16, d32)
            d16, d32 = Ls[i].bwd("g", acts[i], o16, o32, B, T, lengths, want_dx=i > 0,
                                 resid32=d32 if self.residual else None, want32=True)


class Discriminator(Net):
    def __init__(self, handle, d_type="lstm", in_dim=40, cell=256, proj=40, layers=None, units=1024,
                 batch_norm=False, keep_prob=1.0, cat_dim=0, adam=False):
        """cat_dim > 0 (discriminator_dnn only): the discriminator sees tf.concat([conditioning (cat_dim), x (in_dim)], -1)
        as in the frame-level GAN (models/gan.py:159-174, conditioning = centre-frame LPS).  adam: Adam instead of SGD
        for this network (models/gan.py:125 vs models/gan_rnn_placeholder.py:144)."""
        self.d_type, self.in_dim, self.cat_dim = d_type, in_dim, int(cat_dim)
        if cat_dim and d_type != "dnn":
            raise ValueError("a conditioned discriminator input is only defined for discriminator_dnn (models/gan.py)")
        # discriminator_lstm builds normalizer_params / keep_prob but uses neither (models/discriminator_lstm.py:37-52,
        # 64): both are no-ops there, as in the reference
        self.keep_prob = float(keep_prob) if d_type == "dnn" else 1.0
        special = d_type == "dnn" and (batch_norm or self.keep_prob < 1.0)
        if d_type == "lstm":
            # models/discriminator_lstm.py:26-28: cell 256, projection 40, 2 layers, FC -> 1 (no clip, :105)
            L = 2 if layers is None else layers

            def mk(net):
                ls = [LSTMP(net, "d_model/rnn/multi_rnn_cell/cell_%d/lstm_cell/" % i, in_dim if i == 0 else proj,
                            cell, proj) for i in range(L)]
                for i in range(0, L - 1, 2):      # stacked pairs run as one wavefront launch (LSTMP.fwd_wave / bwd_wave)
                    ls[i].wave = ls[i].Cp <= 512
                    ls[i].wave_next = ls[i + 1]
                return ls + [FC(net, "d_model/fully_connected", proj, 1, ACT_NONE)]
        elif d_type == "dnn":
            # models/discriminator_dnn.py:23-24,61-93: 1024 x (1+3) ReLU -> 1, clip_by_value(-0.5, 1.5)
            # (the clip and its gradient mask are applied by rsr_lsgan_mse_losses, clip=1)
            L = 3 if layers is None else layers

            def mk(net):
                dims = [in_dim + self.cat_dim] + [units] * (L + 1)
                name = lambda i: "d_model/fully_connected" + ("" if i == 0 else "_%d" % i)
                cat = lambda i: (self.cat_dim, in_dim) if (i == 0 and self.cat_dim) else None
                if special:
                    ls = [FCBN(net, name(i), dims[i], dims[i + 1], ACT_RELU, batch_norm, i, cat=cat(i))
                          for i in range(L + 1)]
                else:
                    ls = [FC(net, name(i), dims[i], dims[i + 1], ACT_RELU, cat=cat(i)) for i in range(L + 1)]
                return ls + [FC(net, "d_model/fully_connected_%d" % (L + 1), units, 1, ACT_NONE)]
        else:
            raise ValueError("Unrecognized D type {}".format(d_type))
        super(Discriminator, self).__init__(handle, mk, adam=adam)
        self.clip = d_type == "dnn"
        self._ctx = {}
        self.head_fused = {}

    def fwd(self, ctx, x32_tm, B, T, lengths, noise=None, train=True, cat_src=None, head=None):
        """head: a loss specification for FC.head -- when the last layer qualifies (one-unit head of the plain DNN
        discriminator with more than one layer below it), logits, LSGAN loss terms, d loss / d logit and the head's data
        gradient come out of ONE kernel and `self.head_fused[ctx]` is set; otherwise the caller runs rsr_lsgan_mse_losses.
        x32_tm fp32 [T*B, ld] time-major (labels or generator output); noise fp32 (B, in_dim) or None
        (utils/ops.py:19-30: ONE draw per utterance broadcast over time).  cat_src: the conditioning block of a
        conditioned discriminator, an fp32 batch-major (B, T, cat_dim) VIEW of the generator's input (row pitch = its
        stride).  Returns logits32 [T*B, 8] (column 0; pre-clip for the DNN discriminator)."""
        h, ws, rows = self.h, self.ws, T * B
        ip = packing.round_up(self.in_dim + self.cat_dim, 8)
        x16 = ws.get((ctx, "x16", B), rows, ip, h.h16)
        h.stage_input(x32_tm, B, T, self.in_dim, out16=x16, noise=noise, time_major_in=True)
        if self.cat_dim:      # device column order [x | conditioning], see params.fc_w_cat
            h.stage_input(cat_src, B, T, self.cat_dim, out16=x16[:, self.in_dim:], ldx=cat_src.stride(1))
        acts = [x16]
        a = x16
        if self.d_type == "lstm":
            rec, li = self.layers[:-1], 0
            while li < len(rec):
                w = rec[li].fwd_wave(rec[li + 1], ctx, a, B, T, lengths, save=train) if li + 1 < len(rec) else None
                if w is not None:
                    acts += [w[0][B:], w[1][B:]]
                    a = w[1][B:]
                    li += 2
                    continue
                seq, _ = rec[li].fwd(ctx, a, B, T, lengths, save=train)
                a = seq[B:]
                acts.append(a)
                li += 1
        else:
            for l in self.layers[:-1]:
                a, _ = l.fwd(ctx, a, rows)
                acts.append(a)
        last = self.layers[-1]
        fused = (head is not None and self.d_type == "dnn" and len(self.layers) > 1 and last.head_ok()
                 and os.environ.get("RSR_NO_HEAD_FUSION") != "1")
        self.head_fused[ctx] = fused
        if fused:      # (the head always produces its data gradient: every caller of bwd below a fused forward wants it)
            logits = last.head(ctx, a, rows, head, True, ACT_RELU if not self.fcbn else ACT_NONE)
        else:
            _, logits = last.fwd(ctx, a, rows, want16=False, want32=True)
        self._ctx[ctx] = (acts, B, T, lengths)
        return logits

    def bwd(self, ctx, dlogit16, want_dw=True, want_dx=False, resid32=None, pre_last=None):
        """dlogit16 [T*B, 8] (column 0).  want_dx: returns d/d(input) (+ resid32) as 16-bit [T*B, in_pad].
        pre_last: called in front of the first layer's backward (the consumer of resid32): where a caller that computed
        resid32 on the side stream joins it."""
        acts, B, T, lengths = self._ctx[ctx]
        rows, Ls = T * B, self.layers
        if self.d_type == "lstm":
            if pre_last is not None:
                pre_last()
            d16, d32 = Ls[-1].bwd(ctx, acts[-1], dlogit16, rows, want_dw=want_dw, want32=want_dw)
            i = len(Ls) - 2
            while i >= 0:
                last, dout32 = i == 0, d32
                if i >= 1 and Ls[i - 1].wave:     # this layer and the one below it as one wavefront launch
                    low = i - 1 == 0
                    Ls[i].bwd_pre(ctx, d16, B, T)
                    w = Ls[i - 1].bwd_wave(Ls[i], ctx, acts[i - 1], acts[i], dout32, B, T, lengths,
                                           want32=want_dw and not low, want_dw=want_dw, want_dx=(not low) or want_dx,
                                           resid32=resid32 if low else None)
                    if w is not None:
                        d16, d32 = w
                        i -= 2
                        continue
                    # declined (bwd_pre has run): finish this layer the usual way
                    d16, d32 = Ls[i].bwd_main(ctx, d16, B, T, lengths, want_dw=want_dw, want_dx=True, want32=want_dw)
                    if want_dw:
                        Ls[i].bwd_side(ctx, acts[i], dout32, B, T)
                    i -= 1
                    continue
                d16, d32 = Ls[i].bwd(ctx, acts[i], d16, d32, B, T, lengths, want_dw=want_dw,
                                     want_dx=(not last) or want_dx, resid32=resid32 if last else None,
                                     want32=want_dw and not last)
                i -= 1
            return d16
        d = dlogit16
        plain = not self.fcbn
        for i in range(len(Ls) - 1, -1, -1):
            last = i == 0
            kw = dict(dx_done=True) if (i == len(Ls) - 1 and self.head_fused.get(ctx)) else {}
            if last and pre_last is not None:
                pre_last()
            d, _ = Ls[i].bwd(ctx, acts[i], d, rows, want_dw=want_dw, want_dx=(not last) or want_dx,
                             prev_y16=None if last or not plain else acts[i], prev_act=ACT_RELU if plain else ACT_NONE,
                             resid32=resid32 if last else None, **kw)
        return d
