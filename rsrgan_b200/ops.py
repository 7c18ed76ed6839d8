"""Thin Python wrappers over the C ABI (include/rsrgan_b200.h).

torch is only the allocator / stream owner here: every function takes CUDA
tensors, passes raw device pointers + the current torch stream to the library
and returns immediately (asynchronous).  No function has a CPU path.
"""
from __future__ import annotations

import contextlib
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_CLIP, ACT_LRELU, ACT_NONE, ACT_RELU, GemmArgs, WaveArgs, WaveBwdArgs, check  # noqa: F401

H16 = {0: torch.float16, 1: torch.bfloat16}


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Handle(object):
    """Per-rank library handle (rsr_create / rsr_destroy)."""

    def __init__(self, device=0, dtype="f16"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.RsrError("rsrgan_b200 needs an sm_100 GPU; there is no CPU fallback")
        self.dtype_id = {"f16": _lib.RSR_DTYPE_F16, "fp16": _lib.RSR_DTYPE_F16,
                         "bf16": _lib.RSR_DTYPE_BF16}[dtype]
        self.h16 = H16[self.dtype_id]
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        h = C.c_void_p()
        check(self.lib.rsr_create(C.byref(h), device, self.dtype_id), "rsr_create")
        self.h = h
        self.num_sms = self.lib.rsr_num_sms(self.h)
        self.launches = 0          # kernels of librsrgan_sm100.so launched so far (bench accounting)
        self.timing = None         # list of (name, start event, end event) while profiling, else None
        # list of (name, stream tag, start event, end event) with BOTH streams live (scripts/gpu_timeline.py): unlike
        # `timing`, the side stream keeps running, so the events give the true concurrent schedule of one eager pass
        self.timeline = None
        # second stream for work that is independent of the recurrences (which occupy only the SMs of
        # their clusters): D(real) forward/backward, weight-gradient GEMMs.  `overlap = False` serialises.
        self.overlap = True
        self._side = torch.cuda.Stream(device=self.device)
        self._side_dirty = False

    def mark(self):
        """An event at the current tail of the current stream (None when the side stream is off): `side_stream(after=...)`
        orders side work behind this point instead of behind whatever has been enqueued by then."""
        if not self.overlap or self.timing is not None:
            return None
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        return ev

    @contextlib.contextmanager
    def side_stream(self, after=None):
        """Calls made inside run on the side stream, ordered after everything enqueued so far on the
        current stream (or after the `mark()` passed as `after`; `after=False`: no new dependency); `join()` makes the
        current stream wait for them."""
        if not self.overlap or self.timing is not None:
            yield
            return
        if after is False:
            pass                                   # already ordered: continues the side stream's own queue
        elif after is not None:
            self._side.wait_event(after)
        else:
            self._side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._side):
            yield
        self._side_dirty = True

    def join(self):
        if self._side_dirty:
            torch.cuda.current_stream().wait_stream(self._side)
            self._side_dirty = False

    def _call(self, name, n_kernels, *args, work=0.0):
        """One C-ABI call on torch's current stream.  With `self.timing` set (bench.py's kernel-share
        pass) the call is bracketed by CUDA events on that stream; elapsed times are read by
        `timing_summary()` after a synchronize."""
        fn = getattr(self.lib, name)
        if self.timeline is not None:
            cur = torch.cuda.current_stream()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(cur)
            rc = fn(*args)
            e1.record(cur)
            self.timeline.append((name, "side" if cur == self._side else "main", e0, e1, work))
            check(rc, name)
            self.launches += n_kernels
            return
        if self.timing is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            self.timing.append((name, e0, e1, work))
        else:
            rc = fn(*args)
        check(rc, name)
        self.launches += n_kernels

    def timing_summary(self):
        """{call name: (count, total ms, total algorithmic flops)} since `self.timing = []`."""
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1, work in self.timing or []:
            c, t, w = out.get(name, (0, 0.0, 0.0))
            out[name] = (c + 1, t + e0.elapsed_time(e1), w + work)
        return out

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsr_destroy(self.h)
            self.h = None

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, bias=None,
             resid=None, act=ACT_NONE, dact_src=None, dact=ACT_NONE, out32=None, out16=None,
             tile_n=0, lda=None, ldb=None, split_k=0, stats=None):
        """D[M,N] = epi(alpha * A B).  A/B are 2-D h16 tensors (or views with a row stride);
        see rsr_gemm in the header for operand major-ness."""
        a = GemmArgs()
        a.M, a.N, a.K = M, N, K
        a.A, a.lda, a.a_mn = _p(A), (lda if lda is not None else A.stride(0)), int(a_mn)
        a.B, a.ldb, a.b_mn = _p(B), (ldb if ldb is not None else B.stride(0)), int(b_mn)
        a.alpha, a.beta = alpha, beta
        a.bias = _p(bias)
        a.resid, a.ldr = _p(resid), (resid.stride(0) if resid is not None else 0)
        a.act = act
        a.dact_src, a.ldd, a.dact = _p(dact_src), (dact_src.stride(0) if dact_src is not None else 0), dact
        a.out32, a.ldc32 = _p(out32), (out32.stride(0) if out32 is not None else 0)
        a.out16, a.ldc16 = _p(out16), (out16.stride(0) if out16 is not None else 0)
        a.tile_n, a.split_k = tile_n, split_k
        a.stats = _p(stats)       # batch_norm statistics out of the epilogue: (count, mean, M2) per 128-row block and column
        self._call("rsr_gemm", 1, self.h, _stream(), C.byref(a), work=2.0 * M * N * K)

    # --------------------------------------------------------------- staging
    def stage_input(self, x, B, T, D, out16=None, out32=None, mean=None, istd=None, noise=None,
                    time_major_in=False, ldx=None):
        self._call("rsr_stage_input", 1, 
            self.h, _stream(), _p(x), ldx if ldx is not None else (x.stride(0) if time_major_in else D),
            int(time_major_in), B, T, D, _p(mean), _p(istd), _p(noise),
            _p(out16), out16.stride(0) if out16 is not None else 0,
            _p(out32), out32.stride(0) if out32 is not None else 0)

    def unstage_output(self, y_tm, B, T, D, out_bm, mean=None, std=None):
        self._call("rsr_unstage_output", 1, self.h, _stream(), _p(y_tm), y_tm.stride(0), B, T, D,
                                          _p(mean), _p(std), _p(out_bm))

    def cmvn_apply(self, x, mean, std, out):
        n, d = x.shape
        self._call("rsr_cmvn_apply", 1, self.h, _stream(), _p(x), _p(mean), _p(std), n, d, _p(out))

    def cmvn_apply_padded(self, x, lengths, mean64, std64, out):
        """x, out (B, T, D) fp32; lengths (B,) int32; mean64 / std64 (D,) float64 (see include/rsrgan_b200.h)"""
        B, T, D = x.shape
        self._call("rsr_cmvn_apply_padded", 1, self.h, _stream(), _p(x), _p(lengths), _p(mean64), _p(std64), B, T, D, _p(out))

    def cmvn_invert(self, y, mean, std, out):
        n, d = y.shape
        self._call("rsr_cmvn_invert", 1, self.h, _stream(), _p(y), _p(mean), _p(std), n, d, _p(out))

    # ----------------------------------------------------------------- LSTMP
    def lstmp_rec_fwd(self, B, T, Cp, zx, wcT, w_i, w_f, w_o, lengths, mt_seq, save, forget_bias=1.0, work=0.0):
        self._call("rsr_lstmp_rec_fwd", 1, self.h, _stream(), B, T, Cp, _p(zx), _p(wcT), _p(w_i), _p(w_f),
                                         _p(w_o), forget_bias, _p(lengths), _p(mt_seq), _p(save), work=work)

    def lstmp_fused_fwd(self, B, T, I, Cp, x16, kxT, bias, wcT, w_i, w_f, w_o, lengths, mt_seq, save,
                        forget_bias=1.0, work=0.0):
        """Returns False (nothing launched) when the fused variant does not apply to this shape."""
        fn = self.lib.rsr_lstmp_fused_fwd
        args = (self.h, _stream(), B, T, I, Cp, _p(x16), x16.stride(0), _p(kxT), _p(bias), _p(wcT), _p(w_i), _p(w_f),
                _p(w_o), forget_bias, _p(lengths), _p(mt_seq), _p(save))
        if self.timing is not None or self.timeline is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = fn(*args)
            e1.record()
            if rc == 0 and self.timing is not None:
                self.timing.append(("rsr_lstmp_fused_fwd", e0, e1, work))
            if rc == 0 and self.timeline is not None:
                cur = torch.cuda.current_stream()
                self.timeline.append(("rsr_lstmp_fused_fwd", "side" if cur == self._side else "main", e0, e1, work))
        else:
            rc = fn(*args)
        if rc == _lib.RSR_E_RESIDENT:
            return False
        check(rc, "rsr_lstmp_fused_fwd")
        self.launches += 1
        return True

    def lstmp_wave_fwd(self, B, T, Cp, I1, P1, lengths, x16, l1, mt1, save1, wpT1, out1, l2, mt2, save2,
                       forget_bias=1.0, work=0.0):
        """Two stacked LSTMP layers as one wavefront launch (rsr_lstmp_wave_fwd).  l1 / l2 = (kxT, bias, wcT, w_i, w_f,
        w_o) of the layers.  Returns False (nothing launched) when the shape does not apply."""
        a = WaveArgs()
        a.B, a.T, a.Cp, a.I1, a.P1, a.forget_bias = B, T, Cp, I1, P1, forget_bias
        a.lengths = _p(lengths)
        a.x16, a.ldx = _p(x16), x16.stride(0)
        a.kxT1, a.bias1, a.wcT1, a.w_i1, a.w_f1, a.w_o1 = (_p(t) for t in l1)
        a.mt1, a.save1, a.wpT1 = _p(mt1), _p(save1), _p(wpT1)
        a.out1, a.ldo1 = _p(out1), out1.stride(0)
        a.kxT2, a.bias2, a.wcT2, a.w_i2, a.w_f2, a.w_o2 = (_p(t) for t in l2)
        a.mt2, a.save2 = _p(mt2), _p(save2)
        fn = self.lib.rsr_lstmp_wave_fwd
        timed = self.timing is not None or self.timeline is not None
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = fn(self.h, _stream(), C.byref(a))
        if timed:
            e1.record()
            if rc == 0 and self.timing is not None:
                self.timing.append(("rsr_lstmp_wave_fwd", e0, e1, work))
            if rc == 0 and self.timeline is not None:
                cur = torch.cuda.current_stream()
                self.timeline.append(("rsr_lstmp_wave_fwd", "side" if cur == self._side else "main", e0, e1, work))
        if rc == _lib.RSR_E_RESIDENT:
            return False
        check(rc, "rsr_lstmp_wave_fwd")
        self.launches += 1
        return True

    def lstmp_wave_bwd(self, B, T, Cp, lengths, dmt2, l2, save2, dz2, g2, fT, part, l1, save1, dz1, g1, work=0.0,
                       max_nbp=0):
        """Backward of two stacked LSTMP layers as one wavefront launch (rsr_lstmp_wave_bwd).  l = (wc, w_i, w_f, w_o),
        g = (dbias, dw_i, dw_f, dw_o) per layer; part fp32 [T*(B+48), Cp], all zeros (and all zeros again afterwards).  Returns
        False when the shape does not apply."""
        a = WaveBwdArgs()
        a.B, a.T, a.Cp, a.max_nbp, a.lengths = B, T, Cp, int(max_nbp), _p(lengths)
        a.dmt2, a.save2, a.dz2 = _p(dmt2), _p(save2), _p(dz2)
        a.wc2, a.w_i2, a.w_f2, a.w_o2 = (_p(t) for t in l2)
        a.dbias2, a.dw_i2, a.dw_f2, a.dw_o2 = (_p(t) for t in g2)
        a.fT, a.part = _p(fT), _p(part)
        a.save1, a.dz1 = _p(save1), _p(dz1)
        a.wc1, a.w_i1, a.w_f1, a.w_o1 = (_p(t) for t in l1)
        a.dbias1, a.dw_i1, a.dw_f1, a.dw_o1 = (_p(t) for t in g1)
        fn = self.lib.rsr_lstmp_wave_bwd
        timed = self.timing is not None or self.timeline is not None
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        rc = fn(self.h, _stream(), C.byref(a))
        if timed:
            e1.record()
            if rc == 0 and self.timing is not None:
                self.timing.append(("rsr_lstmp_wave_bwd", e0, e1, work))
            if rc == 0 and self.timeline is not None:
                cur = torch.cuda.current_stream()
                self.timeline.append(("rsr_lstmp_wave_bwd", "side" if cur == self._side else "main", e0, e1, work))
        if rc == _lib.RSR_E_RESIDENT:
            return False
        check(rc, "rsr_lstmp_wave_bwd")
        self.launches += 1
        return True

    def transpose16(self, src, rows, cols, dst):
        self._call("rsr_transpose16", 1, self.h, _stream(), _p(src), src.stride(0), rows, cols, _p(dst), dst.stride(0))

    def lstmp_rec_bwd(self, B, T, Cp, dmt, wc, w_i, w_f, w_o, lengths, save, dz16, dbias, dw_i, dw_f, dw_o,
                      work=0.0):
        self._call("rsr_lstmp_rec_bwd", 1, self.h, _stream(), B, T, Cp, _p(dmt), _p(wc), _p(w_i), _p(w_f),
                                         _p(w_o), _p(lengths), _p(save), _p(dz16), _p(dbias), _p(dw_i),
                                         _p(dw_f), _p(dw_o), work=work)

    # ---------------------------------------------------------------- losses
    def lsgan_mse_losses(self, losses, rl=None, fk=None, ld_logit=1, n_logit=0, clip=False, g=None, y=None,
                         n_frames=0, d_out=0, d_real=1.0, d_fake=0.0, lam=0.0, gscale=1.0,
                         d_rl_grad=None, d_fk_grad=None, g_adv_grad=None, ld_grad=1, dg_mse=None):
        self._call("rsr_lsgan_mse_losses", 1, 
            self.h, _stream(), _p(rl), _p(fk), ld_logit, n_logit, int(clip),
            _p(g), g.stride(0) if g is not None else 0, _p(y), y.stride(0) if y is not None else 0,
            n_frames, d_out, d_real, d_fake, lam, gscale, _p(losses), _p(d_rl_grad), _p(d_fk_grad),
            _p(g_adv_grad), ld_grad, _p(dg_mse), dg_mse.stride(0) if dg_mse is not None else 0)

    def colsum16(self, x16, M, N, out, accumulate=False, ld=None):
        self._call("rsr_colsum16", 1, self.h, _stream(), _p(x16), ld if ld is not None else x16.stride(0),
                                    M, N, _p(out), int(accumulate))

    def colsum32(self, x32, M, N, out, accumulate=False, ld=None):
        self._call("rsr_colsum32", 1, self.h, _stream(), _p(x32), ld if ld is not None else x32.stride(0),
                                    M, N, _p(out), int(accumulate))

    # ---------------------------------------------------------------- update
    def seg_sumsq(self, grad, gmul, seg_id, n_seg, sumsq):
        self._call("rsr_seg_sumsq", 1, self.h, _stream(), _p(grad), gmul, _p(seg_id), grad.numel(), n_seg,
                                     _p(sumsq))

    def clip_sgd_ema(self, grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, ema, theta16, n_seg=0):
        """n_seg > 0: the update is skipped (and hyper[7] counts it) when any of sumsq[0..n_seg) is not finite"""
        self._call("rsr_clip_sgd_ema", 1, self.h, _stream(), _p(grad), gmul, _p(seg_id), _p(sumsq), int(n_seg), max_norm,
                                        _p(hyper), ema_decay, theta.numel(), _p(theta), _p(ema), _p(theta16))

    def clip_adam_ema(self, grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, m, v, ema, theta16, n_seg=0):
        self._call("rsr_clip_adam_ema", 3, self.h, _stream(), _p(grad), gmul, _p(seg_id), _p(sumsq), int(n_seg), max_norm,
                                         _p(hyper), ema_decay, theta.numel(), _p(theta), _p(m), _p(v), _p(ema),
                                         _p(theta16))

    def l2_grad(self, grad, theta, seg_id, seg_flag, scale):
        self._call("rsr_l2_grad", 1, self.h, _stream(), _p(grad), _p(theta), _p(seg_id), _p(seg_flag), scale,
                                   grad.numel())

    def add_cast(self, a, b, n, out32=None, out16=None):
        self._call("rsr_add_cast", 1, self.h, _stream(), _p(a), _p(b), n, _p(out32), _p(out16))

    def cast16(self, x, out16):
        self._call("rsr_cast16", 1, self.h, _stream(), _p(x), x.numel(), _p(out16))

    def fill32(self, x, v):
        self._call("rsr_fill32", 1, self.h, _stream(), _p(x), x.numel(), v)

    # ------------------------------------------------------------- 1-D conv glue
    def conv_stage_frames(self, x, B, T, L, S, Cp, out16, mean=None, istd=None, time_major_in=False, ldx=None):
        self._call("rsr_conv_stage_frames", 1, self.h, _stream(), _p(x),
                   ldx if ldx is not None else (x.stride(0) if time_major_in else L), int(time_major_in),
                   B, T, L, S, Cp, _p(mean), _p(istd), _p(out16))

    def conv_mask_rows(self, buf16, frames, S, L, Cp):
        self._call("rsr_conv_mask_rows", 1, self.h, _stream(), _p(buf16), frames, S, L, Cp)

    def conv_w_flip(self, w16, W, cin_p, cout_p, out16):
        self._call("rsr_conv_w_flip", 1, self.h, _stream(), _p(w16), W, cin_p, cout_p, _p(out16))

    def conv_toeplitz_expand(self, w16, kh, W, ci, co, H, cip, cop, out16):
        self._call("rsr_conv_toeplitz_expand", 1, self.h, _stream(), _p(w16), kh, W, ci, co, H, cip, cop, _p(out16))

    def conv_toeplitz_fold(self, dw2, kh, W, ci, co, H, cip, cop, grad):
        self._call("rsr_conv_toeplitz_fold", 1, self.h, _stream(), _p(dw2), kh, W, ci, co, H, cip, cop, _p(grad))

    def vec_tile(self, v, co, H, cop, out):
        self._call("rsr_vec_tile", 1, self.h, _stream(), _p(v), co, H, cop, _p(out))

    def vec_fold(self, t, co, H, grad):
        self._call("rsr_vec_fold", 1, self.h, _stream(), _p(t), co, H, _p(grad))

    def conv_stage_lines(self, x, B, T, H, L, S, Cp, out16, mean=None, istd=None, time_major_in=False, ldx=None):
        self._call("rsr_conv_stage_lines", 1, self.h, _stream(), _p(x),
                   ldx if ldx is not None else (x.stride(0) if time_major_in else H * L), int(time_major_in), B, T, H, L, S,
                   Cp, _p(mean), _p(istd), _p(out16))

    def conv_w_phase(self, w16, W, ap, bp, step, phase, out16):
        self._call("rsr_conv_w_phase", 1, self.h, _stream(), _p(w16), W, ap, bp, step, phase, _p(out16))

    # ------------------------------------------------- virtual batch norm (utils/bnorm.py)
    def vbn_stats(self, z32, rows, N, gamma, beta, coef, scratch, eps=1e-5, batch_weight=1.0, ref_stats=None,
                  stats_out=None):
        self._call("rsr_vbn_stats", 2, self.h, _stream(), _p(z32), z32.stride(0), rows, N, _p(gamma), _p(beta), eps,
                   float(batch_weight), _p(ref_stats), _p(stats_out), _p(coef), _p(scratch))

    def vbn_bwd(self, da16, z32, rows, N, act, batch_weight, coef, dgamma, dbeta, dz16, scratch, dz32=None):
        self._call("rsr_vbn_bwd", 3, self.h, _stream(), _p(da16), da16.stride(0), _p(z32), z32.stride(0), rows, N, act,
                   float(batch_weight), _p(coef), _p(dgamma), _p(dbeta), _p(dz16),
                   dz16.stride(0) if dz16 is not None else 0, _p(dz32), dz32.stride(0) if dz32 is not None else 0,
                   _p(scratch))

    # ------------------------------------------------- one-output fully_connected (discriminator heads)
    def fc1_fwd(self, x16, rows, K, w16, bias, out32):
        self._call("rsr_fc1_fwd", 1, self.h, _stream(), _p(x16), x16.stride(0), rows, K, _p(w16), w16.stride(0),
                   _p(bias), _p(out32), out32.stride(0), work=2.0 * rows * K)

    def fc1_bwd_dx(self, dy16, rows, K, w16, dx16, dact_src=None, dact=ACT_NONE):
        self._call("rsr_fc1_bwd_dx", 1, self.h, _stream(), _p(dy16), dy16.stride(0), rows, K, _p(w16), w16.stride(0),
                   _p(dact_src), dact_src.stride(0) if dact_src is not None else 0, dact, _p(dx16), dx16.stride(0),
                   work=2.0 * rows * K)

    def fc1_head(self, x16, rows, K, w16, bias, which, clip, d_real, d_fake, grad_target, gscale, losses, logit32,
                 dlogit16=None, dact=ACT_NONE, dx16=None):
        """The discriminator head in one pass (rsr_fc1_head): logits, LSGAN loss terms, d loss / d logit, head data gradient."""
        self._call("rsr_fc1_head", 1, self.h, _stream(), _p(x16), x16.stride(0), rows, K, _p(w16), w16.stride(0), _p(bias),
                   int(which), int(clip), float(d_real), float(d_fake), float(grad_target), float(gscale), _p(losses),
                   _p(logit32), logit32.stride(0) if logit32 is not None else 0,
                   _p(dlogit16), dlogit16.stride(0) if dlogit16 is not None else 0, dact,
                   _p(dx16), dx16.stride(0) if dx16 is not None else 0, work=4.0 * rows * K)

    # ------------------------------------------------------ gradient all-reduce over peer memory (rsrgan_b200/peer.py)
    PEER_HEADER_BYTES, PEER_IPC_HANDLE_BYTES = 16384, 64

    def peer_alloc(self, data_bytes):
        """-> (block address, 64-byte IPC handle)"""
        blk, hd = C.c_void_p(), C.create_string_buffer(self.PEER_IPC_HANDLE_BYTES)
        check(self.lib.rsr_peer_alloc(self.h, int(data_bytes), C.byref(blk), C.cast(hd, C.c_void_p)), "rsr_peer_alloc")
        return int(blk.value), hd.raw

    def peer_open(self, ipc_handle):
        blk = C.c_void_p()
        hd = C.create_string_buffer(bytes(ipc_handle), self.PEER_IPC_HANDLE_BYTES)
        check(self.lib.rsr_peer_open(self.h, C.cast(hd, C.c_void_p), C.byref(blk)), "rsr_peer_open")
        return int(blk.value)

    def peer_close(self, block):
        check(self.lib.rsr_peer_close(self.h, C.c_void_p(block)), "rsr_peer_close")

    def peer_free(self, block):
        check(self.lib.rsr_peer_free(self.h, C.c_void_p(block)), "rsr_peer_free")

    def peer_error(self, block):
        e = C.c_int(0)
        check(self.lib.rsr_peer_error(self.h, C.c_void_p(block), C.byref(e)), "rsr_peer_error")
        return int(e.value)

    def peer_allreduce(self, blocks, rank, data_off_bytes, n_floats, max_blocks=0):
        arr = (C.c_void_p * len(blocks))(*blocks)
        self._call("rsr_peer_allreduce", 1 if len(blocks) > 1 else 0, self.h, _stream(), arr, rank, len(blocks),
                   int(data_off_bytes), int(n_floats), int(max_blocks))

    # ------------------------------------------------------ batch_norm(renorm) / dropout
    BN_EPS, BN_DECAY, BN_RENORM_DECAY = 1e-3, 0.999, 0.99     # contrib batch_norm defaults (TF 1.4)

    def bn_train_stats(self, z32, rows, N, gamma, beta, state, coef, scratch, update_state=False):
        self._call("rsr_bn_train_stats", 2, self.h, _stream(), _p(z32), z32.stride(0), rows, N, _p(gamma), _p(beta),
                   self.BN_EPS, _p(state), self.BN_DECAY, self.BN_RENORM_DECAY, int(update_state), _p(coef),
                   _p(scratch))

    BN_STATS_ROWS_MAX = 256 * 128       # row blocks the partial buffer of rsr_bn_train_stats holds

    def bn_train_finish(self, splits, rows, N, gamma, beta, state, coef, scratch, update_state=False):
        """rsr_bn_train_stats without its first pass: `scratch` holds `splits` row-block partials written by the epilogue of
        the GEMM that produced the pre-activation (gemm(..., stats=scratch))."""
        self._call("rsr_bn_train_finish", 1, self.h, _stream(), int(splits), rows, N, _p(gamma), _p(beta), self.BN_EPS,
                   _p(state), self.BN_DECAY, self.BN_RENORM_DECAY, int(update_state), _p(coef), _p(scratch))

    def bn_eval_coef(self, N, gamma, beta, state, coef):
        self._call("rsr_bn_eval_coef", 1, self.h, _stream(), N, _p(gamma), _p(beta), self.BN_EPS, _p(state), _p(coef))

    def affine_act_drop(self, z32, rows, N, A, Bc, act, keep_prob, rng, salt, out16, out32=None):
        self._call("rsr_affine_act_drop", 1, self.h, _stream(), _p(z32), z32.stride(0), rows, N, _p(A), _p(Bc), act,
                   float(keep_prob), _p(rng), int(salt), _p(out16), out16.stride(0) if out16 is not None else 0,
                   _p(out32), out32.stride(0) if out32 is not None else 0)

    def bn_bwd(self, da16, z32, rows, N, act, keep_prob, rng, salt, bn, coef, bias, dgamma, dbeta, dz16, scratch,
               dz32=None):
        n = (2 if (bn or dbeta is not None) else 0) + (1 if (dz16 is not None or dz32 is not None) else 0)
        self._call("rsr_bn_bwd", n, self.h, _stream(), _p(da16), da16.stride(0), _p(z32), z32.stride(0), rows, N, act,
                   float(keep_prob), _p(rng), int(salt), int(bn), _p(coef), _p(bias), _p(dgamma), _p(dbeta),
                   _p(dz16), dz16.stride(0) if dz16 is not None else 0, _p(dz32),
                   dz32.stride(0) if dz32 is not None else 0, _p(scratch))

    # batch_norm behind the convolutions of the frame layout (include/rsrgan_b200.h); state [6, lds] per channel
    def bn_train_stats_lines(self, z32, frames, S, L, H, C, N, gamma, beta, state, coef, scratch, update_state=False):
        self._call("rsr_bn_train_stats_lines", 2, self.h, _stream(), _p(z32), z32.stride(0), frames, S, L, H, C, N,
                   _p(gamma), _p(beta), self.BN_EPS, _p(state), state.stride(0), self.BN_DECAY, self.BN_RENORM_DECAY,
                   int(update_state), _p(coef), _p(scratch))

    def bn_eval_coef_lines(self, N, H, C, gamma, beta, state, coef):
        self._call("rsr_bn_eval_coef_lines", 1, self.h, _stream(), N, H, C, _p(gamma), _p(beta), self.BN_EPS, _p(state),
                   state.stride(0), _p(coef))

    def affine_act_lines(self, z32, frames, S, L, N, A, Bc, act, out16):
        self._call("rsr_affine_act_lines", 1, self.h, _stream(), _p(z32), z32.stride(0), frames, S, L, N, _p(A), _p(Bc),
                   act, _p(out16), out16.stride(0))

    def bn_bwd_lines(self, da16, z32, frames, S, L, H, C, N, act, coef, dgamma, dbeta, dz16, scratch):
        self._call("rsr_bn_bwd_lines", 3, self.h, _stream(), _p(da16), da16.stride(0), _p(z32), z32.stride(0), frames, S,
                   L, H, C, N, act, _p(coef), _p(dgamma), _p(dbeta), _p(dz16), dz16.stride(0), _p(scratch))

    def rng_tick(self, rng):
        self._call("rsr_rng_tick", 1, self.h, _stream(), _p(rng))

    def gauss_noise(self, rng, salt, out, stddev):
        """out (fp32, contiguous) = stddev * N(0, 1) from the counter-based stream {seed, tick} (rsr_gauss_noise)."""
        self._call("rsr_gauss_noise", 1, self.h, _stream(), _p(rng), salt, _p(out), out.numel(), float(stddev))

    # ------------------------------------------------------ Kaldi compressed-matrix decode
    def ark_decompress(self, col_hdr, data, min_value, rng, rows, cols, out64=None, out32=None, mean=None, std=None):
        """col_hdr: device int16/uint16-as-int16 [cols, 4]; data: device uint8 [cols, rows] (both as read from the ark);
        mean/std: device float64 [cols] (CMVN in float64, make_tfrecords.py:84-87) or None."""
        self._call("rsr_ark_decompress", 1, self.h, _stream(), _p(col_hdr), _p(data), float(min_value), float(rng),
                   int(rows), int(cols), _p(out64), out64.stride(0) if out64 is not None else 0, _p(mean), _p(std),
                   _p(out32), out32.stride(0) if out32 is not None else 0)
