"""TEST DOUBLE of rsrgan_b200.ops.Handle: emulates every C-ABI call with torch CPU ops so that the
HOST logic (network wiring, parameter packing, update schedule, data-parallel all-reduce over gloo,
trainer / CLI) can be exercised by the `-m "not gpu"` suite in a container without a GPU.

It lives under tests/ and is injected through GAN_RNN(handle=...).  Nothing under rsrgan_b200/
imports it: the product path only ever talks to librsrgan_sm100.so and fails loudly without it.
The emulation follows the contracts written in include/rsrgan_b200.h (operand major-ness, packed
gate columns, time-major rows, loss scaling), not the CUDA code.
"""
from __future__ import annotations

import contextlib

import torch

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_CLIP = 0, 1, 2, 3


def _act(v, act):
    if act == ACT_RELU:
        return torch.clamp_min(v, 0.0)
    if act == ACT_LRELU:
        return torch.maximum(v, 0.3 * v)
    if act == ACT_CLIP:
        return torch.clamp(v, -0.5, 1.5)
    return v


def _dact(y, act):
    if act == ACT_RELU:
        return (y > 0).to(torch.float32)
    if act == ACT_LRELU:
        return torch.where(y > 0, torch.ones_like(y), torch.full_like(y, 0.3))
    return torch.ones_like(y)


class FakeHandle(object):
    def __init__(self, dtype="f16"):
        self.dtype_id = {"f16": 0, "bf16": 1}[dtype]
        self.h16 = torch.float16 if self.dtype_id == 0 else torch.bfloat16
        self.device = torch.device("cpu")
        self.num_sms = 148
        self.launches = 0

    def close(self):
        pass

    @contextlib.contextmanager
    def side_stream(self, after=None):     # no streams on the CPU double: everything is serial
        yield

    def mark(self):
        return None

    def join(self):
        pass

    def lstmp_fused_fwd(self, *a, **k):      # the double has no fused variant: callers fall back to gemm + rec
        return False

    def lstmp_wave_fwd(self, *a, **k):       # ... and no layer-wavefront launch: the layers run one after the other
        return False

    def lstmp_wave_bwd(self, *a, **k):
        return False

    def transpose16(self, src, rows, cols, dst):
        self.launches += 1
        dst[:cols, :rows] = src[:rows, :cols].t()

    # ------------------------------------------------------------------ GEMM
    def gemm(self, A, B, M, N, K, a_mn=False, b_mn=False, alpha=1.0, beta=0.0, bias=None, resid=None,
             act=ACT_NONE, dact_src=None, dact=ACT_NONE, out32=None, out16=None, tile_n=0, lda=None, ldb=None, split_k=0,
             stats=None):
        self.launches += 1
        Ae = (A[:K, :M].t() if a_mn else A[:M, :K]).float()
        Be = (B[:K, :N] if b_mn else B[:N, :K].t()).float()
        v = alpha * (Ae @ Be)
        if bias is not None:
            v = v + bias[:N]
        if resid is not None:
            v = v + resid[:M, :N]
        v = _act(v, act)
        if dact_src is not None:
            v = v * _dact(dact_src[:M, :N].float(), dact)
        if out16 is not None:
            out16[:M, :N] = v.to(self.h16)
        if out32 is not None:
            out32[:M, :N] = v + (beta * out32[:M, :N] if beta != 0.0 else 0.0)
        if stats is not None:       # rsr_gemm_args.stats: (count, mean, M2) of every 128-row block and column, [block][3][N]
            assert out32 is not None and out16 is None and bias is None and resid is None and act == ACT_NONE \
                and dact_src is None and alpha == 1.0 and beta == 0.0 and N % 32 == 0 and M <= self.BN_STATS_ROWS_MAX
            for blk in range((M + 127) // 128):
                z = v[128 * blk:128 * blk + 128]
                stats[3 * blk, :N] = float(z.shape[0])
                stats[3 * blk + 1, :N] = z.mean(0)
                stats[3 * blk + 2, :N] = ((z - z.mean(0)) ** 2).sum(0)

    # ------------------------------------------------- one-output fully_connected
    def fc1_fwd(self, x16, rows, K, w16, bias, out32):
        self.launches += 1
        out32[:rows, 0] = x16[:rows, :K].float() @ w16[:K, 0].float() + (bias[0] if bias is not None else 0.0)

    def fc1_bwd_dx(self, dy16, rows, K, w16, dx16, dact_src=None, dact=ACT_NONE):
        self.launches += 1
        v = dy16[:rows, :1].float() * w16[:K, 0].float().reshape(1, K)
        if dact_src is not None:
            v = v * _dact(dact_src[:rows, :K].float(), dact)
        dx16[:rows, :K] = v.to(self.h16)

    def fc1_head(self, x16, rows, K, w16, bias, which, clip, d_real, d_fake, grad_target, gscale, losses, logit32,
                 dlogit16=None, dact=ACT_NONE, dx16=None):
        self.launches += 1
        u = x16[:rows, :K].float() @ w16[:K, 0].float() + (bias[0] if bias is not None else 0.0)
        l = torch.clamp(u, -0.5, 1.5) if clip else u
        m = ((u >= -0.5) & (u <= 1.5)).float() if clip else torch.ones_like(u)
        if logit32 is not None:
            logit32[:rows, 0] = u
        if losses is not None:
            if which == 0:
                losses[0] += ((l - d_real) ** 2).mean()
            else:
                losses[1] += ((l - d_fake) ** 2).mean()
                losses[2] += ((l - d_real) ** 2).mean()
        g16 = (gscale * 2 * (l - (d_real if which == 0 else grad_target)) / rows * m).to(self.h16)
        if dlogit16 is not None:
            dlogit16[:rows, 0] = g16
        if dx16 is not None:
            v = g16.float().reshape(rows, 1) * w16[:K, 0].float().reshape(1, K)
            if dact != ACT_NONE:
                v = v * _dact(x16[:rows, :K].float(), dact)
            dx16[:rows, :K] = v.to(self.h16)

    # ------------------------------------------------------------- 1-D conv glue
    def conv_stage_frames(self, x, B, T, L, S, Cp, out16, mean=None, istd=None, time_major_in=False, ldx=None):
        self.launches += 1
        if time_major_in:
            v = x[:T * B, :L].float()
        else:
            v = x.reshape(B, T, -1)[:, :, :L].float().permute(1, 0, 2).reshape(T * B, L)
        if mean is not None:
            v = (v - mean) * istd
        o = out16[:T * B * S].view(T * B, S, Cp)
        o.zero_()
        o[:, :L, 0] = v.to(self.h16)

    def conv_mask_rows(self, buf16, frames, S, L, Cp):
        self.launches += 1
        buf16[:frames * S].view(frames, S, Cp)[:, L:].zero_()

    def conv_w_flip(self, w16, W, cin_p, cout_p, out16):
        self.launches += 1
        out16.view(W, cout_p, cin_p).copy_(w16.reshape(W, cin_p, cout_p).flip(0).transpose(1, 2))

    # --------------------------------------------------------------- staging
    def stage_input(self, x, B, T, D, out16=None, out32=None, mean=None, istd=None, noise=None,
                    time_major_in=False, ldx=None):
        self.launches += 1
        if time_major_in:
            v = x[:T * B, :D].float().reshape(T, B, D)
        else:
            v = x.reshape(B, T, -1)[:, :, :D].float().permute(1, 0, 2)
        if mean is not None:
            v = (v - mean) * istd
        if noise is not None:
            v = v + noise.reshape(1, B, D)
        v = v.reshape(T * B, D)
        if out16 is not None:
            out16[:, :D] = v.to(self.h16)
        if out32 is not None:
            out32[:, :D] = v

    def unstage_output(self, y_tm, B, T, D, out_bm, mean=None, std=None):
        self.launches += 1
        v = y_tm[:T * B, :D].reshape(T, B, D).permute(1, 0, 2)
        if mean is not None:
            v = v * std + mean
        out_bm.copy_(v)

    def cmvn_apply(self, x, mean, std, out):
        self.launches += 1
        out.copy_((x - mean) / std)

    def cmvn_apply_padded(self, x, lengths, mean64, std64, out):
        self.launches += 1
        B, T, D = x.shape
        v = ((x.double() - mean64) / std64).float()
        mask = torch.arange(T)[None, :] < lengths.long()[:, None]
        out.copy_(torch.where(mask[:, :, None], v, torch.zeros_like(v)))

    def cmvn_invert(self, y, mean, std, out):
        self.launches += 1
        out.copy_(y * std + mean)

    # ----------------------------------------------------------------- LSTMP
    @staticmethod
    def _gates(z, Cp):
        """packed [.., 4Cp] -> [.., 4, Cp]  (col = (cell/32)*128 + gate*32 + cell%32)"""
        s = z.shape[:-1]
        return z.reshape(*s, Cp // 32, 4, 32).transpose(-3, -2).reshape(*s, 4, Cp)

    @staticmethod
    def _pack(g, Cp):
        s = g.shape[:-2]
        return g.reshape(*s, 4, Cp // 32, 32).transpose(-3, -2).reshape(*s, 4 * Cp)

    def lstmp_rec_fwd(self, B, T, Cp, zx, wcT, w_i, w_f, w_o, lengths, mt_seq, save, forget_bias=1.0, work=0.0):
        self.launches += 1
        c = torch.zeros(B, Cp)
        W = wcT.float().t()                                  # [Cp, 4Cp] packed columns
        for t in range(T):
            mprev = mt_seq[t * B:(t + 1) * B].float()
            z = self._gates(zx[t * B:(t + 1) * B] + mprev @ W, Cp)
            ig = torch.sigmoid(z[:, 0] + w_i * c)
            fg = torch.sigmoid(z[:, 2] + forget_bias + w_f * c)
            jg = torch.tanh(z[:, 1])
            cn = fg * c + ig * jg
            og = torch.sigmoid(z[:, 3] + w_o * cn)
            mt = og * torch.tanh(cn)
            act = (t < lengths).reshape(B, 1)
            if save is not None:
                save[t * B:(t + 1) * B] = torch.cat([ig, fg, og, jg, cn], 1)
            mt_seq[(t + 1) * B:(t + 2) * B] = torch.where(act, mt, torch.zeros_like(mt)).to(self.h16)
            c = torch.where(act, cn, c)

    def lstmp_rec_bwd(self, B, T, Cp, dmt, wc, w_i, w_f, w_o, lengths, save, dz16, dbias, dw_i, dw_f, dw_o, work=0.0):
        self.launches += 1
        W = wc.float()                                       # [Cp, 4Cp]
        dcar = torch.zeros(B, Cp)
        for t in range(T - 1, -1, -1):
            sv = save[t * B:(t + 1) * B].reshape(B, 5, Cp)
            ig, fg, og, jg, cn = (sv[:, k] for k in range(5))
            cp = save[(t - 1) * B:t * B].reshape(B, 5, Cp)[:, 4] if t > 0 else torch.zeros(B, Cp)
            a = (t < lengths).reshape(B, 1).float()
            dm = dmt[t * B:(t + 1) * B]
            tc = torch.tanh(cn)
            dzo = dm * tc * og * (1 - og)
            dc = dcar + dm * og * (1 - tc * tc) + dzo * w_o
            dzf = dc * cp * fg * (1 - fg)
            dzi = dc * jg * ig * (1 - ig)
            dzj = dc * ig * (1 - jg * jg)
            dzi, dzj, dzf, dzo = dzi * a, dzj * a, dzf * a, dzo * a
            dcar = (dc * fg + dzf * w_f + dzi * w_i) * a
            dw_o += (dzo * cn).sum(0)
            dw_f += (dzf * cp).sum(0)
            dw_i += (dzi * cp).sum(0)
            dz = self._pack(torch.stack([dzi, dzj, dzf, dzo], 1), Cp)
            dbias += dz.sum(0)
            dzh = dz.to(self.h16)
            dz16[t * B:(t + 1) * B] = dzh
            if t > 0:
                dmt[(t - 1) * B:t * B] += dzh.float() @ W.t()

    # ---------------------------------------------------------------- losses
    def lsgan_mse_losses(self, losses, rl=None, fk=None, ld_logit=1, n_logit=0, clip=False, g=None, y=None,
                         n_frames=0, d_out=0, d_real=1.0, d_fake=0.0, lam=0.0, gscale=1.0,
                         d_rl_grad=None, d_fk_grad=None, g_adv_grad=None, ld_grad=1, dg_mse=None):
        self.launches += 1

        def prep(u):
            u = u[:n_logit, 0]
            if clip:
                return torch.clamp(u, -0.5, 1.5), ((u >= -0.5) & (u <= 1.5)).float()
            return u, torch.ones_like(u)

        if rl is not None:
            l, m = prep(rl)
            losses[0] += ((l - d_real) ** 2).mean()
            if d_rl_grad is not None:
                d_rl_grad[:n_logit, 0] = (gscale * 2 * (l - d_real) / n_logit * m).to(self.h16)
        if fk is not None:
            l, m = prep(fk)
            losses[1] += ((l - d_fake) ** 2).mean()
            losses[2] += ((l - d_real) ** 2).mean()
            if d_fk_grad is not None:
                d_fk_grad[:n_logit, 0] = (gscale * 2 * (l - d_fake) / n_logit * m).to(self.h16)
            if g_adv_grad is not None:
                g_adv_grad[:n_logit, 0] = (gscale * 2 * (l - d_real) / n_logit * m).to(self.h16)
        if g is not None:
            e = g[:n_frames, :d_out] - y[:n_frames, :d_out]
            losses[3] += 0.5 * d_out * (e ** 2).mean()
            if dg_mse is not None:
                dg_mse[:n_frames, :d_out] = gscale * lam * e / n_frames

    def colsum16(self, x16, M, N, out, accumulate=False, ld=None):
        self.launches += 1
        s = x16[:M, :N].float().sum(0)
        out[:N] = out[:N] + s if accumulate else s

    colsum32 = colsum16

    # ---------------------------------------------------------------- update
    def seg_sumsq(self, grad, gmul, seg_id, n_seg, sumsq):
        self.launches += 1
        b = ((grad * gmul) ** 2).reshape(-1, 1024).sum(1)
        sumsq.zero_()
        sumsq.index_add_(0, seg_id.long(), b)

    def _clipped(self, grad, gmul, seg_id, sumsq, max_norm):
        nrm = torch.sqrt(sumsq)[seg_id.long()]
        sc = gmul * (max_norm / torch.clamp_min(nrm, max_norm))
        return (grad.reshape(-1, 1024) * sc[:, None]).reshape(-1)

    @staticmethod
    def _skip(sumsq, n_seg, hyper):
        if n_seg and not bool(torch.isfinite(sumsq[:n_seg]).all()):
            hyper[7] += 1.0
            return True
        return False

    def clip_sgd_ema(self, grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, ema, theta16, n_seg=0):
        self.launches += 1
        if self._skip(sumsq, n_seg, hyper):
            return
        theta -= hyper[0] * self._clipped(grad, gmul, seg_id, sumsq, max_norm)
        if ema is not None:
            ema -= (1 - ema_decay) * (ema - theta)
        if theta16 is not None:
            theta16.copy_(theta.to(self.h16))

    def clip_adam_ema(self, grad, gmul, seg_id, sumsq, max_norm, hyper, ema_decay, theta, m, v, ema, theta16, n_seg=0):
        self.launches += 1
        if self._skip(sumsq, n_seg, hyper):
            return
        g = self._clipped(grad, gmul, seg_id, sumsq, max_norm)
        lr, b1, b2, eps, b1p, b2p = (float(hyper[i]) for i in range(6))
        lr_t = lr * (1 - b2p) ** 0.5 / (1 - b1p)
        m.mul_(b1).add_((1 - b1) * g)
        v.mul_(b2).add_((1 - b2) * g * g)
        theta -= lr_t * m / (torch.sqrt(v) + eps)
        hyper[4] *= b1
        hyper[5] *= b2
        if ema is not None:
            ema -= (1 - ema_decay) * (ema - theta)
        if theta16 is not None:
            theta16.copy_(theta.to(self.h16))

    def l2_grad(self, grad, theta, seg_id, seg_flag, scale):
        self.launches += 1
        f = seg_flag[seg_id.long()].float()
        grad += (theta.reshape(-1, 1024) * (scale * f)[:, None]).reshape(-1)

    def add_cast(self, a, b, n, out32=None, out16=None):
        self.launches += 1
        v = a.reshape(-1)[:n] + b.reshape(-1)[:n]
        if out32 is not None:
            out32.reshape(-1)[:n] = v
        if out16 is not None:
            out16.reshape(-1)[:n] = v.to(self.h16)

    def cast16(self, x, out16):
        self.launches += 1
        out16.copy_(x.to(self.h16))

    def fill32(self, x, v):
        self.launches += 1
        x.fill_(v)

    # ------------------------------------------------------ batch_norm(renorm) / dropout
    BN_EPS, BN_DECAY, BN_RENORM_DECAY = 1e-3, 0.999, 0.99

    def bn_train_stats(self, z32, rows, N, gamma, beta, state, coef, scratch, update_state=False):
        self.launches += 2
        z = z32[:rows, :N]
        mean = z.mean(0)
        self._bn_from_moments(mean, ((z - mean) ** 2).mean(0), N, gamma, beta, state, coef, update_state)

    def _bn_from_moments(self, mean, var, N, gamma, beta, state, coef, update_state):
        std = torch.sqrt(var + self.BN_EPS)
        mm, mv, rm, rs, rmw, rsw = (state[i, :N] for i in range(6))
        denom = rs + (1 - rsw) * std
        r, d = std / denom, (mean - (rm + (1 - rmw) * mean)) / denom
        A = r * gamma[:N] / std
        coef[0, :N], coef[1, :N], coef[2, :N], coef[3, :N] = A, d * gamma[:N] + beta[:N] - mean * A, mean, 1.0 / std
        coef[4, :N], coef[5, :N] = r, d
        if update_state:
            k = 1 - self.BN_RENORM_DECAY
            rm -= (rm - mean) * k
            rmw -= (rmw - 1) * k
            rs -= (rs - std) * k
            rsw -= (rsw - 1) * k
            mm -= (mm - rm / rmw) * (1 - self.BN_DECAY)
            mv -= (mv - ((rs / rsw) ** 2 - self.BN_EPS)) * (1 - self.BN_DECAY)

    BN_STATS_ROWS_MAX = 256 * 128

    def bn_train_finish(self, splits, rows, N, gamma, beta, state, coef, scratch, update_state=False):
        """rsr_bn_train_finish: the partials the GEMM epilogue left in `scratch`, merged in block order (Chan), then the same
        coefficients and UPDATE_OPS as bn_train_stats."""
        self.launches += 1
        n = torch.zeros(N)
        mean, m2 = torch.zeros(N), torch.zeros(N)
        for blk in range(splits):
            nb, mb, qb = scratch[3 * blk, :N], scratch[3 * blk + 1, :N], scratch[3 * blk + 2, :N]
            tot = n + nb
            d = mb - mean
            mean = mean + d * nb / tot
            m2 = m2 + qb + d * d * n * nb / tot
            n = tot
        assert float(n[0]) == rows
        self._bn_from_moments(mean, m2 / rows, N, gamma, beta, state, coef, update_state)

    def bn_eval_coef(self, N, gamma, beta, state, coef):
        self.launches += 1
        inv = torch.rsqrt(state[1, :N] + self.BN_EPS)
        coef[0, :N], coef[1, :N] = gamma[:N] * inv, beta[:N] - state[0, :N] * gamma[:N] * inv
        coef[2, :N], coef[3, :N], coef[4, :N], coef[5, :N] = state[0, :N], inv, 1.0, 0.0

    @staticmethod
    def _drop_mask(rng, salt, rows, N, keep_prob):
        import numpy as np
        M = (1 << 64) - 1

        def sm(x):                                     # splitmix64 finaliser on numpy uint64 (wraps)
            x = x ^ (x >> np.uint64(30))
            x = x * np.uint64(0xbf58476d1ce4e5b9)
            x = x ^ (x >> np.uint64(27))
            x = x * np.uint64(0x94d049bb133111eb)
            return x ^ (x >> np.uint64(31))
        seed, tick = (int(v) & M for v in rng.tolist())
        with np.errstate(over="ignore"):
            key = sm(np.uint64((seed + 0x9E3779B97F4A7C15 * (tick * 65536 + salt)) & M))
            hsh = sm(key ^ np.arange(rows * N // 2, dtype=np.uint64))      # one hash per (even, odd) column pair
            bits = np.stack([hsh >> np.uint64(40), (hsh >> np.uint64(16)) & np.uint64(0xffffff)], 1).reshape(rows, N)
        return torch.from_numpy((bits < np.uint64(int(float(np.float32(keep_prob)) * 16777216.0))))

    def affine_act_drop(self, z32, rows, N, A, Bc, act, keep_prob, rng, salt, out16, out32=None):
        self.launches += 1
        y = z32[:rows, :N] * (A[:N] if A is not None else 1.0) + Bc[:N]
        a = _act(y, act)
        if keep_prob < 1.0:
            a = torch.where(self._drop_mask(rng, salt, rows, N, keep_prob), a / keep_prob, torch.zeros_like(a))
        if out16 is not None:
            out16[:rows, :N] = a.to(self.h16)
        if out32 is not None:
            out32[:rows, :N] = a

    def bn_bwd(self, da16, z32, rows, N, act, keep_prob, rng, salt, bn, coef, bias, dgamma, dbeta, dz16, scratch,
               dz32=None):
        self.launches += 3
        z = z32[:rows, :N]
        y = z * coef[0, :N] + coef[1, :N] if bn else z + bias[:N]
        g = da16[:rows, :N].float() * _dact(y, act)
        if keep_prob < 1.0:
            g = torch.where(self._drop_mask(rng, salt, rows, N, keep_prob), g / keep_prob, torch.zeros_like(g))
        s1 = g.sum(0)
        if dbeta is not None:
            dbeta[:N] += s1
        if bn:
            xh = (z - coef[2, :N]) * coef[3, :N]
            s2 = (g * xh).sum(0)
            if dgamma is not None:
                dgamma[:N] += coef[4, :N] * s2 + coef[5, :N] * s1
            g = coef[0, :N] * (g - s1 / rows - xh * (s2 / rows))
        if dz16 is not None:
            dz16[:rows, :N] = g.to(self.h16)
        if dz32 is not None:
            dz32[:rows, :N] = g

    # batch_norm behind the convolutions of the frame layout: channel ch of line l = column l * C + ch, data rows r % S < L
    def _live(self, frames, S, L):
        return (torch.arange(frames * S) % S) < L

    def bn_train_stats_lines(self, z32, frames, S, L, H, C, N, gamma, beta, state, coef, scratch, update_state=False):
        self.launches += 2
        z = z32[:frames * S, :H * C][self._live(frames, S, L)].reshape(-1, C)
        mean = z.mean(0)
        var = ((z - mean) ** 2).mean(0)
        std = torch.sqrt(var + self.BN_EPS)
        mm, mv, rm, rs, rmw, rsw = (state[i, :C] for i in range(6))
        denom = rs + (1 - rsw) * std
        r, d = std / denom, (mean - (rm + (1 - rmw) * mean)) / denom
        A = r * gamma[:C] / std
        k = torch.stack([A, d * gamma[:C] + beta[:C] - mean * A, mean, 1.0 / std, r, d])
        coef[:, :N] = 0.0
        coef[:6, :H * C] = k.repeat(1, H)
        if update_state:
            kk = 1 - self.BN_RENORM_DECAY
            rm -= (rm - mean) * kk
            rmw -= (rmw - 1) * kk
            rs -= (rs - std) * kk
            rsw -= (rsw - 1) * kk
            mm -= (mm - rm / rmw) * (1 - self.BN_DECAY)
            mv -= (mv - ((rs / rsw) ** 2 - self.BN_EPS)) * (1 - self.BN_DECAY)

    def bn_eval_coef_lines(self, N, H, C, gamma, beta, state, coef):
        self.launches += 1
        inv = torch.rsqrt(state[1, :C] + self.BN_EPS)
        k = torch.stack([gamma[:C] * inv, beta[:C] - state[0, :C] * gamma[:C] * inv, state[0, :C], inv,
                         torch.ones_like(inv), torch.zeros_like(inv)])
        coef[:, :N] = 0.0
        coef[:6, :H * C] = k.repeat(1, H)

    def affine_act_lines(self, z32, frames, S, L, N, A, Bc, act, out16):
        self.launches += 1
        rows = frames * S
        a = _act(z32[:rows, :N] * (A[:N] if A is not None else 1.0) + Bc[:N], act)
        out16[:rows, :N] = torch.where(self._live(frames, S, L)[:, None], a, torch.zeros_like(a)).to(self.h16)

    def bn_bwd_lines(self, da16, z32, frames, S, L, H, C, N, act, coef, dgamma, dbeta, dz16, scratch):
        self.launches += 3
        rows, live = frames * S, self._live(frames, S, L)[:, None]
        z = z32[:rows, :N]
        g = da16[:rows, :N].float() * _dact(z * coef[0, :N] + coef[1, :N], act)
        g = torch.where(live, g, torch.zeros_like(g))
        xh = (z - coef[2, :N]) * coef[3, :N]
        s1 = g[:, :H * C].sum(0).reshape(H, C).sum(0)
        s2 = (g * xh)[:, :H * C].sum(0).reshape(H, C).sum(0)
        if dbeta is not None:
            dbeta[:C] += s1
        if dgamma is not None:
            dgamma[:C] += coef[4, :C] * s2 + coef[5, :C] * s1
        n = float(frames * L * H)
        m1, m2 = torch.zeros(N), torch.zeros(N)
        m1[:H * C], m2[:H * C] = (s1 / n).repeat(H), (s2 / n).repeat(H)
        g = coef[0, :N] * (g - m1 - xh * m2)
        dz16[:rows, :N] = torch.where(live, g, torch.zeros_like(g)).to(self.h16)

    def rng_tick(self, rng):
        self.launches += 1
        rng[1] += 1

    def gauss_noise(self, rng, salt, out, stddev):
        from oracle import rsr_oracle as O
        self.launches += 1
        out.copy_(torch.tensor(O.gauss_noise(int(rng[0]), int(rng[1]), salt, out.numel(), stddev)).reshape(out.shape))

    # ------------------------------------------------------ Kaldi compressed-matrix decode
    def ark_decompress(self, col_hdr, data, min_value, rng, rows, cols, out64=None, out32=None, mean=None, std=None):
        import io
        import numpy as np
        from rsrgan_b200.kaldi_io import ArkReader
        self.launches += 1
        buf = io.BytesIO(col_hdr.numpy().view(np.uint16).astype("<u2").tobytes() + data.numpy().tobytes())
        m = ArkReader().read_compress(min_value, rng, rows, cols, buf)
        if out64 is not None:
            out64[:rows, :cols] = torch.from_numpy(m)
        if out32 is not None:
            if mean is not None:
                m = (m - mean.numpy()) / std.numpy()
            out32[:rows, :cols] = torch.from_numpy(m.astype(np.float32))
