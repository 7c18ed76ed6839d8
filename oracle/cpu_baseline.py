"""CPU baseline: the reference's GAN batch schedule restated in PyTorch-CPU fp32 (autograd).
TEST / MEASUREMENT INFRASTRUCTURE ONLY -- imported by bench.py's `cpu_baseline` leg and by
`bench.py --impl reference`, never by rsrgan_b200/.

The reference's own TF-1.4 CPU path cannot run here (python2 + tensorflow 1.4 absent, SURVEY.md
section 8c), so this is kind = "port": the same math as oracle/torch_ref.py, driven through the
schedule of scripts/train_gan_rnn_placeholder.py:72-101 (1 D update + 2 G updates per batch,
G forward recomputed each time, per-tensor clip_by_norm(15), SGD for D, TF-form Adam for G, EMA).
"""
from __future__ import annotations

import math
import os
import time
from collections import OrderedDict

import numpy as np
import torch

from . import torch_ref as R
from . import rsr_oracle as O


def _params(model_cfg, seed):
    rng = np.random.default_rng(seed)
    g_type, d_type = model_cfg["g_type"], model_cfg["d_type"]
    if g_type == "rced":
        gp = O.init_g_rced(rng, dtype=np.float32)
    elif g_type == "dnn":
        gp = O.init_g_dnn(rng, dtype=np.float32)
    elif g_type == "lstm":
        gp = O.init_g_lstm(rng, cell=model_cfg["g_cell"], proj=model_cfg["g_proj"], layers=model_cfg["g_layers"],
                           dtype=np.float32)
    else:
        gp = O.init_g_res_lstm_l(rng, cell=model_cfg["g_cell"], layers=model_cfg["g_layers"], dtype=np.float32)
    if d_type == "lstm":
        dp = O.init_d_lstm(rng, dtype=np.float32)
    else:
        dp = O.init_d_dnn(rng, dtype=np.float32)
    return gp, dp


class CpuGan(object):
    def __init__(self, model_cfg, seed=1234, threads=None):
        torch.set_num_threads(threads or os.cpu_count())
        self.cfg = model_cfg
        gp, dp = _params(model_cfg, seed)
        self.g = R.to_torch(gp, torch.float32, True)
        self.d = R.to_torch(dp, torch.float32, True)
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.g.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.g.items())
        self.g_ema = OrderedDict((k, v.detach().clone()) for k, v in self.g.items())
        self.d_ema = OrderedDict((k, v.detach().clone()) for k, v in self.d.items())
        self.t = 0

    @staticmethod
    def _clip(g, max_norm=15.0):
        return g * (max_norm / torch.clamp_min(g.norm(), max_norm))

    def d_step(self, x, y, lengths, lr, noise_std=0.05):
        B = x.shape[0]
        nz = (lambda: torch.randn(B, 1, y.shape[-1]) * noise_std) if self.cfg["d_type"] == "lstm" else (lambda: None)
        ls, grads, _ = R.grads(self.g, self.d, self.cfg["g_type"], self.cfg["d_type"], x, y, lengths, "d",
                               noise_rl=nz(), noise_fk=nz(), mse_lambda=10.0)
        with torch.no_grad():
            for k, gr in grads.items():
                self.d[k] -= lr * self._clip(gr)
                self.d_ema[k] -= (1 - 0.9999) * (self.d_ema[k] - self.d[k])
        return ls

    def g_step(self, x, y, lengths, lr, noise_std=0.05):
        B = x.shape[0]
        nz = (lambda: torch.randn(B, 1, y.shape[-1]) * noise_std) if self.cfg["d_type"] == "lstm" else (lambda: None)
        ls, grads, _ = R.grads(self.g, self.d, self.cfg["g_type"], self.cfg["d_type"], x, y, lengths, "g",
                               noise_rl=nz(), noise_fk=nz(), mse_lambda=10.0)
        self.t += 1
        lr_t = lr * math.sqrt(1 - 0.999 ** self.t) / (1 - 0.9 ** self.t)
        with torch.no_grad():
            for k, gr in grads.items():
                gr = self._clip(gr)
                self.m[k].mul_(0.9).add_(0.1 * gr)
                self.v[k].mul_(0.999).add_(0.001 * gr * gr)
                self.g[k] -= lr_t * self.m[k] / (self.v[k].sqrt() + 1e-8)
                self.g_ema[k] -= (1 - 0.9999) * (self.g_ema[k] - self.g[k])
        return ls

    def schedule(self, x, y, lengths, disc_updates=1, gen_updates=2):
        for _ in range(disc_updates):
            self.d_step(x, y, lengths, 1e-3)
        for _ in range(gen_updates):
            ls = self.g_step(x, y, lengths, 8e-5)
        return float(ls["g_loss"].detach())


def time_schedule(model_cfg, B, T, steps=1, warmup=0, seed=1234, threads=None):
    """Returns (frames_per_sec, seconds_per_schedule, cores)."""
    gan = CpuGan(model_cfg, seed, threads)
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, 257, generator=g)
    y = torch.randn(B, T, 40, generator=g)
    lengths = torch.full((B,), T, dtype=torch.int64)
    for _ in range(warmup):
        gan.schedule(x, y, lengths)
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        gan.schedule(x, y, lengths)
        ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    return B * T / dt, dt, torch.get_num_threads()
