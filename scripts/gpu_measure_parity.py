"""Measures what the 16-bit tensor-core path actually deviates from the float64 oracle at the BENCHMARKED sequence
lengths (cfg-2: T = 100, cfg-5: T = 200) -- generator-output RMS, losses, per-tensor gradient error of one D and one G
update, generator output after a whole schedule -- and prints one JSON line per case.  The bars in tests/test_gan_gpu.py
(test_cfg2_T100_*, test_cfg5_T200_*) are set from these numbers (2x the measured gradient error).

    python scripts/gpu_measure_parity.py f16|bf16 [cfg2|cfg5|all]        (RSR_FAST_GATES=1: MUFU.TANH gate math)
"""
import json
import os
import sys
from argparse import Namespace
from collections import OrderedDict

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rsr_oracle as O  # noqa: E402  (checker only)
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402

dtype = sys.argv[1] if len(sys.argv) > 1 else "f16"
which = sys.argv[2] if len(sys.argv) > 2 else "all"
tag = dict(dtype=dtype, fast_gates=int(os.environ.get("RSR_FAST_GATES", "0")))


def rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = float(np.sqrt(((a - b) ** 2).mean()))
    return d, d / (float(np.sqrt((b ** 2).mean())) + 1e-30)


def make(g_type, d_type, B, **kw):
    a = dict(g_type=g_type, d_type=d_type, batch_size=B, init_mse_weight=10.0, init_disc_noise_std=0.05,
             g_learning_rate=8e-5, d_learning_rate=1e-3, l2_scale=0.0, seed=3, dtype=dtype, use_graph=False)
    a.update(kw)
    return GAN_RNN(None, Namespace(**a), ["/gpu:0"])


def state(m, g_type, d_type):
    return O.GanState(OrderedDict((k, v.astype(np.float64)) for k, v in m.G.P.export_tf().items()),
                      OrderedDict((k, v.astype(np.float64)) for k, v in m.D.P.export_tf().items()), g_type, d_type)


def out(**kw):
    print(json.dumps(dict(tag, **kw)), flush=True)


def g_output_slice(name, g_type, d_type, B, T, idx, kw):
    m = make(g_type, d_type, B, **kw)
    rng = np.random.default_rng(B + T)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[idx[0]] = T
    st = state(m, g_type, d_type)
    g = m.generate(x, lengths).cpu().numpy()
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(st.g, x[idx].astype(np.float64), lengths[idx])
    a, r = rms(g[idx], g_ref)
    # error growth along the sequence: RMS over the last tenth of the longest utterance
    tail = rms(g[idx[0], -T // 10:], g_ref[0, -T // 10:])[0]
    out(case=name + "/g_output", B=B, T=T, utterances=len(idx), abs_rms=a, rel_rms=r, abs_rms_last_tenth=tail)


def schedule(name, g_type, d_type, B, T, kw):
    m = make(g_type, d_type, B, **kw)
    rng = np.random.default_rng(B * T)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[0] = T
    lstm_d = d_type == "lstm"
    n_rl = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32) if lstm_d else None
    n_fk = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32) if lstm_d else None
    st = state(m, g_type, d_type)
    tower = dict(x=x.astype(np.float64), y=y.astype(np.float64), lengths=lengths,
                 noise_rl=None if n_rl is None else n_rl.astype(np.float64),
                 noise_fk=None if n_fk is None else n_fk.astype(np.float64))
    gs = m._gscale(B * T)
    # raw gradients (learning rates 0 keep the weights)
    m.d_learning_rate, m.g_learning_rate = 0.0, 0.0
    ours = m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    Ld, Gd, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "d", tower["noise_rl"], tower["noise_fk"])
    dg = m.D.P.export_tf("grad")
    d_err = {k: rms(dg[k] / gs, Gd[k])[1] for k in Gd}
    d_loss_err = abs(ours["d_loss"] - Ld["d_loss"]) / abs(Ld["d_loss"])
    ours = m.g_step(x, y, lengths, noise_fk=n_fk)
    Lg, Gg, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "g", tower["noise_rl"], tower["noise_fk"])
    gg = m.G.P.export_tf("grad")
    g_err = {k: rms(gg[k] / gs, Gg[k])[1] for k in Gg}
    g_loss_err = abs(ours["g_loss"] - Lg["g_loss"]) / abs(Lg["g_loss"])
    out(case=name + "/grads", B=B, T=T, d_loss_rel=d_loss_err, g_loss_rel=g_loss_err,
        d_grad_rel_max=max(d_err.values()), d_grad_worst=max(d_err, key=d_err.get),
        g_grad_rel_max=max(g_err.values()), g_grad_worst=max(g_err, key=g_err.get),
        g_grad_rel_median=float(np.median(list(g_err.values()))), d_grad_rel_median=float(np.median(list(d_err.values()))))
    # one whole schedule with the reference learning rates
    m.d_learning_rate, m.g_learning_rate = 1e-3, 8e-5
    import torch
    m.G.P.m.zero_(); m.G.P.v.zero_(); m.G.P.hyper[4:6] = torch.tensor([0.9, 0.999], device=m.h.device)
    m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    O.d_step(st, [tower], 1e-3)
    for _ in range(2):
        m.g_step(x, y, lengths, noise_fk=n_fk)
        O.g_step(st, [tower], 8e-5)
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(st.g, tower["x"], lengths)
    a, r = rms(m.generate(x, lengths).cpu().numpy(), g_ref)
    out(case=name + "/after_schedule", B=B, T=T, abs_rms=a, rel_rms=r)


CFG2 = dict(g_cell=512, g_proj=256, g_layers=2)
CFG5 = dict(g_cell=1024, g_layers=4)
if which in ("cfg2", "all"):
    g_output_slice("cfg2", "lstm", "dnn", 128, 100, [0, 17, 31, 32, 63, 64, 100, 127], CFG2)
    schedule("cfg2", "lstm", "dnn", 24, 100, CFG2)
if which in ("cfg5", "all"):
    g_output_slice("cfg5", "res_lstm_l", "lstm", 64, 200, [0, 21, 42, 63], CFG5)
    schedule("cfg5", "res_lstm_l", "lstm", 4, 200, CFG5)
