"""GPU bring-up probe for the whole GAN step: builds a small GAN_RNN, loads the same weights into
the oracle (float64 numpy), and prints the RMS error of G outputs, losses, gradients and weight
deltas after a D update and a G update.  Diagnostic for `gpurun`; the asserting versions live in
tests/test_gan_gpu.py.

    python scripts/gpu_probe_model.py [lstm|res_lstm_l|res_lstm_base] [lstm|dnn] [B] [T]
"""
import os
import sys
import time
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402
from oracle import rsr_oracle as O  # noqa: E402  (checker only)


def rms(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    d = np.sqrt(((a - b) ** 2).mean())
    return d, d / (np.sqrt((b ** 2).mean()) + 1e-30)


def make(g_type, d_type, B, dtype, small):
    kw = {}
    if small:
        kw = dict(g_cell=96, g_proj=48, d_cell=64, d_units=128)
    args = Namespace(g_type=g_type, d_type=d_type, batch_size=B, init_mse_weight=10.0, init_disc_noise_std=0.05,
                     g_learning_rate=8e-5, d_learning_rate=1e-3, l2_scale=0.0, dtype=dtype, seed=7, **kw)
    return GAN_RNN(None, args, ["/gpu:0"])


def oracle_state(model):
    gp = OrderedDict((k, v.astype(np.float64)) for k, v in model.G.P.export_tf().items())
    dp = OrderedDict((k, v.astype(np.float64)) for k, v in model.D.P.export_tf().items())
    return O.GanState(gp, dp, model.g_type, model.d_type)


def main():
    g_type = sys.argv[1] if len(sys.argv) > 1 else "lstm"
    d_type = sys.argv[2] if len(sys.argv) > 2 else "lstm"
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    T = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    small = os.environ.get("RSR_SMALL", "0") == "1"
    dtype = os.environ.get("RSR_DTYPE", "f16")
    model = make(g_type, d_type, B, dtype, small)
    print("model %s + D-%s  B%d T%d dtype %s  G params %d  D params %d" %
          (g_type, d_type, B, T, dtype, model.G.P.n_params(), model.D.P.n_params()), flush=True)
    rng = np.random.default_rng(3)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    y = rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[0] = T
    n_rl = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32)
    n_fk = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32)
    st = oracle_state(model)

    # ---- forward parity
    g = model.generate(x, lengths).cpu().numpy()
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(st.g, x.astype(np.float64), lengths)
    print("G forward: rms %.3e rel %.3e (ref rms %.3f)" % (rms(g, g_ref) + (np.sqrt((g_ref ** 2).mean()),)), flush=True)
    ev = model.eval_losses(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    tower = dict(x=x.astype(np.float64), y=y.astype(np.float64), lengths=lengths,
                 noise_rl=n_rl.astype(np.float64) if d_type == "lstm" else None,
                 noise_fk=n_fk.astype(np.float64) if d_type == "lstm" else None)
    L, _, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "d", tower["noise_rl"], tower["noise_fk"])
    for k in ("d_rl_loss", "d_fk_loss", "g_adv_loss", "g_mse_loss", "g_loss"):
        print("  %-10s got %.6f ref %.6f" % (k, ev[k], L[k]))

    # ---- D update
    model.d_learning_rate = 1e-3
    model.g_learning_rate = 8e-5
    d0 = model.D.P.export_tf()
    t0 = time.time()
    model.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    torch.cuda.synchronize()
    print("d_step %.1f ms" % (1e3 * (time.time() - t0)))
    gs = model._gscale(B * T)
    dgrad = model.D.P.export_tf("grad")
    _, ref_clipped = O.d_step(st, [tower], 1e-3)
    _, ref_raw, _ = O.tower_losses_and_grads(oracle_state_from(d0, model, st, "d"), tower["x"], tower["y"], lengths, "d",
                                             tower["noise_rl"], tower["noise_fk"])
    d1 = model.D.P.export_tf()
    for k in ref_raw:
        e = rms(dgrad[k] / gs, ref_raw[k])
        w = rms(d1[k] - d0[k], st.d[k] - d0[k].astype(np.float64))
        print("  D grad %-60s rel %.3e | delta rel %.3e" % (k, e[1], w[1]))

    # ---- G update (oracle D already updated by O.d_step above, like ours)
    g0 = model.G.P.export_tf()
    t0 = time.time()
    model.g_step(x, y, lengths, noise_fk=n_fk)
    torch.cuda.synchronize()
    print("g_step %.1f ms" % (1e3 * (time.time() - t0)))
    ggrad = model.G.P.export_tf("grad")
    _, ref_raw, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "g", tower["noise_rl"], tower["noise_fk"])
    O.g_step(st, [tower], 8e-5)
    g1 = model.G.P.export_tf()
    for k in ref_raw:
        e = rms(ggrad[k] / gs, ref_raw[k])
        w = rms(g1[k] - g0[k], st.g[k] - g0[k].astype(np.float64))
        print("  G grad %-60s rel %.3e | delta rel %.3e" % (k, e[1], w[1]))
    g = model.generate(x, lengths).cpu().numpy()
    g_ref, _ = gf(st.g, x.astype(np.float64), lengths)
    print("G forward after updates: rms %.3e rel %.3e" % rms(g, g_ref))
    print("probe done")


def oracle_state_from(d_params, model, st, which):
    s = O.GanState(st.g, OrderedDict((k, v.astype(np.float64)) for k, v in d_params.items()), model.g_type, model.d_type)
    return s


if __name__ == "__main__":
    main()
