"""Writes tests/golden/gan_*.npz: seeded inputs + weights of small GANs and what the float64 oracle
(oracle/rsr_oracle.py) computes for them -- generator output, the seven losses, raw D / G
gradients and the generator output after one batch schedule (1 D + 2 G updates).

These vectors are the ORACLE's outputs (they pin the product to the oracle).  What pins the oracle to the reference is
tests/golden/make_reference_graph_golden.py -> tests/test_reference_graph.py (the reference's model files executed over a
TensorFlow stand-in; rsr_oracle.py header) and the independent torch-autograd statement (oracle/torch_ref.py) in
tests/test_oracle.py; TensorFlow 1.4 itself cannot be run here.

    python -m oracle.make_golden
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from . import rsr_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")

CASES = {
    # name: (g_type, d_type, sizes, B, T)
    "gan_lstm_dlstm": ("lstm", "lstm", dict(g_cell=64, g_proj=32, g_layers=2, d_cell=32), 3, 7),
    "gan_res_ddnn": ("res_lstm_l", "dnn", dict(g_cell=40, g_layers=2, d_units=64), 2, 5),
    # BASELINE.json configs[3] in small: RCED generator (models/rced.py, splice = 1) + discriminator_dnn on frames
    "gan_rced_ddnn": ("rced", "dnn", dict(d_units=64), 3, 2),
}

# MSE-only trainer (models/dnn_trainer_single_gpu.py; BASELINE.json configs[0] in small): name -> (g_type, sizes, frames)
MSE_CASES = {
    "mse_dnn": ("dnn", dict(g_units=64), 24),
    "mse_rced": ("rced", dict(), 6),
}

# the dnn generator as run_dnn_single_gpu.sh trains it (:129-145): batch_norm(renorm) + dropout + l2, UPDATE_OPS with
# every step.  name -> (sizes, frames, keep_prob, dropout seed)
BN_CASES = {
    "mse_dnn_bn": (dict(g_units=64), 40, 0.8, 11),
}


def build(name):
    g_type, d_type, sz, B, T = CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    if g_type == "lstm":
        gp = O.init_g_lstm(rng, cell=sz["g_cell"], proj=sz["g_proj"], layers=sz["g_layers"])
    elif g_type == "rced":
        gp = O.init_g_rced(rng)
    else:
        gp = O.init_g_res_lstm_l(rng, cell=sz["g_cell"], layers=sz["g_layers"])
    dp = O.init_d_lstm(rng, cell=sz["d_cell"]) if d_type == "lstm" else O.init_d_dnn(rng, units=sz["d_units"])
    # non-zero biases so that they are exercised
    for p in (gp, dp):
        for k in p:
            if "bias" in k:
                p[k] = rng.standard_normal(p[k].shape) * 0.1
    x = rng.standard_normal((B, T, 257))
    y = rng.standard_normal((B, T, 40))
    lengths = rng.integers(max(T // 2, 1), T + 1, size=B)
    lengths[0] = T
    n_rl = rng.standard_normal((B, 1, 40)) * 0.05 if d_type == "lstm" else None
    n_fk = rng.standard_normal((B, 1, 40)) * 0.05 if d_type == "lstm" else None
    return g_type, d_type, sz, gp, dp, x, y, lengths, n_rl, n_fk


def compute(name):
    g_type, d_type, sz, gp, dp, x, y, lengths, n_rl, n_fk = build(name)
    st = O.GanState(OrderedDict(gp), OrderedDict(dp), g_type, d_type)
    out = OrderedDict(x=x.astype(np.float32), y=y.astype(np.float32), lengths=lengths.astype(np.int32))
    if n_rl is not None:
        out["noise_rl"], out["noise_fk"] = n_rl.astype(np.float32), n_fk.astype(np.float32)
    for k, v in gp.items():
        out["G/" + k] = v.astype(np.float32)
    for k, v in dp.items():
        out["D/" + k] = v.astype(np.float32)
    # the oracle runs on the float32-rounded inputs/weights the device will also see
    st.g = OrderedDict((k, out["G/" + k].astype(np.float64)) for k in gp)
    st.d = OrderedDict((k, out["D/" + k].astype(np.float64)) for k in dp)
    x64, y64 = out["x"].astype(np.float64), out["y"].astype(np.float64)
    nr = out["noise_rl"].astype(np.float64) if n_rl is not None else None
    nf = out["noise_fk"].astype(np.float64) if n_fk is not None else None
    L, dgr, g_out = O.tower_losses_and_grads(st, x64, y64, lengths, "d", nr, nf)
    _, ggr, _ = O.tower_losses_and_grads(st, x64, y64, lengths, "g", nr, nf)
    out["g_out"] = g_out
    for k in ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_loss"):
        out["loss/" + k] = np.float64(L[k])
    for k, v in dgr.items():
        out["dgrad/" + k] = v.astype(np.float32)
    for k, v in ggr.items():
        out["ggrad/" + k] = v.astype(np.float32)
    tower = dict(x=x64, y=y64, lengths=lengths, noise_rl=nr, noise_fk=nf)
    O.d_step(st, [tower], 1e-3)
    O.g_step(st, [tower], 8e-5)
    O.g_step(st, [tower], 8e-5)
    gf, _ = O.GENERATORS[g_type]
    out["g_out_after"], _ = gf(st.g, x64, lengths)
    return out


def compute_mse(name, l2_scale=1e-4, lr=1e-3, steps=2):
    """x, y, weights; generator output, g_mse / g_l2 / g_loss, raw gradients; losses and output after `steps` Adam updates."""
    g_type, sz, N = MSE_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    gp = O.init_g_dnn(rng, units=sz["g_units"]) if g_type == "dnn" else O.init_g_rced(rng)
    for k in gp:
        if "bias" in k:
            gp[k] = gp[k] + rng.standard_normal(gp[k].shape) * 0.1
    out = OrderedDict(x=rng.standard_normal((N, 257)).astype(np.float32),
                      y=rng.standard_normal((N, 40)).astype(np.float32),
                      l2_scale=np.float64(l2_scale), lr=np.float64(lr), steps=np.int32(steps))
    for k, v in gp.items():
        out["G/" + k] = v.astype(np.float32)
    st = O.MseState(OrderedDict((k, out["G/" + k].astype(np.float64)) for k in gp), g_type)
    x64, y64 = out["x"].astype(np.float64), out["y"].astype(np.float64)
    L, gr, g_out = O.mse_losses_and_grads(st.g, g_type, x64, y64, l2_scale)
    out["g_out"] = g_out
    for k, v in L.items():
        out["loss/" + k] = np.float64(v)
    for k, v in gr.items():
        out["ggrad/" + k] = v.astype(np.float32)
    for _ in range(steps):
        O.mse_step(st, x64, y64, lr, l2_scale)
    L, _, g_out = O.mse_losses_and_grads(st.g, g_type, x64, y64, l2_scale)
    out["g_out_after"] = g_out
    for k, v in L.items():
        out["loss_after/" + k] = np.float64(v)
    return out


def compute_mse_bn(name, l2_scale=1e-4, lr=1e-3, steps=3):
    """Batch-normalised dnn generator: weights (gamma / beta perturbed), training-graph output / losses / raw gradients
    of the first step (dropout tick 0), then `steps` Adam updates with the UPDATE_OPS on fresh minibatches: the
    non-trainable batch_norm variables and the INFERENCE-graph output (moving averages, no dropout) afterwards."""
    sz, N, keep, seed = BN_CASES[name]
    rng = np.random.default_rng(sum(map(ord, name)))
    gp = O.init_g_dnn(rng, units=sz["g_units"], batch_norm=True)
    for k in gp:
        if "bias" in k or "BatchNorm" in k:
            gp[k] = gp[k] + rng.standard_normal(gp[k].shape) * 0.1
    out = OrderedDict(x=rng.standard_normal((steps, N, 257)).astype(np.float32),
                      y=rng.standard_normal((steps, N, 40)).astype(np.float32),
                      l2_scale=np.float64(l2_scale), lr=np.float64(lr), steps=np.int32(steps),
                      keep_prob=np.float64(keep), seed=np.int64(seed))
    for k, v in gp.items():
        out["G/" + k] = v.astype(np.float32)
    st = O.MseState(OrderedDict((k, out["G/" + k].astype(np.float64)) for k in gp), "dnn")
    bst = O.init_bn_state(st.g)
    x64, y64 = out["x"].astype(np.float64), out["y"].astype(np.float64)
    opts = lambda tick, update: dict(bn_state=bst, update=update, keep_prob=keep, rng=(seed, tick))
    L, gr, g_out = O.mse_losses_and_grads(st.g, "dnn", x64[0], y64[0], l2_scale, opts(0, False))
    out["g_out"] = g_out
    for k, v in L.items():
        out["loss/" + k] = np.float64(v)
    for k, v in gr.items():
        out["ggrad/" + k] = v.astype(np.float32)
    for t in range(steps):
        L, _ = O.mse_step(st, x64[t], y64[t], lr, l2_scale, opts(t, True))
        out["loss_step%d/g_mse_loss" % t] = np.float64(L["g_mse_loss"])
    for k, v in bst.items():
        out["BN/" + k] = np.asarray(v, np.float64)
    out["g_out_after"], _ = O.g_dnn_fwd(st.g, x64[0], None, opts=dict(bn_state=bst, train=False))
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    import sys
    only = sys.argv[1:]
    for name in list(CASES) + list(MSE_CASES) + list(BN_CASES):
        if only and name not in only:
            continue
        d = compute(name) if name in CASES else compute_mse(name) if name in MSE_CASES else compute_mse_bn(name)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
        print("wrote", name, os.path.getsize(os.path.join(OUT, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
