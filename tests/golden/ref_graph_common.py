"""Shared by tests/golden/make_reference_graph_golden.py (which executes the reference's model files, here, over
tests/golden/tf_standin.py) and tests/test_reference_graph.py (which replays the oracle against the committed fixtures):
the seeded parameter sets / feeds of every case, and the compact form in which big tensors are stored."""
import zlib
from collections import OrderedDict

import numpy as np

from oracle import rsr_oracle as O

# case -> (generator type, towers, utterances per tower, frames, l2_scale, input scale)
GAN_RNN_CASES = OrderedDict([
    ("lstm_2towers", dict(g_type="lstm", towers=2, B=2, T=5, l2_scale=1e-4, scale=3.0, seed=101)),
    ("res_lstm_l_1tower", dict(g_type="res_lstm_l", towers=1, B=3, T=4, l2_scale=0.0, scale=2.0, seed=102)),
    ("res_lstm_base_1tower", dict(g_type="res_lstm_base", towers=1, B=2, T=4, l2_scale=1e-4, scale=2.0, seed=103)),
])
LR_D, LR_G, NOISE_STD, MSE_LAMBDA = 1e-3, 8e-5, 0.05, 10.0


def gan_rnn_setup(case):
    """Parameters (reference-native sizes, TF names), feeds and unit-variance noise draws of one case, from its seed."""
    c = GAN_RNN_CASES[case]
    rng = np.random.default_rng(c["seed"])
    init = {"lstm": O.init_g_lstm, "res_lstm_l": O.init_g_res_lstm_l, "res_lstm_base": O.init_g_res_lstm_l}[c["g_type"]]
    gp, dp = init(rng), O.init_d_lstm(rng)
    for p in (gp, dp):                      # zero-initialised biases would hide a bias wired to the wrong place
        for k in p:
            if k.endswith("bias") or k.endswith("biases"):
                p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    n = c["towers"] * c["B"]
    x = c["scale"] * rng.standard_normal((n, c["T"], 257))
    y = c["scale"] * rng.standard_normal((n, c["T"], 40))
    lengths = rng.integers(max(c["T"] // 2, 1), c["T"] + 1, size=n)
    lengths[0] = c["T"]
    noise = [rng.standard_normal((c["B"], 1, 40)) for _ in range(1 + 2 * c["towers"])]     # dummy D, then (rl, fk) per tower
    return c, gp, dp, x, y, lengths, noise


def compact(name, a, full_below=600, samples=64):
    """Small tensors whole; big ones as (sum, l2 norm, `samples` entries at indices derived from the name)."""
    a = np.asarray(a, np.float64)
    if a.size <= full_below:
        return {"full": a}
    idx = np.random.default_rng(zlib.crc32(name.encode())).choice(a.size, samples, replace=False)
    return {"sum": np.float64(a.sum()), "l2": np.float64(np.sqrt((a * a).sum())), "idx": idx, "at": a.reshape(-1)[idx]}


def pack(store, prefix, tensors):
    for k, v in tensors.items():
        for f, a in compact(k, v).items():
            store["%s|%s|%s" % (prefix, k, f)] = a


def check(fix, prefix, tensors, rtol=1e-9, atol=1e-12):
    """Every tensor the fixture holds under `prefix` against `tensors` (same compact form).  Returns how many were compared."""
    names = sorted({k.split("|")[1] for k in fix.files if k.startswith(prefix + "|")})
    assert names and set(names) == set(tensors), (prefix, sorted(set(names) ^ set(tensors)))
    for k in names:
        mine = compact(k, tensors[k])
        for f, a in mine.items():
            ref = fix["%s|%s|%s" % (prefix, k, f)]
            if f == "idx":
                assert np.array_equal(ref, a), (prefix, k)
            else:
                scale = float(np.abs(ref).max()) if np.size(ref) else 0.0
                assert np.allclose(a, ref, rtol=rtol, atol=atol + rtol * scale), (prefix, k, f, float(np.abs(a - ref).max()), scale)
    return len(names)


# frame-level trainers: models/gan.py (DNN generator + conditioned discriminator_dnn, Adam for both, no clipping) and
# models/dnn_trainer_single_gpu.py (MSE + l2, Adam) at the reference's layer sizes, splice 5 + 1 + 5
FRAME_CASES = OrderedDict([
    ("gan_dnn", dict(N=6, l2_scale=1e-4, scale=2.0, seed=201, lr_d=1e-3, lr_g=8e-5)),
    ("dnn_trainer", dict(N=5, l2_scale=1e-3, scale=2.0, seed=202, lr_g=1e-3)),
])
LEFT = RIGHT = 5


def frame_setup(case):
    c = FRAME_CASES[case]
    rng = np.random.default_rng(c["seed"])
    in_dim = 257 * (LEFT + 1 + RIGHT)
    gp = O.init_g_dnn(rng, in_dim=in_dim, out_dim=40, units=1024, hidden=3)
    dp = O.init_d_dnn(rng, in_dim=257 + 40, units=1024, hidden=3) if case == "gan_dnn" else OrderedDict()
    for p in (gp, dp):
        for k in p:
            if k.endswith("biases"):
                p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    x = c["scale"] * rng.standard_normal((c["N"], in_dim))
    y = c["scale"] * rng.standard_normal((c["N"], 40))
    return c, gp, dp, x, y


# the convolutional generator (models/rced.py) under the reference's multi-tower MSE trainer (models/dnn_trainer.py):
# splice 1 (BASELINE configs[3]) and a [3, w] case of the 2-D convolutions run_dnn.sh trains with splice 11
RCED_CASES = OrderedDict([
    ("rced_splice1", dict(N=4, ctx=0, l2_scale=1e-3, scale=1.5, seed=301, lr_g=1e-3)),
    ("rced_splice3", dict(N=3, ctx=1, l2_scale=0.0, scale=1.5, seed=302, lr_g=1e-3)),
])


def rced_setup(case):
    c = RCED_CASES[case]
    rng = np.random.default_rng(c["seed"])
    splice = 2 * c["ctx"] + 1
    gp = O.init_g_rced(rng, in_dim=257, out_dim=40, splice=splice)
    for k in gp:
        if k.endswith("biases"):
            gp[k] = gp[k] + 0.1 * rng.standard_normal(gp[k].shape)
    x = c["scale"] * rng.standard_normal((c["N"], splice * 257))
    y = c["scale"] * rng.standard_normal((c["N"], 40))
    return c, gp, x, y


# the reference's own training loop (scripts/train_gan_rnn_placeholder.py train_one_iteration) over three queued
# minibatches, the second one short of an utterance (skipped, :69-70); discriminator noise off
SCHEDULE = dict(g_type="lstm", B=2, T=5, l2_scale=1e-4, scale=2.0, seed=401, batches=(2, 1, 2))


def schedule_setup():
    c = SCHEDULE
    rng = np.random.default_rng(c["seed"])
    gp, dp = O.init_g_lstm(rng), O.init_d_lstm(rng)
    for p in (gp, dp):
        for k in p:
            if k.endswith("bias") or k.endswith("biases"):
                p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    batches = []
    for n in c["batches"]:
        lengths = rng.integers(max(c["T"] // 2, 1), c["T"] + 1, size=n)
        lengths[0] = c["T"]
        batches.append((c["scale"] * rng.standard_normal((n, c["T"], 257)), c["scale"] * rng.standard_normal((n, c["T"], 40)), lengths))
    return c, gp, dp, batches
