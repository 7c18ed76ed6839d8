"""The CUDA path against the reference's OWN graph code: tests/golden/ref_graph_gan_rnn_*.npz hold what
models/gan_rnn_placeholder.py (with lstm.py / res_lstm_l.py / res_lstm_base.py / discriminator_lstm.py) computed when it was
executed in the build container over the TensorFlow stand-in (tests/golden/make_reference_graph_golden.py) -- generator
output, the seven losses and the raw gradients of every tower at the reference-native layer sizes.  Here the product
(GAN_RNN over the C ABI, fp16 operands) replays one tower of each case from the same seeded parameters and feeds."""
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

from test_gan_gpu import make_model, rms

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import ref_graph_common as C  # noqa: E402

pytestmark = pytest.mark.gpu
LOSS_REL = 3e-3              # fp16 operands against float64
GRAD_BAR = 1.2e-2            # per-tensor relative RMS of raw gradients: 2x the measured worst (5.8e-3, a bias; typical 1e-3)


def _grad_dev(fix, prefix, mine, gs):
    """Largest per-tensor deviation of `mine / gs` from the fixture's (compact) gradients: relative RMS over the stored
    entries (the whole tensor, or 64 sampled entries measured against the tensor's RMS from its stored l2 norm)."""
    worst = (0.0, None)
    names = sorted({k.split("|")[1] for k in fix.files if k.startswith(prefix + "|")})
    assert set(names) == set(mine), sorted(set(names) ^ set(mine))
    for k in names:
        g = np.asarray(mine[k], np.float64) / gs
        if prefix + "|" + k + "|full" in fix.files:
            dev = rms(g, fix[prefix + "|" + k + "|full"])[1]
        else:
            idx, at, l2 = fix[prefix + "|" + k + "|idx"], fix[prefix + "|" + k + "|at"], float(fix[prefix + "|" + k + "|l2"])
            typical = l2 / np.sqrt(g.size)
            dev = float(np.sqrt(((g.reshape(-1)[idx] - at) ** 2).mean())) / (typical + 1e-30)
            dev = max(dev, abs(float(np.sqrt((g * g).sum())) - l2) / (l2 + 1e-30))
        if dev > worst[0]:
            worst = (dev, k)
    return worst


@pytest.mark.parametrize("case,tower", [("lstm_2towers", 0), ("lstm_2towers", 1), ("res_lstm_l_1tower", 0),
                                        ("res_lstm_base_1tower", 0)])
def test_cuda_path_against_the_reference_graph(case, tower):
    fix = np.load(os.path.join(GOLD, "ref_graph_gan_rnn_%s.npz" % case))
    c, gp, dp, x, y, lengths, noise = C.gan_rnn_setup(case)
    B, T = c["B"], c["T"]
    sl = slice(B * tower, B * (tower + 1))                                     # gan_rnn_placeholder.py:157-159
    xs, ys, ls = x[sl].astype(np.float32), y[sl].astype(np.float32), lengths[sl]
    n_rl = (C.NOISE_STD * noise[1 + 2 * tower]).astype(np.float32)
    n_fk = (C.NOISE_STD * noise[2 + 2 * tower]).astype(np.float32)
    m = make_model(c["g_type"], "lstm", B, l2_scale=c["l2_scale"], init_mse_weight=C.MSE_LAMBDA, use_graph=False)
    m.load_params(OrderedDict((k, v.astype(np.float32)) for k, v in gp.items()),
                  OrderedDict((k, v.astype(np.float32)) for k, v in dp.items()))
    seen = {}
    a, r = rms(m.generate(xs, ls).cpu().numpy(), fix["fwd|tower%d/g_clean|full" % tower])
    seen["g_out_rel"] = r
    assert r < 1e-3, (a, r)                     # measured 2.4e-4 .. 3.0e-4 (profiles/r2_reference_graph_gpu.txt)
    gs = m._gscale(B * T)
    m.d_learning_rate, m.g_learning_rate = 0.0, 0.0                            # raw gradients of one D and one G update
    ours = m.d_step(xs, ys, ls, noise_rl=n_rl, noise_fk=n_fk)
    for fk, ok in (("d_rl_losses", "d_rl_loss"), ("d_fk_losses", "d_fk_loss"), ("d_losses", "d_loss")):
        assert ours[ok] == pytest.approx(float(fix["loss|" + fk][tower]), rel=LOSS_REL, abs=1e-5), (fk, seen)
    seen["grad_d"] = _grad_dev(fix, "grad_d_tower%d" % tower, m.D.P.export_tf("grad"), gs)
    ours = m.g_step(xs, ys, ls, noise_fk=n_fk)
    for fk, ok in (("g_adv_losses", "g_adv_loss"), ("g_mse_losses", "g_mse_loss"), ("g_l2_losses", "g_l2_loss"),
                   ("g_losses", "g_loss")):
        assert ours[ok] == pytest.approx(float(fix["loss|" + fk][tower]), rel=LOSS_REL, abs=1e-5), (fk, seen)
    seen["grad_g"] = _grad_dev(fix, "grad_g_tower%d" % tower, m.G.P.export_tf("grad"), gs)
    print("reference-graph deviations", case, tower, seen)
    assert seen["grad_d"][0] < GRAD_BAR and seen["grad_g"][0] < GRAD_BAR, seen


def test_train_batch_against_the_reference_training_loop():
    """tests/golden/ref_graph_schedule.npz: the reference's own `train_one_iteration` (scripts/train_gan_rnn_placeholder.py:
    48-133) run over three queued minibatches, the second one short (skipped).  The product's train_batch -- one call per
    full minibatch, the whole 1 D + 2 G schedule on the device with the generator forward shared between the D update and
    the first G update -- must land on the same mean losses, weight changes and generator output."""
    import torch
    fix = np.load(os.path.join(GOLD, "ref_graph_schedule.npz"))
    c, gp, dp, batches = C.schedule_setup()
    B = c["B"]
    m = make_model(c["g_type"], "lstm", B, l2_scale=c["l2_scale"], init_mse_weight=C.MSE_LAMBDA, init_disc_noise_std=0.0,
                   d_learning_rate=C.LR_D, g_learning_rate=C.LR_G)
    m.load_params(OrderedDict((k, v.astype(np.float32)) for k, v in gp.items()),
                  OrderedDict((k, v.astype(np.float32)) for k, v in dp.items()))
    keys = ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_l2_loss", "g_loss")
    acc = OrderedDict((k, []) for k in keys)
    for x, y, ln in batches:
        if x.shape[0] != B:                                   # the trainer CLI skips ragged minibatches like the reference
            continue
        d_all, g_all = m.train_batch(x.astype(np.float32), y.astype(np.float32), ln, all_updates=True)
        for d in d_all:
            for k in keys[:3]:
                acc[k].append(d[k])
        for g in g_all:
            for k in keys[3:]:
                acc[k].append(g[k])
    torch.cuda.synchronize()
    assert len(acc["d_loss"]) == 2 and len(acc["g_loss"]) == 4
    means = np.array([np.mean(v) for v in acc.values()])
    seen = {"loss_rel": float(np.abs(means / fix["means"] - 1.0).max())}
    assert seen["loss_rel"] < 5e-4, (means, fix["means"])            # measured 3.5e-5 (profiles/r2_reference_graph_gpu.txt)
    x0, _, l0 = batches[0]
    a, r = rms(m.generate(x0.astype(np.float32), l0).cpu().numpy(), fix["g_after"])
    seen["g_after_rel"] = r
    assert r < 1e-3, (a, r)                                          # measured 3.4e-4
    # weight CHANGES over the two schedules against the reference's: whole small tensors, sampled entries of the big ones
    worst = (0.0, None)
    for prefix, net, p0 in (("theta_g", m.G, gp), ("theta_d", m.D, dp)):
        th = net.P.export_tf()
        for k, v0 in p0.items():
            mine = th[k].astype(np.float64) - v0.astype(np.float32).astype(np.float64)
            if prefix + "|" + k + "|full" in fix.files:
                ref = fix[prefix + "|" + k + "|full"] - v0
                dev = rms(mine, ref)[1]
            else:
                idx = fix[prefix + "|" + k + "|idx"]
                ref = fix[prefix + "|" + k + "|at"] - v0.reshape(-1)[idx]
                dev = rms(mine.reshape(-1)[idx], ref)[1]
            if dev > worst[0]:
                worst = (dev, k)
    seen["delta"] = worst
    print("reference-loop deviations", seen)
    assert worst[0] < 2e-2, seen                                     # measured 7.2e-3 (relative RMS of theta_after - theta_before)
