"""The oracle against the reference's OWN graph-building code.

tests/golden/ref_graph_*.npz were produced in the build container by tests/golden/make_reference_graph_golden.py, which
imports /root/reference/models/gan_rnn_placeholder.py (and with it lstm.py, res_lstm_l.py, res_lstm_base.py,
discriminator_lstm.py, utils/ops.py) plus models/BNLSTMCell.py, and executes them over an eager float64 stand-in for the
TensorFlow-1.4 calls they make (tests/golden/tf_standin.py: TensorFlow itself cannot be installed here).  So the variable
names and shapes, the layer wiring, the (B, 1, 40) discriminator noise, the loss formulas, tower slicing,
average_gradients, clip_by_norm 15, which optimizer updates which network and the EMA are the REFERENCE's statements;
only TensorFlow's library layers and op kernels are restated (and the LSTM step of that restatement is checked against
models/BNLSTMCell.py:176-213).  Here the oracle replays the same seeded parameters and feeds and must agree to 1e-9."""
import copy
import os
import sys
from collections import OrderedDict

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
sys.path.insert(0, GOLD)
import ref_graph_common as C  # noqa: E402
from oracle import rsr_oracle as O  # noqa: E402

LOSS_KEYS = [("d_rl_losses", "d_rl_loss"), ("d_fk_losses", "d_fk_loss"), ("d_losses", "d_loss"), ("g_adv_losses", "g_adv_loss"),
             ("g_mse_losses", "g_mse_loss"), ("g_l2_losses", "g_l2_loss"), ("g_losses", "g_loss")]


def close(a, b, rtol=1e-9):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return a.shape == b.shape and np.allclose(a, b, rtol=rtol, atol=rtol * (float(np.abs(b).max()) if b.size else 0.0) + 1e-13)


@pytest.mark.parametrize("case", list(C.GAN_RNN_CASES))
def test_gan_rnn_graph_of_the_reference(case):
    fix = np.load(os.path.join(GOLD, "ref_graph_gan_rnn_%s.npz" % case))
    c, gp, dp, x, y, lengths, noise = C.gan_rnn_setup(case)
    B = c["B"]
    # 1. the variables models/gan_rnn_placeholder.py creates are exactly the ones the oracle names, with these shapes
    mine = {"%s %s" % (k, list(v.shape)) for p in (gp, dp) for k, v in p.items()}
    assert set(fix["variables"].tolist()) == mine
    kw = dict(mse_lambda=C.MSE_LAMBDA, l2_scale=c["l2_scale"])
    towers = []
    for i in range(c["towers"]):
        sl = slice(B * i, B * (i + 1))                                   # gan_rnn_placeholder.py:157-159
        towers.append(dict(x=x[sl], y=y[sl], lengths=lengths[sl], noise_rl=C.NOISE_STD * noise[1 + 2 * i],
                           noise_fk=C.NOISE_STD * noise[2 + 2 * i]))
    st = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), c["g_type"], "lstm")
    for i, t in enumerate(towers):
        # 2. forward values and the seven losses of every tower (:191-260)
        Ld, Gd, g_out = O.tower_losses_and_grads(st, t["x"], t["y"], t["lengths"], "d", t["noise_rl"], t["noise_fk"], **kw)
        Lg, Gg, _ = O.tower_losses_and_grads(st, t["x"], t["y"], t["lengths"], "g", t["noise_rl"], t["noise_fk"], **kw)
        assert close(g_out, fix["fwd|tower%d/g_clean|full" % i])
        assert close(O.d_lstm_fwd(st.d, t["y"], t["lengths"], t["noise_rl"])[0], fix["fwd|tower%d/d_real|full" % i])
        assert close(O.d_lstm_fwd(st.d, g_out, t["lengths"], t["noise_fk"])[0], fix["fwd|tower%d/d_fake|full" % i])
        for fk, ok in LOSS_KEYS:
            assert Lg.get(ok, 0.0) == pytest.approx(float(fix["loss|" + fk][i]), rel=1e-10, abs=1e-14), (fk, i)
        assert Ld["d_loss"] == pytest.approx(float(fix["loss|d_losses"][i]), rel=1e-10)
        # 3. raw gradients of d_loss wrt the d_ variables and of g_loss wrt the g_ variables (:169-175)
        assert C.check(fix, "grad_d_tower%d" % i, Gd) == len(dp)
        assert C.check(fix, "grad_g_tower%d" % i, Gg) == len(gp)
    # 4. average over towers, clip each tensor to norm 15, SGD on D / Adam on G, EMA 0.9999 (:177-189)
    sd = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), c["g_type"], "lstm")
    _, clipped_d = O.d_step(sd, towers, C.LR_D, **kw)
    C.check(fix, "applied_d", clipped_d)
    C.check(fix, "theta_d_after_d_opt", sd.d, rtol=1e-10)
    C.check(fix, "ema_d_after_d_opt", sd.d_ema, rtol=1e-10)
    sg = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), c["g_type"], "lstm")
    _, clipped_g = O.g_step(sg, towers, C.LR_G, **kw)
    C.check(fix, "applied_g", clipped_g)
    C.check(fix, "theta_g_after_g_opt", sg.g, rtol=1e-10)
    C.check(fix, "ema_g_after_g_opt", sg.g_ema, rtol=1e-10)
    # the clip bites in these cases (the generator's loss is large), so its order -- after the tower mean -- is exercised
    raw = O.average_gradients([O.tower_losses_and_grads(st, t["x"], t["y"], t["lengths"], "g", t["noise_rl"], t["noise_fk"], **kw)[1]
                               for t in towers])
    assert max(float(np.sqrt((v ** 2).sum())) for v in raw.values()) > 15.0


def test_lstm_step_of_the_reference():
    """models/BNLSTMCell.py:176-213 (the reference's statement of the peephole LSTMP step; batch norms replaced by the
    identity) == oracle lstmp_fwd, state and projected output, over four steps."""
    f = np.load(os.path.join(GOLD, "ref_graph_lstm_cell.npz"))
    out, cache = O.lstmp_fwd(f["x"], np.full(f["x"].shape[0], f["x"].shape[1]), f["K"], f["b"], f["w_i"], f["w_f"], f["w_o"], f["Wp"])
    assert close(out, f["m"], 1e-12)


def test_exponential_decay_of_the_reference():
    """utils/ops.py:378-391 called directly == the oracle's and the trainer CLI's restatements."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "rsr_cli", os.path.join(os.path.dirname(GOLD), os.pardir, "scripts", "train_gan_rnn_placeholder.py"))
    rows = np.load(os.path.join(GOLD, "ref_graph_schedules.npz"))["exponential_decay"]
    cli = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cli)
    for it, jobs, iters, init, mult, want in rows:
        for fn in (O.exponential_decay, cli.exponential_decay):
            assert fn(int(it), int(jobs), int(iters), float(init), bool(mult)) == pytest.approx(want, rel=1e-13)


def test_frame_gan_graph_of_the_reference():
    """models/gan.py (+ dnn.py, discriminator_dnn.py) executed over the stand-in: conditioned discriminator input
    concat(centre LPS frame, MFCC) (:159-174), clip_by_value(-0.5, 1.5) on the logits, g_l2 from the REGULARIZATION_LOSSES of
    g_model only (:209-214), Adam for both networks, no gradient clipping (:139-153)."""
    fix = np.load(os.path.join(GOLD, "ref_graph_frame_gan_dnn.npz"))
    c, gp, dp, x, y = C.frame_setup("gan_dnn")
    assert set(fix["variables"].tolist()) == {"%s %s" % (k, list(v.shape)) for p in (gp, dp) for k, v in p.items()}
    N = c["N"]
    x3, y3, ones = x[:, None], y[:, None], np.ones(N, int)
    kw = dict(mse_lambda=C.MSE_LAMBDA, l2_scale=c["l2_scale"], d_cat=(257 * C.LEFT, 257 * (C.LEFT + 1)), l2_weights_only=True)
    st = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), "dnn", "dnn")
    Ld, Gd, g_out = O.tower_losses_and_grads(st, x3, y3, ones, "d", **kw)
    Lg, Gg, _ = O.tower_losses_and_grads(st, x3, y3, ones, "g", **kw)
    assert close(g_out[:, 0], fix["fwd|g_clean|full"])
    centre = x3[..., 257 * C.LEFT:257 * (C.LEFT + 1)]
    assert close(O.d_dnn_fwd(st.d, np.concatenate([centre, y3], -1))[0][:, 0], fix["fwd|d_real|full"])
    assert close(O.d_dnn_fwd(st.d, np.concatenate([centre, g_out], -1))[0][:, 0], fix["fwd|d_fake|full"])
    for fk, ok in LOSS_KEYS:
        assert Lg.get(ok, 0.0) == pytest.approx(float(fix["loss|" + fk][0]), rel=1e-10, abs=1e-14), fk
    assert C.check(fix, "grad_d", Gd) == len(dp)
    assert C.check(fix, "grad_g", Gg) == len(gp)
    tower = dict(x=x3, y=y3, lengths=ones)
    sd = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), "dnn", "dnn")
    _, applied = O.d_step(sd, [tower], c["lr_d"], max_norm=1e30, adam=True, **kw)
    C.check(fix, "applied_d", applied)
    C.check(fix, "theta_d_after_d_opt", sd.d, rtol=1e-10)
    C.check(fix, "ema_d_after_d_opt", sd.d_ema, rtol=1e-10)
    sg = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), "dnn", "dnn")
    _, applied = O.g_step(sg, [tower], c["lr_g"], max_norm=1e30, **kw)
    C.check(fix, "applied_g", applied)
    C.check(fix, "theta_g_after_g_opt", sg.g, rtol=1e-10)
    C.check(fix, "ema_g_after_g_opt", sg.g_ema, rtol=1e-10)


def test_dnn_trainer_graph_of_the_reference():
    """models/dnn_trainer_single_gpu.py:93-133 (+ dnn.py) executed over the stand-in: 0.5 * 40 * mse + the weights-only l2
    of the contrib regularizer, Adam.minimize."""
    fix = np.load(os.path.join(GOLD, "ref_graph_dnn_trainer.npz"))
    c, gp, _, x, y = C.frame_setup("dnn_trainer")
    assert set(fix["variables"].tolist()) == {"%s %s" % (k, list(v.shape)) for k, v in gp.items()}
    st = O.MseState(copy.deepcopy(gp), "dnn")
    losses, grads = O.mse_step(st, x, y, c["lr_g"], l2_scale=c["l2_scale"])
    for fk, ok in (("g_mse_losses", "g_mse_loss"), ("g_l2_losses", "g_l2_loss"), ("g_losses", "g_loss")):
        assert losses[ok] == pytest.approx(float(fix["loss|" + fk][0]), rel=1e-10), fk
    assert C.check(fix, "grad_g", grads) == len(gp)
    C.check(fix, "theta_g_after_g_opt", st.g, rtol=1e-10)


@pytest.mark.parametrize("case", list(C.RCED_CASES))
def test_rced_graph_of_the_reference(case):
    """models/rced.py under models/dnn_trainer.py executed over the stand-in: (N, splice * 257) frames reshaped to
    (N, splice, 257, 1) (:46-57), nine SAME [splice, w] convolutions with ReLU (:90-101), NHWC flatten into the linear output
    layer with bias 0.1 (:106-113); 0.5 * 40 * mse + weights-only l2, Adam, EMA over all trainable variables."""
    fix = np.load(os.path.join(GOLD, "ref_graph_%s.npz" % case))
    c, gp, x, y = C.rced_setup(case)
    assert set(fix["variables"].tolist()) == {"%s %s" % (k, list(v.shape)) for k, v in gp.items()}
    st = O.MseState(copy.deepcopy(gp), "rced")
    losses, grads = O.mse_step(st, x, y, c["lr_g"], l2_scale=c["l2_scale"])
    for fk, ok in (("g_mse_losses", "g_mse_loss"), ("g_l2_losses", "g_l2_loss"), ("g_losses", "g_loss")):
        assert losses[ok] == pytest.approx(float(fix["loss|" + fk][0]), rel=1e-10, abs=1e-14), fk
    assert C.check(fix, "grad_g", grads) == len(gp)
    C.check(fix, "theta_g_after_g_opt", st.g, rtol=1e-10)
    C.check(fix, "ema_g_after_g_opt", O.ema_update(copy.deepcopy(gp), st.g, 0.9999), rtol=1e-10)


def test_virtual_batch_norm_of_the_reference():
    """utils/bnorm.py:11-69 executed over the stand-in (reference pass, then a live pass blended with weight 1 / (B + 1)):
    outputs, and the gradients autograd takes through the reference's expressions, against vbn_fwd / vbn_bwd."""
    f = np.load(os.path.join(GOLD, "ref_graph_vbn.npz"))
    out, cache = O.vbn_fwd(f["x_ref"], f["gamma"], f["beta"])
    assert close(out, f["out_ref"], 1e-12)
    dx, dg, db = O.vbn_bwd(f["r_ref"], cache)
    assert close(dx, f["dx_ref"], 1e-11) and close(dg, f["dgamma_ref"], 1e-11) and close(db, f["dbeta_ref"], 1e-11)
    out, cache = O.vbn_fwd(f["x"], f["gamma"], f["beta"], ref=O.vbn_reference(f["x_ref"]))
    assert close(out, f["out_live"], 1e-12)
    dx, dg, db = O.vbn_bwd(f["r_live"], cache)
    assert close(dx, f["dx_live"], 1e-11) and close(dg, f["dgamma_live"], 1e-11) and close(db, f["dbeta_live"], 1e-11)


def test_conv_family_of_the_reference():
    """utils/ops.py downconv (:78-98), deconv (:277-310), conv1d (:138-156) and leakyrelu (:120-121) executed over the stand-in
    (tf.nn.conv2d with TensorFlow's SAME rule; conv2d_transpose literally as the gradient of conv2d), even and odd lengths:
    outputs and autograd gradients against the oracle's forward / backward statements."""
    f = np.load(os.path.join(GOLD, "ref_graph_conv_family.npz"))
    for par in ("even", "odd"):
        g = lambda k: f["down_%s|%s" % (par, k)]
        y, cache = O.downconv_fwd(g("x"), g("W")[:, 0], g("b"), pool=2)
        assert close(y, g("y"), 1e-12)
        dx, dW, db = O.downconv_bwd(g("r"), cache)
        assert close(dx, g("dx"), 1e-11) and close(dW, g("dW")[:, 0], 1e-11) and close(db, g("db"), 1e-11)
        g = lambda k: f["de_%s|%s" % (par, k)]
        y, cache = O.deconv_fwd(g("x"), g("W")[:, 0], g("b"), dilation=2)
        assert close(y, g("y"), 1e-12)
        dx, dW, db = O.deconv_bwd(g("r"), cache)
        assert close(dx, g("dx"), 1e-11) and close(dW, g("dW")[:, 0], 1e-11) and close(db, g("db"), 1e-11)
    g = lambda k: f["conv1d|" + k]
    y, cache = O.conv1d_same_fwd(g("x"), g("W")[None], g("b"), O.ACT_NONE)
    assert close(y, g("y"), 1e-12)
    dx, dW, db = O.conv1d_same_bwd(g("r"), cache)
    assert close(dx, g("dx"), 1e-11) and close(dW[0], g("dW"), 1e-11) and close(db, g("db"), 1e-11)
    assert close(O.act_fwd(f["leakyrelu|x"], O.ACT_LRELU), f["leakyrelu|y"], 1e-15)


def test_splice_feats_of_the_reference():
    """io_funcs/tfrecords_io.py:177-204 executed over the stand-in == the loader's splice_feats (edge replication)."""
    from rsrgan_b200.dataset import splice_feats
    f = np.load(os.path.join(GOLD, "ref_graph_splice.npz"))
    n = len([k for k in f.files if k.endswith("|feats")])
    assert n >= 5
    for i in range(n):
        left, right = (int(v) for v in f["case%d|ctx" % i])
        assert np.array_equal(splice_feats(f["case%d|feats" % i], left, right), f["case%d|spliced" % i]), i


def test_training_loop_of_the_reference():
    """scripts/train_gan_rnn_placeholder.py:48-133 `train_one_iteration`, the reference's own loop, was run over a queue of
    three minibatches (the second one an utterance short: skipped, :69-70), every sess.run re-executing the reference's graph
    code on the current variables: per minibatch one d_opt, then two g_opt -- each g_opt seeing the discriminator the d_opt
    before it produced, the second g_opt the generator the first one produced.  The oracle replays d_step, g_step, g_step per
    full minibatch: returned mean losses (over update counts, :124-130), weights, EMA shadows, Adam step count, and the
    generator output afterwards."""
    fix = np.load(os.path.join(GOLD, "ref_graph_schedule.npz"))
    c, gp, dp, batches = C.schedule_setup()
    st = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), c["g_type"], "lstm")
    kw = dict(mse_lambda=C.MSE_LAMBDA, l2_scale=c["l2_scale"])
    acc = OrderedDict((k, []) for k in ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_l2_loss", "g_loss"))
    for x, y, ln in batches:
        if x.shape[0] != c["B"]:
            continue
        tower = dict(x=x, y=y, lengths=ln)
        Ld, _ = O.d_step(st, [tower], C.LR_D, **kw)
        for k in ("d_rl_loss", "d_fk_loss", "d_loss"):
            acc[k].append(Ld[0][k])
        for _ in range(2):
            Lg, _ = O.g_step(st, [tower], C.LR_G, **kw)
            for k in ("g_adv_loss", "g_mse_loss", "g_l2_loss", "g_loss"):
                acc[k].append(Lg[0][k])
    assert np.allclose([np.mean(v) for v in acc.values()], fix["means"], rtol=1e-10)
    assert st.adam_t == int(fix["adam_t"]) == 4
    C.check(fix, "theta_g", st.g, rtol=1e-10)
    C.check(fix, "theta_d", st.d, rtol=1e-10)
    C.check(fix, "ema_g", st.g_ema, rtol=1e-10)
    C.check(fix, "ema_d", st.d_ema, rtol=1e-10)
    x0, _, l0 = batches[0]
    assert close(O.g_lstm_fwd(st.g, x0, l0)[0], fix["g_after"], 1e-9)


def test_update_ops_wiring_of_the_reference():
    """Which batch_norm UPDATE_OPS the reference's graphs hold and which train op runs them (wiring only: the stand-in's
    batch_norm does not restate contrib's arithmetic).  gan_rnn_placeholder.py:163-175: g_opt runs the g_model updates, d_opt
    the d_model ones (none: discriminator_lstm has no normalised layer); gan.py:139-143 and dnn_trainer*.py: every train op
    runs the whole collection.  One update op exists per graph COPY of a layer: tower 0 builds the generator twice (reuse=False,
    then reuse=True, :196-205) and the discriminator three times (dummy, labels, G(x)), so the reference assigns the same
    moving averages two / three times per run, unlocked and in no defined order -- between one and two (three) effective
    updates.  The oracle and the product apply one per pass the loss uses (G once; D on labels, then on G(x)): inside that
    envelope, recorded in DESIGN.md section 2.  dnn_trainer_single_gpu.py builds one copy: no ambiguity there."""
    import json
    with open(os.path.join(GOLD, "ref_graph_update_ops.json")) as f:
        w = json.load(f)
    net = lambda names: sorted({n.split("/")[0] for n in names})
    copies = lambda names, layer: sum(1 for n in names if n.split("#")[0] == layer + "/BatchNorm/AssignMovingAvg")
    r = w["gan_rnn_placeholder lstm, 1 tower"]
    assert r["d_opt_runs"] == [] and net(r["g_opt_runs"]) == ["g_model"] and copies(r["update_ops"], "g_model/fully_connected") == 2
    r = w["gan (frame level) dnn + discriminator_dnn, 1 tower"]
    assert r["d_opt_runs"] == r["g_opt_runs"] == r["update_ops"] and net(r["update_ops"]) == ["d_model", "g_model"]
    assert copies(r["update_ops"], "g_model/fully_connected_1") == 2 and copies(r["update_ops"], "d_model/fully_connected_1") == 3
    r = w["dnn_trainer_single_gpu dnn"]
    assert r["g_opt_runs"] == r["update_ops"] and copies(r["update_ops"], "g_model/fully_connected") == 1
    r = w["dnn_trainer (multi-tower trainer) dnn, 1 tower"]
    assert r["g_opt_runs"] == r["update_ops"] and copies(r["update_ops"], "g_model/fully_connected") == 2
    # the product's association of networks to updates is the reference's (rsrgan_b200/gan_rnn.py::_mode, gan.py, dnn_trainer.py)
    from rsrgan_b200 import gan_rnn, gan as frame_gan
    src = open(gan_rnn.__file__).read() + open(frame_gan.__file__).read()
    assert "g_update" in src and "d_update" in src


def test_cross_validation_and_decode_graphs_of_the_reference():
    """GAN_RNN(cross_validation=True) and GAN_RNN(infer=True) of models/gan_rnn_placeholder.py executed over the stand-in
    (res_lstm_l): the cross-validation losses carry no l2 term although l2_scale > 0 (:253-258), keep_prob is forced to 1
    (:72-75), the discriminator noise is still applied (discriminator_lstm.py:60); the decode graph is the generator alone."""
    f = np.load(os.path.join(GOLD, "ref_graph_cv_infer.npz"))
    c, gp, dp, x, y, lengths, noise = C.gan_rnn_setup("res_lstm_l_1tower")
    st = O.GanState(copy.deepcopy(gp), copy.deepcopy(dp), c["g_type"], "lstm")
    L, _, g_out = O.tower_losses_and_grads(st, x, y, lengths, "g", C.NOISE_STD * noise[1], C.NOISE_STD * noise[2],
                                           mse_lambda=C.MSE_LAMBDA, l2_scale=0.0)
    for fk, ok in LOSS_KEYS:
        assert L.get(ok, 0.0) == pytest.approx(float(f["cv|" + fk][0]), rel=1e-10, abs=1e-14), fk
    assert float(f["cv|g_l2_losses"][0]) == 0.0
    assert close(g_out, f["infer|g_outputs"])
