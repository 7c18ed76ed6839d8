"""batch_norm(renorm) / dropout: the oracle restatement against independent torch statements, and the host wiring
(nets.FCBN inside GAN_RNN / DNNTrainer) through the CPU test double against the oracle.  The CUDA kernels of
csrc/batchnorm.cu are checked by tests/test_batchnorm_gpu.py."""
import copy
import os
import sys
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_handle import FakeHandle  # noqa: E402

from oracle import rsr_oracle as O  # noqa: E402
from oracle import torch_ref as R  # noqa: E402
from rsrgan_b200.dnn_trainer import DNNTrainer  # noqa: E402
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def warm_state(st, rng):
    """renorm averages as after a few updates (r != 1, d != 0)"""
    for k in st:
        if k.endswith("weight"):
            st[k] = np.float64(0.3)
        elif k.endswith("renorm_stddev"):
            st[k] = 0.3 * (1 + 0.2 * rng.random(st[k].shape))
        elif k.endswith("renorm_mean"):
            st[k] = 0.05 * rng.standard_normal(st[k].shape)
    return st


def test_bn_first_batch_is_plain_batch_norm_and_state_update():
    rng = np.random.default_rng(0)
    z = rng.standard_normal((50, 6)) * 3 + 1
    gamma, beta = rng.standard_normal(6), rng.standard_normal(6)
    st = O.bn_init_state(6)
    y, _ = O.bn_renorm_train_fwd(z, gamma, beta, st, update=True)
    # zero-initialised renorm variables: r = 1, d = 0 on the first batch == torch batch_norm in training mode
    yt = torch.nn.functional.batch_norm(torch.tensor(z), None, None, torch.tensor(gamma), torch.tensor(beta), True, 0.0,
                                        O.BN_EPS).numpy()
    assert np.allclose(y, yt, atol=1e-12)
    # after one update the de-biased renorm averages ARE the batch moments; moving averages move by (1 - 0.999)
    mean, std = z.mean(0), np.sqrt(z.var(0) + O.BN_EPS)
    assert np.allclose(st["renorm_mean"] / st["renorm_mean_weight"], mean)
    assert np.allclose(st["renorm_stddev"] / st["renorm_stddev_weight"], std)
    assert float(st["renorm_mean_weight"]) == pytest.approx(0.01)
    assert np.allclose(st["moving_mean"], 0.001 * mean)
    assert np.allclose(st["moving_variance"], 1 + 0.001 * (z.var(0) - 1))
    # second batch now sees r, d != (1, 0): y = (xhat r + d) gamma + beta with the PRE-update averages
    z2 = rng.standard_normal((50, 6)) * 2 - 1
    pre = copy.deepcopy(st)
    y2, (xh, r, d, _, _) = O.bn_renorm_train_fwd(z2, gamma, beta, st, update=True)
    std2 = np.sqrt(z2.var(0) + O.BN_EPS)
    denom = pre["renorm_stddev"] + (1 - pre["renorm_stddev_weight"]) * std2
    assert np.allclose(r, std2 / denom) and not np.allclose(r, 1.0)
    assert np.allclose(d, (z2.mean(0) - (pre["renorm_mean"] + (1 - pre["renorm_mean_weight"]) * z2.mean(0))) / denom)
    # inference: plain affine with the moving averages
    ye = O.bn_eval_fwd(z, gamma, beta, st)
    yte = torch.nn.functional.batch_norm(torch.tensor(z), torch.tensor(st["moving_mean"]),
                                         torch.tensor(st["moving_variance"]), torch.tensor(gamma), torch.tensor(beta),
                                         False, 0.0, O.BN_EPS).numpy()
    assert np.allclose(ye, yte, atol=1e-12)


def test_bn_backward_finite_differences():
    rng = np.random.default_rng(1)
    z = rng.standard_normal((7, 5))
    gamma, beta = 1 + 0.3 * rng.standard_normal(5), rng.standard_normal(5)
    st = warm_state(O.bn_init_state(5), rng)
    w = rng.standard_normal((7, 5))

    def f(z_, g_, b_):
        # r and d are stop_gradient: freeze them at the unperturbed point
        y0, (xh, r, d, _, sd) = O.bn_renorm_train_fwd(z, gamma, beta, copy.deepcopy(st), update=False)
        m, s = z_.mean(0), np.sqrt(z_.var(0) + O.BN_EPS)
        return float((((((z_ - m) / s) * r + d) * g_ + b_) * w).sum())

    _, cache = O.bn_renorm_train_fwd(z, gamma, beta, copy.deepcopy(st), update=False)
    dz, dgamma, dbeta = O.bn_renorm_train_bwd(w, cache)
    eps = 1e-6
    for idx in [(0, 0), (3, 2), (6, 4)]:
        zp, zm = z.copy(), z.copy()
        zp[idx] += eps
        zm[idx] -= eps
        assert (f(zp, gamma, beta) - f(zm, gamma, beta)) / (2 * eps) == pytest.approx(dz[idx], rel=1e-5, abs=1e-8)
    for j in (0, 3):
        gp_, gm_ = gamma.copy(), gamma.copy()
        gp_[j] += eps
        gm_[j] -= eps
        assert (f(z, gp_, beta) - f(z, gm_, beta)) / (2 * eps) == pytest.approx(dgamma[j], rel=1e-5)
        bp_, bm_ = beta.copy(), beta.copy()
        bp_[j] += eps
        bm_[j] -= eps
        assert (f(z, gamma, bp_) - f(z, gamma, bm_)) / (2 * eps) == pytest.approx(dbeta[j], rel=1e-5)


def test_dropout_mask_statistics_and_streams():
    m = O.dropout_mask(7, 0, 3, 400, 256, 0.75)
    assert m.dtype == bool and abs(m.mean() - 0.75) < 0.01
    assert np.array_equal(m, O.dropout_mask(7, 0, 3, 400, 256, 0.75))
    for other in (O.dropout_mask(8, 0, 3, 400, 256, 0.75), O.dropout_mask(7, 1, 3, 400, 256, 0.75),
                  O.dropout_mask(7, 0, 4, 400, 256, 0.75)):
        assert abs((m == other).mean() - (0.75 ** 2 + 0.25 ** 2)) < 0.02     # independent streams
    # the test double's restatement of the same generator (what the wiring tests below rely on)
    fm = FakeHandle._drop_mask(torch.tensor([7, 0]), 3, 400, 256, 0.75).numpy()
    assert np.array_equal(m, fm)
    # known answers of the counter-based generator (pins the C ABI contract in include/rsrgan_b200.h)
    assert O._splitmix64(np.uint64(0)) == np.uint64(0)
    assert int(O._splitmix64(np.uint64(1))) == 0x5692161D100B05E5


@pytest.mark.parametrize("which", ["d", "g"])
def test_numpy_backward_matches_autograd_with_bn_and_dropout(which):
    rng = np.random.default_rng(3)
    gp = O.init_g_dnn(rng, in_dim=24, out_dim=8, units=32, hidden=2, batch_norm=True)
    dp = O.init_d_dnn(rng, in_dim=8, units=32, hidden=2, batch_norm=True)
    for p in (gp, dp):
        for k in p:
            if "BatchNorm" in k:
                p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    gs, ds = warm_state(O.init_bn_state(gp), rng), warm_state(O.init_bn_state(dp), rng)
    x, y, ln = rng.standard_normal((3, 5, 24)), rng.standard_normal((3, 5, 8)), np.array([5, 5, 5])
    st = O.GanState(gp, dp, "dnn", "dnn")
    mk = lambda: (dict(bn_state=copy.deepcopy(gs), keep_prob=0.8, rng=(7, 3)),
                  dict(bn_state=copy.deepcopy(ds), keep_prob=0.7, rng=(9, 3)))
    go, do = mk()
    L, G, g_out = O.tower_losses_and_grads(st, x, y, ln, which, g_opts=go, d_opts=do)
    go, do = mk()
    Lt, Gt, gt = R.grads(R.to_torch(gp, requires_grad=True), R.to_torch(dp, requires_grad=True), "dnn", "dnn",
                         torch.tensor(x), torch.tensor(y), ln, which, g_opts=go, d_opts=do)
    assert np.abs(g_out - gt.detach().numpy()).max() < 1e-12
    for k in Lt:
        assert abs(L[k] - float(Lt[k].detach())) < 1e-10
    for k in G:
        assert np.abs(G[k] - Gt[k].numpy()).max() < 1e-10, k


# ------------------------------------------------------------------ host wiring through the test double
def tf32(p):
    return OrderedDict((k, np.asarray(v, np.float32)) for k, v in p.items())


def test_dnn_trainer_batch_norm_steps_match_oracle():
    """run_dnn_single_gpu.sh trains the dnn generator WITH batch_norm (:129-145): two Adam steps with the
    UPDATE_OPS, then the inference graph on the moving averages."""
    rng = np.random.default_rng(5)
    N, I, U = 48, 40, 32
    args = Namespace(g_type="dnn", batch_size=N, input_dim=I, output_dim=8, g_units=U, g_layers=2, batch_norm=True,
                     keep_prob=0.8, l2_scale=1e-3, g_learning_rate=1e-3, seed=11)
    m = DNNTrainer(None, args, ["/gpu:0"], handle=FakeHandle("f16"))
    assert m.G.fcbn and m.G.keep_prob == 0.8
    gp = O.init_g_dnn(rng, in_dim=I, out_dim=8, units=U, hidden=2, batch_norm=True)
    for k in gp:
        if "BatchNorm" in k:
            gp[k] = gp[k] + 0.1 * rng.standard_normal(gp[k].shape)
    assert list(m.G.P.segs) == list(gp)                    # TF creation order: weights, BatchNorm/beta, BatchNorm/gamma
    m.load_params(tf32(gp))
    bst = O.init_bn_state(gp)
    st = O.MseState(OrderedDict((k, v.copy()) for k, v in gp.items()), "dnn")
    seed = int(m.G.rng[0])
    for step in range(2):
        x, y = rng.standard_normal((N, I)).astype(np.float32), rng.standard_normal((N, 8)).astype(np.float32)
        out = m.train_step(x, y)
        opts = dict(bn_state=bst, update=True, keep_prob=0.8, rng=(seed, step))
        losses, grads = O.mse_step(st, x.astype(np.float64), y.astype(np.float64), 1e-3, l2_scale=1e-3, g_opts=opts)
        assert out["g_mse_loss"] == pytest.approx(losses["g_mse_loss"], rel=3e-3)
        assert out["g_l2_loss"] == pytest.approx(losses["g_l2_loss"], rel=1e-3)
    assert int(m.G.rng[1]) == 2
    th = m.G.P.export_tf()
    for k in gp:
        assert rel(th[k], st.g[k]) < 2e-3, k
    mine = m.G.bn_state_tf()
    assert set(mine) == set(bst)
    for k in bst:          # 16-bit operands in the double: compare in RMS, not element-wise
        assert rel(mine[k], bst[k]) < 5e-3, k
    # inference graph (cross-validation model shares the weights): moving averages, no dropout
    cv = DNNTrainer(None, args, ["/gpu:0"], cross_validation=True, share=m)
    g = cv.generate(x).numpy()
    g_ref, _ = O.g_dnn_fwd(st.g, x.astype(np.float64), None, opts=dict(bn_state=bst, train=False))
    assert rel(g, g_ref) < 3e-3
    # checkpoint carries the non-trainable variables and the dropout stream
    sd = m.state_dict()
    m2 = DNNTrainer(None, args, ["/gpu:0"], handle=FakeHandle("f16"))
    m2.load_state_dict(sd)
    for k, v in m2.G.bn_state_tf().items():
        assert np.array_equal(v, mine[k]), k
    assert int(m2.G.rng[1]) == 2


def test_fcbn_statistics_epilogue_and_two_pass_forms_agree(monkeypatch):
    """nets.FCBN in training mode asks the GEMM for the batch_norm partials (rsr_gemm_args.stats) and finishes them
    (rsr_bn_train_finish); RSR_NO_EPILOGUE_STATS=1, a width that is not a multiple of 32, or the inference graph keep the
    two-pass / moving-average forms.  Same numbers either way."""
    def run(units, no_epi):
        if no_epi:
            monkeypatch.setenv("RSR_NO_EPILOGUE_STATS", "1")
        else:
            monkeypatch.delenv("RSR_NO_EPILOGUE_STATS", raising=False)
        h = FakeHandle("f16")
        calls = {"finish": 0, "two_pass": 0}
        fin, two = h.bn_train_finish, h.bn_train_stats
        h.bn_train_finish = lambda *a, **k: (calls.__setitem__("finish", calls["finish"] + 1), fin(*a, **k))[1]
        h.bn_train_stats = lambda *a, **k: (calls.__setitem__("two_pass", calls["two_pass"] + 1), two(*a, **k))[1]
        args = Namespace(g_type="dnn", batch_size=300, input_dim=40, output_dim=8, g_units=units, g_layers=2, batch_norm=True,
                         keep_prob=1.0, l2_scale=0.0, g_learning_rate=1e-3, seed=11)
        m = DNNTrainer(None, args, ["/gpu:0"], handle=h)
        rng = np.random.default_rng(3)
        outs = [m.train_step(rng.standard_normal((300, 40)).astype(np.float32),
                             rng.standard_normal((300, 8)).astype(np.float32)) for _ in range(2)]
        return calls, outs, m.G.bn_state_tf(), m.G.P.export_tf()
    ca, oa, sa, ta = run(64, False)
    cb, ob, sb, tb = run(64, True)
    n = 2 * sum(1 for l in DNNTrainer(None, Namespace(g_type="dnn", batch_size=300, input_dim=40, output_dim=8, g_units=64,
                                                     g_layers=2, batch_norm=True, seed=11), ["/gpu:0"],
                                      handle=FakeHandle("f16")).G.layers if getattr(l, "bn", False))
    assert n > 0 and ca == {"finish": n, "two_pass": 0} and cb == {"finish": 0, "two_pass": n}   # every layer, both steps
    for a, b in zip(oa, ob):
        assert a["g_mse_loss"] == pytest.approx(b["g_mse_loss"], rel=1e-5)
    for k in sa:
        assert rel(sa[k], sb[k]) < 1e-5, k
    for k in ta:          # two Adam steps from zero-initialised betas are all update: rounding of the moments shows there
        assert rel(ta[k], tb[k]) < 1e-3, k
    cc, _, _, _ = run(40, False)                                   # 40 units: not whole 32-column chunks -> two-pass form
    assert cc == {"finish": 0, "two_pass": n}


def test_gan_with_batch_norm_discriminator_matches_oracle():
    """dnn generator + discriminator_dnn, both batch-normalised, dropout in D: gradients of one D and one G update.
    UPDATE_OPS as models/gan_rnn_placeholder.py:163-175 wires them: the D update assigns the discriminator's statistics
    (both passes) and leaves the generator's alone, the G update the other way round."""
    rng = np.random.default_rng(8)
    B, T, I, U = 6, 16, 40, 64            # enough rows that a single relu' sign flip of a near-zero pre-activation stays below the bar
    args = Namespace(g_type="dnn", d_type="dnn", batch_size=B, input_dim=I, output_dim=8, g_units=U, g_layers=1,
                     d_units=U, d_layers=1, batch_norm=True, keep_prob=0.75, init_mse_weight=10.0, l2_scale=0.0,
                     g_learning_rate=0.0, d_learning_rate=0.0, seed=4)
    m = GAN_RNN(None, args, ["/gpu:0"], handle=FakeHandle("f16"))
    assert m.G.keep_prob == 1.0 and m.D.keep_prob == 0.75    # models/dnn.py:64-68: no l2 -> the generator keeps everything
    gp = O.init_g_dnn(rng, in_dim=I, out_dim=8, units=U, hidden=1, batch_norm=True)
    dp = O.init_d_dnn(rng, in_dim=8, units=U, hidden=1, batch_norm=True)
    for p in (gp, dp):
        for k in p:
            if "BatchNorm" in k:
                p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    m.load_params(tf32(gp), tf32(dp))
    gbs, dbs = warm_state(O.init_bn_state(gp), rng), warm_state(O.init_bn_state(dp), rng)
    m.G.load_bn_state_tf(gbs)
    m.D.load_bn_state_tf(dbs)
    x = rng.standard_normal((B, T, I)).astype(np.float32)
    y = rng.standard_normal((B, T, 8)).astype(np.float32)
    ln = np.full(B, T)
    st = O.GanState(gp, dp, "dnn", "dnn")
    seed = int(m.D.rng[0])
    gs = m._gscale(B * T)
    assert m.update_bn_stats and m.bn_update_scope == "own"
    g_run, d_run = copy.deepcopy(gbs), copy.deepcopy(dbs)      # the oracle's running statistics
    for tick, which in enumerate("dg"):
        go = dict(bn_state=g_run)
        do = dict(bn_state=d_run, keep_prob=0.75, rng=(seed, tick))
        # time-major rows inside the library: the oracle must draw its mask over the same (t, b) row order
        xt, yt = x.transpose(1, 0, 2).astype(np.float64), y.transpose(1, 0, 2).astype(np.float64)
        L, G, _ = O.tower_losses_and_grads(st, xt, yt, ln, which, g_opts=go, d_opts=do, update_ops="own")
        out = (m.d_step if which == "d" else m.g_step)(x, y, ln)
        if which == "d":        # the generator's statistics did not move during the D update
            for k, v in m.G.bn_state_tf().items():
                assert np.allclose(v, gbs[k], rtol=1e-6), k
        net, keys = (m.D, ("d_rl_loss", "d_fk_loss")) if which == "d" else (m.G, ("g_adv_loss", "g_mse_loss"))
        for k in keys:
            assert out[k] == pytest.approx(L[k], rel=3e-3, abs=1e-5), k
        mine = net.P.export_tf("grad")
        for k in G:
            # (relu' and the output clip are step functions: one near-zero pre-activation that rounds to the other side
            #  in 16-bit moves a per-column gradient of this small network by a few per cent)
            assert rel(mine[k] / gs, G[k]) < 3e-2, (which, k)
    moved = 0
    for net, ref, init in ((m.G, g_run, gbs), (m.D, d_run, dbs)):
        for k, v in net.bn_state_tf().items():
            assert np.allclose(v, ref[k], rtol=2e-3, atol=2e-5), k
            moved += int(not np.allclose(ref[k], init[k], rtol=1e-6))
    assert moved >= 8                                        # ... and they did move in the update that owns them


def test_lstm_generator_first_layer_batch_norm_and_noops():
    """models/lstm.py:61-67,82-87: only the first fully_connected of the lstm generator is normalised;
    res_lstm_l and discriminator_lstm build normalizer_params and never use them."""
    a = Namespace(g_type="lstm", d_type="lstm", batch_size=2, g_cell=40, g_proj=24, g_layers=1, d_cell=32,
                  batch_norm=True, seed=2)
    m = GAN_RNN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    names = list(m.G.P.segs)
    assert "g_model/fully_connected/BatchNorm/gamma" in names and "g_model/fully_connected/biases" not in names
    assert "g_model/fully_connected_1/biases" in names and not m.D.fcbn
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 5, 257)).astype(np.float32), rng.standard_normal((2, 5, 40)).astype(np.float32)
    out = m.train_batch(x, y, np.array([5, 4]))
    assert all(np.isfinite(v) for v in out.values())
    g0 = m.G.P.export_tf("grad")["g_model/fully_connected/BatchNorm/gamma"]
    assert np.abs(g0).max() > 0
    b = Namespace(**dict(vars(a), g_type="res_lstm_l"))
    m2 = GAN_RNN(None, b, ["/gpu:0"], handle=FakeHandle("f16"))
    assert not m2.G.fcbn and not m2.D.fcbn


@pytest.mark.parametrize("g_type", ["lstm", "res_lstm_l", "res_lstm_base"])
def test_lstm_generator_dropout_wrapper_matches_oracle(g_type):
    """DropoutWrapper(output_keep_prob) on every LSTM layer of the generator (models/lstm.py:99-102,
    models/res_lstm_l.py:96-99), batch_norm on the lstm generator's first layer: losses and raw gradients of a G update."""
    rng = np.random.default_rng(8)
    B, T = 3, 6
    a = Namespace(g_type=g_type, d_type="lstm", batch_size=B, g_cell=40, g_proj=24, g_layers=2, d_cell=32,
                  batch_norm=g_type == "lstm", keep_prob=0.8, init_mse_weight=10.0, init_disc_noise_std=0.0,
                  g_learning_rate=0.0, d_learning_rate=0.0, seed=5)
    m = GAN_RNN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    assert m.G.keep_prob == 0.8 and m.D.keep_prob == 1.0
    gp, dp = m.G.P.export_tf(dtype=np.float64), m.D.P.export_tf(dtype=np.float64)
    if g_type == "lstm":
        for k in gp:
            if "BatchNorm" in k:
                gp[k] = gp[k] + 0.1 * rng.standard_normal(gp[k].shape)
        m.load_params(tf32(gp), None)
        gp = m.G.P.export_tf(dtype=np.float64)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    y = rng.standard_normal((B, T, 40)).astype(np.float32)
    ln = np.array([T, T - 2, T - 1])
    st = O.GanState(gp, dp, g_type, "lstm")
    go = dict(bn_state=O.init_bn_state(gp), keep_prob=0.8, rng=(5, 0))
    L, G, _ = O.tower_losses_and_grads(st, x.astype(np.float64), y.astype(np.float64), ln, "g", g_opts=go)
    out = m.g_step(x, y, ln)
    gs = m._gscale(B * T)
    for k in ("g_adv_loss", "g_mse_loss"):
        assert out[k] == pytest.approx(L[k], rel=3e-3, abs=1e-5), k
    mine = m.G.P.export_tf("grad")
    for k in G:
        assert rel(mine[k] / gs, G[k]) < 3e-2, k
    assert int(m.G.rng[1]) == 1
    # the cross-validation graph keeps everything (lstm.keep_prob = 1.0 when not training, models/lstm.py:72-73)
    cv = GAN_RNN(None, a, ["/gpu:0"], cross_validation=True, share=m)
    g_cv = cv.generate(x, ln).numpy()
    g_ref, _ = GEN_FWD[g_type](gp, x.astype(np.float64), ln, opts=dict(bn_state=O.init_bn_state(gp), train=False))
    assert rel(g_cv, g_ref) < 3e-3


GEN_FWD = {k: v[0] for k, v in O.GENERATORS.items()}


def _run_golden_mse_dnn_bn(handle, tol_w, tol_state, tol_out):
    """tests/golden/mse_dnn_bn.npz (oracle/make_golden.py): batch-normalised dnn generator with dropout + l2 under
    DNNTrainer, three Adam steps with the UPDATE_OPS, then the inference graph."""
    import os
    from rsrgan_b200.dnn_trainer import DNNTrainer
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mse_dnn_bn.npz"))
    gp = OrderedDict((k[2:], z[k]) for k in z.files if k.startswith("G/"))
    N = z["x"].shape[1]
    args = Namespace(g_type="dnn", batch_size=N, g_units=64, batch_norm=True, keep_prob=float(z["keep_prob"]),
                     l2_scale=float(z["l2_scale"]), g_learning_rate=float(z["lr"]), seed=int(z["seed"]), dtype="f16")
    m = DNNTrainer(None, args, ["/gpu:0"], **({"handle": handle} if handle is not None else {}))
    m.load_params(gp)
    # raw gradients of the first step (learning rate 0 keeps the weights; the UPDATE_OPS are switched off so that
    # the statistics stay at their initial values for the steps below)
    m.g_learning_rate, m.update_bn_stats = 0.0, False
    out = m.train_step(z["x"][0], z["y"][0])
    assert out["g_mse_loss"] == pytest.approx(float(z["loss/g_mse_loss"]), rel=3e-3)
    assert out["g_l2_loss"] == pytest.approx(float(z["loss/g_l2_loss"]), rel=1e-3)
    gs = m._gscale(N)
    mine = m.G.P.export_tf("grad")
    for k in gp:
        assert rel(mine[k] / gs, z["ggrad/" + k]) < 5e-2, k
    # the same model again from tick 0: Adam state and dropout stream reset by a fresh trainer
    m = DNNTrainer(None, args, ["/gpu:0"], **({"handle": handle} if handle is not None else {}))
    m.load_params(gp)
    for t in range(int(z["steps"])):
        out = m.train_step(z["x"][t], z["y"][t])
        assert out["g_mse_loss"] == pytest.approx(float(z["loss_step%d/g_mse_loss" % t]), rel=tol_w), t
    st = m.G.bn_state_tf()
    for k in st:
        assert rel(np.asarray(st[k]) + 1.0, z["BN/" + k] + 1.0) < tol_state, k
    cv = DNNTrainer(None, args, ["/gpu:0"], cross_validation=True, share=m)
    g = cv.generate(z["x"][0])
    g = g.cpu().numpy() if hasattr(g, "cpu") else np.asarray(g)
    assert rel(g, z["g_out_after"]) < tol_out


def test_golden_mse_dnn_bn_host_wiring():
    _run_golden_mse_dnn_bn(FakeHandle("f16"), 5e-3, 1e-3, 5e-2)


def _double(operands):
    fh = FakeHandle("f16")
    if operands == "f32":
        fh.h16 = torch.float32        # the double with fp32 GEMM operands: isolates the wiring from 16-bit rounding
    return fh


@pytest.mark.parametrize("operands,gbar", [("f32", 1e-3), ("f16", 2e-1)])
def test_rced_batch_norm_host_wiring_matches_oracle(operands, gbar):
    """models/rced.py:63-71,94-97: normalizer_fn=batch_norm on the nine convolutions (no biases; moments pooled over
    frames and positions of a channel).  nets.ConvBN through the CPU test double: loss, every raw gradient, the
    UPDATE_OPS and the inference graph against the oracle.  The gradient of a normalised layer is orthogonal to (1, x_hat),
    so the weight gradients behind it are sums with heavy cancellation: with 16-bit operands nine such layers measure
    2e-2 (Conv_8) .. 1.6e-1 (Conv_1) relative, with fp32 operands the same wiring is exact to < 1e-4."""
    rng = np.random.default_rng(9)
    N, bins = 12, 24
    args = Namespace(g_type="rced", batch_size=N, input_dim=bins, output_dim=8, batch_norm=True, g_learning_rate=0.0, seed=3)
    m = DNNTrainer(None, args, ["/gpu:0"], handle=_double(operands))
    assert m.G.has_bn_state and not m.G.fcbn
    gp = O.init_g_rced(rng, in_dim=bins, out_dim=8, batch_norm=True)
    for k in gp:
        if "BatchNorm" in k:
            gp[k] = gp[k] + 0.1 * rng.standard_normal(gp[k].shape)
    assert list(m.G.P.segs) == list(gp) and "g_model/Conv/biases" not in gp
    m.load_params(tf32(gp))
    bst = warm_state(O.init_bn_state(gp), rng)
    m.G.load_bn_state_tf(bst)
    x, y = rng.standard_normal((N, bins)).astype(np.float32), rng.standard_normal((N, 8)).astype(np.float32)
    opts = dict(bn_state=bst, update=True)
    L, G, _ = O.mse_losses_and_grads(gp, "rced", x.astype(np.float64), y.astype(np.float64), g_opts=opts)
    out = m.train_step(x, y)
    assert out["g_mse_loss"] == pytest.approx(L["g_mse_loss"], rel=3e-3)
    gs = m._gscale(N)
    gg = m.G.P.export_tf("grad")
    for k in G:
        assert rel(gg[k] / gs, G[k]) < gbar, k
    mine = m.G.bn_state_tf()
    assert set(mine) == set(bst)
    for k in bst:
        assert rel(np.asarray(mine[k]) + 1.0, np.asarray(bst[k]) + 1.0) < 2e-3, k
    cv = DNNTrainer(None, args, ["/gpu:0"], cross_validation=True, share=m)
    g_ref, _ = O.g_rced_fwd(gp, x.astype(np.float64), None, opts=dict(bn_state=bst, train=False))
    assert rel(cv.generate(x).numpy(), g_ref) < 5e-3
    sd = m.state_dict()
    m2 = DNNTrainer(None, args, ["/gpu:0"], handle=_double(operands))
    m2.load_state_dict(sd)
    for k, v in m2.G.bn_state_tf().items():
        assert np.array_equal(v, mine[k]), k
