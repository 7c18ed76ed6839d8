"""Pins the CPU oracle (oracle/rsr_oracle.py): against the independent torch-autograd statement,
finite differences, torch.nn.LSTM(proj_size) and the committed golden vectors.  (The oracle against the reference's own
model code: tests/test_reference_graph.py.  TF-1.4 itself is not runnable here -- see the oracle header.)"""
import os
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import make_golden, rsr_oracle as O, torch_ref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def small_gan(g_type="lstm", d_type="lstm", seed=0):
    rng = np.random.default_rng(seed)
    if g_type == "rced":
        gp = O.init_g_rced(rng)
    elif g_type == "dnn":
        gp = O.init_g_dnn(rng, units=32, hidden=2)
    else:
        gp = O.init_g_lstm(rng, cell=24, proj=12, layers=2) if g_type == "lstm" else O.init_g_res_lstm_l(rng, cell=16, layers=2)
    dp = O.init_d_lstm(rng, cell=16) if d_type == "lstm" else O.init_d_dnn(rng, units=32, hidden=2)
    for p in (gp, dp):
        for k in p:
            if "bias" in k:
                p[k] = rng.standard_normal(p[k].shape) * 0.1
    B, T = 3, 5
    x, y = rng.standard_normal((B, T, 257)), rng.standard_normal((B, T, 40))
    lengths = np.array([5, 3, 4])
    nz = (rng.standard_normal((B, 1, 40)) * 0.05, rng.standard_normal((B, 1, 40)) * 0.05) if d_type == "lstm" else (None, None)
    return gp, dp, x, y, lengths, nz


@pytest.mark.parametrize("g_type,d_type", [("lstm", "lstm"), ("res_lstm_l", "dnn"), ("res_lstm_base", "lstm"),
                                           ("rced", "dnn"), ("dnn", "dnn")])
@pytest.mark.parametrize("which", ["d", "g"])
def test_numpy_backward_matches_torch_autograd(g_type, d_type, which):
    gp, dp, x, y, lengths, (n1, n2) = small_gan(g_type, d_type)
    st = O.GanState(gp, dp, g_type, d_type)
    L, grads, g_out = O.tower_losses_and_grads(st, x, y, lengths, which, n1, n2)
    tg, td = R.to_torch(gp, requires_grad=True), R.to_torch(dp, requires_grad=True)
    t = lambda a: None if a is None else torch.tensor(a)
    Lt, gt, g_t = R.grads(tg, td, g_type, d_type, t(x), t(y), torch.tensor(lengths), which, noise_rl=t(n1), noise_fk=t(n2))
    assert np.allclose(g_out, g_t.detach().numpy(), atol=1e-12)
    for k in ("d_loss", "g_loss", "g_mse_loss", "g_adv_loss"):
        assert abs(L[k] - float(Lt[k])) < 1e-10
    for k in grads:
        assert np.allclose(grads[k], gt[k].numpy(), atol=1e-10, rtol=1e-8), k


def test_lstmp_matches_torch_nn_lstm_without_peepholes():
    """tf LSTMCell gate order i,j,f,o with forget_bias added at run time  ==  torch.nn.LSTM(proj_size)
    gate order i,f,g,o with the forget bias folded into b_ih, when the peepholes are zero."""
    rng = np.random.default_rng(1)
    B, T, I, C, P = 2, 6, 5, 7, 3
    x = rng.standard_normal((B, T, I))
    K = rng.standard_normal((I + P, 4 * C)) * 0.3
    b = rng.standard_normal(4 * C) * 0.1
    Wp = rng.standard_normal((C, P)) * 0.3
    z = np.zeros(C)
    out, _ = O.lstmp_fwd(x, np.full(B, T), K, b, z, z, z, Wp, forget_bias=1.0)
    lstm = torch.nn.LSTM(I, C, batch_first=True, proj_size=P).double()
    i_, j_, f_, o_ = (K[:, k * C:(k + 1) * C] for k in range(4))
    Kt = np.concatenate([i_, f_, j_, o_], 1)
    bi, bj, bf, bo = (b[k * C:(k + 1) * C] for k in range(4))
    bt = np.concatenate([bi, bf + 1.0, bj, bo])
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(torch.tensor(Kt[:I].T))
        lstm.weight_hh_l0.copy_(torch.tensor(Kt[I:].T))
        lstm.bias_ih_l0.copy_(torch.tensor(bt))
        lstm.bias_hh_l0.zero_()
        lstm.weight_hr_l0.copy_(torch.tensor(Wp.T))
        ref, _ = lstm(torch.tensor(x))
    assert np.allclose(out, ref.numpy(), atol=1e-12)


def test_lstmp_finite_difference_with_peepholes_and_ragged_lengths():
    rng = np.random.default_rng(2)
    B, T, I, C, P = 2, 4, 3, 5, 2
    x = rng.standard_normal((B, T, I))
    prm = [rng.standard_normal((I + P, 4 * C)) * 0.4, rng.standard_normal(4 * C) * 0.1, rng.standard_normal(C) * 0.5,
           rng.standard_normal(C) * 0.5, rng.standard_normal(C) * 0.5, rng.standard_normal((C, P)) * 0.4]
    lengths = np.array([4, 2])
    w = rng.standard_normal((B, T, P))
    f = lambda: float((O.lstmp_fwd(x, lengths, *prm)[0] * w).sum())
    out, cache = O.lstmp_fwd(x, lengths, *prm)
    assert np.all(out[1, 2:] == 0.0)                     # dynamic_rnn zeroes outputs past sequence_length
    dx, g = O.lstmp_bwd(w, cache)
    names = ["kernel", "bias", "w_i_diag", "w_f_diag", "w_o_diag", "proj"]
    eps = 1e-6
    for arr, name in zip(prm, names):
        idx = tuple(rng.integers(0, s) for s in arr.shape)
        old = arr[idx]
        arr[idx] = old + eps; fp = f()
        arr[idx] = old - eps; fm = f()
        arr[idx] = old
        assert abs((fp - fm) / (2 * eps) - g[name][idx]) < 1e-6, name
    idx = (0, 1, 2)
    old = x[idx]
    x[idx] = old + eps; fp = f()
    x[idx] = old - eps; fm = f()
    x[idx] = old
    assert abs((fp - fm) / (2 * eps) - dx[idx]) < 1e-6
    assert np.all(dx[1, 2:] == 0.0)


def test_losses_include_padded_frames_and_formulas():
    """models/gan_rnn_placeholder.py:244-260: means over every element, g_mse = 0.5*mse*output_dim."""
    rl, fk = np.array([[0.5], [2.0]]), np.array([[0.25], [-1.0]])
    g, y = np.ones((2, 40)), np.zeros((2, 40))
    L = O.lsgan_mse_losses(rl, fk, g, y, mse_lambda=10.0)
    assert L["d_rl_loss"] == pytest.approx((0.25 + 1.0) / 2)
    assert L["d_fk_loss"] == pytest.approx((0.0625 + 1.0) / 2)
    assert L["g_adv_loss"] == pytest.approx((0.5625 + 4.0) / 2)
    assert L["g_mse_loss"] == pytest.approx(0.5 * 1.0 * 40)
    assert L["g_loss"] == pytest.approx(L["g_adv_loss"] + 10.0 * 20.0)


def test_update_rules():
    g = np.full(100, 3.0)                                  # ||g|| = 30 -> clipped to 15
    c = O.clip_by_norm(g, 15.0)
    assert np.linalg.norm(c) == pytest.approx(15.0)
    assert np.array_equal(O.clip_by_norm(np.ones(4)), np.ones(4))
    p = OrderedDict(a=np.array([1.0, 2.0]))
    gr = OrderedDict(a=np.array([0.5, -0.25]))
    m, v = OrderedDict(a=np.zeros(2)), OrderedDict(a=np.zeros(2))
    pn, m, v, t = O.adam_update_tf(p, gr, m, v, 0, 1e-3)
    lr_t = 1e-3 * np.sqrt(1 - 0.999) / (1 - 0.9)
    assert np.allclose(pn["a"], p["a"] - lr_t * (0.1 * gr["a"]) / (np.sqrt(0.001 * gr["a"] ** 2) + 1e-8))
    # torch.optim.Adam puts eps inside the bias-corrected denominator: must differ from the TF form for tiny grads
    tiny = OrderedDict(a=np.array([1e-9, 1e-9]))
    pt, _, _, _ = O.adam_update_tf(p, tiny, OrderedDict(a=np.zeros(2)), OrderedDict(a=np.zeros(2)), 0, 1e-3)
    tp = torch.tensor(p["a"], requires_grad=True)
    opt = torch.optim.Adam([tp], lr=1e-3)
    tp.grad = torch.tensor(tiny["a"])
    opt.step()
    assert not np.allclose(pt["a"], tp.detach().numpy(), rtol=0, atol=1e-6)
    s = O.ema_update(OrderedDict(a=np.zeros(2)), OrderedDict(a=np.ones(2)))
    assert np.allclose(s["a"], 1e-4)


def test_exponential_decay_schedule():
    """utils/ops.py:378-391 with the x num_gpu factor of train...py:525-533."""
    assert O.exponential_decay(0, 2, 10, 1e-3) == pytest.approx(2e-3)
    assert O.exponential_decay(5, 2, 10, 1e-3) == pytest.approx(2e-3 * np.exp(5 * np.log(1e-4) / 10))
    assert O.exponential_decay(9, 2, 10, 1e-3) == pytest.approx(2e-7)
    assert O.exponential_decay(3, 4, 10, 0.05, multiply_jobs=False) == pytest.approx(0.05 * np.exp(3 * np.log(1e-4) / 10))


@pytest.mark.parametrize("name", sorted(make_golden.CASES))
def test_oracle_reproduces_golden_vectors(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    out = make_golden.compute(name)
    for k in ("g_out", "g_out_after"):
        assert np.allclose(out[k], gold[k], atol=1e-10), k
    for k in gold.files:
        if k.startswith("loss/"):
            assert float(out[k]) == pytest.approx(float(gold[k]), rel=1e-10)
        if k.startswith(("dgrad/", "ggrad/")):
            assert np.allclose(out[k], gold[k], atol=1e-7, rtol=1e-5), k


@pytest.mark.parametrize("name", sorted(make_golden.MSE_CASES))
def test_oracle_reproduces_mse_golden_vectors(name):
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    out = make_golden.compute_mse(name)
    for k in ("g_out", "g_out_after"):
        assert np.allclose(out[k], gold[k], atol=1e-10), k
    for k in gold.files:
        if k.startswith(("loss/", "loss_after/")):
            assert float(out[k]) == pytest.approx(float(gold[k]), rel=1e-10)
        if k.startswith("ggrad/"):
            assert np.allclose(out[k], gold[k], atol=1e-7, rtol=1e-5), k


def test_conv1d_same_is_tf_conv2d_same_and_its_gradient():
    """models/rced.py:94-101: conv2d([1, w], SAME, stride 1) == torch conv2d with padding w//2 (odd w), NHWC<->NCHW and
    HWIO<->OIHW transposed; hand-written backward == autograd; finite difference on one tap."""
    import torch.nn.functional as F
    rng = np.random.default_rng(4)
    N, L, ci, co, w = 3, 19, 5, 4, 7
    x, W, b = rng.standard_normal((N, L, ci)), rng.standard_normal((1, w, ci, co)) * 0.3, rng.standard_normal(co) * 0.1
    y, cache = O.conv1d_same_fwd(x, W, b)
    xt, Wt, bt = (torch.tensor(a, requires_grad=True) for a in (x, W, b))
    yt = torch.relu(F.conv2d(xt.permute(0, 2, 1).unsqueeze(2), Wt.permute(3, 2, 0, 1), bt, padding=(0, w // 2)))
    yt = yt.squeeze(2).permute(0, 2, 1)
    assert np.allclose(y, yt.detach().numpy(), atol=1e-12)
    dy = rng.standard_normal(y.shape)
    (yt * torch.tensor(dy)).sum().backward()
    dx, dW, db = O.conv1d_same_bwd(dy, cache)
    assert np.allclose(dx, xt.grad.numpy(), atol=1e-12) and np.allclose(dW, Wt.grad.numpy(), atol=1e-12)
    assert np.allclose(db, bt.grad.numpy(), atol=1e-12)
    f = lambda: float((O.conv1d_same_fwd(x, W, b)[0] * dy).sum())
    old, eps = W[0, 2, 1, 3], 1e-6
    W[0, 2, 1, 3] = old + eps; fp = f()
    W[0, 2, 1, 3] = old - eps; fm = f()
    W[0, 2, 1, 3] = old
    assert abs((fp - fm) / (2 * eps) - dW[0, 2, 1, 3]) < 1e-6


def test_mse_trainer_gradients_match_torch_autograd():
    """models/dnn_trainer_single_gpu.py:106-115: 0.5 * output_dim * mse + l2 on weights only."""
    rng = np.random.default_rng(5)
    for g_type, gp in (("dnn", O.init_g_dnn(rng, units=16, hidden=2)), ("rced", O.init_g_rced(rng))):
        x, y = rng.standard_normal((5, 257)), rng.standard_normal((5, 40))
        L, grads, g = O.mse_losses_and_grads(gp, g_type, x, y, l2_scale=1e-2)
        tp = R.to_torch(gp, requires_grad=True)
        gt = R.GEN[g_type](tp, torch.tensor(x), None)
        mse = 0.5 * 40 * ((gt - torch.tensor(y)) ** 2).mean()
        l2 = sum(1e-2 * 0.5 * (v ** 2).sum() for k, v in tp.items() if k.endswith("weights"))
        (mse + l2).backward()
        assert L["g_mse_loss"] == pytest.approx(float(mse), rel=1e-12) and L["g_l2_loss"] == pytest.approx(float(l2), rel=1e-12)
        for k in gp:
            assert np.allclose(grads[k], tp[k].grad.numpy(), atol=1e-10, rtol=1e-8), k


def test_conv2d_same_against_torch_and_toeplitz_equivalence():
    """The [splice, w] convolution of models/rced.py:90-101 for splice > 1: (i) the numpy statement against
    torch.nn.functional.conv2d (forward and autograd), (ii) the block-Toeplitz 1-D form the product's overlapped-view
    GEMM will use gives the same output, input gradient and -- folded over the tied copies -- the
    same filter gradient."""
    import torch.nn.functional as F
    rng = np.random.default_rng(4)
    N, H, L, ci, co, kh, kw = 3, 5, 17, 2, 3, 5, 7
    x, W, b = rng.standard_normal((N, H, L, ci)), rng.standard_normal((kh, kw, ci, co)) * 0.3, rng.standard_normal(co)
    w_out = rng.standard_normal((N, H, L, co))
    y, cache = O.conv2d_same_fwd(x, W, b)
    dx, dW, db = O.conv2d_same_bwd(w_out, cache)
    xt = torch.tensor(x, requires_grad=True)
    Wt, bt = torch.tensor(W, requires_grad=True), torch.tensor(b, requires_grad=True)
    yt = torch.relu(F.conv2d(xt.permute(0, 3, 1, 2), Wt.permute(3, 2, 0, 1), bt, padding=(kh // 2, kw // 2)))
    yt = yt.permute(0, 2, 3, 1)
    assert np.abs(y - yt.detach().numpy()).max() < 1e-12
    gx, gW, gb = torch.autograd.grad((yt * torch.tensor(w_out)).sum(), (xt, Wt, bt))
    assert np.abs(dx - gx.numpy()).max() < 1e-12 and np.abs(dW - gW.numpy()).max() < 1e-11
    assert np.abs(db - gb.numpy()).max() < 1e-11
    # splice = 1 is the 1-D statement the product runs today
    y1, _ = O.conv1d_same_fwd(x[:, 0], W[kh // 2:kh // 2 + 1], b)
    y1b, _ = O.conv2d_same_fwd(x[:, :1], W[kh // 2:kh // 2 + 1], b)
    assert np.array_equal(y1, y1b[:, 0])
    # Toeplitz form: lines as channels
    x2 = x.transpose(0, 2, 1, 3).reshape(N, L, H * ci)                   # (n, p, (h, ci))
    W2 = O.toeplitz_taps(W, H)
    y2, c2 = O.conv1d_same_fwd(x2, W2, np.tile(b, H))
    assert np.abs(y2.reshape(N, L, H, co).transpose(0, 2, 1, 3) - y).max() < 1e-12
    dx2, dW2, db2 = O.conv1d_same_bwd(w_out.transpose(0, 2, 1, 3).reshape(N, L, H * co), c2)
    assert np.abs(dx2.reshape(N, L, H, ci).transpose(0, 2, 1, 3) - dx).max() < 1e-12
    assert np.abs(O.toeplitz_fold_grad(dW2, kh, ci, co) - dW).max() < 1e-11
    assert np.abs(db2.reshape(H, co).sum(0) - db).max() < 1e-11
    dense = (W2 != 0).mean()
    assert 0.5 < dense < 1.0                                              # 19 of 25 line pairs at H = kh = 5


@pytest.mark.parametrize("B,L,ci,co,k", [(2, 64, 3, 5, 31), (3, 17, 4, 2, 5), (1, 32, 1, 16, 31)])
def test_strided_conv_family_against_torch(B, L, ci, co, k):
    """utils/ops.py `downconv` (conv2d, stride 2, SAME) and `deconv` (conv2d_transpose, stride 2): the oracle's statement
    against torch.nn.functional.conv1d / conv_transpose1d on the explicitly SAME-padded tensors, values and gradients."""
    import torch.nn.functional as F
    rng = np.random.default_rng(L)
    x, W, b = rng.standard_normal((B, L, ci)), rng.standard_normal((k, ci, co)) * 0.1, rng.standard_normal(co)
    y, c = O.downconv_fwd(x, W, b)
    _, pl, pr = O.same_pad(L, k, 2)
    xt, Wt, bt = (torch.tensor(v, requires_grad=True) for v in (x, W, b))
    yt = F.conv1d(F.pad(xt.permute(0, 2, 1), (pl, pr)), Wt.permute(2, 1, 0), bt, stride=2).permute(0, 2, 1)
    dy = rng.standard_normal(y.shape)
    yt.backward(torch.tensor(dy))
    dx, dW, db = O.downconv_bwd(dy, c)
    for a, r in ((y, yt.detach()), (dx, xt.grad), (dW, Wt.grad), (db, bt.grad)):
        assert np.abs(a - r.numpy()).max() < 1e-11
    x2, W2, b2 = rng.standard_normal((B, L, co)), rng.standard_normal((k, ci, co)) * 0.1, rng.standard_normal(ci)
    y2, c2 = O.deconv_fwd(x2, W2, b2)
    assert y2.shape == (B, 2 * L, ci)
    x2t, W2t, b2t = (torch.tensor(v, requires_grad=True) for v in (x2, W2, b2))
    _, pl2, _ = O.same_pad(2 * L, k, 2)
    full = F.conv_transpose1d(x2t.permute(0, 2, 1), W2t.permute(2, 1, 0), None, stride=2)
    full = F.pad(full, (0, max(0, pl2 + 2 * L - full.shape[2])))
    y2t = (full[:, :, pl2:pl2 + 2 * L] + b2t[None, :, None]).permute(0, 2, 1)
    dy2 = rng.standard_normal(y2.shape)
    y2t.backward(torch.tensor(dy2))
    dx2, dW2, db2 = O.deconv_bwd(dy2, c2)
    for a, r in ((y2, y2t.detach()), (dx2, x2t.grad), (dW2, W2t.grad), (db2, b2t.grad)):
        assert np.abs(a - r.numpy()).max() < 1e-11


def test_virtual_batch_norm_against_autograd():
    """utils/bnorm.py: reference pass and live pass (statistics blended with weight 1 / (batch + 1)) against an
    independent torch statement differentiated by autograd."""
    rng = np.random.default_rng(9)
    B, L, C = 5, 12, 7
    xr, xl = rng.standard_normal((B, L, C)), rng.standard_normal((B, L, C)) * 2 + 1
    gamma, beta = 1 + 0.1 * rng.standard_normal(C), rng.standard_normal(C)
    y_ref, _ = O.vbn_fwd(xr, gamma, beta)
    ref = O.vbn_reference(xr)
    y, cache = O.vbn_fwd(xl, gamma, beta, ref=ref)
    dy = rng.standard_normal(y.shape)
    dx, dg, db = O.vbn_bwd(dy, cache)

    def torch_vbn(x, g, b_, m_ref=None, q_ref=None):
        m, q = x.mean((0, 1)), (x ** 2).mean((0, 1))
        if m_ref is not None:
            w = 1.0 / (B + 1.0)
            m, q = w * m + (1 - w) * m_ref, w * q + (1 - w) * q_ref
        return (x - m) / torch.sqrt(1e-5 + q - m ** 2) * g + b_
    assert np.abs(y_ref - torch_vbn(torch.tensor(xr), torch.tensor(gamma), torch.tensor(beta)).numpy()).max() < 1e-12
    xt, gt, bt = (torch.tensor(v, requires_grad=True) for v in (xl, gamma, beta))
    yt = torch_vbn(xt, gt, bt, torch.tensor(ref[0]), torch.tensor(ref[1]))
    yt.backward(torch.tensor(dy))
    for a, r in ((y, yt.detach()), (dx, xt.grad), (dg, gt.grad), (db, bt.grad)):
        assert np.abs(a - r.numpy()).max() < 1e-11


def test_gauss_noise_moments():
    """The counter-based Gaussian stream of rsr_gauss_noise (discriminator input noise, utils/ops.py:19-30): unit
    moments, no correlation between draws that differ by tick or salt."""
    a = O.gauss_noise(1234, 0, 0x4e01, 200000, 1.0).astype(np.float64)
    b = O.gauss_noise(1234, 1, 0x4e01, 200000, 1.0).astype(np.float64)
    c = O.gauss_noise(1234, 0, 0x4e02, 200000, 1.0).astype(np.float64)
    for v in (a, b, c):
        assert abs(v.mean()) < 0.01 and abs(v.std() - 1.0) < 0.01
        assert abs((v ** 3).mean()) < 0.03 and abs((v ** 4).mean() - 3.0) < 0.1
    assert abs((a * b).mean()) < 0.01 and abs((a * c).mean()) < 0.01
    assert np.abs(O.gauss_noise(1234, 0, 0x4e01, 16, 0.5) * 2 - a[:16]).max() < 1e-6
