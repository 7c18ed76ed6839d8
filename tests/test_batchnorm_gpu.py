"""csrc/batchnorm.cu on the GPU, through the C ABI: batch_norm(renorm) statistics / normalisation / backward and the
counter-based dropout against the float64 oracle, then nets.FCBN inside DNNTrainer and GAN_RNN.

Tolerances: the kernels are fp32 streams over an fp32 pre-activation -> statistics and coefficients 1e-5 relative RMS
against float64; 16-bit outputs within half an ulp of the oracle value (relative RMS 1e-3 f16 / 8e-3 bf16); the
dropout mask is integer arithmetic and must be BIT-EXACT.  Model level: as tests/test_frame_models_gpu.py."""
import copy
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import rsr_oracle as O

pytestmark = pytest.mark.gpu
F32 = torch.float32


@pytest.fixture(scope="module", params=["f16", "bf16"])
def h(request):
    from rsrgan_b200 import ops
    hd = ops.Handle(0, request.param)
    yield hd
    hd.close()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def tol(h, f16, bf16):
    return f16 if h.dtype_id == 0 else bf16


def state_arrays(st, N):
    a = np.zeros((6, N), np.float32)
    for i, k in enumerate(O.BN_STATE_KEYS):
        a[i] = st[k]
    return a


def warm(st, rng):
    st["renorm_mean_weight"] = np.float64(0.26)
    st["renorm_stddev_weight"] = np.float64(0.26)
    st["renorm_stddev"] = 0.26 * (1.5 + 0.3 * rng.random(st["renorm_stddev"].shape))
    st["renorm_mean"] = 0.26 * (0.5 + 0.2 * rng.standard_normal(st["renorm_mean"].shape))
    st["moving_mean"] = 0.1 * rng.standard_normal(st["moving_mean"].shape)
    st["moving_variance"] = 1 + 0.2 * rng.random(st["moving_variance"].shape)
    return st


# rows: one split / ragged splits / empty trailing split / the cfg-2 size; N: one float4 ... several column tiles
@pytest.mark.parametrize("rows,N", [(1, 8), (64, 40), (300, 280), (4097, 1024), (12800, 1024), (100, 136)])
@pytest.mark.parametrize("update", [False, True])
def test_bn_train_stats_and_state(h, rows, N, update):
    dev, rng = h.device, np.random.default_rng(rows + N)
    ld = N + 4
    z = (rng.standard_normal((rows, N)) * (0.5 + rng.random(N)) + 3.0 * rng.standard_normal(N)).astype(np.float32)
    gamma, beta = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32), rng.standard_normal(N).astype(np.float32)
    st = warm(O.bn_init_state(N), rng)
    zt = torch.zeros(rows, ld, dtype=F32, device=dev)
    zt[:, :N] = torch.tensor(z)
    state = torch.tensor(state_arrays(st, N), device=dev)
    coef = torch.zeros(8, N, dtype=F32, device=dev)
    scratch = torch.zeros(768, N, dtype=F32, device=dev)
    h.bn_train_stats(zt, rows, N, torch.tensor(gamma, device=dev), torch.tensor(beta, device=dev), state, coef, scratch,
                     update_state=update)
    torch.cuda.synchronize()
    st_ref = copy.deepcopy(st)
    y_ref, (xh, r, d, _, std) = O.bn_renorm_train_fwd(z.astype(np.float64), gamma.astype(np.float64),
                                                        beta.astype(np.float64), st_ref, update=update)
    c = coef.cpu().numpy().astype(np.float64)
    mean = z.astype(np.float64).mean(0)
    assert rel(c[2], mean) < 1e-5 and rel(c[3], 1.0 / std) < 1e-5
    assert rel(c[4], r) < 1e-5 and rel(c[5] + 1.0, d + 1.0) < 1e-5
    # y = z A + B reproduces the oracle's normalised value
    assert rel(z * c[0] + c[1], y_ref) < 2e-5
    s = state.cpu().numpy()
    for i, k in enumerate(O.BN_STATE_KEYS):
        want = np.broadcast_to(st_ref[k], (N,))
        assert rel(s[i] + 1.0, want + 1.0) < 1e-5, k
        if not update:
            assert np.array_equal(s[i], state_arrays(st, N)[i]), k


def test_bn_eval_coef(h):
    dev, rng, N = h.device, np.random.default_rng(1), 280
    st = warm(O.bn_init_state(N), rng)
    gamma, beta = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32), rng.standard_normal(N).astype(np.float32)
    coef = torch.zeros(8, N, dtype=F32, device=dev)
    h.bn_eval_coef(N, torch.tensor(gamma, device=dev), torch.tensor(beta, device=dev),
                   torch.tensor(state_arrays(st, N), device=dev), coef)
    z = rng.standard_normal((50, N))
    c = coef.cpu().numpy().astype(np.float64)
    assert rel(z * c[0] + c[1], O.bn_eval_fwd(z, gamma.astype(np.float64), beta.astype(np.float64), st)) < 1e-5
    assert np.all(c[4] == 1.0) and np.all(c[5] == 0.0)


@pytest.mark.parametrize("rows,N", [(7, 8), (300, 280), (12800, 1024)])
@pytest.mark.parametrize("act", [O.ACT_RELU, O.ACT_LRELU, O.ACT_NONE])
@pytest.mark.parametrize("keep", [1.0, 0.8])
def test_affine_act_drop_and_mask_bit_exact(h, rows, N, act, keep):
    dev, rng = h.device, np.random.default_rng(rows + N + act)
    z = rng.standard_normal((rows, N)).astype(np.float32)
    A, Bc = (1 + 0.3 * rng.standard_normal(N)).astype(np.float32), rng.standard_normal(N).astype(np.float32)
    seed, tick, salt = 99, 5, 513
    rngbuf = torch.tensor([seed, tick], dtype=torch.int64, device=dev)
    out = torch.zeros(rows, N + 8, dtype=h.h16, device=dev)
    h.affine_act_drop(torch.tensor(z, device=dev), rows, N, torch.tensor(A, device=dev), torch.tensor(Bc, device=dev),
                      act, keep, rngbuf, salt, out)
    torch.cuda.synchronize()
    y = O.act_fwd(z.astype(np.float64) * A + Bc, act)
    got = out[:, :N].float().cpu().numpy()
    if keep < 1.0:
        mask = O.dropout_mask(seed, tick, salt, rows, N, keep)
        y = np.where(mask, y / keep, 0.0)
        nz = np.abs(y) > 1e-3                      # away from values that round to zero anyway
        assert np.array_equal(got[nz] != 0, mask[nz]) and not got[~mask].any()
    assert rel(got, y) < tol(h, 1e-3, 8e-3)
    assert not out[:, N:].any()
    # the tick kernel advances the stream: a different mask
    if keep < 1.0 and rows > 100:
        h.rng_tick(rngbuf)
        out2 = torch.zeros_like(out)
        h.affine_act_drop(torch.tensor(z, device=dev), rows, N, torch.tensor(A, device=dev),
                          torch.tensor(Bc, device=dev), act, keep, rngbuf, salt, out2)
        assert rngbuf.tolist() == [seed, tick + 1]
        m2 = O.dropout_mask(seed, tick + 1, salt, rows, N, keep)
        assert not out2[:, :N].float().cpu().numpy()[~m2].any() and not np.array_equal(m2, mask)


@pytest.mark.parametrize("rows,N", [(5, 8), (300, 280), (12800, 1024)])
@pytest.mark.parametrize("bn,act,keep", [(1, O.ACT_RELU, 1.0), (1, O.ACT_RELU, 0.8), (1, O.ACT_LRELU, 1.0),
                                         (0, O.ACT_RELU, 0.7)])
def test_bn_backward(h, rows, N, bn, act, keep):
    dev, rng = h.device, np.random.default_rng(rows * 3 + N + act)
    z = (rng.standard_normal((rows, N)) + rng.standard_normal(N)).astype(np.float32)
    gamma, beta = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32), (0.3 * rng.standard_normal(N)).astype(np.float32)
    st = warm(O.bn_init_state(N), rng)
    seed, tick, salt = 5, 2, 258
    rngbuf = torch.tensor([seed, tick], dtype=torch.int64, device=dev)
    zt = torch.tensor(z, device=dev)
    gam_t, bet_t = torch.tensor(gamma, device=dev), torch.tensor(beta, device=dev)
    coef = torch.zeros(8, N, dtype=F32, device=dev)
    scratch = torch.zeros(768, N, dtype=F32, device=dev)
    da = (rng.standard_normal((rows, N)) * 0.1).astype(np.float32)
    da16 = torch.tensor(da, device=dev).to(h.h16)
    da_r = da16.double().cpu().numpy()                 # the 16-bit values the kernel actually reads
    dgam = torch.full((N,), 0.5, dtype=F32, device=dev)
    dbet = torch.full((N,), -0.25, dtype=F32, device=dev)
    dz16 = torch.zeros(rows, N, dtype=h.h16, device=dev)
    z64 = z.astype(np.float64)
    if bn:
        h.bn_train_stats(zt, rows, N, gam_t, bet_t, torch.tensor(state_arrays(st, N), device=dev), coef, scratch)
        y, cache = O.bn_renorm_train_fwd(z64, gamma.astype(np.float64), beta.astype(np.float64), copy.deepcopy(st),
                                         update=False)
    else:
        y = z64 + beta
    h.bn_bwd(da16, zt, rows, N, act, keep, rngbuf, salt, bool(bn), coef if bn else None, None if bn else bet_t,
             dgam if bn else None, dbet, dz16, scratch)
    torch.cuda.synchronize()
    g = da_r.copy()
    if keep < 1.0:
        g = np.where(O.dropout_mask(seed, tick, salt, rows, N, keep), g / keep, 0.0)
    dy = O.act_bwd(y, g, act)
    if bn:
        dz, dgamma, dbeta = O.bn_renorm_train_bwd(dy, cache)
        assert rel(dgam.cpu().numpy() - 0.5, dgamma) < 1e-4        # accumulated INTO the gradient buffer
    else:
        dz, dbeta = dy, dy.sum(0)
    if rows > 1:
        assert rel(dbet.cpu().numpy() + 0.25, dbeta) < 1e-4
    got = dz16.float().cpu().numpy()
    assert rel(got, dz) < tol(h, 1e-3, 8e-3)
    if bn and rows >= 300 and keep == 1.0:
        # size-independent property of the batch-norm gradient: dz is orthogonal to 1 and to x_hat, column by column
        xh = cache[0]
        scale = np.abs(dz).sum(0) + 1e-30
        assert np.abs(dz.sum(0) / scale).max() < 1e-9
        assert np.abs(got.astype(np.float64).sum(0) / scale).max() < tol(h, 2e-3, 2e-2)
        assert np.abs((got * xh).sum(0) / scale).max() < tol(h, 2e-3, 2e-2)


@pytest.mark.parametrize("M,N,K,update", [(12800, 1024, 1024, True),     # cfg-2 discriminator layer: two-CTA GEMM
                                          (12800, 1024, 40, False),      # its first layer: one-CTA GEMM, K = 40
                                          (1000, 96, 64, True),          # ragged last row block, N = 3 chunks
                                          (130, 32, 264, False)])        # two row blocks, the second with 2 rows
def test_gemm_epilogue_batch_norm_statistics(h, M, N, K, update):
    """rsr_gemm(stats=...) leaves the (count, mean, M2) of every 128-row block and column of its fp32 output in the partial
    buffer, and rsr_bn_train_finish on those partials == rsr_bn_train_stats on the output (coefficients and UPDATE_OPS)."""
    dev, rng = h.device, np.random.default_rng(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, dtype=h.h16, device=dev)
    A[:, :K] = torch.tensor(rng.standard_normal((M, K)).astype(np.float32), device=dev).to(h.h16)
    W = torch.zeros(Kp, N, dtype=h.h16, device=dev)
    W[:K] = torch.tensor((rng.standard_normal((K, N)) * (0.05 + 0.1 * rng.random(N))).astype(np.float32), device=dev).to(h.h16)
    z = torch.zeros(M, N, dtype=F32, device=dev)
    scratch = torch.zeros(768, N, dtype=F32, device=dev)
    h.gemm(A, W, M, N, Kp, b_mn=True, out32=z, stats=scratch)
    torch.cuda.synchronize()
    zz = z.double().cpu().numpy()
    assert rel(zz, A.double().cpu().numpy() @ W.double().cpu().numpy()) < 1e-5
    nb = (M + 127) // 128
    part = scratch[:3 * nb].cpu().numpy().reshape(nb, 3, N).astype(np.float64)
    for b in range(nb):
        blk = zz[128 * b:128 * b + 128]
        assert np.all(part[b, 0] == blk.shape[0])
        assert np.abs(part[b, 1] - blk.mean(0)).max() < 1e-5 * (np.abs(blk).max() + 1)
        m2 = ((blk - blk.mean(0)) ** 2).sum(0)
        assert np.abs(part[b, 2] - m2).max() < 1e-4 * (m2.max() + 1e-6)
    gamma, beta = (1 + 0.2 * rng.standard_normal(N)).astype(np.float32), rng.standard_normal(N).astype(np.float32)
    st = warm(O.bn_init_state(N), rng)
    gam_t, bet_t = torch.tensor(gamma, device=dev), torch.tensor(beta, device=dev)
    state_a, state_b = (torch.tensor(state_arrays(st, N), device=dev) for _ in range(2))
    coef_a, coef_b = (torch.zeros(8, N, dtype=F32, device=dev) for _ in range(2))
    h.bn_train_finish(nb, M, N, gam_t, bet_t, state_a, coef_a, scratch, update_state=update)
    h.bn_train_stats(z, M, N, gam_t, bet_t, state_b, coef_b, torch.zeros(768, N, dtype=F32, device=dev), update_state=update)
    torch.cuda.synchronize()
    assert rel(coef_a[:6].cpu().numpy(), coef_b[:6].cpu().numpy()) < 1e-5
    assert rel(state_a.cpu().numpy(), state_b.cpu().numpy()) < 1e-5
    from rsrgan_b200._lib import RsrError
    with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # statistics want a plain fp32 output
        h.gemm(A, W, M, N, Kp, b_mn=True, out32=z, bias=bet_t, stats=scratch)


def test_bn_shape_errors(h):
    from rsrgan_b200._lib import RsrError
    dev = h.device
    z = torch.zeros(8, 10, dtype=F32, device=dev)
    v = torch.zeros(16, dtype=F32, device=dev)
    with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # N not a multiple of 4
        h.bn_train_stats(z, 8, 10, v, v, torch.zeros(6, 10, device=dev), torch.zeros(8, 10, device=dev),
                         torch.zeros(768, 10, device=dev))
    with pytest.raises(RsrError, match="RSR_E_ARG"):             # dropout without a stream
        h.affine_act_drop(torch.zeros(8, 8, device=dev), 8, 8, None, v, 1, 0.5, None, 0,
                          torch.zeros(8, 8, dtype=h.h16, device=dev))


# frame layout: (frames, S, L), H lines x C channels, N columns (>= H C: zero-coefficient padding columns)
@pytest.mark.parametrize("frames,S,L,H,C,N", [(3, 8, 5, 1, 12, 16), (256, 264, 257, 1, 32, 32), (48, 48, 40, 11, 12, 136),
                                              (700, 48, 40, 11, 20, 224), (1, 16, 16, 2, 4, 8)])
@pytest.mark.parametrize("update", [False, True])
def test_bn_lines_entries_against_oracle(h, frames, S, L, H, C, N, update):
    """batch_norm behind the convolutions: channel ch of line l = column l * C + ch, data rows r % S < L.  Statistics,
    UPDATE_OPS, normalise (+ zeroed padding rows), backward and the eval coefficients against the float64 oracle run on
    the compacted (frames * L * H, C) matrix."""
    dev, rng = h.device, np.random.default_rng(frames + S + H * C)
    rows, HC = frames * S, H * C
    z = (rng.standard_normal((rows, N)) * 0.7 + 2.0).astype(np.float32)          # padding rows / columns hold junk
    z[:, :HC] = (rng.standard_normal((rows, HC)) * np.tile(0.5 + rng.random(C), H) + np.tile(rng.standard_normal(C), H))
    live = (np.arange(rows) % S) < L
    zc = z[live][:, :HC].astype(np.float64).reshape(-1, C)                        # (frame, position, line) x channel
    gamma, beta = (1 + 0.2 * rng.standard_normal(C)).astype(np.float32), (0.3 * rng.standard_normal(C)).astype(np.float32)
    st = warm(O.bn_init_state(C), rng)
    Cs = -(-C // 8) * 8
    state = torch.zeros(6, Cs, dtype=F32, device=dev)
    state[:, :C] = torch.tensor(state_arrays(st, C), device=dev)
    state0 = state.clone()
    zt = torch.tensor(z, device=dev)
    gam_t, bet_t = torch.tensor(gamma, device=dev), torch.tensor(beta, device=dev)
    coef = torch.full((8, N), 7.0, dtype=F32, device=dev)
    scratch = torch.zeros(768, N, dtype=F32, device=dev)
    h.bn_train_stats_lines(zt, frames, S, L, H, C, N, gam_t, bet_t, state, coef, scratch, update_state=update)
    st_ref = copy.deepcopy(st)
    y_ref, cache = O.bn_renorm_train_fwd(zc, gamma.astype(np.float64), beta.astype(np.float64), st_ref, update=update)
    c = coef.cpu().numpy().astype(np.float64)
    for l in range(H):                                   # every line of a channel holds the channel's coefficients
        assert np.array_equal(c[:6, l * C:(l + 1) * C], c[:6, :C])
    assert np.all(c[:6, HC:] == 0.0)
    assert rel(c[2, :C], zc.mean(0)) < 1e-5 and rel(c[4, :C], cache[1]) < 1e-5
    assert rel(zc * c[0, :C] + c[1, :C], y_ref) < 2e-5
    s = state.cpu().numpy()
    for i, k in enumerate(O.BN_STATE_KEYS):
        assert rel(s[i, :C] + 1.0, np.broadcast_to(st_ref[k], (C,)) + 1.0) < 1e-5, k
    if not update:
        assert torch.equal(state, state0)
    # normalise + ReLU, padding rows zeroed
    y16 = torch.full((rows, N), 3.0, dtype=h.h16, device=dev)
    h.affine_act_lines(zt, frames, S, L, N, coef[0], coef[1], 1, y16)
    y = y16.float().cpu().numpy()
    assert np.all(y[~live] == 0.0) and np.all(y[:, HC:] == 0.0)
    assert rel(y[live][:, :HC].reshape(-1, C), np.maximum(y_ref, 0.0)) < tol(h, 1e-3, 8e-3)
    # backward: da holds junk on the padding rows
    da = (rng.standard_normal((rows, N)) * 0.1).astype(np.float32)
    da16 = torch.tensor(da, device=dev).to(h.h16)
    dac = da16.double().cpu().numpy()[live][:, :HC].reshape(-1, C)
    dgam = torch.full((Cs,), 0.5, dtype=F32, device=dev)
    dbet = torch.full((Cs,), -0.25, dtype=F32, device=dev)
    dz16 = torch.full((rows, N), 3.0, dtype=h.h16, device=dev)
    h.bn_bwd_lines(da16, zt, frames, S, L, H, C, N, 1, coef, dgam, dbet, dz16, scratch)
    torch.cuda.synchronize()
    dz, dgamma, dbeta = O.bn_renorm_train_bwd(O.act_bwd(y_ref, dac, O.ACT_RELU), cache)
    assert rel(dgam.cpu().numpy()[:C] - 0.5, dgamma) < 1e-4 and rel(dbet.cpu().numpy()[:C] + 0.25, dbeta) < 1e-4
    got = dz16.float().cpu().numpy()
    assert np.all(got[~live] == 0.0) and np.all(got[:, HC:] == 0.0)
    assert rel(got[live][:, :HC].reshape(-1, C), dz) < tol(h, 1e-3, 8e-3)
    # inference coefficients from the moving averages
    h.bn_eval_coef_lines(N, H, C, gam_t, bet_t, state, coef)
    c = coef.cpu().numpy().astype(np.float64)
    st_now = OrderedDict((k, s[i, :C].astype(np.float64)) for i, k in enumerate(O.BN_STATE_KEYS))
    assert rel(zc * c[0, :C] + c[1, :C], O.bn_eval_fwd(zc, gamma.astype(np.float64), beta.astype(np.float64), st_now)) < 1e-5
    assert np.array_equal(c[:6, (H - 1) * C:HC], c[:6, :C]) and np.all(c[:6, HC:] == 0.0)


def test_bn_lines_shape_errors(h):
    from rsrgan_b200._lib import RsrError
    dev = h.device
    z = torch.zeros(16, 8, dtype=F32, device=dev)
    v = torch.zeros(8, dtype=F32, device=dev)
    st, coef, scr = torch.zeros(6, 8, device=dev), torch.zeros(8, 8, device=dev), torch.zeros(768, 8, device=dev)
    with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # fewer columns than H * C
        h.bn_train_stats_lines(z, 2, 8, 5, 3, 4, 8, v, v, st, coef, scr)
    with pytest.raises(RsrError, match="RSR_E_SHAPE"):           # S < L
        h.bn_train_stats_lines(z, 2, 4, 5, 1, 8, 8, v, v, st, coef, scr)


# ------------------------------------------------------------------ model level
def tf32(p):
    return OrderedDict((k, np.asarray(v, np.float32)) for k, v in p.items())


def perturb_bn(p, rng):
    for k in p:
        if "BatchNorm" in k:
            p[k] = p[k] + 0.1 * rng.standard_normal(p[k].shape)
    return p


def test_dnn_trainer_with_batch_norm_reference_driver_shape():
    """What run_dnn_single_gpu.sh trains (:129-145): splice 11 x 257 inputs, batch 256, batch_norm on, dropout + l2;
    three Adam steps with the UPDATE_OPS against the oracle, then the inference graph on the moving averages."""
    from rsrgan_b200.dnn_trainer import DNNTrainer
    rng = np.random.default_rng(21)
    N, I, U = 256, 257 * 11, 1024
    args = Namespace(g_type="dnn", batch_size=N, input_dim=257, left_context=5, right_context=5, output_dim=40,
                     g_units=U, batch_norm=True, keep_prob=0.8, l2_scale=1e-5, g_learning_rate=1e-3, seed=11,
                     dtype="f16")
    m = DNNTrainer(None, args, ["/gpu:0"])
    gp = perturb_bn(O.init_g_dnn(rng, in_dim=I, out_dim=40, units=U, hidden=3, batch_norm=True), rng)
    m.load_params(tf32(gp))
    bst = O.init_bn_state(gp)
    st = O.MseState(OrderedDict((k, v.copy()) for k, v in gp.items()), "dnn")
    n0 = m.h.launches
    for step in range(3):
        x, y = rng.standard_normal((N, I)).astype(np.float32), rng.standard_normal((N, 40)).astype(np.float32)
        out = m.train_step(x, y)
        opts = dict(bn_state=bst, update=True, keep_prob=0.8, rng=(11, step))
        losses, _ = O.mse_step(st, x.astype(np.float64), y.astype(np.float64), 1e-3, l2_scale=1e-5, g_opts=opts)
        assert out["g_mse_loss"] == pytest.approx(losses["g_mse_loss"], rel=3e-3)
        assert out["g_l2_loss"] == pytest.approx(losses["g_l2_loss"], rel=1e-3)
    assert m.h.launches > n0
    mine = m.G.bn_state_tf()
    for k in bst:
        assert rel(np.asarray(mine[k]) + 1.0, np.asarray(bst[k]) + 1.0) < 1e-3, k
    # Adam divides by sqrt(v): in the first steps every weight moves by ~lr whatever the size of its gradient, so
    # entries whose gradient is within 16-bit rounding of zero move the other way -- compare the weights in RMS
    th = m.G.P.export_tf()
    for k in gp:
        assert rel(th[k], st.g[k]) < 5e-2, k
    # the inference graph (moving averages, no dropout) on the DEVICE's own weights and statistics
    cv = DNNTrainer(None, args, ["/gpu:0"], cross_validation=True, share=m)
    g = cv.generate(x).cpu().numpy()
    p64 = OrderedDict((k, v.astype(np.float64)) for k, v in th.items())
    bs64 = OrderedDict((k, np.asarray(v, np.float64)) for k, v in mine.items())
    g_ref, _ = O.g_dnn_fwd(p64, x.astype(np.float64), None, opts=dict(bn_state=bs64, train=False))
    d = float(np.sqrt(((g - g_ref) ** 2).mean()))
    assert d < 1e-3, d                                     # north_star: generator output within 1e-3 RMS


@pytest.mark.parametrize("graph", [False, True])
def test_gan_batch_norm_and_dropout(graph):
    """dnn generator + discriminator_dnn, batch_norm in both, dropout in D: gradients of one D and one G update against
    the oracle; then the whole schedule (eager and as a CUDA graph: the tick kernel and the statistics are captured)."""
    from rsrgan_b200.gan_rnn import GAN_RNN
    rng = np.random.default_rng(6)
    B, T, I, U = 8, 50, 257, 256
    args = Namespace(g_type="dnn", d_type="dnn", batch_size=B, input_dim=I, output_dim=40, g_units=U, g_layers=1,
                     d_units=U, d_layers=1, batch_norm=True, keep_prob=0.75, init_mse_weight=10.0, l2_scale=0.0,
                     g_learning_rate=0.0, d_learning_rate=0.0, seed=4, dtype="f16", use_graph=graph)
    m = GAN_RNN(None, args, ["/gpu:0"])
    gp = perturb_bn(O.init_g_dnn(rng, in_dim=I, out_dim=40, units=U, hidden=1, batch_norm=True), rng)
    dp = perturb_bn(O.init_d_dnn(rng, in_dim=40, units=U, hidden=1, batch_norm=True), rng)
    m.load_params(tf32(gp), tf32(dp))
    x = rng.standard_normal((B, T, I)).astype(np.float32)
    y = rng.standard_normal((B, T, 40)).astype(np.float32)
    ln = np.full(B, T)
    xt, yt = x.transpose(1, 0, 2).astype(np.float64), y.transpose(1, 0, 2).astype(np.float64)   # library rows = t*B + b
    st = O.GanState(gp, dp, "dnn", "dnn")
    gbs, dbs = O.init_bn_state(gp), O.init_bn_state(dp)
    # UPDATE_OPS as models/gan_rnn_placeholder.py:163-175 wires them: the D update assigns the discriminator's statistics
    # (both passes), the G update the generator's -- the oracle's running statistics follow the same rule
    g_run, d_run = copy.deepcopy(gbs), copy.deepcopy(dbs)
    if not graph:
        gs = m._gscale(B * T)
        for tick, which in enumerate("dg"):
            go, do = dict(bn_state=g_run), dict(bn_state=d_run, keep_prob=0.75, rng=(4, tick))
            L, G, _ = O.tower_losses_and_grads(st, xt, yt, ln, which, g_opts=go, d_opts=do, update_ops="own")
            out = (m.d_step if which == "d" else m.g_step)(x, y, ln)
            net, keys = (m.D, ("d_rl_loss", "d_fk_loss")) if which == "d" else (m.G, ("g_adv_loss", "g_mse_loss"))
            for k in keys:
                assert out[k] == pytest.approx(L[k], rel=3e-3, abs=1e-5), k
            mine = net.P.export_tf("grad")
            for k in G:
                assert rel(mine[k] / gs, G[k]) < 5e-2, (which, k)
        for net, ref in ((m.G, g_run), (m.D, d_run)):          # ... and the statistics moved exactly as the oracle's
            for k, v in net.bn_state_tf().items():
                assert np.allclose(v, ref[k], rtol=2e-3, atol=2e-5), k
        return
    # three schedules through train_batch: the third replays the captured graph; the dropout stream advances 3 per schedule
    outs = [m.train_batch(x, y, ln) for _ in range(4)]
    torch.cuda.synchronize()
    assert int(m.D.rng[1]) == 12
    # learning rates are 0, so every schedule sees the same weights: losses differ only through the dropout masks of D
    vals = [o["d_fk_loss"] for o in outs]
    assert all(np.isfinite(v) for v in vals) and len(set(vals)) == len(vals)
    for tick0, o in zip((0, 3, 6, 9), outs):
        for k, which in enumerate("dgg"):                      # the schedule: one D update, two G updates
            do = dict(bn_state=d_run, keep_prob=0.75, rng=(4, tick0 + k))
            L, _, _ = O.tower_losses_and_grads(st, xt, yt, ln, which, g_opts=dict(bn_state=g_run), d_opts=do,
                                               update_ops="own")
            if which == "d":
                assert o["d_rl_loss"] == pytest.approx(L["d_rl_loss"], rel=3e-3) and \
                    o["d_fk_loss"] == pytest.approx(L["d_fk_loss"], rel=3e-3, abs=1e-5)
    for net, ref in ((m.G, g_run), (m.D, d_run)):
        for k, v in net.bn_state_tf().items():
            assert np.allclose(v, ref[k], rtol=5e-3, atol=5e-5), k


@pytest.mark.parametrize("g_type,kw", [("lstm", dict(g_cell=128, g_proj=64, g_layers=2, batch_norm=True)),
                                       ("res_lstm_l", dict(g_cell=128, g_layers=2)),
                                       ("res_lstm_base", dict(g_cell=128, g_layers=2))])
def test_lstm_generator_dropout_wrapper(g_type, kw):
    """DropoutWrapper(output_keep_prob) on every LSTM layer of the generator (models/lstm.py:99-102,
    models/res_lstm_l.py:96-99; P = 257 exercises the padded row pitch 264 of the mask), batch_norm on the lstm
    generator's first layer, ragged lengths: losses and raw gradients of one G update, then the whole schedule twice
    (the generator forward is NOT shared between the D and the first G update when it drops)."""
    from rsrgan_b200.gan_rnn import GAN_RNN
    rng = np.random.default_rng(8)
    B, T = 6, 20
    a = Namespace(g_type=g_type, d_type="lstm", batch_size=B, d_cell=64, keep_prob=0.8, init_mse_weight=10.0,
                  init_disc_noise_std=0.0, g_learning_rate=0.0, d_learning_rate=0.0, seed=5, dtype="f16",
                  use_graph=False, **kw)
    m = GAN_RNN(None, a, ["/gpu:0"])
    gp, dp = m.G.P.export_tf(dtype=np.float64), m.D.P.export_tf(dtype=np.float64)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    y = rng.standard_normal((B, T, 40)).astype(np.float32)
    ln = np.array([T, T - 7, T - 1, T // 2, T, 3])
    st = O.GanState(gp, dp, g_type, "lstm")
    go = dict(bn_state=O.init_bn_state(gp), keep_prob=0.8, rng=(5, 0))
    L, G, _ = O.tower_losses_and_grads(st, x.astype(np.float64), y.astype(np.float64), ln, "g", g_opts=go)
    out = m.g_step(x, y, ln)
    gs = m._gscale(B * T)
    for k in ("g_adv_loss", "g_mse_loss"):
        assert out[k] == pytest.approx(L[k], rel=3e-3, abs=1e-5), k
    mine = m.G.P.export_tf("grad")
    for k in G:
        assert rel(mine[k] / gs, G[k]) < 5e-2, k
    n0 = int(m.G.rng[1])
    o1 = m.train_batch(x, y, ln)
    o2 = m.train_batch(x, y, ln)
    assert int(m.G.rng[1]) == n0 + 6 and o1["g_mse_loss"] != o2["g_mse_loss"]       # new masks every update
    go = dict(bn_state=O.init_bn_state(gp), keep_prob=0.8, rng=(5, n0 + 5))          # last G update of the 2nd schedule
    L2, _, _ = O.tower_losses_and_grads(st, x.astype(np.float64), y.astype(np.float64), ln, "g", g_opts=go)
    assert o2["g_mse_loss"] == pytest.approx(L2["g_mse_loss"], rel=3e-3)


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


def test_golden_mse_dnn_bn():
    # weights after three Adam steps carry the sign noise of near-zero gradients (see the reference-driver-shape test),
    # hence the loose bar on the inference output; the statistics and the per-step losses are tight
    _run_golden_mse_dnn_bn(None, 5e-3, 1e-3, 5e-2)


@pytest.mark.parametrize("splice,N,bins", [(1, 256, 257), (11, 48, 40)])
def test_rced_batch_norm_against_oracle(splice, N, bins):
    """models/rced.py:63-71,94-97 under DNNTrainer (run_dnn.sh:129-140 shape for splice 11): batch_norm on the nine
    convolutions.  Loss, UPDATE_OPS and the inference graph tightly; raw gradients at the bar measured for nine stacked
    normalised layers with 16-bit operands (tests/test_batchnorm_host.py shows the wiring exact with fp32 operands)."""
    from rsrgan_b200.dnn_trainer import DNNTrainer
    rng = np.random.default_rng(31 + splice)
    ctx = (splice - 1) // 2
    args = Namespace(g_type="rced", batch_size=N, input_dim=bins, left_context=ctx, right_context=ctx, output_dim=40,
                     batch_norm=True, g_learning_rate=0.0, seed=11, dtype="f16")
    m = DNNTrainer(None, args, ["/gpu:0"])
    gp = perturb_bn(O.init_g_rced(rng, in_dim=bins, out_dim=40, splice=splice, batch_norm=True), rng)
    assert list(m.G.P.segs) == list(gp)
    m.load_params(tf32(gp))
    bst = O.init_bn_state(gp)
    x = rng.standard_normal((N, splice * bins)).astype(np.float32)
    y = rng.standard_normal((N, 40)).astype(np.float32)
    # two steps with lr = 0: the second one sees warmed renorm averages (r != 1, d != 0)
    for step in range(2):
        L, G, _ = O.mse_losses_and_grads(gp, "rced", x.astype(np.float64), y.astype(np.float64),
                                         g_opts=dict(bn_state=bst, update=True))
        out = m.train_step(x, y)
        assert out["g_mse_loss"] == pytest.approx(L["g_mse_loss"], rel=3e-3)
    gs = m._gscale(N)
    gg = m.G.P.export_tf("grad")
    for k in G:
        assert rel(gg[k] / gs, G[k]) < 2e-1, k
    mine = m.G.bn_state_tf()
    assert set(mine) == set(bst)
    for k in bst:
        assert rel(np.asarray(mine[k]) + 1.0, np.asarray(bst[k]) + 1.0) < 1e-3, k
    cv = DNNTrainer(None, args, ["/gpu:0"], cross_validation=True, share=m)
    g = cv.generate(x).cpu().numpy()
    g_ref, _ = O.g_rced_fwd(gp, x.astype(np.float64), None, opts=dict(bn_state=bst, train=False))
    d = float(np.sqrt(((g - g_ref) ** 2).mean()))
    assert d < 1e-3 and rel(g, g_ref) < 5e-3, (d, rel(g, g_ref))
    # a real Adam step moves the weights and lowers the loss
    m.g_learning_rate = 1e-3
    losses = [m.train_step(x, y)["g_mse_loss"] for _ in range(4)]
    assert np.isfinite(losses).all() and losses[-1] < losses[0], losses
