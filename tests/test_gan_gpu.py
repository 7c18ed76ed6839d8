"""End-to-end parity of the CUDA GAN step (GAN_RNN over librsrgan_sm100.so) with the oracle and the
committed golden vectors, plus size-independent properties at BASELINE.json's full batch size.

Tolerances (fp16 tensor-core operands -- the PRODUCT dtype --, fp32 accumulate / cell state; north_star bar:
generator output within 1e-3 RMS of the reference).  The bars below are 2x what scripts/gpu_measure_parity.py measured
on a B200 (profiles/r2_parity_measured_v0.jsonl), never looser than that:
    generator output  : absolute RMS < 1e-3 (north_star) at every size incl. the benchmarked T = 100 / T = 200
                        (measured 5.8e-5 at cfg-2 B = 128 x T = 100, 4.3e-4 at cfg-5 T = 200)
    losses            : relative 2e-3 (measured 3e-6)
    raw gradients     : relative RMS per tensor 2e-2 (measured worst tensor 9.7e-3: the first fully_connected of a
                        network, whose input is the 16-bit rounded feature itself; median 1e-3), weight deltas likewise
bf16 operands (BASELINE.json configs[4] names bf16) are supported but do NOT meet the 1e-3 bar at cfg-5 (3.5e-3 at
T = 200; 4.6e-4 at cfg-2): test_cfg5_T200_against_oracle[bf16] records that as an expected failure.
"""
import os
from argparse import Namespace
from collections import OrderedDict

import numpy as np
import pytest
import torch

from oracle import rsr_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
GRAD_BAR = 2e-2            # per-tensor relative RMS of raw gradients, fp16 operands (2x measured)
# weight DELTAS of one SGD step (theta_after - theta_before in fp32): lr * g is ~1e-6 against weights of ~1e-1, so the
# difference of two fp32 weights carries their storage rounding (6e-9 / 1e-6 = a per cent or two) on top of the
# gradient error: measured up to 3.7e-2
DELTA_BAR = 6e-2


def make_model(g_type, d_type, B, **kw):
    from rsrgan_b200.gan_rnn import GAN_RNN
    a = dict(g_type=g_type, d_type=d_type, batch_size=B, init_mse_weight=10.0, init_disc_noise_std=0.05,
             g_learning_rate=8e-5, d_learning_rate=1e-3, l2_scale=0.0, seed=3, dtype="f16")
    a.update(kw)
    return GAN_RNN(None, Namespace(**a), ["/gpu:0"])


def rms(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = float(np.sqrt(((a - b) ** 2).mean()))
    return d, d / (float(np.sqrt((b ** 2).mean())) + 1e-30)


def load_gold(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    gp = OrderedDict((k[2:], z[k]) for k in z.files if k.startswith("G/"))
    dp = OrderedDict((k[2:], z[k]) for k in z.files if k.startswith("D/"))
    return z, gp, dp


@pytest.mark.parametrize("name,kw", [
    ("gan_lstm_dlstm", dict(g_type="lstm", d_type="lstm", g_cell=64, g_proj=32, g_layers=2, d_cell=32)),
    ("gan_res_ddnn", dict(g_type="res_lstm_l", d_type="dnn", g_cell=40, g_layers=2, d_units=64)),
])
def test_golden_vectors(name, kw):
    z, gp, dp = load_gold(name)
    B, T = z["x"].shape[:2]
    kw = dict(kw)
    m = make_model(kw.pop("g_type"), kw.pop("d_type"), B, **kw)
    m.load_params(gp, dp)
    nz = dict(noise_rl=z["noise_rl"], noise_fk=z["noise_fk"]) if "noise_rl" in z.files else {}
    a, r = rms(m.generate(z["x"], z["lengths"]).cpu().numpy(), z["g_out"])
    assert a < 1e-3 and r < 3e-3, (a, r)
    ev = m.eval_losses(z["x"], z["y"], z["lengths"], **nz)
    for k in ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_loss"):
        assert ev[k] == pytest.approx(float(z["loss/" + k]), rel=2e-3, abs=1e-5), k
    m.d_learning_rate, m.g_learning_rate = 0.0, 0.0
    gs = m._gscale(B * T)
    m.d_step(z["x"], z["y"], z["lengths"], **nz)
    dg = m.D.P.export_tf("grad")
    for k in dp:
        assert rms(dg[k] / gs, z["dgrad/" + k])[1] < GRAD_BAR, k
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=nz.get("noise_fk"))
    gg = m.G.P.export_tf("grad")
    for k in gp:
        assert rms(gg[k] / gs, z["ggrad/" + k])[1] < GRAD_BAR, k
    # the whole batch schedule with the reference learning rates
    m.load_params(gp, dp)
    m.G.P.m.zero_(); m.G.P.v.zero_(); m.G.P.hyper[4:6] = torch.tensor([0.9, 0.999], device=m.h.device)
    m.d_learning_rate, m.g_learning_rate = 1e-3, 8e-5
    m.d_step(z["x"], z["y"], z["lengths"], **nz)
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=nz.get("noise_fk"))
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=nz.get("noise_fk"))
    a, r = rms(m.generate(z["x"], z["lengths"]).cpu().numpy(), z["g_out_after"])
    assert a < 1e-3 and r < 5e-3, (a, r)


@pytest.mark.parametrize("g_type,d_type,B,T", [
    ("lstm", "lstm", 8, 30),              # reference-native sizes: models/lstm.py:43-45 + discriminator_lstm.py:26-28
    ("res_lstm_l", "lstm", 8, 24),        # what run_gan_rnn_placeholder.sh:124 trains
    ("res_lstm_base", "dnn", 4, 12),
])
def test_reference_native_sizes_against_oracle(g_type, d_type, B, T):
    m = make_model(g_type, d_type, B)
    rng = np.random.default_rng(B * T)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[0] = T
    n_rl, n_fk = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32), (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32)
    st = O.GanState(OrderedDict((k, v.astype(np.float64)) for k, v in m.G.P.export_tf().items()),
                    OrderedDict((k, v.astype(np.float64)) for k, v in m.D.P.export_tf().items()), g_type, d_type)
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(st.g, x.astype(np.float64), lengths)
    a, r = rms(m.generate(x, lengths).cpu().numpy(), g_ref)
    assert a < 1e-3 and r < 3e-3, (a, r)
    lstm_d = d_type == "lstm"
    tower = dict(x=x.astype(np.float64), y=y.astype(np.float64), lengths=lengths,
                 noise_rl=n_rl.astype(np.float64) if lstm_d else None, noise_fk=n_fk.astype(np.float64) if lstm_d else None)
    d0 = m.D.P.export_tf()
    ours = m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    ref_losses, _ = O.d_step(st, [tower], 1e-3)
    assert ours["d_loss"] == pytest.approx(ref_losses[0]["d_loss"], rel=2e-3)
    d1 = m.D.P.export_tf()
    for k in d0:
        assert rms(d1[k] - d0[k], st.d[k] - d0[k].astype(np.float64))[1] < DELTA_BAR, k
    ours = m.g_step(x, y, lengths, noise_fk=n_fk)
    ref_losses, _ = O.g_step(st, [tower], 8e-5)
    assert ours["g_loss"] == pytest.approx(ref_losses[0]["g_loss"], rel=2e-3)
    g_ref, _ = gf(st.g, x.astype(np.float64), lengths)
    a, r = rms(m.generate(x, lengths).cpu().numpy(), g_ref)
    assert a < 1e-3 and r < 5e-3, (a, r)


def test_full_size_properties_cfg2():
    """BASELINE.json configs[1] (B=128 x T=100, 2xLSTMP-512/256 G + DNN D): properties that need no oracle run."""
    B, T = 128, 100
    m = make_model("lstm", "dnn", B, g_cell=512, g_proj=256, g_layers=2)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[5] = T
    g1 = m.generate(x, lengths).cpu().numpy()
    b_out = m.G.P.export_tf()["g_model/fully_connected_1/biases"]
    # (1) padded frames: dynamic_rnn zero output -> y = b_out exactly; (2) causality + batch independence:
    #     utterance b, frames < t depend only on x[b, :t]
    for b in (0, 17, 127):
        assert np.allclose(g1[b, lengths[b]:], b_out, atol=1e-6)
    x2 = x.copy()
    x2[5, 60:] += 1.0
    x2[9] *= -1.0
    g2 = m.generate(x2, lengths).cpu().numpy()
    keep = [b for b in range(B) if b not in (5, 9)]
    assert np.array_equal(g1[keep], g2[keep])
    assert np.array_equal(g1[5, :60], g2[5, :60]) and not np.array_equal(g1[5, 60:lengths[5]], g2[5, 60:lengths[5]])
    # (3) determinism of the whole schedule (D-DNN has no noise): two models, same seed, same batch -> same bits
    m2 = make_model("lstm", "dnn", B, g_cell=512, g_proj=256, g_layers=2)
    o1, o2 = m.train_batch(x, y, lengths), m2.train_batch(x, y, lengths)
    assert o1["g_mse_loss"] == pytest.approx(o2["g_mse_loss"], rel=1e-5)
    # (4) loss identities: d_loss = d_rl + d_fk, g_loss = g_adv + lambda g_mse ; a G step lowers g_mse on the same batch
    assert o1["d_loss"] == pytest.approx(o1["d_rl_loss"] + o1["d_fk_loss"], rel=1e-6)
    before = m.eval_losses(x, y, lengths)["g_mse_loss"]
    for _ in range(3):
        m.g_step(x, y, lengths)
    assert m.eval_losses(x, y, lengths)["g_mse_loss"] < before
    # (5) padding of every parameter tensor is still exactly zero after the updates
    for net in (m.G, m.D):
        flat = net.P.theta.cpu().numpy()
        from rsrgan_b200 import params
        for s in net.P.segs.values():
            n = int(np.prod(s.dev_shape))
            d = flat[s.off:s.off + n]
            assert np.array_equal(params.to_dev_layout(s, params.from_dev_layout(s, d)).reshape(-1), d), s.name


@pytest.mark.parametrize("g_type,d_type,kw", [
    ("lstm", "dnn", dict(g_cell=512, g_proj=256, g_layers=2)),            # cluster kernels + DNN discriminator
    ("lstm", "lstm", dict(g_cell=256, g_proj=64, g_layers=1, d_cell=256)),  # recurrences on both streams, D input noise
])
def test_cuda_graph_and_stream_overlap_match_eager_serial(g_type, d_type, kw):
    """The captured schedule (CUDA graph, D(real) and weight-gradient GEMMs on the side stream) trains to the
    same weights as the eager, single-stream schedule: only the order of fp32 atomic accumulation differs."""
    B, T = 24, 20
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    ma = make_model(g_type, d_type, B, init_disc_noise_std=0.0, **kw)
    mb = make_model(g_type, d_type, B, init_disc_noise_std=0.0, use_graph=False, **kw)
    mb.h.overlap = False
    assert ma.use_graph and ma.h.overlap
    for i in range(5):                                  # calls 1-2 eager (workspace allocation), 3 captures, 4-5 replay
        oa, ob = ma.train_batch(x, y, lengths), mb.train_batch(x, y, lengths)
    assert any(st["graph"] is not None for st in ma._graphs.values())
    for k in ob:
        assert oa[k] == pytest.approx(ob[k], rel=2e-3, abs=1e-6), k
    for na, nb in ((ma.G, mb.G), (ma.D, mb.D)):
        ta, tb = na.P.theta.cpu().numpy(), nb.P.theta.cpu().numpy()
        assert rms(ta, tb)[1] < 1e-4
    # a changed by-value scalar (the decayed noise std, train...py:529-533) re-captures instead of replaying stale values
    n_graphs = len(ma._graphs)
    ma.disc_noise_std = 0.01 if d_type == "lstm" else 0.0
    ma.mse_lambda = 5.0
    mb.mse_lambda = 5.0
    oa, ob = ma.train_batch(x, y, lengths), mb.train_batch(x, y, lengths)
    assert len(ma._graphs) == n_graphs + 1
    if d_type != "lstm":
        assert oa["g_loss"] == pytest.approx(ob["g_loss"], rel=2e-3)


def test_prefetch_double_buffering_feeds_the_right_batches():
    """GAN_RNN.prefetch starts the H2D copy of the next minibatch on a copy stream while the current schedule
    runs; training with it must consume exactly the batches it was given, in order."""
    B, T = 16, 12
    rng = np.random.default_rng(4)
    kw = dict(g_cell=256, g_proj=64, g_layers=1)
    batches = []
    for _ in range(3):
        x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
        ln = rng.integers(T // 2, T + 1, size=B).astype(np.float32)      # the reference feeds lengths as float32
        batches.append(tuple(torch.from_numpy(v).pin_memory() for v in (x, y, ln)))
    ma, mb = make_model("lstm", "dnn", B, **kw), make_model("lstm", "dnn", B, **kw)
    order = [0, 1, 2, 0, 1, 2, 1]
    for i in order:
        ma.train_batch(*batches[i])
    mb.prefetch(*batches[order[0]])
    for n, i in enumerate(order):
        out = mb.train_batch(*batches[i], sync=False)
        if n + 1 < len(order):
            mb.prefetch(*batches[order[n + 1]])
    torch.cuda.synchronize()
    for na, nb in ((ma.G, mb.G), (ma.D, mb.D)):
        assert rms(na.P.theta.cpu().numpy(), nb.P.theta.cpu().numpy())[1] < 1e-4


@pytest.mark.parametrize("dtype,abs_bar,rel_bar", [("f16", 1e-3, 3e-3), ("bf16", 8e-3, 2.5e-2)])
def test_cfg5_res_lstm_l_1024_against_oracle(dtype, abs_bar, rel_bar):
    """BASELINE.json configs[4]: res_lstm_l generator, 4 layers x C = 1024 (P = 257), discriminator_lstm, B = 64 per GPU
    (T shortened from 200 so the float64 oracle finishes in seconds; the full-length run is bench.py --config cfg5).
    C = 1024 takes the L2-exchange recurrence with the weight slab split between shared memory and TMEM.
    fp16 operands meet the north_star 1e-3 RMS bar; with bf16 operands (8 mantissa bits) the measured generator RMS
    is ~3e-3, so that arm is held to the bf16 rounding bar instead."""
    B, T = 64, 10
    m = make_model("res_lstm_l", "lstm", B, g_cell=1024, g_layers=4, dtype=dtype)
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[0] = T
    n_rl, n_fk = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32), (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32)
    st = O.GanState(OrderedDict((k, v.astype(np.float64)) for k, v in m.G.P.export_tf().items()),
                    OrderedDict((k, v.astype(np.float64)) for k, v in m.D.P.export_tf().items()), "res_lstm_l", "lstm")
    assert st.g["g_model/lstm_cell_4/rnn/lstm_cell/kernel"].shape == (257 + 257, 4096)
    g_ref, _ = O.g_res_lstm_l_fwd(st.g, x.astype(np.float64), lengths)
    a, r = rms(m.generate(x, lengths).cpu().numpy(), g_ref)
    assert a < abs_bar and r < rel_bar, (a, r)
    tower = dict(x=x.astype(np.float64), y=y.astype(np.float64), lengths=lengths,
                 noise_rl=n_rl.astype(np.float64), noise_fk=n_fk.astype(np.float64))
    ours = m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    ref, _ = O.d_step(st, [tower], 1e-3)
    assert ours["d_loss"] == pytest.approx(ref[0]["d_loss"], rel=2e-3 if dtype == "f16" else 2e-2)
    ours = m.g_step(x, y, lengths, noise_fk=n_fk)
    ref, grads = O.g_step(st, [tower], 8e-5)
    assert ours["g_loss"] == pytest.approx(ref[0]["g_loss"], rel=2e-3 if dtype == "f16" else 2e-2)


# ---------------------------------------------------------------------------------------------------------------------
# parity at the BENCHMARKED sequence lengths (the float64 oracle on a slice of the utterances: the generator treats
# utterances independently; one whole update schedule at a batch the oracle finishes in seconds)
# ---------------------------------------------------------------------------------------------------------------------
def _oracle_state(m, g_type, d_type):
    return O.GanState(OrderedDict((k, v.astype(np.float64)) for k, v in m.G.P.export_tf().items()),
                      OrderedDict((k, v.astype(np.float64)) for k, v in m.D.P.export_tf().items()), g_type, d_type)


def _g_slice(g_type, d_type, B, T, idx, dtype, kw):
    m = make_model(g_type, d_type, B, dtype=dtype, **kw)
    rng = np.random.default_rng(B + T)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[idx[0]] = T
    g = m.generate(x, lengths).cpu().numpy()
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(_oracle_state(m, g_type, d_type).g, x[idx].astype(np.float64), lengths[idx])
    return rms(g[idx], g_ref)


def _schedule(g_type, d_type, B, T, dtype, kw, grad_bar, out_bar):
    m = make_model(g_type, d_type, B, dtype=dtype, use_graph=False, **kw)
    rng = np.random.default_rng(B * T)
    x, y = rng.standard_normal((B, T, 257)).astype(np.float32), rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    lengths[0] = T
    lstm_d = d_type == "lstm"
    n_rl = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32) if lstm_d else None
    n_fk = (rng.standard_normal((B, 1, 40)) * 0.05).astype(np.float32) if lstm_d else None
    f64 = lambda v: None if v is None else v.astype(np.float64)
    st = _oracle_state(m, g_type, d_type)
    tower = dict(x=f64(x), y=f64(y), lengths=lengths, noise_rl=f64(n_rl), noise_fk=f64(n_fk))
    gs = m._gscale(B * T)
    m.d_learning_rate, m.g_learning_rate = 0.0, 0.0           # raw gradients of one D and one G update
    ours = m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    L, G, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "d", tower["noise_rl"], tower["noise_fk"])
    assert ours["d_loss"] == pytest.approx(L["d_loss"], rel=2e-3)
    mine = m.D.P.export_tf("grad")
    for k in G:
        assert rms(mine[k] / gs, G[k])[1] < grad_bar, ("d", k)
    ours = m.g_step(x, y, lengths, noise_fk=n_fk)
    L, G, _ = O.tower_losses_and_grads(st, tower["x"], tower["y"], lengths, "g", tower["noise_rl"], tower["noise_fk"])
    assert ours["g_loss"] == pytest.approx(L["g_loss"], rel=2e-3)
    mine = m.G.P.export_tf("grad")
    for k in G:
        assert rms(mine[k] / gs, G[k])[1] < grad_bar, ("g", k)
    # the whole schedule (1 D + 2 G updates) with the reference learning rates, then the generator again
    m.d_learning_rate, m.g_learning_rate = 1e-3, 8e-5
    m.G.P.m.zero_(); m.G.P.v.zero_(); m.G.P.hyper[4:6] = torch.tensor([0.9, 0.999], device=m.h.device)
    m.d_step(x, y, lengths, noise_rl=n_rl, noise_fk=n_fk)
    O.d_step(st, [tower], 1e-3)
    for _ in range(2):
        m.g_step(x, y, lengths, noise_fk=n_fk)
        O.g_step(st, [tower], 8e-5)
    gf, _ = O.GENERATORS[g_type]
    g_ref, _ = gf(st.g, tower["x"], lengths)
    a, r = rms(m.generate(x, lengths).cpu().numpy(), g_ref)
    assert a < out_bar, (a, r)
    assert m.skipped_updates() == (0, 0)


CFG2 = dict(g_cell=512, g_proj=256, g_layers=2)
CFG5 = dict(g_cell=1024, g_layers=4)


@pytest.mark.parametrize("dtype,out_bar,grad_bar", [("f16", 1.2e-4, 2e-2), ("bf16", 1e-3, 6e-2)])
def test_cfg2_T100_against_oracle(dtype, out_bar, grad_bar):
    """BASELINE.json configs[1] at the benchmarked shape (B = 128 x T = 100, ragged lengths): generator output of eight
    utterances spread over the four utterance groups against the float64 oracle; then one D update, one G update (raw
    gradients per tensor) and a whole schedule at B = 24 x T = 100.  out_bar: 2x the measured RMS for fp16 (5.8e-5), the
    north_star bar 1e-3 for bf16 (measured 4.6e-4)."""
    a, r = _g_slice("lstm", "dnn", 128, 100, [0, 17, 31, 32, 63, 64, 100, 127], dtype, CFG2)
    assert a < out_bar <= 1e-3, (a, r)
    _schedule("lstm", "dnn", 24, 100, dtype, CFG2, grad_bar, out_bar)


@pytest.mark.parametrize("dtype", ["f16", "bf16"])
def test_cfg5_T200_against_oracle(dtype):
    """BASELINE.json configs[4] at its sequence length (res_lstm_l 4 x 1024, T = 200, B = 64 per GPU): generator output
    of four utterances against the float64 oracle, one D / G update and a whole schedule at B = 4.  fp16 (the product
    dtype) meets the 1e-3 north_star bar (measured 4.3e-4); bf16, which configs[4] names, does NOT (measured 3.5e-3):
    that arm checks its own regression bound and then reports the miss as an expected failure."""
    a, r = _g_slice("res_lstm_l", "lstm", 64, 200, [0, 21, 42, 63], dtype, CFG5)
    if dtype == "f16":
        assert a < 9e-4, (a, r)
        _schedule("res_lstm_l", "lstm", 4, 200, dtype, CFG5, 2e-3, 1e-3)
        return
    assert a < 7e-3, (a, r)
    _schedule("res_lstm_l", "lstm", 4, 200, dtype, CFG5, 1.2e-2, 7e-3)
    if a >= 1e-3:
        pytest.xfail("bf16 operands: generator output RMS %.1e at cfg-5 T = 200 is above the 1e-3 parity bar "
                     "(fp16 is the product dtype; DESIGN.md section 2)" % a)


@pytest.mark.parametrize("g_type,d_type,T,kw,bar", [("lstm", "dnn", 1000, CFG2, 1.5e-4),
                                                    ("res_lstm_l", "lstm", 400, {}, 9.5e-4)])
def test_decode_of_one_long_utterance_against_oracle(g_type, d_type, T, kw, bar):
    """The reference decodes whole utterances one at a time (batch_size = 1, scripts/train_gan_rnn_placeholder.py:262-300):
    generator output of one 1000-frame utterance through the cfg-2 generator, and of a 400-frame one through the network
    the shipped driver trains (res_lstm_l 4 x 760, the L2-exchange kernels), against the float64 oracle.  fp16, the
    product dtype; bars = 2x the measured RMS (7.0e-5 / 4.6e-4, profiles/r2_long_T_v1.jsonl), both inside 1e-3."""
    a, r = _g_slice(g_type, d_type, 1, T, [0], "f16", kw)
    assert a < bar <= 1e-3, (a, r)


def test_graph_replay_survives_workspace_growth():
    """A captured schedule holds raw workspace addresses.  A longer batch (train_batch) or a longer cross-validation
    utterance (eval_losses on the model that shares the workspace) replaces workspace buffers: the graphs captured
    before must be dropped and re-captured, never replayed on freed memory.  Twin model, eager, same sequence."""
    from rsrgan_b200.gan_rnn import GAN_RNN
    kw = dict(g_cell=256, g_proj=64, g_layers=1, init_disc_noise_std=0.0)
    ma = make_model("lstm", "dnn", 16, **kw)
    mb = make_model("lstm", "dnn", 16, use_graph=False, **kw)
    cv = GAN_RNN(None, Namespace(g_type="lstm", d_type="dnn", batch_size=16, init_mse_weight=10.0, dtype="f16"), ["/gpu:0"],
                 cross_validation=True, share=ma)
    rng = np.random.default_rng(2)

    def batch(T):
        return (rng.standard_normal((16, T, 257)).astype(np.float32), rng.standard_normal((16, T, 40)).astype(np.float32),
                rng.integers(T // 2, T + 1, size=16))
    b12, b20, b36 = batch(12), batch(20), batch(36)
    for _ in range(4):                                   # capture at T = 12
        ma.train_batch(*b12); mb.train_batch(*b12)
    assert any(st["graph"] is not None for st in ma._graphs.values())
    gen0 = ma._ws_generation()
    for _ in range(4):                                   # a longer batch grows the workspace and gets its own graph
        ma.train_batch(*b20); mb.train_batch(*b20)
    cv.eval_losses(*b36)                                 # longer still, eager, on the SHARED workspace
    assert ma._ws_generation() != gen0
    for b in (b12, b20, b12, b12, b12, b20):             # back to the shapes whose graphs are stale now
        oa, ob = ma.train_batch(*b), mb.train_batch(*b)
        for k in ob:
            assert oa[k] == pytest.approx(ob[k], rel=2e-3, abs=1e-6), k
    assert any(st["graph"] is not None for st in ma._graphs.values())      # ... and they were captured again
    for na, nb in ((ma.G, mb.G), (ma.D, mb.D)):
        assert rms(na.P.theta.cpu().numpy(), nb.P.theta.cpu().numpy())[1] < 1e-4


@pytest.mark.parametrize("g_type,d_type,B,T,kw", [
    ("lstm", "dnn", 96, 24, dict(g_cell=512, g_proj=256, g_layers=2)),      # cfg-2 generator: forward AND backward wavefront
    ("lstm", "dnn", 128, 16, dict(g_cell=512, g_proj=256, g_layers=2)),     # ... benchmarked batch: forward only (48 per cluster)
    ("lstm", "lstm", 40, 20, dict(g_cell=256, g_proj=64, g_layers=3)),      # three layers (pair + single), LSTM discriminator
])
def test_layer_wavefront_matches_layer_by_layer(monkeypatch, g_type, d_type, B, T, kw):
    """One whole schedule (1 D + 2 G updates) with the wavefront launches == the same schedule with every LSTMP layer
    launched on its own (RSR_NO_WAVE=1): same losses, same generator afterwards.  The two paths differ only in where 16-bit
    roundings fall (the projection stage against the projection GEMM; dz2 (W_p1 K_x2)^T folded against two GEMMs)."""
    rng = np.random.default_rng(B + T)
    x = rng.standard_normal((B, T, 257)).astype(np.float32)
    y = rng.standard_normal((B, T, 40)).astype(np.float32)
    lengths = rng.integers(T // 2, T + 1, size=B)
    nz = {}
    outs = []
    for no_wave in (False, True):
        if no_wave:
            monkeypatch.setenv("RSR_NO_WAVE", "1")
        m = make_model(g_type, d_type, B, use_graph=False, init_disc_noise_std=0.0, **kw)
        n0 = m.h.launches
        losses = m.train_batch(x, y, lengths)
        g = m.generate(x, lengths).cpu().numpy()
        grads = m.G.P.grad.clone().cpu().numpy()
        outs.append((losses, g, grads, m.h.launches - n0))
    (l_w, g_w, gr_w, n_w), (l_s, g_s, gr_s, n_s) = outs
    assert n_w < n_s                                   # the wavefront path really ran (fewer launches)
    for k in l_s:
        assert l_w[k] == pytest.approx(l_s[k], rel=2e-3, abs=1e-6), k
    a, r = rms(g_w, g_s)
    assert r < 2e-3, (a, r)
    a, r = rms(gr_w, gr_s)
    assert r < 1e-2, (a, r)
