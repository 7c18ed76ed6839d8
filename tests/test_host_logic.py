"""HOST logic on CPU: the GAN_RNN wiring (layer order, packing, residuals, D/G update schedule,
checkpoints, data-parallel all-reduce) driven through the test double tests/fake_handle.py and
checked against the oracle.  The CUDA kernels themselves are checked by the `-m gpu` tests."""
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
from rsrgan_b200 import packing, params  # noqa: E402
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def make_model(g_type, d_type, B, **kw):
    a = dict(g_type=g_type, d_type=d_type, batch_size=B, init_mse_weight=10.0, init_disc_noise_std=0.05,
             g_learning_rate=8e-5, d_learning_rate=1e-3, l2_scale=0.0, seed=3)
    a.update(kw)
    return GAN_RNN(None, Namespace(**a), ["/gpu:0"], handle=FakeHandle("f16"))


def load_gold(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    gp = OrderedDict((k[2:], z[k]) for k in z.files if k.startswith("G/"))
    dp = OrderedDict((k[2:], z[k]) for k in z.files if k.startswith("D/"))
    return z, gp, dp


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def test_packed_gate_columns_roundtrip():
    C = 40
    a = np.arange(3 * 4 * C, dtype=np.float32).reshape(3, 4 * C)
    p = packing.pack_cols(a, C)
    assert p.shape == (3, 4 * packing.cell_pad(C))
    assert np.array_equal(packing.unpack_cols(p, C), a)
    # gate g of cell c sits at (c//32)*128 + g*32 + c%32
    assert p[0, 1 * 128 + 2 * 32 + 3] == a[0, 2 * C + 35]
    for s in params.lstm_cell("x/", 257, 760, 257) + [params.fc_w("w", 257, 40), params.fc_b("b", 1),
                                                      params.conv_w("c", 13, 12, 20), params.fc_w_frames("f", 257, 12, 40)]:
        t = np.random.default_rng(0).standard_normal(s.tf_shape).astype(np.float32)
        assert np.array_equal(params.from_dev_layout(s, params.to_dev_layout(s, t)), t)


@pytest.mark.parametrize("name,kw", [
    ("gan_lstm_dlstm", dict(g_type="lstm", d_type="lstm", g_cell=64, g_proj=32, g_layers=2, d_cell=32)),
    ("gan_res_ddnn", dict(g_type="res_lstm_l", d_type="dnn", g_cell=40, g_layers=2, d_units=64)),
    ("gan_rced_ddnn", dict(g_type="rced", d_type="dnn", d_units=64)),
])
def test_wiring_matches_golden(name, kw):
    z, gp, dp = load_gold(name)
    B = z["x"].shape[0]
    g_type, d_type = kw.pop("g_type"), kw.pop("d_type")
    m = make_model(g_type, d_type, B, **kw)
    m.load_params(gp, dp)
    assert m.G.P.n_params() == sum(v.size for v in gp.values())
    nz = dict(noise_rl=z["noise_rl"], noise_fk=z["noise_fk"]) if "noise_rl" in z.files else {}
    g = m.generate(z["x"], z["lengths"]).numpy()
    assert rel(g, z["g_out"]) < 3e-3
    ev = m.eval_losses(z["x"], z["y"], z["lengths"], **nz)
    for k in ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_loss"):
        # (d_fk of the rced case is the square of a ~0.1 logit computed from a ~0.03-sized generator output)
        assert ev[k] == pytest.approx(float(z["loss/" + k]), rel=5e-3 if g_type == "rced" else 2e-3, abs=1e-5), k
    # raw gradients of one D update and one G update
    m.d_learning_rate, m.g_learning_rate = 0.0, 0.0             # keep weights fixed: compare gradients only
    gs = m._gscale(B * z["x"].shape[1])
    m.d_step(z["x"], z["y"], z["lengths"], **nz)
    dg = m.D.P.export_tf("grad")
    for k in dp:
        assert rel(dg[k] / gs, z["dgrad/" + k]) < 2e-2, k
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=nz.get("noise_fk"))
    gg = m.G.P.export_tf("grad")
    # nine stacked ReLU layers: a pre-activation within rounding distance of zero flips its mask and with it a whole
    # gradient entry (a fraction f of flipped entries is an RMS error of sqrt(f)), so the bar is looser for rced
    for k in gp:
        assert rel(gg[k] / gs, z["ggrad/" + k]) < (1e-1 if g_type == "rced" else 2e-2), k


def test_batch_schedule_matches_oracle_after_updates():
    z, gp, dp = load_gold("gan_lstm_dlstm")
    B = z["x"].shape[0]
    m = make_model("lstm", "lstm", B, g_cell=64, g_proj=32, g_layers=2, d_cell=32)
    m.load_params(gp, dp)
    nz = dict(noise_rl=z["noise_rl"], noise_fk=z["noise_fk"])
    m.d_step(z["x"], z["y"], z["lengths"], **nz)
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=z["noise_fk"])
    m.g_step(z["x"], z["y"], z["lengths"], noise_fk=z["noise_fk"])
    g = m.generate(z["x"], z["lengths"]).numpy()
    assert rel(g, z["g_out_after"]) < 5e-3
    # padded frames: dynamic_rnn emits zeros -> y = b_out
    b_out = m.G.P.export_tf()["g_model/fully_connected_1/biases"]
    for b, n in enumerate(z["lengths"]):
        assert np.allclose(g[b, n:], b_out, atol=1e-6)
    # EMA shadow moved by (1-decay) of the update
    th, ema = m.G.P.export_tf(), m.G.P.export_tf("ema")
    k = "g_model/fully_connected_1/weights"
    assert not np.array_equal(th[k], gp[k]) and rel(ema[k], gp[k]) < 1e-3 and not np.array_equal(ema[k], gp[k])


def test_padding_stays_zero_and_l2():
    m = make_model("lstm", "dnn", 2, g_cell=40, g_proj=24, g_layers=1, d_units=32, l2_scale=1e-3)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 4, 257)).astype(np.float32), rng.standard_normal((2, 4, 40)).astype(np.float32)
    before = m.G.P.export_tf()
    out = m.train_batch(x, y, np.array([4, 3]))
    l2_0 = O.l2_loss_g(before, 1e-3)                      # value at the first G update; weights barely move
    assert out["g_l2_loss"] == pytest.approx(l2_0, rel=1e-2)
    assert out["g_l2_loss"] > 0 and out["g_loss"] == pytest.approx(out["g_adv_loss"] + 10 * out["g_mse_loss"] + out["g_l2_loss"], rel=1e-6)
    for net in (m.G, m.D):
        P = net.P
        for s in P.segs.values():
            n = int(np.prod(s.dev_shape))
            d = P.theta[s.off:s.off + s.size].numpy()
            back = params.to_dev_layout(s, params.from_dev_layout(s, d[:n]))
            assert np.array_equal(back.reshape(-1), d[:n]), s.name       # padded entries still exactly zero
            assert not d[n:].any()


def test_checkpoint_roundtrip(tmp_path):
    m = make_model("lstm", "lstm", 2, g_cell=40, g_proj=24, g_layers=1, d_cell=32)
    rng = np.random.default_rng(0)
    x, y, ln = rng.standard_normal((2, 4, 257)).astype(np.float32), rng.standard_normal((2, 4, 40)).astype(np.float32), np.array([4, 2])
    m.train_batch(x, y, ln)
    m.save(str(tmp_path), 3)
    assert os.path.exists(str(tmp_path / "GAN_RNN-3.pt")) and open(str(tmp_path / "checkpoint")).read().strip() == "GAN_RNN-3.pt"
    m2 = make_model("lstm", "lstm", 2, g_cell=40, g_proj=24, g_layers=1, d_cell=32, seed=99)
    assert m2.load(str(tmp_path))
    for a, b in ((m.G.P, m2.G.P), (m.D.P, m2.D.P)):
        assert torch.equal(a.theta, b.theta) and torch.equal(a.ema, b.ema) and torch.equal(a.hyper, b.hyper)
    assert torch.equal(m.G.P.m, m2.G.P.m)
    nz = np.zeros((2, 1, 40), np.float32)
    assert m.eval_losses(x, y, ln, noise_rl=nz, noise_fk=nz) == m2.eval_losses(x, y, ln, noise_rl=nz, noise_fk=nz)
    assert not make_model("lstm", "lstm", 2, g_cell=40, g_proj=24, g_layers=1, d_cell=32).load(str(tmp_path / "nope"))


def _dp_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z, gp, dp = load_gold("gan_lstm_dlstm")
    # two towers: utterances [0:2] and [1:3] of the fixture (B = 2 per rank)
    sl = slice(rank, rank + 2)
    m = make_model("lstm", "lstm", 2, g_cell=64, g_proj=32, g_layers=2, d_cell=32, num_gpu=world)
    m.load_params(gp, dp)
    assert m.world == world
    m.d_step(z["x"][sl], z["y"][sl], z["lengths"][sl], noise_rl=z["noise_rl"][sl], noise_fk=z["noise_fk"][sl])
    m.g_step(z["x"][sl], z["y"][sl], z["lengths"][sl], noise_fk=z["noise_fk"][sl])
    q.put((rank, m.D.P.export_tf(), m.G.P.export_tf()))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_two_ranks_gloo_matches_two_tower_oracle():
    """world_size 2 over gloo: all-reduce(sum) x 1/N == utils/ops.py:343-376 average_gradients, clip AFTER the mean."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    z, gp, dp = load_gold("gan_lstm_dlstm")
    for k in gp:                                               # both ranks hold identical weights after the update
        assert np.array_equal(res[0][2][k], res[1][2][k]), k
    st = O.GanState(OrderedDict((k, v.astype(np.float64)) for k, v in gp.items()),
                    OrderedDict((k, v.astype(np.float64)) for k, v in dp.items()), "lstm", "lstm")
    f = lambda a: a.astype(np.float64)
    towers = [dict(x=f(z["x"][s]), y=f(z["y"][s]), lengths=z["lengths"][s], noise_rl=f(z["noise_rl"][s]),
                   noise_fk=f(z["noise_fk"][s])) for s in (slice(0, 2), slice(1, 3))]
    O.d_step(st, towers, 1e-3)
    O.g_step(st, towers, 8e-5)
    for k in dp:
        assert rel(res[0][1][k] - dp[k], st.d[k] - dp[k]) < 3e-2, k
    k = "g_model/fully_connected_1/weights"
    assert rel(res[0][2][k] - gp[k], st.g[k] - gp[k]) < 5e-2


@pytest.mark.parametrize("name,kw", [("mse_dnn", dict(g_type="dnn", g_units=64)), ("mse_rced", dict(g_type="rced"))])
def test_dnn_trainer_wiring_matches_golden(name, kw):
    """DNNTrainer (models/dnn_trainer_single_gpu.py:52-133) through the test double vs the oracle's golden vectors."""
    from rsrgan_b200.dnn_trainer import DNNTrainer
    z, gp, _ = load_gold(name)
    N = z["x"].shape[0]
    a = dict(batch_size=N, g_learning_rate=float(z["lr"]), l2_scale=float(z["l2_scale"]), seed=3, **kw)
    m = DNNTrainer(None, Namespace(**a), ["/gpu:0"], handle=FakeHandle("f16"))
    m.load_params(gp)
    assert m.G.P.n_params() == sum(v.size for v in gp.values())
    g = m.generate(z["x"]).numpy()
    assert g.shape == (N, 40) and rel(g, z["g_out"]) < 3e-3
    ev = m.eval_losses(z["x"], z["y"])
    assert ev["g_mse_loss"] == pytest.approx(float(z["loss/g_mse_loss"]), rel=2e-3)
    assert ev["g_l2_loss"] == pytest.approx(float(z["loss/g_l2_loss"]), rel=1e-3)
    cv = DNNTrainer(None, Namespace(**a), ["/gpu:0"], cross_validation=True, share=m)
    assert cv.eval_losses(z["x"], z["y"])["g_l2_loss"] == 0.0   # dnn_trainer_single_gpu.py:111: l2 only when training
    m.g_learning_rate = 0.0
    out = m.train_step(z["x"], z["y"])
    for k in ("g_mse_loss", "g_l2_loss", "g_loss"):
        assert out[k] == pytest.approx(float(z["loss/" + k]), rel=2e-3), k
    gs = m._gscale(N)
    gg = m.G.P.export_tf("grad")
    for k in gp:
        assert rel(gg[k] / gs, z["ggrad/" + k]) < (1e-1 if kw["g_type"] == "rced" else 5e-2), k   # ReLU-mask flips, see above
    # `steps` Adam updates from the golden start point (no clip_by_norm on this trainer)
    m.load_params(gp)
    m.G.P.m.zero_(); m.G.P.v.zero_(); m.G.P.hyper[4:6] = torch.tensor([0.9, 0.999])
    m.g_learning_rate = float(z["lr"])
    for _ in range(int(z["steps"])):
        m.train_step(z["x"], z["y"])
    ev = m.eval_losses(z["x"], z["y"])
    assert ev["g_mse_loss"] == pytest.approx(float(z["loss_after/g_mse_loss"]), rel=5e-3)
    with pytest.raises(AttributeError):
        m.d_step(z["x"], z["y"], None)


def test_overflow_guard_skips_the_update_and_backs_the_loss_scale_off():
    """fp16 loss scale: a non-finite gradient norm makes both update kernels leave weights, slots, shadows and the Adam
    powers untouched and count the event; GAN_RNN.check_overflow then halves the scale (and re-keys the CUDA graphs)."""
    from argparse import Namespace
    from fake_handle import FakeHandle
    from rsrgan_b200.gan_rnn import GAN_RNN
    a = Namespace(g_type="lstm", d_type="lstm", batch_size=2, g_cell=40, g_proj=24, g_layers=1, d_cell=32, seed=2,
                  init_mse_weight=10.0)
    m = GAN_RNN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 5, 257)).astype(np.float32), rng.standard_normal((2, 5, 40)).astype(np.float32)
    ln = np.array([5, 4])
    m.train_batch(x, y, ln)
    assert m.skipped_updates() == (0, 0) and m.check_overflow() == 0
    before = {k: v.clone() for k, v in (("theta", m.G.P.theta), ("m", m.G.P.m), ("v", m.G.P.v), ("ema", m.G.P.ema))}
    hyper = m.G.P.hyper.clone()
    gs0, key0 = m._gscale(10), m._graph_key(2, 5)
    # poison the generator's gradient path: an inf where the backward pass accumulates
    orig = m.h.seg_sumsq

    def poisoned(grad, gmul, seg_id, n_seg, sumsq):
        orig(grad, gmul, seg_id, n_seg, sumsq)
        if grad.data_ptr() == m.G.P.grad.data_ptr():
            sumsq[1] = float("inf")
    m.h.seg_sumsq = poisoned
    m.g_step(x, y, ln)
    m.h.seg_sumsq = orig
    for k, v in before.items():
        assert torch.equal(getattr(m.G.P, k), v), k
    assert torch.equal(m.G.P.hyper[:6], hyper[:6])
    assert m.skipped_updates() == (1, 0)
    assert m.check_overflow() == 1 and m.scale_backoff == 1 and m.check_overflow() == 0
    assert m._gscale(10) == gs0 / 2 and m._graph_key(2, 5) != key0
    m.g_step(x, y, ln)                                   # and training goes on
    assert not torch.equal(m.G.P.theta, before["theta"])
