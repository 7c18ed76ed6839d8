"""Executes the reference's OWN model-building code (/root/reference/models/*.py, utils/ops.py) over the eager float64
stand-in for TensorFlow in tests/golden/tf_standin.py and writes what it produced to tests/golden/ref_graph_*.npz:
variable names and shapes, generator outputs, discriminator logits, the seven losses of every tower, the raw gradients
of every tower (torch autograd through the reference's graph), the tower-averaged and clipped gradients the reference
hands to apply_gradients, and the weights / EMA shadows after one `d_opt` and one `g_opt`.  tests/test_reference_graph.py
replays the oracle (oracle/rsr_oracle.py) on the same seeded parameters and feeds against these files.

Run in the build container only (needs /root/reference):   python tests/golden/make_reference_graph_golden.py
The parameter sets are not stored (5.8 M values): both sides rebuild them from the case's seed (ref_graph_common.py); the
stand-in refuses any variable the reference graph creates that the oracle's parameter set does not name, with that shape,
and the script fails if the oracle names a variable the reference graph never created.
"""
import io
import os
import sys
from argparse import Namespace
from collections import OrderedDict
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path[:0] = [HERE, ROOT]
import tf_standin                                            # noqa: E402
import ref_graph_common as C                                 # noqa: E402
from oracle import rsr_oracle as O                           # noqa: E402

tf, S = tf_standin.install()
sys.path.insert(0, REF)                                      # `models`, `utils` of the reference
import models.gan_rnn_placeholder as ref_gan                 # noqa: E402
import models.BNLSTMCell as ref_cell                         # noqa: E402
import utils.ops as ref_ops                                  # noqa: E402


class Sess(object):
    graph = None


def named(gv):
    return OrderedDict((v.name[:-2], g.numpy()) for g, v in gv)


def gan_rnn_case(case):
    c, gp, dp, x, y, lengths, noise = C.gan_rnn_setup(case)
    S.reset()
    S.init = dict(gp, **dp)
    S.feeds = {"inputs": x, "labels": y, "lengths": lengths.astype(np.float64)}     # lengths are fed as float32 (:102-104)
    S.noise = [n.copy() for n in noise]
    args = Namespace(keep_prob=1.0, batch_norm=False, batch_size=c["B"], num_gpu=c["towers"], save_dir="/tmp/ref_graph",
                     l2_scale=c["l2_scale"], input_dim=257, output_dim=40, left_context=0, right_context=0,
                     disc_updates=1, gen_updates=2, init_mse_weight=C.MSE_LAMBDA, init_disc_noise_std=C.NOISE_STD,
                     d_learning_rate=C.LR_D, g_learning_rate=C.LR_G, g_type=c["g_type"])
    log = io.StringIO()
    with redirect_stdout(log):
        m = ref_gan.GAN_RNN(Sess(), args, ["gpu:%d" % i for i in range(c["towers"])])
    assert S.init_used == set(S.init), ("the oracle names variables the reference graph never created",
                                        sorted(set(S.init) - S.init_used))
    assert not S.noise, "the reference drew fewer noise tensors than expected"
    out = {"variables": np.array(["%s %s" % (k, list(v.v.shape)) for k, v in S.vars.items() if not k.startswith("__anon__/")])}
    # forward values: tf.summary.histogram calls of build_model_single_gpu (:224-228) in tower order
    hist = [(n, t) for n, t in S.summaries if n in ("d_real", "d_fake", "g_clean")]
    assert len(hist) == 3 * c["towers"]
    for i in range(c["towers"]):
        for n, t in hist[3 * i:3 * i + 3]:
            out["fwd|tower%d/%s|full" % (i, n)] = t.numpy()
    for k in ("d_rl_losses", "d_fk_losses", "d_losses", "g_adv_losses", "g_mse_losses", "g_l2_losses", "g_losses"):
        out["loss|" + k] = np.array([float(tf_standin._raw(t).detach()) for t in getattr(m, k)])
    # compute_gradients calls: per tower d then g (:169-175)
    assert len(S.grad_log) == 2 * c["towers"]
    for i in range(c["towers"]):
        C.pack(out, "grad_d_tower%d" % i, named(S.grad_log[2 * i][1]))
        C.pack(out, "grad_g_tower%d" % i, named(S.grad_log[2 * i + 1][1]))
    # what apply_gradients receives: average_gradients (utils/ops.py:343-376) then clip_by_norm 15 per tensor (:177-182)
    assert len(S.apply_log) == 2
    C.pack(out, "applied_d", named(S.apply_log[0][1]))
    C.pack(out, "applied_g", named(S.apply_log[1][1]))
    raw_norm = max(float(np.sqrt((g ** 2).sum())) for g in named(S.grad_log[0][1]).values())
    ema = S.emas[0]
    m.d_opt()                                                 # sess.run(d_opt): SGD step + EMA of the D variables
    C.pack(out, "theta_d_after_d_opt", OrderedDict((k, S.vars[k].numpy()) for k in dp))
    C.pack(out, "ema_d_after_d_opt", OrderedDict((k, ema.shadow[S.vars[k]].numpy().copy()) for k in dp))
    m.g_opt()                                                 # sess.run(g_opt) (gradients of the graph built above)
    C.pack(out, "theta_g_after_g_opt", OrderedDict((k, S.vars[k].numpy()) for k in gp))
    C.pack(out, "ema_g_after_g_opt", OrderedDict((k, ema.shadow[S.vars[k]].numpy().copy()) for k in gp))
    np.savez_compressed(os.path.join(HERE, "ref_graph_gan_rnn_%s.npz" % case), **out)
    print("%-22s %3d variables, d_loss %s g_loss %s, largest raw D-gradient norm %.1f (clip at 15)"
          % (case, len(out["variables"]), out["loss|d_losses"], out["loss|g_losses"], raw_norm))


def lstm_cell_case():
    """models/BNLSTMCell.py:176-213 -- the reference's own statement of the peephole LSTMP step -- with its three
    batch_norm calls replaced by the identity, over a few steps; against it: the stand-in's LSTMCell (checked here) and the
    oracle's lstmp_fwd (checked in the test).  W_xh / W_hh are the two row blocks of LSTMCell's kernel."""
    rng = np.random.default_rng(7)
    B, T, I, Cc, P = 3, 4, 6, 8, 5
    K = 0.5 * rng.standard_normal((I + P, 4 * Cc))
    b = 0.3 * rng.standard_normal(4 * Cc)
    wf, wi, wo = (0.5 * rng.standard_normal(Cc) for _ in range(3))
    Wp = 0.5 * rng.standard_normal((Cc, P))
    x = rng.standard_normal((B, T, I))
    S.reset()
    S.unknown_time = False
    S.init = {"c/input_kernel": K[:I], "c/state_kernel": K[I:], "c/bias": b, "c/W_F_diag": wf, "c/W_I_diag": wi, "c/W_O_diag": wo,
              "c/projection/kernel": Wp}
    saved = ref_cell.batch_norm
    ref_cell.batch_norm = lambda inputs, name_scope, is_training, **k: inputs
    try:
        cell = ref_cell.BNLSTMCell(Cc, use_peepholes=True, num_proj=P, forget_bias=1.0)
        c_, h_ = tf_standin.TT(torch.zeros(B, Cc, dtype=torch.float64)), tf_standin.TT(torch.zeros(B, P, dtype=torch.float64))
        ms, cs = [], []
        for t in range(T):
            with tf.variable_scope("c", reuse=t > 0):
                h_, (c_, h_) = cell.call(tf_standin.TT(torch.tensor(x[:, t])), (c_, h_))
            ms.append(h_.numpy()); cs.append(c_.numpy())
    finally:
        ref_cell.batch_norm = saved
    assert S.init_used == set(S.init)
    ref_m, ref_c = np.stack(ms, 1), np.stack(cs, 1)
    # the stand-in's LSTMCell (what the reference's generators / discriminator are built from in the cases above)
    S.reset()
    S.unknown_time = False
    S.init = {"rnn/lstm_cell/kernel": K, "rnn/lstm_cell/bias": b, "rnn/lstm_cell/w_f_diag": wf, "rnn/lstm_cell/w_i_diag": wi,
              "rnn/lstm_cell/w_o_diag": wo, "rnn/lstm_cell/projection/kernel": Wp}
    cell = tf_standin.LSTMCell(Cc, use_peepholes=True, num_proj=P, forget_bias=1.0, activation=tf.tanh)
    outs, _ = tf_standin.dynamic_rnn(cell, tf_standin.TT(torch.tensor(x)), sequence_length=None,
                                     initial_state=cell.zero_state(B, None))
    assert np.abs(outs.numpy() - ref_m).max() < 1e-14, "stand-in LSTMCell != models/BNLSTMCell.py without its batch norms"
    np.savez_compressed(os.path.join(HERE, "ref_graph_lstm_cell.npz"), K=K, b=b, w_f=wf, w_i=wi, w_o=wo, Wp=Wp, x=x, m=ref_m, c=ref_c)
    print("lstm cell              BNLSTMCell.call (batch norms -> identity) == stand-in LSTMCell; %d steps stored" % T)


def schedules_case():
    """utils/ops.py:378-391 exponential_decay, called the way scripts/train_gan_rnn_placeholder.py:458-461,525-533 does."""
    rows = []
    for it, n_jobs, n_iters, init, mult in [(0, 2, 20, 1e-3, True), (7, 2, 20, 1e-3, True), (19, 2, 20, 1e-3, True),
                                            (25, 2, 20, 1e-3, True), (3, 1, 10, 8e-5, True), (3, 4, 10, 0.05, False),
                                            (9, 4, 10, 0.05, False), (0, 8, 1, 1e-3, True)]:
        rows.append((it, n_jobs, n_iters, init, float(mult), ref_ops.exponential_decay(it, n_jobs, n_iters, init, mult)))
    np.savez_compressed(os.path.join(HERE, "ref_graph_schedules.npz"), exponential_decay=np.array(rows, np.float64))
    print("schedules              %d exponential_decay calls" % len(rows))


if __name__ == "__main__":
    lstm_cell_case()
    schedules_case()
    for case in C.GAN_RNN_CASES:
        gan_rnn_case(case)
