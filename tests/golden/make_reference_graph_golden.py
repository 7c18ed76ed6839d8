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


def frame_gan_case():
    """models/gan.py: DNN generator on the spliced frame, discriminator_dnn on concat(centre LPS frame, MFCC), LSGAN + MSE +
    REGULARIZATION_LOSSES of g_model, Adam for D and for G, gradients applied unclipped."""
    import models.gan as ref_frame_gan
    c, gp, dp, x, y = C.frame_setup("gan_dnn")
    S.reset()
    S.unknown_time = False
    S.init = dict(gp, **dp)
    args = Namespace(keep_prob=1.0, batch_norm=False, batch_size=c["N"], save_dir="/tmp/ref_graph", l2_scale=c["l2_scale"],
                     input_dim=257, output_dim=40, left_context=C.LEFT, right_context=C.RIGHT, disc_updates=1, gen_updates=2,
                     init_mse_weight=C.MSE_LAMBDA, init_disc_noise_std=C.NOISE_STD, d_learning_rate=c["lr_d"],
                     g_learning_rate=c["lr_g"], g_type="dnn")
    with redirect_stdout(io.StringIO()):
        m = ref_frame_gan.GAN(Sess(), args, ["gpu:0"], tf_standin.TT(torch.tensor(x)), tf_standin.TT(torch.tensor(y)))
    assert S.init_used == set(S.init), sorted(set(S.init) - S.init_used)
    out = {"variables": np.array(["%s %s" % (k, list(v.v.shape)) for k, v in S.vars.items() if not k.startswith("__anon__/")])}
    for n, t in S.summaries:
        if n in ("d_real", "d_fake", "g_clean"):
            out["fwd|%s|full" % n] = t.numpy()
    for k in ("d_rl_losses", "d_fk_losses", "d_losses", "g_adv_losses", "g_mse_losses", "g_l2_losses", "g_losses"):
        out["loss|" + k] = np.array([float(tf_standin._raw(t).detach()) for t in getattr(m, k)])
    assert len(S.grad_log) == 2 and len(S.apply_log) == 2
    C.pack(out, "grad_d", named(S.grad_log[0][1]))
    C.pack(out, "grad_g", named(S.grad_log[1][1]))
    C.pack(out, "applied_d", named(S.apply_log[0][1]))
    C.pack(out, "applied_g", named(S.apply_log[1][1]))
    ema = S.emas[0]
    m.d_opt()
    C.pack(out, "theta_d_after_d_opt", OrderedDict((k, S.vars[k].numpy()) for k in dp))
    C.pack(out, "ema_d_after_d_opt", OrderedDict((k, ema.shadow[S.vars[k]].numpy().copy()) for k in dp))
    m.g_opt()
    C.pack(out, "theta_g_after_g_opt", OrderedDict((k, S.vars[k].numpy()) for k in gp))
    C.pack(out, "ema_g_after_g_opt", OrderedDict((k, ema.shadow[S.vars[k]].numpy().copy()) for k in gp))
    np.savez_compressed(os.path.join(HERE, "ref_graph_frame_gan_dnn.npz"), **out)
    print("frame gan (models/gan.py) %d variables, d_loss %s g_loss %s g_l2 %s" % (len(out["variables"]), out["loss|d_losses"],
                                                                               out["loss|g_losses"], out["loss|g_l2_losses"]))


def dnn_trainer_case():
    """models/dnn_trainer_single_gpu.py: 0.5 * 40 * mse + REGULARIZATION_LOSSES, Adam.minimize on the g_ variables; two steps
    are not possible in an eagerly built graph, so: losses, gradients and the weights after the one `g_opt`."""
    import models.dnn_trainer_single_gpu as ref_trainer
    c, gp, _, x, y = C.frame_setup("dnn_trainer")
    S.reset()
    S.unknown_time = False
    S.init = dict(gp)
    args = Namespace(keep_prob=1.0, batch_norm=False, batch_size=c["N"], save_dir="/tmp/ref_graph", l2_scale=c["l2_scale"],
                     input_dim=257, output_dim=40, left_context=C.LEFT, right_context=C.RIGHT, g_learning_rate=c["lr_g"],
                     g_type="dnn")
    with redirect_stdout(io.StringIO()):
        m = ref_trainer.DNNTrainer(Sess(), args, ["gpu:0"], tf_standin.TT(torch.tensor(x)), tf_standin.TT(torch.tensor(y)))
    assert S.init_used == set(S.init)
    out = {"variables": np.array(["%s %s" % (k, list(v.v.shape)) for k, v in S.vars.items() if not k.startswith("__anon__/")])}
    for k in ("g_mse_losses", "g_l2_losses", "g_losses"):
        out["loss|" + k] = np.array([float(tf_standin._raw(getattr(m, k)).detach())])
    assert len(S.grad_log) == 1 and len(S.apply_log) == 1
    C.pack(out, "grad_g", named(S.grad_log[0][1]))
    m.g_opt()
    C.pack(out, "theta_g_after_g_opt", OrderedDict((k, S.vars[k].numpy()) for k in gp))
    np.savez_compressed(os.path.join(HERE, "ref_graph_dnn_trainer.npz"), **out)
    print("dnn trainer            %d variables, g_mse %s g_l2 %s" % (len(out["variables"]), out["loss|g_mse_losses"], out["loss|g_l2_losses"]))


def rced_case(case):
    """models/rced.py under models/dnn_trainer.py (the multi-tower MSE trainer: average_gradients, Adam, EMA over
    tf.trainable_variables()): nine [splice, w] SAME convolutions on (N, splice, 257, 1), NHWC flatten, linear output."""
    import models.dnn_trainer as ref_mt
    c, gp, x, y = C.rced_setup(case)
    S.reset()
    S.unknown_time = False
    S.init = dict(gp)
    args = Namespace(keep_prob=1.0, batch_norm=False, batch_size=c["N"], save_dir="/tmp/ref_graph", l2_scale=c["l2_scale"],
                     input_dim=257, output_dim=40, left_context=c["ctx"], right_context=c["ctx"], g_learning_rate=c["lr_g"],
                     g_type="rced")
    with redirect_stdout(io.StringIO()):
        m = ref_mt.DNNTrainer(Sess(), args, ["gpu:0"], tf_standin.TT(torch.tensor(x)), tf_standin.TT(torch.tensor(y)))
    assert S.init_used == set(S.init), sorted(set(S.init) - S.init_used)
    out = {"variables": np.array(["%s %s" % (k, list(v.v.shape)) for k, v in S.vars.items() if not k.startswith("__anon__/")])}
    for k in ("g_mse_losses", "g_l2_losses", "g_losses"):
        out["loss|" + k] = np.array([float(tf_standin._raw(t).detach()) for t in getattr(m, k)])
    assert len(S.grad_log) == 1 and len(S.apply_log) == 1
    C.pack(out, "grad_g", named(S.grad_log[0][1]))
    ema = S.emas[0]
    m.g_opt()
    C.pack(out, "theta_g_after_g_opt", OrderedDict((k, S.vars[k].numpy()) for k in gp))
    C.pack(out, "ema_g_after_g_opt", OrderedDict((k, ema.shadow[S.vars[k]].numpy().copy()) for k in gp))
    np.savez_compressed(os.path.join(HERE, "ref_graph_%s.npz" % case), **out)
    print("%-22s %d variables, g_mse %s g_l2 %s" % (case, len(out["variables"]), out["loss|g_mse_losses"], out["loss|g_l2_losses"]))


def vbn_case():
    """utils/bnorm.py:11-69 (virtual batch norm): the reference pass and a live pass, outputs and -- by autograd through the
    reference's own expressions -- the gradients of a weighted sum of the outputs wrt the input, gamma and beta."""
    import utils.bnorm as ref_bnorm
    rng = np.random.default_rng(11)
    B, L, Cn = 3, 5, 4
    x_ref, x = 1.5 * rng.standard_normal((B, L, Cn)) + 0.3, 0.8 * rng.standard_normal((B, L, Cn)) - 0.2
    gamma, beta = 1.0 + 0.2 * rng.standard_normal(Cn), 0.1 * rng.standard_normal(Cn)
    r_ref, r_live = rng.standard_normal((B, L, Cn)), rng.standard_normal((B, L, Cn))
    S.reset()
    S.unknown_time = False
    S.init = {"d_vbn/gamma": gamma, "d_vbn/beta": beta}
    xr = tf_standin.TT(torch.tensor(x_ref, requires_grad=True))
    xl = tf_standin.TT(torch.tensor(x, requires_grad=True))
    vb = ref_bnorm.VBN(xr, "d_vbn")
    with tf.variable_scope(tf.get_variable_scope(), reuse=True):
        live = vb(xl)
    g, b = S.vars["d_vbn/gamma"], S.vars["d_vbn/beta"]
    gr = torch.autograd.grad((vb.reference_output.v * torch.tensor(r_ref)).sum(), [xr.v, g.v, b.v], retain_graph=True)
    # live pass: the reference batch's statistics are constants of this pass (they belong to the reference pass)
    vb.mean, vb.mean_sq = tf_standin.TT(vb.mean.v.detach()), tf_standin.TT(vb.mean_sq.v.detach())
    with tf.variable_scope(tf.get_variable_scope(), reuse=True):
        live = vb(xl)
    gl = torch.autograd.grad((live.v * torch.tensor(r_live)).sum(), [xl.v, g.v, b.v])
    np.savez_compressed(os.path.join(HERE, "ref_graph_vbn.npz"), x_ref=x_ref, x=x, gamma=gamma, beta=beta, r_ref=r_ref, r_live=r_live,
                        out_ref=vb.reference_output.numpy(), out_live=live.numpy(),
                        dx_ref=gr[0].numpy(), dgamma_ref=gr[1].numpy(), dbeta_ref=gr[2].numpy(),
                        dx_live=gl[0].numpy(), dgamma_live=gl[1].numpy(), dbeta_live=gl[2].numpy())
    print("vbn                    utils/bnorm.py reference + live pass, outputs and gradients")


def conv_family_case():
    """utils/ops.py:78-98 downconv (k = 31, stride 2, bias), :138-156 conv1d (k = 31), :277-310 deconv (k = 31, dilation 2,
    bias): outputs and the autograd gradients of a weighted sum of the outputs wrt the input, filter and bias."""
    rng = np.random.default_rng(13)
    out = {}
    B, L, Ci, Co, k = 2, 14, 3, 5, 31
    for name, odd in (("even", 0), ("odd", 1)):
        Lx = L + odd
        x = rng.standard_normal((B, Lx, Ci))
        # downconv
        S.reset(); S.unknown_time = False
        W, b = 0.2 * rng.standard_normal((k, 1, Ci, Co)), 0.1 * rng.standard_normal(Co)
        S.init = {"dc/W": W, "dc/b": b}
        xt = tf_standin.TT(torch.tensor(x, requires_grad=True))
        y = ref_ops.downconv(xt, Co, kwidth=k, pool=2, bias_init=tf.constant_initializer(0.), name="dc")
        r = rng.standard_normal(tuple(y.v.shape))
        g = torch.autograd.grad((y.v * torch.tensor(r)).sum(), [xt.v, S.vars["dc/W"].v, S.vars["dc/b"].v])
        out.update({"down_%s|%s" % (name, kk): v for kk, v in dict(x=x, W=W, b=b, r=r, y=y.numpy(), dx=g[0].numpy(), dW=g[1].numpy(),
                                                                  db=g[2].numpy()).items()})
        # deconv: (B, Lx, Ci) -> (B, 2 Lx, Co)
        S.reset(); S.unknown_time = False
        W, b = 0.2 * rng.standard_normal((k, 1, Co, Ci)), 0.1 * rng.standard_normal(Co)
        S.init = {"de/W": W, "de/b": b}
        xt = tf_standin.TT(torch.tensor(x, requires_grad=True))
        y = ref_ops.deconv(xt, [B, 2 * Lx, Co], kwidth=k, dilation=2, bias_init=0.0, name="de")
        r = rng.standard_normal(tuple(y.v.shape))
        g = torch.autograd.grad((y.v * torch.tensor(r)).sum(), [xt.v, S.vars["de/W"].v, S.vars["de/b"].v])
        out.update({"de_%s|%s" % (name, kk): v for kk, v in dict(x=x, W=W, b=b, r=r, y=y.numpy(), dx=g[0].numpy(), dW=g[1].numpy(),
                                                                db=g[2].numpy()).items()})
    # conv1d k = 31, stride 1
    S.reset(); S.unknown_time = False
    x = rng.standard_normal((B, L, Ci))
    W, b = 0.2 * rng.standard_normal((k, Ci, 1)), np.array([0.3])
    S.init = {"c1/W": W, "c1/b": b}
    xt = tf_standin.TT(torch.tensor(x, requires_grad=True))
    y = ref_ops.conv1d(xt, kwidth=k, num_kernels=1, bias_init=0.3, name="c1")
    r = rng.standard_normal(tuple(y.v.shape))
    g = torch.autograd.grad((y.v * torch.tensor(r)).sum(), [xt.v, S.vars["c1/W"].v, S.vars["c1/b"].v])
    out.update({"conv1d|%s" % kk: v for kk, v in dict(x=x, W=W, b=b, r=r, y=y.numpy(), dx=g[0].numpy(), dW=g[1].numpy(),
                                                      db=g[2].numpy()).items()})
    lr = ref_ops.leakyrelu(tf_standin.TT(torch.tensor(x)))
    out["leakyrelu|x"], out["leakyrelu|y"] = x, lr.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_graph_conv_family.npz"), **out)
    print("conv family            utils/ops.py downconv / deconv (even and odd lengths), conv1d, leakyrelu")


def splice_case():
    """io_funcs/tfrecords_io.py:177-204 splice_feats (slice + one-row SYMMETRIC pads), the loader's context splicing."""
    import io_funcs.tfrecords_io as ref_io
    rng = np.random.default_rng(17)
    out = {}
    for i, (rows, cols, left, right) in enumerate([(9, 4, 5, 5), (6, 3, 2, 0), (6, 3, 0, 3), (7, 2, 1, 1), (12, 257, 5, 5)]):
        feats = rng.standard_normal((rows, cols))
        out["case%d|feats" % i], out["case%d|ctx" % i] = feats, np.array([left, right])
        out["case%d|spliced" % i] = ref_io.splice_feats(tf_standin.TT(torch.tensor(feats)), left, right).numpy()
    np.savez_compressed(os.path.join(HERE, "ref_graph_splice.npz"), **out)
    print("splice                 io_funcs/tfrecords_io.py splice_feats, %d cases" % (i + 1))


def schedule_case():
    """scripts/train_gan_rnn_placeholder.py:48-133 train_one_iteration -- the reference's own loop -- over a queue of three
    minibatches (the second one ragged: skipped): per minibatch one sess.run of d_opt, then two of g_opt, each re-executing
    the reference's graph code on the current variables (tf_standin.Session)."""
    import importlib.util
    import queue
    spec = importlib.util.spec_from_file_location("ref_train_script", os.path.join(REF, "scripts", "train_gan_rnn_placeholder.py"))
    ref_script = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref_script)
    c, gp, dp, batches = C.schedule_setup()
    B = c["B"]
    S.reset()
    S.init = dict(gp, **dp)
    S.noise_fn = lambda shape: np.zeros(shape) if shape == [B, 1, 40] else None        # std 0; the SHAPE is still checked
    args = Namespace(keep_prob=1.0, batch_norm=False, batch_size=B, num_gpu=1, save_dir="/tmp/ref_graph", l2_scale=c["l2_scale"],
                     input_dim=257, output_dim=40, left_context=0, right_context=0, disc_updates=1, gen_updates=2,
                     init_mse_weight=C.MSE_LAMBDA, init_disc_noise_std=0.0, d_learning_rate=C.LR_D, g_learning_rate=C.LR_G,
                     g_type=c["g_type"])

    def rebuild(feeds):
        S.begin_retrace(feeds)
        with redirect_stdout(io.StringIO()):
            return ref_gan.GAN_RNN(sess, args, ["gpu:0"])
    sess = tf_standin.Session(rebuild)
    x0, y0, l0 = batches[0]
    S.feeds = {"inputs": x0, "labels": y0, "lengths": l0.astype(np.float64)}
    with redirect_stdout(io.StringIO()):
        model = ref_gan.GAN_RNN(sess, args, ["gpu:0"])
    assert S.init_used == set(S.init)
    sess.bind(model)
    ref_script.FLAGS = Namespace(num_gpu=1, batch_size=B)
    q = queue.Queue()
    for i, (x, y, ln) in enumerate(batches):
        q.put(["utt%d" % i, x, y, ln.astype(np.float64)])
    means = ref_script.train_one_iteration(sess, model, len(batches), 0, q)
    assert sess.log == [["d_opt", "d_rl_losses", "d_fk_losses", "d_losses"]] + 2 * [["g_opt", "g_adv_losses", "g_mse_losses",
                                                                               "g_l2_losses", "g_losses"]] + \
        [["summaries"]] * 0 + [["d_opt", "d_rl_losses", "d_fk_losses", "d_losses"]] + 2 * [["g_opt", "g_adv_losses", "g_mse_losses",
                                                                                        "g_l2_losses", "g_losses"]], sess.log
    out = {"means": np.array(means, np.float64), "adam_t": np.array(S.persist[("adam", 0)]["t"])}
    ema = S.persist[("ema", 0)]
    C.pack(out, "theta_g", OrderedDict((k, S.vars[k].numpy()) for k in gp))
    C.pack(out, "theta_d", OrderedDict((k, S.vars[k].numpy()) for k in dp))
    C.pack(out, "ema_g", OrderedDict((k, ema[S.vars[k]].numpy().copy()) for k in gp))
    C.pack(out, "ema_d", OrderedDict((k, ema[S.vars[k]].numpy().copy()) for k in dp))
    sess.run([model.d_losses], feed_dict={model.inputs: x0, model.labels: y0, model.lengths: l0.astype(np.float64)})
    out["g_after"] = [t for n, t in S.summaries if n == "g_clean"][0].numpy()
    np.savez_compressed(os.path.join(HERE, "ref_graph_schedule.npz"), **out)
    print("schedule               train_one_iteration over %d queued minibatches (%d sess.run of d_opt / g_opt), mean losses %s"
          % (len(batches), len(sess.log) - 1, np.round(out["means"], 4)))


def update_ops_case():
    """WIRING ONLY (tf_standin.batch_norm): with batch_norm = True, which batch_norm UPDATE_OPS exist in the reference's graphs
    -- one per graph COPY of a batch-normalised layer -- and which of them each optimizer's train op runs."""
    import json
    import models.gan as ref_frame_gan
    import models.dnn_trainer as ref_mt
    import models.dnn_trainer_single_gpu as ref_st
    rng = np.random.default_rng(23)
    res = OrderedDict()

    def record(tag, opt_names):
        ops = [o.name for o in S.collections.get(tf.GraphKeys.UPDATE_OPS, [])]
        entry = OrderedDict(update_ops=ops)
        for (opt, _), nm in zip(S.grad_log, opt_names):
            entry[nm + "_runs"] = [o.name for o in opt.deps]
        res[tag] = entry

    base = dict(keep_prob=1.0, batch_norm=True, save_dir="/tmp/ref_graph", l2_scale=0.0, input_dim=257, output_dim=40,
                disc_updates=1, gen_updates=2, init_mse_weight=10.0, init_disc_noise_std=0.0, d_learning_rate=1e-3,
                g_learning_rate=8e-5)
    # recurrent GAN: batch_norm sits on the lstm generator's first fully_connected only (models/lstm.py:61-67,84-85)
    S.reset(); S.allow_bn = True
    S.noise_fn = lambda shape: np.zeros(shape)
    B, T = 2, 3
    S.feeds = {"inputs": rng.standard_normal((B, T, 257)), "labels": rng.standard_normal((B, T, 40)), "lengths": np.full(B, float(T))}
    with redirect_stdout(io.StringIO()):
        ref_gan.GAN_RNN(Sess(), Namespace(batch_size=B, num_gpu=1, left_context=0, right_context=0, g_type="lstm", **base), ["gpu:0"])
    record("gan_rnn_placeholder lstm, 1 tower", ["d_opt", "g_opt"])
    # frame-level GAN: batch_norm on every hidden layer of the DNN generator and of discriminator_dnn
    S.reset(); S.allow_bn = True; S.unknown_time = False
    N = 4
    x, y = rng.standard_normal((N, 257 * 11)), rng.standard_normal((N, 40))
    with redirect_stdout(io.StringIO()):
        ref_frame_gan.GAN(Sess(), Namespace(batch_size=N, left_context=5, right_context=5, g_type="dnn", **base), ["gpu:0"],
                          tf_standin.TT(torch.tensor(x)), tf_standin.TT(torch.tensor(y)))
    record("gan (frame level) dnn + discriminator_dnn, 1 tower", ["d_opt", "g_opt"])
    for tag, mod in (("dnn_trainer_single_gpu dnn", ref_st), ("dnn_trainer (multi-tower trainer) dnn, 1 tower", ref_mt)):
        S.reset(); S.allow_bn = True; S.unknown_time = False
        with redirect_stdout(io.StringIO()):
            mod.DNNTrainer(Sess(), Namespace(batch_size=N, left_context=5, right_context=5, g_type="dnn", **base), ["gpu:0"],
                           tf_standin.TT(torch.tensor(x)), tf_standin.TT(torch.tensor(y)))
        record(tag, ["g_opt"])
    with open(os.path.join(HERE, "ref_graph_update_ops.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k, v in res.items():
        print("update ops  %-50s %d in the graph; %s" % (k, len(v["update_ops"]), {n: len(o) for n, o in v.items() if n != "update_ops"}))


def cv_and_infer_case():
    """The two other graphs scripts/train_gan_rnn_placeholder.py builds from models/gan_rnn_placeholder.py: the
    cross-validation model (cross_validation=True: losses only, no l2 term even with l2_scale > 0, discriminator noise still
    applied, :253-258 and discriminator_lstm.py:60) and the decode model (infer=True: generator only, :133-135)."""
    c, gp, dp, x, y, lengths, noise = C.gan_rnn_setup("res_lstm_l_1tower")
    out = {}
    for tag, kw in (("cv", dict(cross_validation=True)), ("infer", dict(cross_validation=True, infer=True))):
        S.reset()
        S.init = dict(gp, **dp) if tag == "cv" else dict(gp)
        S.feeds = {"inputs": x, "labels": y, "lengths": lengths.astype(np.float64)}
        S.noise = [n.copy() for n in noise]
        args = Namespace(keep_prob=0.7, batch_norm=False, batch_size=c["B"], num_gpu=1, save_dir="/tmp/ref_graph", l2_scale=1e-4,
                         input_dim=257, output_dim=40, left_context=0, right_context=0, disc_updates=1, gen_updates=2,
                         init_mse_weight=C.MSE_LAMBDA, init_disc_noise_std=C.NOISE_STD, d_learning_rate=C.LR_D,
                         g_learning_rate=C.LR_G, g_type=c["g_type"])
        with redirect_stdout(io.StringIO()):
            m = ref_gan.GAN_RNN(Sess(), args, ["gpu:0"], **kw)
        assert S.init_used == set(S.init)
        if tag == "cv":
            assert not S.grad_log and not S.apply_log and not hasattr(m, "d_opt")          # no optimizer in this graph
            for k in ("d_rl_losses", "d_fk_losses", "d_losses", "g_adv_losses", "g_mse_losses", "g_l2_losses", "g_losses"):
                out["cv|" + k] = np.array([float(tf_standin._raw(t).detach()) for t in getattr(m, k)])
        else:
            assert len(S.noise) == len(noise)                      # the decode graph has no discriminator: nothing drawn
            out["infer|g_outputs"] = m.g_outputs.numpy()
    np.savez_compressed(os.path.join(HERE, "ref_graph_cv_infer.npz"), **out)
    print("cv / infer             cross-validation losses %s, decode output %s" % (np.round(out["cv|g_losses"], 3), out["infer|g_outputs"].shape))


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
    vbn_case()
    conv_family_case()
    splice_case()
    for case in C.GAN_RNN_CASES:
        gan_rnn_case(case)
    frame_gan_case()
    dnn_trainer_case()
    schedule_case()
    update_ops_case()
    cv_and_infer_case()
    for case in C.RCED_CASES:
        rced_case(case)
