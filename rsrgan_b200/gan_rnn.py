"""GAN_RNN -- host-side mirror of the reference trainer object.

Reference: models/gan_rnn_placeholder.py:63-317 (class GAN_RNN), driven by
scripts/train_gan_rnn_placeholder.py:48-201.  Same constructor signature and attribute names
(`.inputs/.labels/.lengths` feed names, `.d_real/.d_fake`, `.g_learning_rate/.d_learning_rate`,
`.disc_noise_std`, `.mse_lambda`, `.disc_updates/.gen_updates`, `.save/.load`), but instead of
`sess.run([model.d_opt, ...], feed_dict)` the caller invokes

    d_step(inputs, labels, lengths)   ==  sess.run([d_opt, d_rl_losses, d_fk_losses, d_losses])   (train...py:76-82)
    g_step(inputs, labels, lengths)   ==  sess.run([g_opt, g_adv_losses, g_mse_losses, ...])     (train...py:95-101)
    eval_losses(...)                  ==  the two loss-only sess.run of eval_one_iteration          (train...py:154-172)
    generate(inputs, lengths)         ==  sess.run(model.g_outputs)                                 (train...py:277-281)

One process drives ONE GPU; the reference's in-graph towers (gan_rnn_placeholder.py:152-175)
become ranks of torch.distributed and `average_gradients` (utils/ops.py:343-376) becomes one
NCCL all-reduce of the flat gradient buffer per update, with the 1/num_gpu folded into the
clip+optimizer kernel.  All arithmetic runs in librsrgan_sm100.so; there is no CPU path.
"""
from __future__ import annotations

import contextlib
import math
import os
from collections import OrderedDict

import numpy as np
import torch

from . import nets, ops, tf_checkpoint

F32 = torch.float32

LOSS_NAMES = ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_l2_loss", "g_loss")


def _arg(args, name, default):
    return getattr(args, name, default) if args is not None else default


class Model(object):
    """models/gan_rnn_placeholder.py:20-60 -- save / load with the TF Saver conventions
    (files <save_dir>/<name>-<step>, a `checkpoint` index naming the latest, max_to_keep=10).
    Two containers: torch.save of the state dict (`<name>-<step>.pt`, default) and TensorFlow's own checkpoint-V2
    tensor bundle (`<name>-<step>.index` + `.data-00000-of-00001`, `ckpt_format = "tf"`, rsrgan_b200/tf_checkpoint.py);
    `load` takes whichever the directory holds, so a checkpoint directory written by the reference can be resumed or
    decoded here."""

    ckpt_format = "pt"

    def __init__(self, name="BaseModel"):
        self.name = name

    def _ckpt_path(self, save_dir, step):
        return os.path.join(save_dir, "%s-%d.pt" % (self.name, step))

    @staticmethod
    def _kept(save_dir):
        """checkpoint names listed by the `checkpoint` file, oldest first (our one-name-per-line list or TF's
        text-format CheckpointState)."""
        latest, kept = tf_checkpoint.read_checkpoint_state(save_dir)
        if latest is not None:
            kept = [os.path.basename(k) for k in kept] or [os.path.basename(latest)]
            return kept, True
        index = os.path.join(save_dir, "checkpoint")
        if not os.path.exists(index):
            return [], False
        with open(index) as f:
            return [l.strip() for l in f if l.strip()], False

    def save(self, save_dir, step):
        os.makedirs(save_dir, exist_ok=True)
        kept, _ = self._kept(save_dir)
        if self.ckpt_format == "tf":
            base = "%s-%d" % (self.name, step)
            path = os.path.join(save_dir, base)
            tf_checkpoint.write_bundle(path, tf_checkpoint.state_to_tensors(self.state_dict()))
        else:
            path = self._ckpt_path(save_dir, step)
            torch.save(self.state_dict(), path)
            base = os.path.basename(path)
        kept = [k for k in kept if k != base] + [base]
        while len(kept) > 10:                       # tf.train.Saver(max_to_keep=10), :32
            old = kept.pop(0)
            for suffix in ("", ".index", ".data-00000-of-00001"):
                try:
                    os.remove(os.path.join(save_dir, old + suffix))
                except OSError:
                    pass
        if self.ckpt_format == "tf":
            tf_checkpoint.write_checkpoint_state(save_dir, base, kept)
        else:
            with open(os.path.join(save_dir, "checkpoint"), "w") as f:
                f.write("\n".join(kept) + "\n")
        return path

    def load(self, save_dir, model_file=None, moving_average=False):
        if not os.path.exists(save_dir):
            print("[!] Checkpoints path does not exist...")
            return False
        print("[*] Reading checkpoints...")
        if model_file is None:
            kept, _ = self._kept(save_dir)
            if not kept:
                return False
            ckpt_name = kept[-1]
        else:
            ckpt_name = model_file
        path = os.path.join(save_dir, ckpt_name)
        if os.path.exists(path + ".index"):         # a TensorFlow checkpoint-V2 bundle (the reference's own files)
            sd = self.state_dict()
            missing = tf_checkpoint.tensors_to_state(tf_checkpoint.read_bundle(path), sd)
            if missing:
                print("[*] %d optimizer / average variables not in the checkpoint, kept as initialised" % len(missing))
        else:
            sd = torch.load(path, map_location="cpu", weights_only=False)
        self.load_state_dict(sd, moving_average=moving_average)
        print("[*] Read {}".format(ckpt_name))
        return True


class GAN_RNN(Model):
    """Generative Adversarial Network for speech dereverberation (257-d LPS -> 40-d MFCC)."""

    def __init__(self, sess, args, devices, cross_validation=False, infer=False, name="GAN_RNN",
                 handle=None, share=None):
        super(GAN_RNN, self).__init__(name)
        self.sess = sess                                   # unused (no TF session); kept for the call signature
        self.cross_validation = cross_validation
        self.infer = infer
        self.MOVING_AVERAGE_DECAY = 0.9999                 # :69
        self.max_grad_norm = 15                            # :70
        self.keep_prob = 1.0 if cross_validation else _arg(args, "keep_prob", 1.0)
        self.batch_norm = _arg(args, "batch_norm", False)
        # contrib batch_norm(renorm) on the fully_connected layers and tf.nn.dropout behind them are nets.FCBN
        # (csrc/batchnorm.cu); DropoutWrapper on the LSTM generators is nets.Generator._drop_fwd / _drop_bwd.
        # UPDATE_OPS (moving mean / variance and the renorm statistics of every batch-normalised layer): :163-175
        # collects the g_model and the d_model update ops separately and makes d_opt's compute_gradients depend on the
        # d_model ones, g_opt's on the g_model ones.  So a D update assigns D's statistics (both passes, D(labels) and
        # D(G(x))) and leaves G's alone; a G update assigns G's and leaves D's alone.  `bn_update_scope = "all"` is the
        # frame-level GAN of models/gan.py:139-143, whose two optimizers both depend on the WHOLE collection.
        self.update_bn_stats = True
        self.bn_update_scope = "own"
        self.ckpt_format = _arg(args, "ckpt_format", "pt")   # "tf": TensorFlow checkpoint-V2 bundles (tf_checkpoint.py)
        self.batch_size = _arg(args, "batch_size", 8)
        self.devices = devices
        self.num_gpu = _arg(args, "num_gpu", 1)
        self.save_dir = _arg(args, "save_dir", "exp/gan_rnn")
        self.l2_scale = _arg(args, "l2_scale", 0.0)
        self.input_dim = _arg(args, "input_dim", 257)
        self.output_dim = _arg(args, "output_dim", 40)
        self.left_context = _arg(args, "left_context", 0)
        self.right_context = _arg(args, "right_context", 0)
        self.inputs, self.labels, self.lengths = "inputs", "labels", "lengths"   # feed names (:94-104)
        self.g_disturb_weights = False
        self.d_clip_weights = False
        self.disc_updates = _arg(args, "disc_updates", 1)
        self.gen_updates = _arg(args, "gen_updates", 2)
        self.mse_lambda = float(_arg(args, "init_mse_weight", 1.0))
        self.disc_noise_std = float(_arg(args, "init_disc_noise_std", 0.0))
        self.d_real, self.d_fake = 1.0, 0.0                # :122-123
        self.g_type = _arg(args, "g_type", "lstm")
        self.d_type = _arg(args, "d_type", "lstm")        # reference hard-wires discriminator_lstm (:117)

        self.world = 1
        self.dist = None
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.dist = torch.distributed
            self.world = torch.distributed.get_world_size()
        # the discriminator's input noise: one counter-based stream per rank (every tower of the reference owns its
        # random_normal op), state {seed, draws so far} on the device
        self._noise_seed = int(_arg(args, "seed", 1234)) * 7919 + 17 + (self.dist.get_rank() if self.dist else 0)
        self._noise_rng = None

        if share is not None:
            # train and CV models share weights (scripts/train_gan_rnn_placeholder.py:431-437) but own their
            # scalars
            self.h, self.G, self.D = share.h, share.G, share.D
        else:
            dev = 0
            if devices:
                d0 = str(devices[0])
                dev = int(d0.split(":")[-1]) if ":" in d0 else 0
            self.h = handle if handle is not None else ops.Handle(dev, _arg(args, "dtype", "f16"))
            if _arg(args, "dtype", "f16") == "bf16":
                # fp16 operands are the product dtype: 8 more mantissa bits keep the generator output within 1e-3 RMS of the
                # fp32 reference at every benchmarked length; bf16 does not (profiles/r2_parity_measured_v0.jsonl)
                import warnings
                warnings.warn("rsrgan_b200: bf16 operands deviate from the fp32 reference by up to 3.5e-3 RMS in the "
                              "generator output (4 x 1024 res_lstm_l, T = 200) -- outside the 1e-3 parity bar that "
                              "dtype='f16' (the default) meets", stacklevel=2)
            in_dim = self.input_dim * (self.left_context + 1 + self.right_context)
            # models/dnn.py:64-68: the dnn generator drops out only when it is also regularised (else keep_prob = 1)
            g_keep = self.keep_prob
            if self.g_type == "dnn" and not _arg(args, "l2_scale", 0.0) > 0.0:
                g_keep = 1.0
            gk = dict(in_dim=in_dim, out_dim=self.output_dim, batch_norm=bool(self.batch_norm), keep_prob=g_keep)
            for k_arg, k in (("g_cell", "cell"), ("g_proj", "proj"), ("g_layers", "layers"), ("g_units", "units")):
                v = _arg(args, k_arg, None)
                if v is not None:
                    gk[k] = v
            if self.g_type in ("res_lstm_l", "res_lstm_base"):
                gk.pop("proj", None)
                gk.pop("units", None)
            if self.g_type == "rced":              # models/rced.py:46-57: (batch, splice, input_dim, 1) frames
                gk["splice"] = self.left_context + 1 + self.right_context
            self.G = nets.Generator(self.h, self.g_type, **gk)
            self.D = None
            if not infer:
                # d_cat_dim / d_adam: the frame-level GAN of models/gan.py (rsrgan_b200/gan.py) conditions
                # discriminator_dnn on the centre LPS frame and trains it with Adam
                dk = dict(in_dim=self.output_dim, batch_norm=bool(self.batch_norm), keep_prob=self.keep_prob,
                          cat_dim=_arg(args, "d_cat_dim", 0), adam=bool(_arg(args, "d_adam", False)))
                for k_arg, k in (("d_cell", "cell"), ("d_proj", "proj"), ("d_layers", "layers"), ("d_units", "units")):
                    v = _arg(args, k_arg, None)
                    if v is not None:
                        dk[k] = v
                self.D = nets.Discriminator(self.h, self.d_type, **dk)
            self.init_weights(_arg(args, "seed", 1234))
            for net in (self.G, self.D):
                if net is not None:
                    net.rng[0] = int(_arg(args, "seed", 1234))
        # average_gradients (utils/ops.py:343-376) over NVLink peer memory: the gradient stores move into a block the
        # other ranks of the node have mapped, and the all-reduce is one kernel of the schedule (rsrgan_b200/peer.py);
        # None -> torch.distributed.all_reduce (CPU test double, other world sizes, no peer access)
        self.peer = share.peer if share is not None else None
        if share is None and self.world > 1 and not cross_validation:
            from . import peer
            stores = [n.P for n in (self.G, self.D) if n is not None]
            self.peer = peer.try_create(self.h, self.dist, [P.n for P in stores])
            if self.peer is not None:
                for i, P in enumerate(stores):
                    P.grad = self.peer.buffer(i)
        self.g_learning_rate = float(_arg(args, "g_learning_rate", 0.0003))
        self.d_learning_rate = float(_arg(args, "d_learning_rate", 0.001))
        dev = self.h.device
        self._losses = torch.zeros(8, dtype=F32, device=dev)
        self._l2 = torch.zeros(1, dtype=F32, device=dev)
        self._pin = {}
        # CUDA graph of the whole batch schedule (one launch instead of ~150): single-GPU training only;
        # RSR_NO_GRAPH=1 or use_graph=False runs every kernel eagerly
        self.use_graph = (_arg(args, "use_graph", True) and os.environ.get("RSR_NO_GRAPH", "0") != "1"
                          and dev.type == "cuda")
        self.scale_backoff, self._skipped_seen = 0, 0       # fp16 loss-scale back-off (check_overflow)
        self._cmvn = None                                   # device CMVN of the fed minibatches (set_cmvn)
        self._graphs, self._graphs_gen = {}, None
        self._copy_stream, self._prefetched, self._prefetch_bufs = None, None, {}
        # With several ranks the schedule is captured as one graph SEGMENT per update; the NCCL all-reduce of the
        # flat gradient buffer runs eagerly between segments (capturing NCCL itself hung on 2 x B200 with
        # torch 2.11 / NCCL 2.28.9).  RSR_GRAPH_DDP=0: fully eager with several ranks.
        self.graph_ddp = os.environ.get("RSR_GRAPH_DDP", "1") != "0"
        # RSR_GRAPH_NCCL=1: capture the all-reduces INTO the graph (one graph per schedule, no segment cuts)
        self.graph_nccl = os.environ.get("RSR_GRAPH_NCCL", "0") == "1"
        self._cap = None
        self.g_outputs = None
        # tf.summary.FileWriter(save_dir/train | eval) (:82-86); `summaries` = the tags of the merged scalar summary
        # (:270-298).  Created on first use by write_summaries(); rank 0 only.
        self.summaries = ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss", "g_l2_loss", "g_loss")
        self.writer = None

    # ------------------------------------------------------------------ scalars on device
    @property
    def g_learning_rate(self):
        return self._g_lr

    @g_learning_rate.setter
    def g_learning_rate(self, v):
        self._g_lr = float(v)
        if not self.cross_validation:
            self.G.P.set_lr(v)

    @property
    def d_learning_rate(self):
        return self._d_lr

    @d_learning_rate.setter
    def d_learning_rate(self, v):
        self._d_lr = float(v)
        if self.D is not None and not self.cross_validation:
            self.D.P.set_lr(v)

    # ------------------------------------------------------------------ weights
    def init_weights(self, seed=1234):
        """xavier_initializer() / zeros / truncated-normal as the reference builds its variables
        (models/lstm.py:86,92; models/discriminator_dnn.py:26-27).  Host numpy RNG, then one upload."""
        rng = np.random.default_rng(seed)

        def init(net):
            p = OrderedDict()
            for s in net.P.segs.values():
                if s.name.endswith("/BatchNorm/gamma") or s.name.endswith("/BatchNorm/beta"):
                    p[s.name] = np.full(s.tf_shape, 1.0 if s.name.endswith("gamma") else 0.0, np.float32)
                elif "bias" in s.name:
                    # zeros everywhere except the RCED output layer (models/rced.py:112 constant_initializer(0.1))
                    rced_out = self.g_type == "rced" and s.name == "g_model/fully_connected/biases"
                    p[s.name] = np.full(s.tf_shape, 0.1 if rced_out else 0.0, np.float32)
                elif self.d_type == "dnn" and s.name.startswith("d_model") and s.tf_shape[-1] != 1:
                    std = math.sqrt(2.0 / s.tf_shape[1])
                    v = rng.standard_normal(s.tf_shape)
                    bad = np.abs(v) > 2
                    while bad.any():
                        v[bad] = rng.standard_normal(int(bad.sum()))
                        bad = np.abs(v) > 2
                    p[s.name] = (v * std).astype(np.float32)
                else:
                    if len(s.tf_shape) == 4:               # conv filter (h, w, C_in, C_out)
                        rf = s.tf_shape[0] * s.tf_shape[1]
                        fi, fo = rf * s.tf_shape[2], rf * s.tf_shape[3]
                    else:
                        fi, fo = (s.tf_shape[0], s.tf_shape[0]) if len(s.tf_shape) == 1 else s.tf_shape
                    lim = math.sqrt(6.0 / (fi + fo))
                    p[s.name] = rng.uniform(-lim, lim, s.tf_shape).astype(np.float32)
            net.load_tf(p)

        init(self.G)
        if self.D is not None:
            init(self.D)

    def load_params(self, g_params=None, d_params=None):
        """dicts TF-variable-name -> array (TF layout)."""
        if g_params is not None:
            self.G.load_tf(g_params)
        if d_params is not None:
            self.D.load_tf(d_params)

    def state_dict(self):
        sd = OrderedDict(G=self.G.P.state_dict())
        if self.D is not None:
            sd["D"] = self.D.P.state_dict()
        for net, key in ((self.G, "G"), (self.D, "D")):
            if net is not None and net.has_bn_state:     # non-trainable batch_norm variables + dropout stream
                sd[key]["bn_state"] = net.bn_state_tf()
                sd[key]["rng"] = net.rng.cpu().numpy()
        sd["scalars"] = dict(mse_lambda=self.mse_lambda, disc_noise_std=self.disc_noise_std,
                             d_learning_rate=self.d_learning_rate, g_learning_rate=self.g_learning_rate,
                             d_real=self.d_real, d_fake=self.d_fake)
        return sd

    def load_state_dict(self, sd, moving_average=False):
        for net, key in ((self.G, "G"), (self.D, "D")):
            if net is None or key not in sd:
                continue
            net.P.load_state_dict(sd[key])
            if "bn_state" in sd[key]:
                net.load_bn_state_tf(sd[key]["bn_state"])
                net.rng.copy_(torch.as_tensor(np.asarray(sd[key]["rng"], np.int64)))
            if moving_average:                              # :47-53 restore the EMA shadows as the weights
                net.P.theta.copy_(net.P.ema)
            net.P.refresh16()
            net.refresh()
        for k, v in sd.get("scalars", {}).items():
            setattr(self, k, v)

    # ------------------------------------------------------------------ feeding
    def _to_dev(self, name, a, dtype):
        """host numpy / torch -> device tensor through a pinned staging buffer (async H2D)."""
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.to(dtype).contiguous()
        t = torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a.contiguous()
        t = t.to(dtype)
        if self.h.device.type != "cuda":                   # host-logic tests drive a test double on CPU tensors
            return t.clone()
        key = (name, tuple(t.shape), dtype)
        if key not in self._pin:
            self._pin[key] = (torch.empty(t.shape, dtype=dtype, pin_memory=True),
                              torch.empty(t.shape, dtype=dtype, device=self.h.device))
        pin, dev = self._pin[key]
        if t.is_pinned():                                  # caller already staged the batch in pinned memory
            dev.copy_(t, non_blocking=True)
        else:
            pin.copy_(t)
            dev.copy_(pin, non_blocking=True)
        return dev

    def set_cmvn(self, cmvn):
        """Global CMVN of the loader (io_funcs/make_tfrecords.py:84-87), applied ON THE DEVICE to every fed minibatch
        instead of per utterance on the host: `cmvn` holds mean_inputs / stddev_inputs / mean_labels / stddev_labels
        (train_cmvn.npz, io_funcs/convert_cmvn_to_numpy.py:43-47); inputs spliced with left / right context see the
        statistics tiled per context frame.  The feeds must then be RAW features, zero-padded (dataset.get_padded_batch
        with cmvn_on_device=True).  None switches it off."""
        if cmvn is None:
            self._cmvn = None
            return
        dev, ctx = self.h.device, self.left_context + 1 + self.right_context
        t64 = lambda k, rep: torch.as_tensor(np.tile(np.asarray(cmvn[k], np.float64).reshape(-1), rep), device=dev)
        self._cmvn = dict(mx=t64("mean_inputs", ctx), sx=t64("stddev_inputs", ctx),
                          my=t64("mean_labels", 1), sy=t64("stddev_labels", 1))

    def _normalise(self, name, a, ln, mean, std):
        B, T, D = a.shape
        out = self.G.ws.get(("feed", name, B), B * T, D, F32).view(B, T, D)
        self.h.cmvn_apply_padded(a, ln, mean, std, out)
        return out

    def _feed(self, inputs, labels, lengths):
        x = self._to_dev("x", inputs, F32)
        B, T = int(x.shape[0]), int(x.shape[1])
        # lengths are fed as float32 and cast to int32 (gan_rnn_placeholder.py:102-104)
        ln = self._to_dev("len", lengths, torch.int32)
        if self._cmvn is not None:
            x = self._normalise("x_cmvn", x, ln, self._cmvn["mx"], self._cmvn["sx"])
        y_tm = None
        if labels is not None:
            y = self._to_dev("y", labels, F32)
            if self._cmvn is not None:
                y = self._normalise("y_cmvn", y, ln, self._cmvn["my"], self._cmvn["sy"])
            y_tm = self.G.ws.get(("feed", "y_tm", B), T * B, self.output_dim, F32)
            self.h.stage_input(y, B, T, self.output_dim, out32=y_tm)
        return x, y_tm, ln, B, T

    def _cat(self, x):
        """Conditioning block of the discriminator input (models/gan.py:159-160: tf.slice(inputs, [0, input_dim *
        left_context], [-1, input_dim]) -- the centre frame of the spliced LPS input), as a view of the fed batch."""
        if self.D is None or not self.D.cat_dim:
            return None
        c0 = self.input_dim * self.left_context
        return x[:, :, c0:c0 + self.D.cat_dim]

    def _noise(self, B, given, slot="rl"):
        if self.d_type != "lstm":
            return None                                    # discriminator_dnn has no noise layer
        if given is not None:
            return self._to_dev("noise_" + slot, np.asarray(given, np.float32).reshape(B, -1), F32)
        if self.disc_noise_std <= 0.0:
            return None
        # utils/ops.py:19-30: tf.random_normal of shape (B, 1, D) -- one draw per utterance, from the library's
        # counter-based stream (the tick advances with every draw, on the device: CUDA-graph safe)
        if self._noise_rng is None:
            self._noise_rng = torch.tensor([self._noise_seed, 0], dtype=torch.int64, device=self.h.device)
        out = self.D.ws.get(("noise", slot), B, self.output_dim, F32)
        self.h.gauss_noise(self._noise_rng, 0x4e01 if slot == "rl" else 0x4e02, out, self.disc_noise_std)
        self.h.rng_tick(self._noise_rng)
        return out

    def _gscale(self, rows):
        """Loss scale keeping the 16-bit gradient tensors of the backward pass in range (fp16 operands): the gradients
        of a mean over `rows` frames are O(1 / rows), so they are multiplied by ~rows / lambda (a power of two) on the
        way down and divided again inside the update sweep.  `scale_backoff` halves it once per overflow the update
        kernels reported (check_overflow)."""
        if self.h.dtype_id == ops._lib.RSR_DTYPE_BF16:
            return 1.0
        return float(2.0 ** (round(math.log2(max(rows / max(self.mse_lambda, 1.0), 1.0))) - self.scale_backoff))

    def skipped_updates(self):
        """(G, D) numbers of updates the overflow guard skipped so far (device counters hyper[7]; synchronises)."""
        return tuple(int(n.P.hyper[7].item()) if n is not None else 0 for n in (self.G, self.D))

    def check_overflow(self):
        """fp16 overflow handling, host half: when the update kernels skipped an update since the last call (a gradient
        norm was inf / NaN: include/rsrgan_b200.h, rsr_clip_adam_ema) the loss scale is halved for the following
        batches -- skip-and-back-off as in mixed-precision training practice; the fp32 reference has no counterpart
        because it cannot overflow there.  Returns the number of newly skipped updates.  Synchronises; the trainer
        calls it where it reads the losses anyway."""
        if self.peer is not None and self.peer.error():
            raise RuntimeError("peer all-reduce: a barrier timed out (a rank died or issued a different call sequence)")
        if self.h.dtype_id == ops._lib.RSR_DTYPE_BF16:
            return 0
        now = sum(self.skipped_updates())
        new = now - self._skipped_seen
        if new > 0:
            self._skipped_seen = now
            self.scale_backoff += 1
            print("[!] %d update(s) skipped: non-finite 16-bit gradients; loss scale halved (back-off %d)"
                  % (new, self.scale_backoff))
        return new

    # ------------------------------------------------------------------ updates
    def _mode(self, training, g_update=False, d_update=False):
        """is_training of the graph about to run (batch_norm statistics, dropout): False on the cross-validation /
        inference models (models/dnn.py:50, models/discriminator_dnn.py:30).  g_update / d_update: which network's
        UPDATE_OPS the optimizer op about to run depends on (see __init__)."""
        every = self.bn_update_scope == "all" and (g_update or d_update)
        for net, upd in ((self.G, g_update), (self.D, d_update)):
            if net is not None:
                net.training = bool(training) and not self.cross_validation
                net.bn_update = bool(self.update_bn_stats) and net.training and (upd or every)

    def _update(self, net, gscale, adam):
        P, h = net.P, self.h
        h.join()                                           # weight gradients computed on the side stream
        for n in (self.G, self.D):                         # next update draws new dropout masks in both networks; only
            if n is not None:                              # after the join: backward passes on the side stream regenerate
                n.tick()                                   # their masks from the current tick
        if self.world > 1:
            # utils/ops.py:343-376 average_gradients: sum over ranks here, 1/N folded into the update kernel
            if self.peer is not None:
                self.peer.all_reduce(P.grad)     # one kernel over peer memory; part of the captured schedule
            elif self._cap is not None and not self.graph_nccl:
                # graph capture in progress: the collective stays OUTSIDE the graphs -- close the segment, run the
                # all-reduce eagerly (keeps every rank's collective sequence aligned), open the next segment
                self._cap_end(P.grad)
                self.dist.all_reduce(P.grad)
                self._cap_begin()
            else:
                self.dist.all_reduce(P.grad)     # eager, or captured into the one graph of the schedule (graph_nccl)
        gmul = 1.0 / (self.world * gscale)
        h.seg_sumsq(P.grad, gmul, P.seg_id, len(P.segs), P.sumsq)
        if adam:
            h.clip_adam_ema(P.grad, gmul, P.seg_id, P.sumsq, float(self.max_grad_norm), P.hyper,
                            self.MOVING_AVERAGE_DECAY, P.theta, P.m, P.v, P.ema, P.theta16, n_seg=len(P.segs))
        else:
            h.clip_sgd_ema(P.grad, gmul, P.seg_id, P.sumsq, float(self.max_grad_norm), P.hyper,
                           self.MOVING_AVERAGE_DECAY, P.theta, P.ema, P.theta16, n_seg=len(P.segs))
        net.refresh()

    def _loss_dict(self, vals, which):
        d_rl, d_fk, g_adv, g_mse, g_l2 = (float(v) for v in vals[:5])
        out = OrderedDict()
        if which in ("d", "both"):
            out.update(d_rl_loss=d_rl, d_fk_loss=d_fk, d_loss=d_rl + d_fk)
        if which in ("g", "both"):
            out.update(g_adv_loss=g_adv, g_mse_loss=g_mse, g_l2_loss=g_l2,
                       g_loss=g_adv + self.mse_lambda * g_mse + g_l2)
        return out

    def _l2_loss(self):
        """l2_scale * sum_{non-bias} 0.5 ||v||^2   (gan_rnn_placeholder.py:253-258)"""
        if self.l2_scale <= 0.0 or self.cross_validation:
            self.h.fill32(self._losses[4:5], 0.0)
            return
        P = self.G.P
        self.h.seg_sumsq(P.theta, 1.0, P.seg_id, len(P.segs), P.sumsq)
        self._losses[4:5] = 0.5 * self.l2_scale * (P.sumsq * P.seg_l2.to(F32)).sum()

    def d_step(self, inputs, labels, lengths, noise_rl=None, noise_fk=None, sync=True, _feed=None, _g32=None,
               _g_train=False, _g_reuse=False):
        """One discriminator update: L_D = mean((D(y)-d_real)^2) + mean((D(G(x))-d_fake)^2),
        gradients wrt theta_D only, tower mean, per-tensor clip 15, SGD(lr_d), EMA."""
        x, y_tm, ln, B, T = _feed if _feed is not None else self._feed(inputs, labels, lengths)
        h, G, D, rows = self.h, self.G, self.D, T * B
        # _g_reuse: the generator forward computed here is the one the first G update of the schedule reuses, so it is
        # the forward that G update's UPDATE_OPS belong to
        self._mode(True, g_update=_g_reuse and _g32 is None, d_update=True)
        gs = self._gscale(rows)
        d_rl16 = D.ws.get(("loss", "d_rl16"), rows, 8, h.h16)
        d_fk16 = D.ws.get(("loss", "d_fk16"), rows, 8, h.h16)
        n_rl, n_fk = self._noise(B, noise_rl), self._noise(B, noise_fk, "fk")
        h.fill32(self._losses, 0.0)
        h.fill32(D.P.grad, 0.0)
        kw = dict(n_logit=rows, clip=D.clip, d_real=self.d_real, d_fake=self.d_fake, lam=self.mse_lambda, gscale=gs,
                  ld_grad=8)
        # D(labels) does not depend on the generator: its forward, loss and backward run on the side stream
        # while the generator recurrences (which occupy only the SMs of their clusters) run on this one.
        # (a batch-normalised D assigns its moving averages in both passes: those two then stay on one stream,
        # D(labels) first, so that neither read-modify-write is lost)
        serial = D.fcbn and D.bn_update
        # (D.fwd(head=...): where the discriminator's head qualifies, logits, loss terms, d loss / d logit and the head's
        #  data gradient come out of one kernel -- rsr_fc1_head -- and only the MSE term is left for rsr_lsgan_mse_losses)
        hd = dict(clip=D.clip, d_real=self.d_real, d_fake=self.d_fake, gscale=gs, losses=self._losses)
        with (contextlib.nullcontext() if serial else h.side_stream()):
            lg_rl = D.fwd("rl", y_tm, B, T, ln, noise=n_rl, cat_src=self._cat(x),
                          head=dict(hd, which=0, grad_target=self.d_real, dlogit16=d_rl16))
            if not D.head_fused["rl"]:
                h.lsgan_mse_losses(self._losses, rl=lg_rl, ld_logit=lg_rl.stride(0), d_rl_grad=d_rl16, **kw)
            D.bwd("rl", d_rl16)
        g32 = _g32 if _g32 is not None else G.fwd(x, B, T, ln, train=_g_train)
        self._last_g32 = g32
        lg_fk = D.fwd("fk", g32, B, T, ln, noise=n_fk, cat_src=self._cat(x),
                      head=dict(hd, which=1, grad_target=self.d_fake, dlogit16=d_fk16))
        if D.head_fused["fk"]:
            mse = dict(kw, n_logit=0)
            h.lsgan_mse_losses(self._losses, g=g32, y=y_tm, n_frames=rows, d_out=self.output_dim, **mse)
        else:
            h.lsgan_mse_losses(self._losses, fk=lg_fk, ld_logit=lg_fk.stride(0), g=g32, y=y_tm, n_frames=rows,
                               d_out=self.output_dim, d_fk_grad=d_fk16, **kw)
        D.bwd("fk", d_fk16)
        self._update(D, gs, adam=D.P.adam)
        return self._loss_dict(self._losses.tolist(), "d") if sync else self._losses

    def g_step(self, inputs, labels, lengths, noise_fk=None, sync=True, _feed=None, _g32=None, _x_staged=False):
        """One generator update: L_G = mean((D(G(x))-d_real)^2) + lambda*0.5*40*mean((G(x)-y)^2) [+ l2],
        gradients wrt theta_G only (through D, D frozen), tower mean, clip 15, Adam(lr_g), EMA."""
        x, y_tm, ln, B, T = _feed if _feed is not None else self._feed(inputs, labels, lengths)
        h, G, D, rows = self.h, self.G, self.D, T * B
        self._mode(True, g_update=_g32 is None, d_update=False)
        gs = self._gscale(rows)
        g32 = _g32 if _g32 is not None else G.fwd(x, B, T, ln, train=True, reuse_staged=_x_staged)
        if D.fcbn and D.bn_update:
            # frame-level GAN: g_opt depends on the whole UPDATE_OPS collection (models/gan.py:139-143), which holds the
            # assignments of the D(labels) pass too -- run that pass for its statistics (its output is not needed)
            D.fwd("rl", y_tm, B, T, ln, noise=self._noise(B, None), cat_src=self._cat(x))
        g_adv16 = D.ws.get(("loss", "g_adv16"), rows, 8, h.h16)
        # d(lambda g_mse)/dg is added to the discriminator's input gradient by its last GEMM (resid): same width as
        # that input -- the generator's columns come first, the conditioning columns of a conditioned D stay zero
        dw = g32.shape[1] if not D.cat_dim else (D.in_dim + D.cat_dim + 7) // 8 * 8
        dg32 = D.ws.get(("loss", "dg32"), rows, dw, F32)
        h.fill32(self._losses, 0.0)
        lg_fk = D.fwd("fk", g32, B, T, ln, noise=self._noise(B, noise_fk, "fk"), cat_src=self._cat(x),
                      head=dict(clip=D.clip, d_real=self.d_real, d_fake=self.d_fake, gscale=gs, losses=self._losses,
                                which=1, grad_target=self.d_real, dlogit16=g_adv16))
        if D.head_fused["fk"]:
            # only the MSE term and its gradient are left; nothing reads them before the discriminator's LAST data-gradient
            # GEMM (resid), so they leave the critical chain: side stream, joined in front of that GEMM
            with h.side_stream():
                h.lsgan_mse_losses(self._losses, n_logit=0, clip=D.clip, g=g32, y=y_tm, n_frames=rows, d_out=self.output_dim,
                                   d_real=self.d_real, d_fake=self.d_fake, lam=self.mse_lambda, gscale=gs, ld_grad=8,
                                   dg_mse=dg32)
        else:
            h.lsgan_mse_losses(self._losses, fk=lg_fk, ld_logit=lg_fk.stride(0), n_logit=rows, clip=D.clip,
                               g=g32, y=y_tm, n_frames=rows, d_out=self.output_dim, d_real=self.d_real,
                               d_fake=self.d_fake, lam=self.mse_lambda, gscale=gs, g_adv_grad=g_adv16,
                               ld_grad=8, dg_mse=dg32)
        dg16 = D.bwd("fk", g_adv16, want_dw=False, want_dx=True, resid32=dg32, pre_last=h.join)
        h.fill32(G.P.grad, 0.0)
        G.bwd(dg16[:, :g32.shape[1]] if D.cat_dim else dg16)
        self._l2_loss()
        if self.l2_scale > 0.0:
            h.l2_grad(G.P.grad, G.P.theta, G.P.seg_id, G.P.seg_l2, self.l2_scale * gs)
        self._update(G, gs, adam=True)
        return self._loss_dict(self._losses.tolist(), "g") if sync else self._losses

    def _schedule(self, feed):
        """disc_updates x D update then gen_updates x G update on one fed minibatch (device work only)."""
        x, y_tm, ln, B, T = feed
        # The reference recomputes G(x) in every sess.run; the generator weights do not
        # change until the first G update, so the D updates and that first G update all see the SAME G(x):
        # it is computed once (with the activations the G backward needs) and reused -- same numbers, 2 of the
        # 3 generator forwards of the schedule.
        g32 = None
        d_all, g_all = [], []
        share = self.G.keep_prob >= 1.0                    # a generator with dropout draws a new mask per sess.run
        for _ in range(self.disc_updates):
            d = self.d_step(None, None, None, sync=False, _feed=feed, _g32=g32, _g_train=self.gen_updates > 0,
                            _g_reuse=share and self.gen_updates > 0)
            g32 = self._last_g32 if share else None
            d_all.append(d[:2].clone())
        for k in range(self.gen_updates):
            # the generator has already staged this minibatch (16-bit, time-major) in an earlier forward of the schedule
            g = self.g_step(None, None, None, sync=False, _feed=feed, _g32=g32 if k == 0 else None,
                            _x_staged=k > 0 or g32 is not None)
            g_all.append(g.clone())
        return d_all, g_all

    def _graph_key(self, B, T):
        # everything a captured kernel receives BY VALUE; learning rates and Adam powers live on the device
        return (B, T, self.disc_updates, self.gen_updates, self.d_real, self.d_fake, self.mse_lambda,
                self.disc_noise_std, self.l2_scale, self.world, self.scale_backoff)

    def prefetch(self, inputs, labels, lengths):
        """Starts the host->device copy of the NEXT minibatch on a copy stream, so that it overlaps the batch
        schedule still running (the reference fills a queue from reader threads for the same reason,
        scripts/train_gan_rnn_placeholder.py:469-478).  A later train_batch() called with the same objects picks
        the device copies up instead of copying again.  Host tensors should be pinned."""
        if self.h.device.type != "cuda":
            return
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.h.device)
        srcs = []
        for a, dt in ((inputs, F32), (labels, F32), (lengths, torch.int32)):
            t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
            srcs.append((t, dt))
        key = tuple((tuple(t.shape), dt) for t, dt in srcs)
        if self._prefetch_bufs.get("key") != key:           # two buffer sets, used alternately
            self._prefetch_bufs = {"key": key, "sets": [[torch.empty(t.shape, dtype=dt, device=self.h.device)
                                                         for t, dt in srcs] for _ in range(2)],
                                   "free": [None, None], "next": 0}
        pb = self._prefetch_bufs
        k = pb["next"]
        pb["next"] = k ^ 1
        cs = self._copy_stream
        if pb["free"][k] is not None:
            cs.wait_event(pb["free"][k])                    # the schedule that read this set (two batches ago) has finished
        with torch.cuda.stream(cs):
            for (t, dt), d in zip(srcs, pb["sets"][k]):
                d.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cs)
        self._prefetched = ((id(inputs), id(labels), id(lengths)), k, ev)

    def _take_prefetched(self, inputs, labels, lengths):
        pf = self._prefetched
        if pf is None or pf[0] != (id(inputs), id(labels), id(lengths)):
            return inputs, labels, lengths, None
        self._prefetched = None
        torch.cuda.current_stream().wait_event(pf[2])
        x, y, ln = self._prefetch_bufs["sets"][pf[1]]
        return x, y, ln, pf[1]

    def _ws_generation(self):
        """Replacement counters of the workspaces a captured schedule addresses (nets.Workspace.generation)."""
        return (self.G.ws.generation, self.D.ws.generation if self.D is not None else 0)

    def _schedule_graphed(self, inputs, labels, lengths):
        """Copies the minibatch into static device buffers and replays the captured schedule.  The first two
        calls for a given shape / scalar set run eagerly (they allocate the workspace), the third captures.
        A graph holds raw workspace addresses: when a workspace buffer has been replaced since the capture (a longer
        batch, or the cross-validation model running a longer utterance on the shared workspace) every captured
        schedule is dropped and captured again after one eager call."""
        B, T = int(inputs.shape[0]), int(inputs.shape[1])
        key = self._graph_key(B, T)
        gen = self._ws_generation()
        if self._graphs and self._graphs_gen != gen:
            for other in self._graphs.values():
                if other["graph"] is not None:
                    other["graph"], other["calls"] = None, 1
        self._graphs_gen = gen
        st = self._graphs.get(key)
        if st is None:
            dev = self.h.device
            st = dict(calls=0, graph=None, launches=0,
                      x=torch.empty(B, T, inputs.shape[2], dtype=F32, device=dev),
                      y=torch.empty(B, T, self.output_dim, dtype=F32, device=dev),
                      ln=torch.empty(B, dtype=torch.int32, device=dev))
            if len(self._graphs) > 8:
                self._graphs.clear()
            self._graphs[key] = st
        for dst, src in ((st["x"], inputs), (st["y"], labels), (st["ln"], lengths)):
            src = src if isinstance(src, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(src))
            dst.copy_(src, non_blocking=True)              # H2D from pinned memory, or D2D
        st["calls"] += 1
        if st["graph"] is None and st["calls"] <= 2:
            feed = self._feed(st["x"], st["y"], st["ln"])
            return self._schedule(feed)
        if st["graph"] is None:
            self._capture(st)
            if self._ws_generation() != gen:                # (not expected: the eager calls above sized every buffer)
                for other in self._graphs.values():
                    if other is not st and other["graph"] is not None:
                        other["graph"], other["calls"] = None, 1
                self._graphs_gen = self._ws_generation()
        for g, ar in st["graph"]:
            g.replay()
            if ar is not None:
                self.dist.all_reduce(ar)
        self.h.launches += st["launches"]
        return st["out"]

    # -- segmented CUDA-graph capture ---------------------------------------------------------------
    def _cap_begin(self):
        g = torch.cuda.CUDAGraph()
        # thread_local: the NCCL watchdog thread and NVML samplers may touch the device while we capture
        g.capture_begin(pool=self._cap["pool"], capture_error_mode="thread_local")
        self._cap["g"] = g

    def _cap_end(self, allreduce_after):
        g = self._cap["g"]
        g.capture_end()
        self._cap["segs"].append((g, allreduce_after))
        self._cap["g"] = None

    def _capture(self, st):
        torch.cuda.synchronize()
        cs = torch.cuda.Stream(device=self.h.device)
        cs.wait_stream(torch.cuda.current_stream())
        self._cap = dict(pool=torch.cuda.graph_pool_handle(), segs=[], g=None)
        n0 = self.h.launches
        try:
            with torch.cuda.stream(cs):
                self._cap_begin()
                feed = self._feed(st["x"], st["y"], st["ln"])
                st["out"] = self._schedule(feed)
                self._cap_end(None)
            st["graph"] = self._cap["segs"]
        finally:
            self._cap = None
        torch.cuda.current_stream().wait_stream(cs)
        st["launches"] = self.h.launches - n0
        self.h.launches = n0

    def train_batch(self, inputs, labels, lengths, sync=True, all_updates=False):
        """The per-batch schedule of train_one_iteration (scripts/train_gan_rnn_placeholder.py:72-101):
        disc_updates x D update then gen_updates x G update on the SAME minibatch, which is fed to the
        device once.  Returns the losses of the last D and the last G update."""
        inputs, labels, lengths, pf_set = self._take_prefetched(inputs, labels, lengths)
        graphable = (self.use_graph and (self.world == 1 or self.graph_ddp) and self.h.timing is None
                     and self.D is not None and isinstance(inputs, (torch.Tensor, np.ndarray)))
        if graphable:
            d_all, g_all = self._schedule_graphed(inputs, labels, lengths)
        else:
            d_all, g_all = self._schedule(self._feed(inputs, labels, lengths))
        if pf_set is not None:                              # this prefetch set may be overwritten once the schedule is done
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream())
            self._prefetch_bufs["free"][pf_set] = ev
        self._last_all = (d_all, g_all)
        if not sync:
            return (d_all[-1] if d_all else None), (g_all[-1] if g_all else None)
        if all_updates:
            return self.last_update_losses()
        out = OrderedDict()
        if d_all:
            out.update(self._loss_dict(d_all[-1].tolist() + [0.0] * 3, "d"))
        if g_all:
            out.update(self._loss_dict(g_all[-1].tolist(), "g"))
        return out

    def write_summaries(self, losses, counter):
        """`summary = sess.run(model.summaries); model.writer.add_summary(summary, counter)`
        (scripts/train_gan_rnn_placeholder.py:117-122): the loss scalars as a TensorBoard event."""
        if self.dist is not None and self.dist.get_rank() != 0:
            return
        if self.writer is None:
            from .summary import FileWriter
            self.writer = FileWriter(os.path.join(self.save_dir, "eval" if self.cross_validation else "train"))
        self.writer.add_scalars({k: losses[k] for k in self.summaries if k in losses}, counter)
        self.writer.flush()

    def last_update_losses(self):
        """Losses of EVERY update of the last train_batch() -- ([D-update dicts], [G-update dicts]) -- as the
        reference accumulates them per sess.run (scripts/train_gan_rnn_placeholder.py:85-111).  Synchronises."""
        d_all, g_all = self._last_all
        return ([self._loss_dict(v.tolist() + [0.0] * 3, "d") for v in d_all],
                [self._loss_dict(v.tolist(), "g") for v in g_all])

    def eval_losses(self, inputs, labels, lengths, noise_rl=None, noise_fk=None, sync=True):
        """Loss-only pass of eval_one_iteration (scripts/train_gan_rnn_placeholder.py:154-172)."""
        x, y_tm, ln, B, T = self._feed(inputs, labels, lengths)
        h, G, D, rows = self.h, self.G, self.D, T * B
        self._mode(False)
        g32 = G.fwd(x, B, T, ln, train=False)
        lg_rl = D.fwd("rl", y_tm, B, T, ln, noise=self._noise(B, noise_rl), train=False, cat_src=self._cat(x))
        lg_fk = D.fwd("fk", g32, B, T, ln, noise=self._noise(B, noise_fk, "fk"), train=False, cat_src=self._cat(x))
        h.fill32(self._losses, 0.0)
        h.lsgan_mse_losses(self._losses, rl=lg_rl, fk=lg_fk, ld_logit=lg_rl.stride(0), n_logit=rows, clip=D.clip,
                           g=g32, y=y_tm, n_frames=rows, d_out=self.output_dim, d_real=self.d_real,
                           d_fake=self.d_fake, lam=self.mse_lambda, gscale=1.0)
        return self._loss_dict(self._losses.tolist(), "both") if sync else self._losses

    def generate(self, inputs, lengths, mean=None, std=None):
        """G(inputs) as a (B, T, output_dim) fp32 device tensor; with mean/std the decode-time inverse
        CMVN y*std+mean (scripts/train_gan_rnn_placeholder.py:286-287) is fused into the un-staging."""
        x, _, ln, B, T = self._feed(inputs, None, lengths)
        self._mode(False)
        y32 = self.G.fwd(x, B, T, ln, train=False)
        out = torch.empty(B, T, self.output_dim, dtype=F32, device=self.h.device)
        m = None if mean is None else self._to_dev("cm_mean", mean, F32)
        s = None if std is None else self._to_dev("cm_std", std, F32)
        self.h.unstage_output(y32, B, T, self.output_dim, out, mean=m, std=s)
        self.g_outputs = out
        return out
