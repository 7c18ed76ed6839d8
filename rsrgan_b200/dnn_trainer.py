"""DNNTrainer -- host-side mirror of the reference's MSE-only generator trainer.

Reference: models/dnn_trainer_single_gpu.py:52-133 (BASELINE.json configs[0]: the `dnn` generator,
257 -> 40, trained on frames with Adam on 0.5 * output_dim * mse + l2) and its multi-tower sibling
models/dnn_trainer.py:54-160 (same loss; also accepts `rced`; tower-mean of the gradients;
ExponentialMovingAverage shadows).  Same constructor arguments and attribute names; instead of
`sess.run([model.g_opt, model.g_mse_losses, ...])` (scripts/train_dnn_single_gpu.py:47-55) the caller
invokes

    train_step(inputs, labels)   ==  sess.run([g_opt, g_mse_losses, g_l2_losses, g_losses])
    eval_losses(inputs, labels)  ==  the loss-only sess.run of the cross-validation model
    generate(inputs)             ==  sess.run(model.generator outputs)

`inputs` are frames: (N, input_dim * splice) or (B, T, input_dim * splice); labels likewise with
output_dim.  One process drives one GPU; towers are ranks of torch.distributed (one NCCL all-reduce of
the flat gradient buffer per step).  Differences from GAN_RNN, as in the reference: no discriminator,
no per-tensor clip_by_norm (`minimize` / plain apply_gradients), learning rate not scaled here.
All arithmetic runs in librsrgan_sm100.so; there is no CPU path.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np
import torch

from .gan_rnn import F32, GAN_RNN, _arg


class DNNTrainer(GAN_RNN):
    def __init__(self, sess, args, devices, inputs=None, labels=None, cross_validation=False, name="DNNTrainer",
                 handle=None, share=None):
        g_type = _arg(args, "g_type", "dnn")
        if g_type not in ("dnn", "rced"):
            # dnn_trainer_single_gpu.py:86-89 / dnn_trainer.py:94-101 (`cnn` = models/cnn.py is not on this path)
            raise ValueError("Unrecognized G type {}".format(g_type))
        if not hasattr(args, "g_type"):          # GAN_RNN's default g_type is "lstm"; this trainer's is "dnn"
            from argparse import Namespace
            args = Namespace(**dict(vars(args) if args is not None else {}, g_type=g_type))
        super(DNNTrainer, self).__init__(sess, args, devices, cross_validation=cross_validation, infer=True,
                                         name=name, handle=handle, share=share)
        self.infer = False
        if share is None:
            # contrib l2_regularizer is attached to the layer WEIGHTS only (models/dnn.py:64-67,85-86), so BatchNorm
            # beta / gamma are not regularised here (the GAN's rule is `"bias" not in name`, gan_rnn_placeholder.py:254)
            P = self.G.P
            P.seg_l2 = torch.tensor(np.array([1 if s.name.endswith("weights") else 0 for s in P.segs.values()],
                                             np.int32), device=P.seg_l2.device)
        self.update_bn_stats = True              # UPDATE_OPS run with the step (dnn_trainer_single_gpu.py:101-104)
        self.max_grad_norm = 1e30                # no clip_by_norm on this trainer (the update kernel's clip is a no-op)
        self.mse_lambda = 1.0
        self.g_learning_rate = float(_arg(args, "g_learning_rate", 0.001))
        self._feed_names = (inputs, labels)      # the reference wires queue tensors here; kept for the call signature

    # ------------------------------------------------------------------ feeding
    def _frames(self, a, dim):
        a = a if isinstance(a, torch.Tensor) else np.asarray(a)
        return a.reshape(-1, 1, a.shape[-1]) if a.ndim == 2 else a

    def _fwd_loss(self, inputs, labels, train, want_grad):
        x3, y3 = self._frames(inputs, self.input_dim), self._frames(labels, self.output_dim)
        B, T = int(x3.shape[0]), int(x3.shape[1])
        x, y_tm, ln, B, T = self._feed(x3, y3, np.full(B, T, np.int32))
        h, G, rows = self.h, self.G, T * B
        self._mode(train, g_update=train)     # UPDATE_OPS run with every training step (dnn_trainer_single_gpu.py:101-104)
        gs = self._gscale(rows) if want_grad else 1.0
        g32 = G.fwd(x, B, T, ln, train=train)
        dg32 = G.ws.get(("loss", "dg32"), rows, g32.shape[1], F32) if want_grad else None
        h.fill32(self._losses, 0.0)
        # g_mse = 0.5 * output_dim * mean((g - y)^2)   (dnn_trainer_single_gpu.py:109-110)
        h.lsgan_mse_losses(self._losses, g=g32, y=y_tm, n_frames=rows, d_out=self.output_dim, lam=1.0, gscale=gs,
                           dg_mse=dg32)
        return g32, dg32, gs, rows

    def _dict(self, vals):
        g_mse, g_l2 = float(vals[3]), float(vals[4])
        return OrderedDict(g_mse_loss=g_mse, g_l2_loss=g_l2, g_loss=g_mse + g_l2)

    # ------------------------------------------------------------------ steps
    def train_step(self, inputs, labels, sync=True):
        """One Adam update on g_mse + g_l2 (dnn_trainer_single_gpu.py:93-104)."""
        h, G = self.h, self.G
        g32, dg32, gs, rows = self._fwd_loss(inputs, labels, True, True)
        dg16 = G.ws.get(("loss", "dg16"), rows, g32.shape[1], h.h16)
        h.cast16(dg32, dg16)
        h.fill32(G.P.grad, 0.0)
        G.bwd(dg16)
        self._l2_loss()
        if self.l2_scale > 0.0:
            h.l2_grad(G.P.grad, G.P.theta, G.P.seg_id, G.P.seg_l2, self.l2_scale * gs)
        self._update(G, gs, adam=True)
        return self._dict(self._losses.tolist()) if sync else self._losses

    def eval_losses(self, inputs, labels, sync=True):
        self._fwd_loss(inputs, labels, False, False)
        self._l2_loss()
        return self._dict(self._losses.tolist()) if sync else self._losses

    def generate(self, inputs, lengths=None, mean=None, std=None):
        x3 = self._frames(inputs, self.input_dim)
        B, T = int(x3.shape[0]), int(x3.shape[1])
        out = super(DNNTrainer, self).generate(x3, np.full(B, T, np.int32), mean=mean, std=std)
        nd = inputs.ndim if hasattr(inputs, "ndim") else np.asarray(inputs).ndim
        return out.reshape(B, self.output_dim) if nd == 2 else out

    # the adversarial entry points do not exist on this trainer
    def d_step(self, *a, **k):
        raise AttributeError("DNNTrainer has no discriminator")

    g_step = train_batch = d_step
