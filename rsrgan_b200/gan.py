"""GAN -- host-side mirror of the reference's FRAME-LEVEL adversarial trainer, models/gan.py:60-307.

What it computes (models/gan.py): the `dnn` generator (models/dnn.py) maps spliced LPS frames to MFCC frames; the
discriminator is `discriminator_dnn` on tf.concat([centre LPS frame, MFCC], -1) (297-d, :159-174); LSGAN losses with
hard targets 1 / 0 (:200-204), g_mse = 0.5 * output_dim * mse (:207-208), g_l2 from the layers' l2 regularisers (weights
only, :209-214); Adam for BOTH networks (:125-126), tower-mean gradients applied WITHOUT clipping (:146-151), batch_norm
UPDATE_OPS run with every step (:139-140), ExponentialMovingAverage shadows (:128-129,152-156).

Same constructor as the reference (`GAN(sess, args, devices, inputs, labels, cross_validation)`); the fetch lists of
scripts/train_gan_dnn.py become the methods inherited from GAN_RNN -- d_step / g_step / train_batch / eval_losses /
generate -- fed with frames: (N, input_dim * splice) inputs and (N, output_dim) labels (or (B, T, .) stacks of frames).
One process per GPU; towers are ranks of torch.distributed.  Everything runs in librsrgan_sm100.so through
rsrgan_b200.nets; there is no CPU path.
"""
from __future__ import annotations

from argparse import Namespace

import numpy as np
import torch

from .gan_rnn import GAN_RNN, _arg


class GAN(GAN_RNN):
    def __init__(self, sess, args, devices, inputs=None, labels=None, cross_validation=False, name="GAN",
                 handle=None, share=None):
        g_type = _arg(args, "g_type", "dnn")
        if g_type != "dnn":
            raise ValueError("Unrecognized G type {}".format(g_type))       # models/gan.py:109-112
        a = dict(vars(args)) if args is not None else {}
        a.update(g_type="dnn", d_type="dnn", d_cat_dim=_arg(args, "input_dim", 257), d_adam=True)
        super(GAN, self).__init__(sess, Namespace(**a), devices, cross_validation=cross_validation, infer=False,
                                  name=name, handle=handle, share=share)
        self.max_grad_norm = 1e30            # apply_gradients(avg_grads) with no clip_by_norm (:146-151)
        self.update_bn_stats = True          # control_dependencies(UPDATE_OPS) around compute_gradients (:139-143):
        self.bn_update_scope = "all"         # ... the WHOLE collection, for both optimizers
        self.d_real, self.d_fake = 1.0, 0.0  # squared_difference(logits, 1.) / (logits, 0.) (:200-202)
        self.disc_noise_std = 0.0            # the noise layer is commented out in discriminator_dnn (:58)
        if share is None:
            # l2 comes from the layers' weights_regularizer: weights only (:209-214; models/dnn.py:64-67)
            P = self.G.P
            P.seg_l2 = torch.tensor(np.array([1 if s.name.endswith("weights") else 0 for s in P.segs.values()],
                                             np.int32), device=P.seg_l2.device)
        self._feed_names = (inputs, labels)  # the reference wires queue tensors here; kept for the call signature

    @staticmethod
    def _frames(a):
        a = a if isinstance(a, torch.Tensor) else np.asarray(a)
        return a.reshape(-1, 1, a.shape[-1]) if a.ndim == 2 else a

    def _lengths(self, x3, lengths):
        return np.full(int(x3.shape[0]), int(x3.shape[1]), np.int32) if lengths is None else lengths

    # frames in, same steps: lengths are implied (every frame is real)
    def d_step(self, inputs, labels, lengths=None, **kw):
        if kw.get("_feed") is not None:
            return super(GAN, self).d_step(inputs, labels, lengths, **kw)
        x3, y3 = self._frames(inputs), self._frames(labels)
        return super(GAN, self).d_step(x3, y3, self._lengths(x3, lengths), **kw)

    def g_step(self, inputs, labels, lengths=None, **kw):
        if kw.get("_feed") is not None:
            return super(GAN, self).g_step(inputs, labels, lengths, **kw)
        x3, y3 = self._frames(inputs), self._frames(labels)
        return super(GAN, self).g_step(x3, y3, self._lengths(x3, lengths), **kw)

    def train_batch(self, inputs, labels, lengths=None, **kw):
        x3, y3 = self._frames(inputs), self._frames(labels)
        return super(GAN, self).train_batch(x3, y3, self._lengths(x3, lengths), **kw)

    def eval_losses(self, inputs, labels, lengths=None, **kw):
        x3, y3 = self._frames(inputs), self._frames(labels)
        return super(GAN, self).eval_losses(x3, y3, self._lengths(x3, lengths), **kw)

    def generate(self, inputs, lengths=None, mean=None, std=None):
        x3 = self._frames(inputs)
        out = super(GAN, self).generate(x3, self._lengths(x3, lengths), mean=mean, std=std)
        nd = inputs.ndim if hasattr(inputs, "ndim") else np.asarray(inputs).ndim
        return out.reshape(int(x3.shape[0]), self.output_dim) if nd == 2 else out
