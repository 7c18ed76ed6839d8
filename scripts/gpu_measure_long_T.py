"""Deviation of the recurrence kernels from the float64 oracle at utterance-scale lengths (the reference decodes whole
utterances, batch_size = 1, up to ~1000 frames; scripts/train_gan_rnn_placeholder.py:262-300) -- the numbers the bars of
tests/test_kernels_gpu.py::test_lstmp_recurrence_long_utterances are set from.  One JSON line per (dtype, shape)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import test_kernels_gpu as TK          # noqa: E402
from rsrgan_b200 import ops            # noqa: E402

SHAPES = [(2, 1000, 40, 256, 40, True),        # discriminator_lstm layer, two long ragged utterances
          (1, 1200, 257, 760, 257, False),     # decode of one utterance through a res_lstm_l layer (L2-exchange kernels)
          (8, 800, 256, 512, 256, True)]       # cfg-2 layer at the reference's batch of 8
for dt in ("f16", "bf16"):
    h = ops.Handle(0, dt)
    for (B, T, I, C, P, ragged) in SHAPES:
        r = TK._rec_case(h, B, T, I, C, P, ragged, seed=B + T)
        print(json.dumps(dict(dtype=dt, B=B, T=T, I=I, C=C, P=P, ragged=ragged, **{k: float("%.3g" % v) for k, v in r.items()})), flush=True)
    h.close()

# model level: decode of ONE whole utterance (GAN_RNN.generate, batch_size = 1) -- absolute and relative RMS of the
# generator output against the float64 oracle
import test_gan_gpu as TG              # noqa: E402
for dt in ("f16", "bf16"):
    for name, g_type, d_type, T, kw in (("cfg2 lstm 2x512", "lstm", "dnn", 1000, TG.CFG2),
                                        ("reference driver res_lstm_l 4x760", "res_lstm_l", "lstm", 400, {})):
        a, r = TG._g_slice(g_type, d_type, 1, T, [0], dt, kw)
        print(json.dumps(dict(dtype=dt, model=name, B=1, T=T, g_out_rms=float("%.3g" % a), g_out_rel=float("%.3g" % r))), flush=True)
