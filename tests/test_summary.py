"""TensorBoard event files (rsrgan_b200/summary.py) read back with TensorBoard's own loader, and the shared checksum
code cross-checked against TensorBoard's independent pure-Python masked_crc32c."""
import os
import sys
from argparse import Namespace

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_handle import FakeHandle  # noqa: E402

from rsrgan_b200 import tf_checkpoint as T  # noqa: E402
from rsrgan_b200.summary import FileWriter  # noqa: E402

tb_loader = pytest.importorskip("tensorboard.backend.event_processing.event_file_loader")


def test_masked_crc32c_matches_tensorboards_implementation():
    from tensorboard.compat.tensorflow_stub.pywrap_tensorflow import masked_crc32c
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 8, 9, 63, 1000, 4097):
        b = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert T.mask_crc(T.crc32c(b)) == masked_crc32c(b), n


def test_event_file_read_back_by_tensorboard(tmp_path):
    w = FileWriter(str(tmp_path / "train"))
    w.add_scalars({"d_rl_loss": 0.25, "g_loss": 123.5}, 300)
    w.add_scalars({"d_rl_loss": 0.125, "g_loss": 61.75}, 600)
    w.close()
    events = list(tb_loader.LegacyEventFileLoader(w.path).Load())
    assert events[0].file_version == "brain.Event:2"
    got = [(e.step, {v.tag: v.simple_value for v in e.summary.value}) for e in events[1:]]
    assert got == [(300, {"d_rl_loss": 0.25, "g_loss": 123.5}), (600, {"d_rl_loss": 0.125, "g_loss": 61.75})]
    assert all(e.wall_time > 1e9 for e in events)


def test_model_writes_train_and_eval_summaries(tmp_path):
    from rsrgan_b200.gan_rnn import GAN_RNN
    a = Namespace(g_type="lstm", d_type="lstm", batch_size=2, g_cell=40, g_proj=24, g_layers=1, d_cell=32,
                  save_dir=str(tmp_path / "exp"), seed=1)
    m = GAN_RNN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    cv = GAN_RNN(None, a, ["/gpu:0"], cross_validation=True, share=m)
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 4, 257)).astype(np.float32), rng.standard_normal((2, 4, 40)).astype(np.float32)
    out = m.train_batch(x, y, np.array([4, 3]))
    m.write_summaries(out, 3)
    cv.write_summaries(cv.eval_losses(x, y, np.array([4, 3])), 3)
    for sub, model in (("train", m), ("eval", cv)):
        files = os.listdir(os.path.join(a.save_dir, sub))
        assert len(files) == 1 and files[0].startswith("events.out.tfevents.")
        ev = list(tb_loader.LegacyEventFileLoader(model.writer.path).Load())
        tags = {v.tag: v.simple_value for v in ev[1].summary.value}
        assert set(tags) == set(m.summaries) and ev[1].step == 3
        assert tags["d_loss"] == pytest.approx(tags["d_rl_loss"] + tags["d_fk_loss"], rel=1e-5)
