"""Minibatch pipeline (rsrgan_b200/dataset.py): the contract of io_funcs/tfrecords_dataset.py:53-180 over Kaldi
scp/ark -- bucketing, zero padding, CMVN, ragged tail batches -- plus the threading behaviour of the loader."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_trainer_cli_gpu import _make_data  # noqa: E402  (fixture builder only; no GPU needed)

from rsrgan_b200.dataset import Prefetcher, get_batch, get_padded_batch, read_list  # noqa: E402


def _all_list(d):
    with open(os.path.join(d, "all.scp"), "w") as f:
        f.write(open(os.path.join(d, "tr.scp")).read() + open(os.path.join(d, "cv.scp")).read())
    with open(os.path.join(d, "all.list"), "w") as f:
        f.write(os.path.join(d, "all.scp") + "\n")
    return os.path.join(d, "all.list")


def test_padded_batches_contract_and_thread_safety(tmp_path):
    d = str(tmp_path)
    lens = _make_data(d, n_utt=30)
    lst = _all_list(d)
    cm = np.load(os.path.join(d, "train_cmvn.npz"))          # NpzFile: lazy zip reads, shared by 16 loader threads
    seen = {}
    for rep in range(5):
        for ids, x, y, ln in Prefetcher(get_padded_batch(read_list(lst), 4, 257, 40, 0, 0, 16, 1, cmvn=cm, seed=rep)):
            assert x.shape[0] == y.shape[0] == len(ids) == len(ln) and x.shape[1] == y.shape[1] == int(ln.max())
            for i, u in enumerate(ids):
                T = int(ln[i])
                assert T == lens[int(u[3:])]
                assert not x[i, T:].any() and not y[i, T:].any()          # zero padding to the batch maximum
                # all utterances of a batch share the bucket min(20, (len - 200) // 50)  (floor division)
                assert (T - 200) // 50 == (int(ln[0]) - 200) // 50
                seen[u] = (x[i, :T].copy(), y[i, :T].copy())
    assert len(seen) == 30
    # CMVN applied in float64 then stored as float32: (x - 1) / 2 and (y + 1) / 3 of the N(1, 2) / N(-1, 3) fixtures
    xs = np.concatenate([v[0] for v in seen.values()])
    assert abs(xs.mean()) < 0.02 and abs(xs.std() - 1.0) < 0.02
    # decode-time reader: file order, batch of one, no labels needed
    test_list = os.path.join(d, "test.list")
    got = [ids[0] for ids, x, y, ln in get_batch(read_list(test_list), 1, 257, 40, 0, 0, 2, 1, infer=True, cmvn=cm)]
    assert got == ["utt08", "utt09"]


def test_loader_errors_reach_the_consumer(tmp_path):
    d = str(tmp_path)
    _make_data(d, n_utt=12)
    with open(os.path.join(d, "bad.scp"), "w") as f:
        f.write("uttX %s:999999 %s:0\n" % (os.path.join(d, "inputs.ark"), os.path.join(d, "labels.ark")))
    with open(os.path.join(d, "bad.list"), "w") as f:
        f.write(os.path.join(d, "bad.scp") + "\n")
    with pytest.raises(RuntimeError, match="uttX"):
        list(Prefetcher(get_padded_batch(read_list(os.path.join(d, "bad.list")), 4, 257, 40, 0, 0, 2, 1)))


def test_device_cmvn_of_the_fed_minibatch_is_bit_identical_to_the_host_loader(tmp_path):
    """cmvn_on_device=True: the loader hands out RAW zero-padded features and GAN_RNN.set_cmvn normalises the fed batch
    through rsr_cmvn_apply_padded (float64, masked by the lengths) -- same bits as the host path, which follows
    io_funcs/make_tfrecords.py:84-87, and padded frames stay exact zeros.  (FakeHandle here; the CUDA kernel is held
    to the same statement in tests/test_kernels_gpu.py::test_cmvn_kernels.)"""
    from argparse import Namespace
    import torch
    from fake_handle import FakeHandle
    from rsrgan_b200.gan_rnn import GAN_RNN
    d = str(tmp_path)
    _make_data(d, n_utt=12)
    lst = _all_list(d)
    cm = {k: v for k, v in np.load(os.path.join(d, "train_cmvn.npz")).items()}
    host = list(get_padded_batch(read_list(lst), 4, 257, 40, 1, 1, 2, 1, cmvn=cm, seed=3))
    raw = list(get_padded_batch(read_list(lst), 4, 257, 40, 1, 1, 2, 1, cmvn=cm, seed=3, cmvn_on_device=True))
    a = Namespace(g_type="lstm", d_type="lstm", batch_size=4, g_cell=40, g_proj=24, g_layers=1, d_cell=32, seed=2,
                  left_context=1, right_context=1)
    m = GAN_RNN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    m.set_cmvn(cm)
    n = 0
    for (ids_h, xh, yh, lh), (ids_r, xr, yr, lr) in zip(host, raw):
        assert ids_h == ids_r and not np.array_equal(xh, xr)
        x, y_tm, ln, B, T = m._feed(xr, yr, lr)
        assert np.array_equal(x.numpy(), xh)
        y = y_tm.numpy().reshape(T, B, 40).transpose(1, 0, 2)          # staged time-major
        assert np.array_equal(y, yh)
        for b in range(B):
            assert not x.numpy()[b, int(lh[b]):].any()
        n += 1
    assert n >= 2
