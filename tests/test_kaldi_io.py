"""Kaldi ark / CMVN surface against fixtures produced by the REFERENCE's own reader
(tests/golden/make_kaldi_golden.py imports io_funcs/kaldi_io.py from /root/reference)."""
import os
import struct

import numpy as np

from oracle import rsr_oracle as O
from rsrgan_b200 import kaldi_io

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def cm_utterance_bytes(seed=7, rows=500, cols=257):
    """Archive body ('\\0BCM ' + GlobalHeader + PerColHeaders + bytes) of a full-utterance compressed matrix, rebuilt from
    the seed (so the 130 KB archive need not be committed); every byte value and all three char_to_float branches occur."""
    rng = np.random.default_rng(seed)
    hdr = np.sort(rng.integers(0, 65536, size=(cols, 4)), axis=1).astype("<u2")
    data = rng.integers(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :256] = np.arange(256, dtype=np.uint8)
    return b"\0BCM " + struct.pack("<ffii", -17.3125, 40.75, rows, cols) + hdr.tobytes() + data.tobytes(), (rows, cols)


def _entries():
    r = kaldi_io.ArkReader()
    cwd = os.getcwd()
    os.chdir(GOLD)
    try:
        r(os.path.join(GOLD, "kaldi_small.scp"))
        return r.utt_ids, [r.read_utt_data_from_index(i) for i in range(len(r.utt_ids))]
    finally:
        os.chdir(cwd)


def test_reader_matches_reference_reader_bit_exact():
    exp = np.load(os.path.join(GOLD, "kaldi_small_expected.npz"))
    ids, mats = _entries()
    assert ids == ["utt_fm", "utt_dm", "utt_cm", "utt_1x1"]
    for k, m in zip(ids, mats):
        assert m.dtype == exp[k].dtype and m.shape == exp[k].shape, k
        assert np.array_equal(m, exp[k]), k          # float / double / compressed: identical bits


def test_writer_bytes_and_roundtrip(tmp_path):
    """ArkWriter.write_next_utt layout (io_funcs/kaldi_io.py:260-278): key, '\\0BFM ', '\\4' rows, '\\4' cols, fp32."""
    ark, scp = str(tmp_path / "feats.ark"), str(tmp_path / "feats.scp")
    w = kaldi_io.ArkWriter(scp)
    a = np.arange(12, dtype=np.float64).reshape(3, 4) / 7
    b = np.random.default_rng(0).standard_normal((2, 40))
    w.write_next_utt(ark, "uttA", a)
    w.write_next_utt(ark, "uttB", b)
    w.close()
    raw = open(ark, "rb").read()
    assert raw[:4] == b"uttA" and raw[4:9] == b"\0BFM " and raw[9:14] == struct.pack("<bi", 4, 3) and raw[14:19] == struct.pack("<bi", 4, 4)
    assert raw[19:19 + 48] == a.astype(np.float32).tobytes()
    lines = open(scp).read().splitlines()
    assert lines[0] == "uttA %s:4" % ark
    r = kaldi_io.ArkReader()
    r(scp)
    assert np.array_equal(r.read_utt_data_from_id("uttA"), a.astype(np.float32))
    assert np.array_equal(r.read_utt_data_from_id("uttB"), b.astype(np.float32))
    uid, mat, looped = r.read_next_utt()
    assert uid == "uttA" and not looped and mat.shape == (3, 4)


def test_cmvn_conversion_and_apply(tmp_path):
    exp = np.load(os.path.join(GOLD, "global_cmvn_expected.npz"))
    f = os.path.join(GOLD, "global.cmvn")
    name = kaldi_io.convert_cmvn_to_numpy(f, f, str(tmp_path))
    got = np.load(name)
    assert sorted(got.files) == ["mean_inputs", "mean_labels", "stddev_inputs", "stddev_labels"]
    assert np.array_equal(got["mean_inputs"], exp["mean"]) and np.array_equal(got["stddev_labels"], exp["std"])
    m, s = O.cmvn_from_stats(kaldi_io.read_binary_file(f))
    assert np.array_equal(m, exp["mean"]) and np.array_equal(s, exp["std"])
    x = np.random.default_rng(1).standard_normal((6, 5)).astype(np.float32)
    z = O.cmvn_apply(x, m, s)
    assert z.dtype == np.float32 and np.allclose(O.cmvn_invert(z, m, s), x, atol=1e-5)


def test_read_ark_device_host_double():
    """read_ark_device through the CPU test double: same bits as read_ark + float64 CMVN (the CUDA kernel is checked in
    tests/test_kernels_gpu.py::test_ark_decompress_bit_exact)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from fake_handle import FakeHandle
    exp = np.load(os.path.join(GOLD, "kaldi_small_expected.npz"))
    r = kaldi_io.ArkReader()
    cwd = os.getcwd()
    os.chdir(GOLD)
    try:
        r(os.path.join(GOLD, "kaldi_small.scp"))
        h = FakeHandle("f16")
        for k in ("utt_cm", "utt_fm"):
            path, off = r.scp_data[r.utt_ids.index(k)]
            cols = exp[k].shape[1]
            mean, std = np.linspace(-1, 1, cols), np.linspace(0.5, 2, cols)
            got = r.read_ark_device(h, path, off, mean, std).numpy()
            want = ((exp[k].astype(np.float64) - mean) / std).astype(np.float32)
            assert got.dtype == np.float32 and np.array_equal(got, want), k
            assert np.array_equal(r.read_ark_device(h, path, off).numpy(), exp[k].astype(np.float32)), k
        raw = kaldi_io.ArkReader.read_compressed_raw(*r.scp_data[r.utt_ids.index("utt_cm")])
        assert raw[2:4] == exp["utt_cm"].shape and raw[4].shape == (raw[3], 4) and raw[5].shape == (raw[3], raw[2])
        assert kaldi_io.ArkReader.read_compressed_raw(*r.scp_data[r.utt_ids.index("utt_dm")]) is None
    finally:
        os.chdir(cwd)


def test_compressed_utterance_matches_reference_reader_sha256(tmp_path):
    """500 x 257 compressed matrix: our vectorised reader returns the very float64 bits the REFERENCE's per-element
    reader returned for the same bytes (tests/golden/make_kaldi_golden.py computed the digest by importing it)."""
    import hashlib
    exp = np.load(os.path.join(GOLD, "kaldi_cm_utt_expected.npz"))
    body, (rows, cols) = cm_utterance_bytes()
    ark = str(tmp_path / "big.ark")
    with open(ark, "wb") as f:
        f.write(b"utt_big ")
        pos = f.tell()
        f.write(body)
    m = kaldi_io.ArkReader().read_ark(ark, pos)
    assert m.shape == tuple(exp["shape"]) == (rows, cols) and m.dtype == np.float64
    assert np.array_equal(m[[0, 249, 499]], exp["rows_0_249_499"])
    assert hashlib.sha256(m.tobytes()).digest() == exp["sha256"].tobytes()
