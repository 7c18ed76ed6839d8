"""Property tests (hypothesis) of the byte-level host formats: Kaldi ark write -> read, compressed-matrix decode
structure, TensorFlow bundle write -> read, device parameter layouts.  CPU only."""
import io
import struct

import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

from rsrgan_b200 import kaldi_io, packing, params
from rsrgan_b200 import tf_checkpoint as T

FAST = settings(max_examples=25, deadline=None)


@FAST
@given(rows=st.integers(1, 40), cols=st.integers(1, 300), seed=st.integers(0, 2 ** 31 - 1))
def test_ark_write_read_roundtrip(tmp_path_factory, rows, cols, seed):
    d = tmp_path_factory.mktemp("ark")
    rng = np.random.default_rng(seed)
    mats = {"utt_%d" % i: rng.standard_normal((rows + i, cols)) * 10.0 ** int(rng.integers(-3, 4)) for i in range(3)}
    w = kaldi_io.ArkWriter(str(d / "f.scp"))
    for k, m in mats.items():
        w.write_next_utt(str(d / "f.ark"), k, m)
    w.close()
    r = kaldi_io.ArkReader()
    r(str(d / "f.scp"))
    assert r.utt_ids == list(mats)
    for k, m in mats.items():
        got = r.read_utt_data_from_id(k)
        assert got.dtype == np.float32 and np.array_equal(got, m.astype(np.float32))      # stored as fp32 (:269)


@FAST
@given(rows=st.integers(1, 64), cols=st.integers(1, 20), seed=st.integers(0, 2 ** 31 - 1))
def test_compressed_decode_is_monotone_and_hits_the_quartiles(rows, cols, seed):
    """char_to_float (io_funcs/kaldi_io.py:128-137) is piecewise linear and increasing in the byte, and maps 0 / 64 /
    192 / 255 to the four per-column quartiles."""
    rng = np.random.default_rng(seed)
    hdr = np.sort(rng.integers(0, 65536, (cols, 4)), axis=1).astype("<u2")
    data = np.sort(rng.integers(0, 256, (cols, rows)), axis=1).astype(np.uint8)
    data[:, :1] = 0
    mn, rg = np.float32(-5.0), np.float32(20.0)
    m = kaldi_io.ArkReader().read_compress(mn, rg, rows, cols, io.BytesIO(hdr.tobytes() + data.tobytes()))
    assert m.shape == (rows, cols) and (np.diff(m, axis=0) >= -1e-12).all()
    q = np.float64(mn) + np.float64(rg) * 1.52590218966964e-05 * hdr.astype(np.float64)
    assert np.allclose(m[0], q[:, 0])
    probe = np.tile(np.array([0, 64, 192, 255], np.uint8), (cols, 1))
    pm = kaldi_io.ArkReader().read_compress(mn, rg, 4, cols, io.BytesIO(hdr.tobytes() + probe.tobytes()))
    assert np.allclose(pm.T, q, rtol=1e-12, atol=1e-12)


names = st.text(alphabet="abcdefghijklmnopqrstuvwxyz_/0123456789", min_size=1, max_size=40)
dtypes = st.sampled_from([np.float32, np.float64, np.int32, np.int64, np.bool_, np.float16, np.uint8])
shapes = st.lists(st.integers(0, 6), min_size=0, max_size=3)


@FAST
@given(entries=st.dictionaries(names, st.tuples(dtypes, shapes, st.integers(0, 2 ** 31 - 1)), min_size=1, max_size=40))
def test_bundle_roundtrip_any_names_dtypes_shapes(tmp_path_factory, entries):
    d = tmp_path_factory.mktemp("ckpt")
    tensors = {}
    for name, (dt, shape, seed) in entries.items():
        rng = np.random.default_rng(seed)
        tensors[name] = (rng.standard_normal(shape) * 100).astype(dt)
    prefix = str(d / "model-1")
    T.write_bundle(prefix, tensors)
    back = T.read_bundle(prefix)
    assert list(back) == sorted(tensors, key=lambda n: n.encode())
    for k, v in tensors.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v), k
    # footer: 48 bytes ending in the table magic; every key of the index is reachable
    raw = open(prefix + ".index", "rb").read()
    assert struct.unpack("<Q", raw[-8:])[0] == T.MAGIC
    assert [k for k, _ in T.read_table(prefix + ".index")] == [b""] + sorted(n.encode() for n in tensors)


@FAST
@given(v=st.integers(0, 2 ** 64 - 1))
def test_varint_roundtrip(v):
    b = T.put_varint(v)
    assert T.get_varint(b, 0) == (v, len(b)) and len(b) == max(1, (v.bit_length() + 6) // 7)


@FAST
@given(n_a=st.integers(1, 300), n_b=st.integers(1, 64), n_out=st.integers(1, 70), seed=st.integers(0, 2 ** 31 - 1))
def test_device_layouts_roundtrip(n_a, n_b, n_out, seed):
    rng = np.random.default_rng(seed)
    for seg in (params.fc_w_cat("w", n_a, n_b, n_out), params.fc_w("w", n_a, n_out), params.fc_b("b", n_out),
                params.conv_w("c", 2 * (n_b % 6) + 1, n_b, n_out)):
        t = rng.standard_normal(seg.tf_shape).astype(np.float32)
        d = params.to_dev_layout(seg, t)
        assert d.shape == seg.dev_shape and all(s % 8 == 0 for s in d.shape[-1:])
        assert np.array_equal(params.from_dev_layout(seg, d), t)
        assert abs(float(d.sum()) - float(t.sum())) <= 1e-3 * (1 + abs(float(t.sum())))      # padding is exact zeros
    C = n_b
    a = rng.standard_normal((3, 4 * C)).astype(np.float32)
    assert np.array_equal(packing.unpack_cols(packing.pack_cols(a, C), C), a)
