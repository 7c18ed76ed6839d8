"""TensorFlow checkpoint-V2 tensor bundles (rsrgan_b200/tf_checkpoint.py): CRC-32C known answers, the table / bundle
bytes against a hand assembly of the published format, round trips, corruption detection, and GAN_RNN save / load in
that container (through the CPU test double).  No TensorFlow here: these pin the format as restated, not TF itself."""
import os
import struct
import sys
from argparse import Namespace

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from fake_handle import FakeHandle  # noqa: E402

from rsrgan_b200 import tf_checkpoint as T  # noqa: E402
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402


def test_crc32c_known_answers_and_masking():
    # RFC 3720 B.4 test vectors (the ones tensorflow/core/lib/hash/crc32c_test.cc uses)
    assert T.crc32c(b"123456789") == 0xE3069283
    assert T.crc32c(bytes(32)) == 0x8A9136AA
    assert T.crc32c(b"\xff" * 32) == 0x62A8AB43
    assert T.crc32c(bytes(range(32))) == 0x46DD794E
    assert T.crc32c(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert T.crc32c(b"6789", T.crc32c(b"12345")) == 0xE3069283            # incremental
    big = np.random.default_rng(0).integers(0, 256, 100003, dtype=np.uint8).tobytes()
    assert T.crc32c(big[50001:], T.crc32c(big[:50001])) == T.crc32c(big)
    c = T.crc32c(b"foo")
    assert T.mask_crc(c) != c and T.unmask_crc(T.mask_crc(c)) == c
    assert T.mask_crc(T.mask_crc(c)) != c and T.unmask_crc(T.unmask_crc(T.mask_crc(T.mask_crc(c)))) == c


def test_varints_and_protos_known_bytes():
    assert T.put_varint(0) == b"\x00" and T.put_varint(300) == b"\xac\x02" and T.get_varint(b"\xac\x02", 0) == (300, 2)
    assert T.encode_header(1) == bytes.fromhex("0801" "1a02" "0801")
    e = T.encode_entry(1, (2, 3), 0, 24, 0x01020304)
    # dtype=DT_FLOAT, shape{dim{size:2} dim{size:3}}, size=24, crc32c fixed32 (offset 0 and shard 0 are defaults: omitted)
    assert e == bytes.fromhex("0801" "1208" "12020802" "12020803" "2818" "35" "04030201")
    d = T.decode_entry(e)
    assert d["dtype"] == 1 and d["shape"] == (2, 3) and d["offset"] == 0 and d["size"] == 24 and d["crc32c"] == 0x01020304
    s = T.encode_entry(3, (), 24, 4, 7)                     # a scalar: an empty shape message is still written
    assert s.startswith(bytes.fromhex("0803" "1200" "2018" "2804"))
    assert T.decode_entry(s)["shape"] == () and T.decode_entry(s)["offset"] == 24


def test_table_bytes_match_hand_assembly(tmp_path):
    """A two-entry table assembled by hand from the LevelDB table format."""
    path = str(tmp_path / "t.index")
    T.write_table(path, [(b"", b"H"), (b"ab", b"xyz"), (b"abc", b"")])
    got = open(path, "rb").read()

    def block(body):
        return body + b"\x00" + struct.pack("<I", T.mask_crc(T.crc32c(body + b"\x00")))
    # data block: (shared, non_shared, value_len, key suffix, value) x 3, one restart at 0
    data = bytes([0, 0, 1]) + b"H" + bytes([0, 2, 3]) + b"ab" + b"xyz" + bytes([2, 1, 0]) + b"c" + \
        struct.pack("<II", 0, 1)
    meta = struct.pack("<II", 0, 1)
    meta_off = len(data) + 5
    idx_off = meta_off + len(meta) + 5
    index = bytes([0, 3, 2]) + b"abc" + bytes([0, len(data)]) + struct.pack("<II", 0, 1)
    footer = bytes([meta_off, len(meta), idx_off, len(index)])
    want = block(data) + block(meta) + block(index) + footer + bytes(40 - len(footer)) + struct.pack("<Q", T.MAGIC)
    assert got == want
    assert T.read_table(path) == [(b"", b"H"), (b"ab", b"xyz"), (b"abc", b"")]
    with pytest.raises(ValueError):
        T.write_table(path, [(b"b", b""), (b"a", b"")])


def test_bundle_roundtrip_many_keys_and_corruption(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {"g_model/fully_connected/weights": rng.standard_normal((257, 40)).astype(np.float32),
               "g_model/fully_connected/biases": np.zeros(40, np.float32),
               "model/Variable": np.float32(10.0), "global_step": np.int64(7), "flags": np.array([True, False]),
               "empty": np.zeros((0, 3), np.float32), "d": rng.standard_normal((3, 2, 2))}
    for i in range(6000):                                    # > 256 KB of index: several data blocks, many restarts
        tensors["scope_%05d/some/long/variable/name/weights" % i] = np.float32(i)
    prefix = str(tmp_path / "GAN_RNN-3")
    T.write_bundle(prefix, tensors)
    assert os.path.getsize(prefix + ".index") > T.BLOCK_SIZE
    back = T.read_bundle(prefix)
    assert list(back) == sorted(tensors) and len(back) == len(tensors)
    for k, v in tensors.items():
        a = np.asarray(v)
        assert back[k].dtype == a.dtype and back[k].shape == a.shape and np.array_equal(back[k], a), k
    # data file = tensors back to back in key order
    assert os.path.getsize(prefix + ".data-00000-of-00001") == sum(np.asarray(v).nbytes for v in tensors.values())
    raw = bytearray(open(prefix + ".data-00000-of-00001", "rb").read())
    raw[100] ^= 0x40
    open(prefix + ".data-00000-of-00001", "wb").write(bytes(raw))
    with pytest.raises(ValueError, match="checksum"):
        T.read_bundle(prefix)
    assert len(T.read_bundle(prefix, verify=False)) == len(tensors)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(ValueError, match="checksum"):
        T.read_table(prefix + ".index")


def model(**kw):
    a = dict(g_type="lstm", d_type="dnn", batch_size=2, g_cell=40, g_proj=24, g_layers=1, d_units=32, d_layers=1,
             batch_norm=True, init_mse_weight=10.0, init_disc_noise_std=0.05, seed=3, ckpt_format="tf")
    a.update(kw)
    return GAN_RNN(None, Namespace(**a), ["/gpu:0"], handle=FakeHandle("f16"))


def test_gan_save_load_as_tf_bundle(tmp_path):
    m = model()
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 5, 257)).astype(np.float32), rng.standard_normal((2, 5, 40)).astype(np.float32)
    m.train_batch(x, y, np.array([5, 4]))
    d = str(tmp_path / "exp")
    m.save(d, 1)
    m.train_batch(x, y, np.array([5, 4]))
    m.save(d, 2)
    assert sorted(os.listdir(d)) == ["GAN_RNN-1.data-00000-of-00001", "GAN_RNN-1.index", "GAN_RNN-2.data-00000-of-00001",
                                     "GAN_RNN-2.index", "checkpoint"]
    assert T.read_checkpoint_state(d) == ("GAN_RNN-2", ["GAN_RNN-1", "GAN_RNN-2"])
    names = T.read_bundle(os.path.join(d, "GAN_RNN-2"))
    # the names the reference's Saver holds (SURVEY.md App. B + optimizer slots, EMA shadows, batch_norm statistics)
    for n in ("g_model/rnn/multi_rnn_cell/cell_0/lstm_cell/kernel", "g_model/fully_connected/BatchNorm/gamma",
              "g_model/fully_connected/BatchNorm/moving_variance", "g_model/fully_connected/weights/Adam",
              "g_model/fully_connected/weights/Adam_1", "g_model/fully_connected/weights/ExponentialMovingAverage",
              "d_model/fully_connected/weights", "d_model/fully_connected/weights/ExponentialMovingAverage",
              "model/beta1_power", "model/Variable", "model/Variable_5"):
        assert n in names, n
    assert names["g_model/rnn/multi_rnn_cell/cell_0/lstm_cell/kernel"].shape == (257 // 257 * 24 + 24, 160)
    assert float(names["model/beta1_power"]) == pytest.approx(0.9 ** 5, rel=1e-5)     # 4 Adam steps: beta1^(t+1)
    m2 = model(seed=99)
    assert m2.load(d) is True
    a, b = m.state_dict(), m2.state_dict()
    for key in ("G", "D"):
        for buf in ("theta", "ema", "m", "v", "bn_state"):
            for n in a[key].get(buf, {}):
                assert np.array_equal(np.asarray(a[key][buf][n]), np.asarray(b[key][buf][n])), (key, buf, n)
        assert np.allclose(a[key]["hyper"][:6], b[key]["hyper"][:6])
    assert b["scalars"]["mse_lambda"] == 10.0 and b["scalars"]["disc_noise_std"] == pytest.approx(0.05)
    g1, g2 = m.generate(x, np.array([5, 4])).numpy(), m2.generate(x, np.array([5, 4])).numpy()
    assert np.array_equal(g1, g2)


def test_load_reference_style_checkpoint_with_weights_only(tmp_path):
    """A checkpoint that holds the trainable variables plus slots under a name-scope prefix TensorFlow may have added:
    weights load by exact name, the rest by suffix or stay as initialised; the moving-average restore path
    (models/gan_rnn_placeholder.py:47-53) takes the EMA shadows."""
    m = model(batch_norm=False, ckpt_format="pt")
    sd = m.state_dict()
    rng = np.random.default_rng(5)
    tensors = {}
    for key in ("G", "D"):
        for n, a in sd[key]["theta"].items():
            tensors[n] = rng.standard_normal(a.shape).astype(np.float32)
            tensors[n + "/ExponentialMovingAverage"] = (tensors[n] * 0.5).astype(np.float32)
    tensors["model/device_0/beta1_power"] = np.float32(0.5)
    d = str(tmp_path / "ref")
    os.makedirs(d)
    T.write_bundle(os.path.join(d, "GAN_RNN-7"), tensors)
    T.write_checkpoint_state(d, "GAN_RNN-7", ["GAN_RNN-7"])
    m2 = model(batch_norm=False, ckpt_format="pt", seed=8)
    assert m2.load(d)
    th = m2.G.P.export_tf()
    for n in th:
        assert np.array_equal(th[n], tensors[n]), n
    assert float(m2.G.P.hyper[4]) == 0.5 and float(m2.G.P.hyper[5]) == pytest.approx(0.999)
    m3 = model(batch_norm=False, ckpt_format="pt", seed=9)
    assert m3.load(d, moving_average=True)
    assert np.array_equal(m3.G.P.export_tf()["g_model/fully_connected/weights"],
                          tensors["g_model/fully_connected/weights/ExponentialMovingAverage"])
    del tensors["d_model/fully_connected/weights"]
    T.write_bundle(os.path.join(d, "GAN_RNN-8"), tensors)
    with pytest.raises(KeyError):
        model(batch_norm=False, ckpt_format="pt").load(d, model_file="GAN_RNN-8")


def test_sub_messages_parse_with_tensorflows_own_protos():
    """TensorBoard ships TensorFlow's compiled framework protos: the shape / version sub-messages we emit parse with
    them, and the dtype numbers are TensorFlow's DataType enum."""
    pb_shape = pytest.importorskip("tensorboard.compat.proto.tensor_shape_pb2")
    from tensorboard.compat.proto import types_pb2, versions_pb2
    e = T.encode_entry(1, (257, 3040), 4096, 257 * 3040 * 4, 0xdeadbeef)
    fields = {num: v for num, _, v in T._parse(e)}
    shape = pb_shape.TensorShapeProto.FromString(fields[2])
    assert [d.size for d in shape.dim] == [257, 3040] and not shape.unknown_rank
    assert pb_shape.TensorShapeProto.FromString({n: v for n, _, v in T._parse(T.encode_entry(1, (), 0, 4, 1))}[2]).dim == []
    hdr = {num: v for num, _, v in T._parse(T.encode_header(1))}
    assert versions_pb2.VersionDef.FromString(hdr[3]).producer == 1
    for name, np_dt in (("DT_FLOAT", "<f4"), ("DT_DOUBLE", "<f8"), ("DT_INT32", "<i4"), ("DT_INT64", "<i8"),
                        ("DT_BOOL", "bool"), ("DT_HALF", "<f2"), ("DT_UINT8", "u1"), ("DT_INT16", "<i2"), ("DT_INT8", "i1")):
        assert T.DTYPES[getattr(types_pb2, name)] == np.dtype(np_dt), name
    # and the other way round: a shape serialized by the real proto decodes through our parser
    real = pb_shape.TensorShapeProto(dim=[pb_shape.TensorShapeProto.Dim(size=7), pb_shape.TensorShapeProto.Dim(size=1)])
    entry = T._field(1, 0, T.put_varint(1)) + T._field(2, 2, T.put_varint(len(real.SerializeToString())) +
                                                      real.SerializeToString())
    assert T.decode_entry(entry)["shape"] == (7, 1)


def test_convert_checkpoint_script_both_directions(tmp_path, capsys):
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "convert_checkpoint", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts",
                                           "convert_checkpoint.py"))
    conv = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(conv)
    m = model(ckpt_format="pt")
    rng = np.random.default_rng(0)
    x, y = rng.standard_normal((2, 5, 257)).astype(np.float32), rng.standard_normal((2, 5, 40)).astype(np.float32)
    m.train_batch(x, y, np.array([5, 4]))
    d = str(tmp_path / "exp")
    pt = m.save(d, 4)
    assert conv.main(["--to", "tf", pt]) == 0
    prefix = pt[:-3]
    assert os.path.exists(prefix + ".index") and os.path.exists(prefix + ".data-00000-of-00001")
    assert conv.main(["--list", prefix]) == 0
    assert "g_model/fully_connected/BatchNorm/gamma" in capsys.readouterr().out
    other = model(ckpt_format="pt", seed=77)
    like = other.save(str(tmp_path / "other"), 1)
    back = str(tmp_path / "back.pt")
    assert conv.main(["--to", "pt", prefix, "--like", like, "--out", back]) == 0
    import torch
    a, b = torch.load(pt, weights_only=False), torch.load(back, weights_only=False)
    for key in ("G", "D"):
        for buf in ("theta", "ema", "m", "v", "bn_state"):
            for n in a[key].get(buf, {}):
                assert np.array_equal(np.asarray(a[key][buf][n]), np.asarray(b[key][buf][n])), (key, buf, n)


def test_two_adam_networks_keep_their_own_beta_powers(tmp_path):
    """Frame-level GAN (models/gan.py:125-126,148-151): Adam for D and for G.  D is applied first, so TensorFlow names
    its accumulators beta{1,2}_power and G's beta{1,2}_power_1; the two networks take different numbers of steps
    (disc_updates = 1, gen_updates = 2), so after a save / resume each must get its own bias correction back."""
    from rsrgan_b200.gan import GAN
    a = Namespace(batch_size=4, input_dim=24, output_dim=8, left_context=0, right_context=0, g_units=16, g_layers=1,
                  d_units=16, d_layers=1, batch_norm=False, init_mse_weight=10.0, seed=2, ckpt_format="tf",
                  disc_updates=1, gen_updates=2)
    m = GAN(None, a, ["/gpu:0"], handle=FakeHandle("f16"))
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal((4, 24)).astype(np.float32), rng.standard_normal((4, 8)).astype(np.float32)
    m.train_batch(x, y)
    m.train_batch(x, y)
    d = str(tmp_path / "exp")
    m.save(d, 2)
    names = T.read_bundle(os.path.join(d, "GAN-2"))
    assert float(names["model/beta1_power"]) == pytest.approx(0.9 ** 3, rel=1e-5)        # D: 2 steps
    assert float(names["model/beta1_power_1"]) == pytest.approx(0.9 ** 5, rel=1e-5)      # G: 4 steps
    m2 = GAN(None, Namespace(**dict(vars(a), seed=7)), ["/gpu:0"], handle=FakeHandle("f16"))
    assert m2.load(d)
    for net, net2 in ((m.G, m2.G), (m.D, m2.D)):
        assert np.allclose(net.P.hyper[4:6].numpy(), net2.P.hyper[4:6].numpy())
    assert float(m2.G.P.hyper[4]) != float(m2.D.P.hyper[4])
    # a bundle written by the reference's placeholder GAN (one Adam network) still feeds G from beta{1,2}_power
    sd = model(batch_norm=False, ckpt_format="pt").state_dict()
    t = T.state_to_tensors(sd)
    assert "model/beta1_power" in t and "model/beta1_power_1" not in t
