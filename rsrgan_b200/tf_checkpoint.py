"""TensorFlow checkpoint-V2 ("tensor bundle") reader / writer -- the container behind the reference's
`tf.train.Saver().save(sess, save_dir/GAN_RNN, global_step)` (models/gan_rnn_placeholder.py:26-60), so that a
`GAN_RNN-<step>.{index,data-00000-of-00001}` pair written by the reference can be loaded here and vice versa.

TensorFlow itself is absent from this image, so this module restates the published on-disk format
(tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/{table_builder,block_builder,format}.cc, which are
LevelDB's table format) and is checked for self-consistency and against hand-assembled bytes
(tests/test_tf_checkpoint.py); it has NOT been cross-read by a real TensorFlow.

  <prefix>.index                  an SSTable: sorted (key -> value) entries in prefix-compressed blocks
      key ""           -> BundleHeaderProto {num_shards = 1, endianness = LITTLE, version {producer = 1}}
      key <var name>   -> BundleEntryProto  {dtype, shape, shard_id, offset, size, crc32c (masked)}
      block            := entries, restart offsets (uint32 each), restart count (uint32)
      entry            := varint shared, varint non_shared, varint value_len, key suffix, value
      block trailer    := compression type (0 = none; TF writes bundles uncompressed) + masked crc32c(block + type)
      footer (48 B)    := metaindex handle, index handle (varint offset, varint size), zero padding, magic
  <prefix>.data-00000-of-00001    the tensors' little-endian bytes, back to back in key order

crc32c comes from the C-ABI library (`rsr_crc32c_host`, a host function); everything else is plain Python.
"""
from __future__ import annotations

import ctypes
import os
import re
import struct
from collections import OrderedDict

import numpy as np

from . import _lib

MAGIC = 0xdb4775248b80fb57
BLOCK_SIZE = 262144            # tensorflow/core/lib/io/table_options.h
RESTART_INTERVAL = 16
MASK_DELTA = 0xa282ead8

# tensorflow/core/framework/types.proto
DTYPES = {1: np.dtype("<f4"), 2: np.dtype("<f8"), 3: np.dtype("<i4"), 4: np.dtype("u1"), 5: np.dtype("<i2"),
          6: np.dtype("i1"), 9: np.dtype("<i8"), 10: np.dtype("bool"), 19: np.dtype("<f2")}
DTYPE_IDS = {v: k for k, v in DTYPES.items()}


# ---------------------------------------------------------------------------------- checksums / varints
def crc32c(data, crc=0):
    data = bytes(data)
    if not data:
        return crc
    return int(_lib.load().rsr_crc32c_host(ctypes.c_char_p(data), len(data), crc))


def mask_crc(crc):
    """crc32c::Mask -- stored CRCs are rotated and offset so that a CRC of a CRC is not degenerate."""
    return (((crc >> 15) | (crc << 17)) + MASK_DELTA) & 0xffffffff


def unmask_crc(masked):
    rot = (masked - MASK_DELTA) & 0xffffffff
    return ((rot >> 17) | (rot << 15)) & 0xffffffff


def put_varint(v):
    out = bytearray()
    v &= (1 << 64) - 1
    while v >= 0x80:
        out.append((v & 0x7f) | 0x80)
        v >>= 7
    out.append(v)
    return bytes(out)


def get_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7f) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 63:
            raise ValueError("malformed varint")


# ---------------------------------------------------------------------------------- protobuf (the three messages used)
def _field(num, wire, payload):
    return put_varint((num << 3) | wire) + payload


def encode_header(num_shards=1):
    version = _field(1, 0, put_varint(1))                      # VersionDef.producer = 1
    return _field(1, 0, put_varint(num_shards)) + _field(3, 2, put_varint(len(version)) + version)


def encode_entry(dtype_id, shape, offset, size, crc_masked, shard_id=0):
    dims = b""
    for d in shape:
        dim = _field(1, 0, put_varint(int(d)))                 # TensorShapeProto.Dim.size
        dims += _field(2, 2, put_varint(len(dim)) + dim)
    out = _field(1, 0, put_varint(dtype_id)) + _field(2, 2, put_varint(len(dims)) + dims)
    if shard_id:
        out += _field(3, 0, put_varint(shard_id))
    if offset:
        out += _field(4, 0, put_varint(offset))
    if size:
        out += _field(5, 0, put_varint(size))
    return out + _field(6, 5, struct.pack("<I", crc_masked))


def _parse(buf):
    """-> list of (field number, wire type, value) of one message (varint / fixed32 / fixed64 / bytes)."""
    pos, out = 0, []
    while pos < len(buf):
        tag, pos = get_varint(buf, pos)
        num, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = get_varint(buf, pos)
        elif wire == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wire == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wire == 2:
            n, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + n])
            pos += n
        else:
            raise ValueError("unsupported protobuf wire type %d" % wire)
        out.append((num, wire, v))
    return out


def decode_entry(buf):
    e = dict(dtype=0, shape=(), shard_id=0, offset=0, size=0, crc32c=None, slices=False)
    for num, _, v in _parse(buf):
        if num == 1:
            e["dtype"] = v
        elif num == 2:
            dims = []
            for n2, _, v2 in _parse(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _parse(v2):
                        if n3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    dims.append(size)
            e["shape"] = tuple(dims)
        elif num == 3:
            e["shard_id"] = v
        elif num == 4:
            e["offset"] = v
        elif num == 5:
            e["size"] = v
        elif num == 6:
            e["crc32c"] = v
        elif num == 7:
            e["slices"] = True
    return e


def decode_header(buf):
    h = dict(num_shards=0, endianness=0)
    for num, _, v in _parse(buf):
        if num == 1:
            h["num_shards"] = v
        elif num == 2:
            h["endianness"] = v
    return h


# ---------------------------------------------------------------------------------- SSTable
class _BlockBuilder(object):
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf, self.restarts, self.count, self.last = bytearray(), [0], 0, b""

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last))
            while shared < m and key[shared] == self.last[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value))
        self.buf += key[shared:] + value
        self.last = key
        self.count += 1

    def size(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + \
            struct.pack("<I", len(self.restarts))


def _write_block(f, contents):
    """-> (offset, size) of the block; trailer = type 0 + masked crc32c(contents + type)."""
    off = f.tell()
    f.write(contents)
    f.write(b"\x00" + struct.pack("<I", mask_crc(crc32c(contents + b"\x00"))))
    return off, len(contents)


def write_table(path, items):
    """items: iterable of (key bytes, value bytes), keys strictly increasing."""
    with open(path, "wb") as f:
        index = _BlockBuilder(1)
        data, last, prev = _BlockBuilder(RESTART_INTERVAL), None, None
        for key, value in items:
            if prev is not None and key <= prev:
                raise ValueError("keys must be strictly increasing")
            prev = key
            data.add(key, value)
            last = key
            if data.size() >= BLOCK_SIZE:
                off, size = _write_block(f, data.finish())
                index.add(last, put_varint(off) + put_varint(size))
                data, last = _BlockBuilder(RESTART_INTERVAL), None
        if last is not None:
            off, size = _write_block(f, data.finish())
            index.add(last, put_varint(off) + put_varint(size))
        meta = _write_block(f, _BlockBuilder(RESTART_INTERVAL).finish())
        idx = _write_block(f, index.finish())
        footer = put_varint(meta[0]) + put_varint(meta[1]) + put_varint(idx[0]) + put_varint(idx[1])
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", MAGIC))


def _read_block(buf, off, size, verify=True):
    contents = buf[off:off + size]
    ctype = buf[off + size]
    if verify:
        stored = struct.unpack_from("<I", buf, off + size + 1)[0]
        if unmask_crc(stored) != crc32c(bytes(contents) + bytes([ctype])):
            raise ValueError("block checksum mismatch at offset %d" % off)
    if ctype != 0:
        raise ValueError("compressed table block (type %d): TensorFlow writes bundle indexes uncompressed" % ctype)
    n_restarts = struct.unpack_from("<I", contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = get_varint(contents, pos)
        non_shared, pos = get_varint(contents, pos)
        vlen, pos = get_varint(contents, pos)
        key = key[:shared] + bytes(contents[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(contents[pos:pos + vlen])))
        pos += vlen
    return out


def read_table(path, verify=True):
    """-> list of (key, value) in key order."""
    buf = open(path, "rb").read()
    if len(buf) < 48 or struct.unpack_from("<Q", buf, len(buf) - 8)[0] != MAGIC:
        raise ValueError("%s is not a TensorFlow table file (bad magic)" % path)
    footer = buf[len(buf) - 48:]
    _, pos = get_varint(footer, 0)
    _, pos = get_varint(footer, pos)
    ioff, pos = get_varint(footer, pos)
    isize, pos = get_varint(footer, pos)
    out = []
    for _, handle in _read_block(buf, ioff, isize, verify):
        off, p = get_varint(handle, 0)
        size, _ = get_varint(handle, p)
        out += _read_block(buf, off, size, verify)
    return out


# ---------------------------------------------------------------------------------- bundles
def write_bundle(prefix, tensors):
    """tensors: {variable name: numpy array}.  Writes <prefix>.index and <prefix>.data-00000-of-00001."""
    names = sorted(tensors, key=lambda n: n.encode())
    entries, off = [], 0
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.asarray(tensors[n])
            dt = a.dtype.newbyteorder("<") if a.dtype.byteorder == ">" else a.dtype
            if np.dtype(dt) not in DTYPE_IDS:
                raise TypeError("%s: dtype %s has no TensorFlow mapping here" % (n, a.dtype))
            raw = np.ascontiguousarray(a, dtype=dt).tobytes()
            f.write(raw)
            entries.append((n.encode(), encode_entry(DTYPE_IDS[np.dtype(dt)], a.shape, off, len(raw),
                                                     mask_crc(crc32c(raw)))))
            off += len(raw)
    write_table(prefix + ".index", [(b"", encode_header(1))] + entries)


def read_bundle(prefix, verify=True):
    """-> OrderedDict {variable name: numpy array} of every tensor in the checkpoint."""
    items = read_table(prefix + ".index", verify)
    if not items or items[0][0] != b"":
        raise ValueError("%s.index has no bundle header" % prefix)
    hdr = decode_header(items[0][1])
    if hdr["endianness"] != 0:
        raise ValueError("big-endian bundle")
    shards = {}
    out = OrderedDict()
    for key, value in items[1:]:
        e = decode_entry(value)
        if e["slices"]:
            raise ValueError("%s: partitioned (sliced) variables are not supported" % key.decode())
        if e["dtype"] not in DTYPES:
            raise TypeError("%s: TensorFlow dtype %d is not supported" % (key.decode(), e["dtype"]))
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = open("%s.data-%05d-of-%05d" % (prefix, sid, hdr["num_shards"]), "rb")
        f = shards[sid]
        f.seek(e["offset"])
        raw = f.read(e["size"])
        if verify and e["crc32c"] is not None and unmask_crc(e["crc32c"]) != crc32c(raw):
            raise ValueError("%s: tensor checksum mismatch" % key.decode())
        out[key.decode()] = np.frombuffer(raw, dtype=DTYPES[e["dtype"]]).reshape(e["shape"]).copy()
    for f in shards.values():
        f.close()
    return out


# ---------------------------------------------------------------------------------- `checkpoint` state file
def write_checkpoint_state(save_dir, latest, all_paths):
    """The text-format CheckpointState that tf.train.Saver maintains next to the bundles."""
    with open(os.path.join(save_dir, "checkpoint"), "w") as f:
        f.write('model_checkpoint_path: "%s"\n' % latest)
        for p in all_paths:
            f.write('all_model_checkpoint_paths: "%s"\n' % p)


def read_checkpoint_state(save_dir):
    """-> (latest, [all]) from a TF `checkpoint` file, or (None, []) when it is not in that format."""
    path = os.path.join(save_dir, "checkpoint")
    if not os.path.exists(path):
        return None, []
    text = open(path).read()
    m = re.search(r'^model_checkpoint_path:\s*"([^"]*)"', text, re.M)
    return (m.group(1) if m else None), re.findall(r'^all_model_checkpoint_paths:\s*"([^"]*)"', text, re.M)


# ---------------------------------------------------------------------------------- GAN state <-> TF variable names
SLOT_M, SLOT_V, SLOT_EMA = "/Adam", "/Adam_1", "/ExponentialMovingAverage"
# tf.Variable(..., trainable=False) without a name, created under tf.name_scope('model') in this order
# (models/gan_rnn_placeholder.py:112-123, scripts/train_gan_rnn_placeholder.py:427-431): best-effort names
SCALARS = ("mse_lambda", "disc_noise_std", "d_learning_rate", "g_learning_rate", "d_real", "d_fake")


def state_to_tensors(sd):
    """GAN_RNN.state_dict() -> {TF variable name: array} as the reference's Saver would hold them: weights under
    their names, Adam slots `<var>/Adam`, `<var>/Adam_1`, `beta{1,2}_power`, EMA shadows
    `<var>/ExponentialMovingAverage`, batch_norm statistics, and the six scalar variables."""
    out = OrderedDict()
    both_adam = all(k in sd and "m" in sd[k] for k in ("G", "D"))
    for key in ("G", "D"):
        if key not in sd:
            continue
        net = sd[key]
        for n, a in net["theta"].items():
            out[n] = np.asarray(a, np.float32)
        for buf, suffix in (("m", SLOT_M), ("v", SLOT_V), ("ema", SLOT_EMA)):
            for n, a in net.get(buf, {}).items():
                out[n + suffix] = np.asarray(a, np.float32)
        for n, a in net.get("bn_state", {}).items():
            out[n] = np.asarray(a, np.float32)
        if "m" in net:
            # every AdamOptimizer creates its own beta powers when apply_gradients runs: with one Adam network (the
            # placeholder GAN, G only) they are beta{1,2}_power; with two (models/gan.py:148-151 applies d_opt first, then
            # g_opt) TensorFlow uniquifies the second pair to beta{1,2}_power_1
            sfx = "_1" if (both_adam and key == "G") else ""
            out["model/beta1_power" + sfx] = np.float32(net["hyper"][4])
            out["model/beta2_power" + sfx] = np.float32(net["hyper"][5])
    for i, k in enumerate(SCALARS):
        if k in sd.get("scalars", {}):
            out["model/Variable" + ("_%d" % i if i else "")] = np.float32(sd["scalars"][k])
    return out


def tensors_to_state(tensors, sd):
    """Fills a state dict shaped like `sd` (from GAN_RNN.state_dict()) from checkpoint tensors.  Weights must be
    present under their exact names; optimizer slots, EMA shadows, beta powers and scalars are taken when found
    (matched by suffix, so a name-scope prefix TensorFlow may have added does not matter) and left as they are
    otherwise.  Returns the names that were missing."""
    missing = []
    both_adam = all(k in sd and "m" in sd[k] for k in ("G", "D"))

    def find(name):
        if name in tensors:
            return tensors[name]
        hits = [k for k in tensors if k.endswith("/" + name)]
        return tensors[hits[0]] if len(hits) == 1 else None

    for key in ("G", "D"):
        if key not in sd:
            continue
        net = sd[key]
        for n in list(net["theta"]):
            if n not in tensors:
                raise KeyError("checkpoint has no variable %s" % n)
            net["theta"][n] = np.asarray(tensors[n], np.float32).reshape(net["theta"][n].shape)
        for buf, suffix in (("m", SLOT_M), ("v", SLOT_V), ("ema", SLOT_EMA)):
            for n in list(net.get(buf, {})):
                a = find(n + suffix)
                if a is None:
                    missing.append(n + suffix)
                    if buf == "ema":
                        net[buf][n] = net["theta"][n].copy()
                else:
                    net[buf][n] = np.asarray(a, np.float32).reshape(net[buf][n].shape)
        for n in list(net.get("bn_state", {})):
            a = find(n)
            if a is None:
                missing.append(n)
            else:
                net["bn_state"][n] = np.asarray(a, np.float32).reshape(np.shape(net["bn_state"][n]))
        if "m" in net:
            hyper = np.array(net["hyper"], np.float32)
            sfx = "_1" if (both_adam and key == "G") else ""
            for idx, name in ((4, "beta1_power" + sfx), (5, "beta2_power" + sfx)):
                a = find(name)
                if a is None:
                    missing.append(name)
                else:
                    hyper[idx] = float(a)
            net["hyper"] = hyper
    for i, k in enumerate(SCALARS):
        a = tensors.get("model/Variable" + ("_%d" % i if i else ""))
        if a is not None and np.ndim(a) == 0:
            sd.setdefault("scalars", {})[k] = float(a)
    return missing
