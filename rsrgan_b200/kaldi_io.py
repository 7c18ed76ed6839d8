"""Kaldi ark/scp I/O and global-CMVN conversion -- the reference's data-format surface.

Same classes / method names as io_funcs/kaldi_io.py (ArkReader :41-251, ArkWriter :254-282) and
io_funcs/convert_cmvn_to_numpy.py (:18-74), re-implemented for Python 3 (the reference writer
and CMVN reader only run under Python 2: SURVEY.md 8c) with numpy-vectorised decoding of Kaldi's
compressed `CM` matrices (the reference loops per element with struct.unpack, :143-161).

Byte layout of one archive entry (binary mode):
    <utt_id> ' ' '\\0' 'B' <type token> ...
  type 'FM ' / 'DM ' : '\\4' int32 rows '\\4' int32 cols, then rows*cols float32 / float64, row-major
  type 'CM '         : GlobalHeader{float32 min_value, float32 range, int32 rows, int32 cols},
                       cols x PerColHeader{uint16 p0, p25, p75, p100}, then cols*rows uint8 column-major
"""
from __future__ import annotations

import os
import random
import struct
import sys

import numpy as np


class ArkReader(object):
    """Reads matrices addressed by scp lines `utt_id path:offset` (io_funcs/kaldi_io.py:41-251)."""

    def __init__(self, name="ArkReader"):
        self.name = name
        self.utt_ids, self.scp_data, self.scp_position = [], [], 0

    def __call__(self, scp_path):
        self.scp_position = 0
        self.utt_ids, self.scp_data = [], []
        with open(scp_path, "r") as fin:
            for line in fin:
                line = line.rstrip("\n")
                if not line:
                    continue
                utt_id, path_pos = line.split(" ")
                path, pos = path_pos.rsplit(":", 1)
                self.utt_ids.append(utt_id)
                self.scp_data.append((path, pos))

    def shuffle(self):
        zipped = list(zip(self.utt_ids, self.scp_data))
        random.shuffle(zipped)
        self.utt_ids, self.scp_data = (list(t) for t in zip(*zipped)) if zipped else ([], [])
        self.scp_position = 0

    @staticmethod
    def uint16_to_float(min_value, rng, value):
        # io_funcs/kaldi_io.py:121-126 (the constant is 1/65535)
        return min_value + rng * 1.52590218966964e-05 * value

    @staticmethod
    def char_to_float(p0, p25, p75, p100, value):
        """io_funcs/kaldi_io.py:128-137, vectorised over `value` (uint8 array) with per-column quartiles."""
        v = value.astype(np.float64)
        lo = p0 + (p25 - p0) * v * (1 / 64.0)
        mid = p25 + (p75 - p25) * (v - 64) * (1 / 128.0)
        hi = p75 + (p100 - p75) * (v - 192) * (1 / 63.0)
        return np.where(value < 64, lo, np.where(value <= 192, mid, hi))

    def read_compress(self, min_value, rng, rows, cols, buf):
        """io_funcs/kaldi_io.py:139-161: float64 (rows, cols) matrix."""
        hdr = np.frombuffer(buf.read(8 * cols), dtype="<u2").reshape(cols, 4).astype(np.float64)
        q = self.uint16_to_float(np.float64(np.float32(min_value)), np.float64(np.float32(rng)), hdr)
        data = np.frombuffer(buf.read(rows * cols), dtype=np.uint8).reshape(cols, rows)
        out = self.char_to_float(q[:, 0:1], q[:, 1:2], q[:, 2:3], q[:, 3:4], data)
        return np.ascontiguousarray(out.T)

    def read_ark(self, ark_file, ark_offset=0):
        """io_funcs/kaldi_io.py:81-119."""
        with open(ark_file, "rb") as f:
            f.seek(int(ark_offset), 0)
            header = struct.unpack("<xcccc", f.read(5))
            if header[0] != b"B":
                print(ark_file)
                print("Input .ark file is not binary")
                sys.exit(1)
            if header[1] == b"C":
                if header[2] == b"M" and header[3] != b"2":
                    min_value, rng, rows, cols = struct.unpack("<ffii", f.read(16))
                    if cols == 0:
                        print("Empty matrix.")
                        sys.exit(1)
                    return self.read_compress(min_value, rng, rows, cols, f)
                print("Unsupport format.")
                print("Maybe because of the matrices with 8 or fewer rows.")
                sys.exit(1)
            _, rows = struct.unpack("<bi", f.read(5))
            _, cols = struct.unpack("<bi", f.read(5))
            if header[1] == b"F":
                mat = np.frombuffer(f.read(rows * cols * 4), dtype=np.float32)
            elif header[1] == b"D":
                mat = np.frombuffer(f.read(rows * cols * 8), dtype=np.float64)
            else:
                print("Unsupport format.")
                sys.exit(1)
            return np.reshape(mat, (rows, cols))

    # -- device-side decode of compressed entries -------------------------------------------
    @staticmethod
    def read_compressed_raw(ark_file, ark_offset=0):
        """The undecoded pieces of a `CM` entry: (min_value, range, rows, cols, col_hdr uint16 [cols, 4],
        data uint8 [cols, rows]) exactly as on disk (io_funcs/kaldi_io.py:88-102,139-161); None for any other type."""
        with open(ark_file, "rb") as f:
            f.seek(int(ark_offset), 0)
            header = struct.unpack("<xcccc", f.read(5))
            if header[0] != b"B" or header[1] != b"C" or header[2] != b"M" or header[3] == b"2":
                return None
            min_value, rng, rows, cols = struct.unpack("<ffii", f.read(16))
            hdr = np.frombuffer(f.read(8 * cols), dtype="<u2").reshape(cols, 4)
            data = np.frombuffer(f.read(rows * cols), dtype=np.uint8).reshape(cols, rows)
            return min_value, rng, rows, cols, hdr, data

    def read_ark_device(self, handle, ark_file, ark_offset=0, mean=None, std=None):
        """read_ark + the CMVN of io_funcs/make_tfrecords.py:84-87 with the decode on the GPU: a `CM` entry travels to
        the device as BYTES (1 byte per element instead of 4) and `rsr_ark_decompress` produces
        float32((x - mean) / std) evaluated in float64, bit-identical to the host path.  Returns a float32 device
        tensor (rows, cols).  Uncompressed entries are decoded on the host (they are plain arrays) and normalised by
        the same float64 expression before the upload."""
        import torch
        raw = self.read_compressed_raw(ark_file, ark_offset)
        dev = handle.device
        if raw is None:
            m = self.read_ark(ark_file, ark_offset).astype(np.float64)
            if mean is not None:
                m = (m - np.asarray(mean, np.float64)) / np.asarray(std, np.float64)
            return torch.from_numpy(np.ascontiguousarray(m.astype(np.float32))).to(dev)
        min_value, rng, rows, cols, hdr, data = raw
        hdr_d = torch.from_numpy(hdr.view(np.int16).copy()).to(dev)      # the 16 raw bits; torch has no uint16 arithmetic
        data_d = torch.from_numpy(data.copy()).to(dev)
        out = torch.empty(rows, cols, dtype=torch.float32, device=dev)
        mean_d = std_d = None
        if mean is not None:
            mean_d = torch.from_numpy(np.ascontiguousarray(mean, dtype=np.float64)).to(dev)
            std_d = torch.from_numpy(np.ascontiguousarray(std, dtype=np.float64)).to(dev)
        handle.ark_decompress(hdr_d, data_d, min_value, rng, rows, cols, out32=out, mean=mean_d, std=std_d)
        return out

    def read_next_utt(self):
        if len(self.scp_data) == 0:
            return None, None, True
        looped = False
        if self.scp_position >= len(self.scp_data):
            looped = True
            self.scp_position = 0
        self.scp_position += 1
        return self.utt_ids[self.scp_position - 1], self.read_utt_data_from_index(self.scp_position - 1), looped

    def read_next_scp(self):
        if self.scp_position >= len(self.scp_data):
            self.scp_position = 0
        self.scp_position += 1
        return self.utt_ids[self.scp_position - 1]

    def read_utt_data_from_id(self, utt_id):
        return self.read_utt_data_from_index(self.utt_ids.index(utt_id))

    def read_utt_data_from_index(self, index):
        return self.read_ark(self.scp_data[index][0], self.scp_data[index][1])


class ArkWriter(object):
    """Writes float32 matrices as binary `FM` entries plus the scp index (io_funcs/kaldi_io.py:254-282)."""

    def __init__(self, scp_path, kaldi_separator=False):
        """kaldi_separator=False reproduces the reference byte for byte: the key is NOT followed by the
        space Kaldi's own archives carry (only the scp offsets make the entries addressable);
        True inserts it so that Kaldi binaries can also stream the .ark directly."""
        self.scp_path = scp_path
        self.kaldi_separator = kaldi_separator
        self.scp_file_write = open(self.scp_path, "w")

    def write_next_utt(self, ark_path, utt_id, utt_mat):
        utt_mat = np.ascontiguousarray(np.asarray(utt_mat, dtype=np.float32))
        rows, cols = utt_mat.shape
        key = utt_id.encode() if isinstance(utt_id, str) else bytes(utt_id)
        with open(ark_path, "ab") as f:
            # the reference writes the key with no separator and points the scp entry at the byte
            # after it (:271-272); struct '<xcccc' then emits the '\0' pad, so the entry reads as
            # key '\0' 'B' 'F' 'M' ' ' exactly as read_ark expects at `pos`.
            f.write(key + (b" " if self.kaldi_separator else b""))
            pos = f.tell()
            f.write(struct.pack("<xcccc", b"B", b"F", b"M", b" "))
            f.write(struct.pack("<bi", 4, rows))
            f.write(struct.pack("<bi", 4, cols))
            f.write(utt_mat.tobytes())
        self.scp_file_write.write("%s %s:%s\n" % (key.decode(), ark_path, pos))
        self.scp_file_write.flush()

    def close(self):
        self.scp_file_write.close()


def read_binary_file(filename, offset=0):
    """io_funcs/convert_cmvn_to_numpy.py:52-80: one uncompressed binary matrix."""
    return ArkReader().read_ark(filename, offset)


def cmvn_from_stats(stats):
    """Kaldi global CMVN stats (2, D+1): row 0 = sums | frame count, row 1 = sums of squares | 0
    -> (mean, stddev)   (io_funcs/convert_cmvn_to_numpy.py:29-41)."""
    stats = np.asarray(stats)
    n = stats[0][-1]
    s = np.hsplit(stats, [stats.shape[1] - 1])[0]
    mean = s[0] / n
    std = np.sqrt(s[1] / n - mean ** 2)
    return mean, std


def convert_cmvn_to_numpy(inputs_cmvn, labels_cmvn, save_dir):
    """io_funcs/convert_cmvn_to_numpy.py:18-49 -> <save_dir>/train_cmvn.npz with keys
    mean_inputs, stddev_inputs, mean_labels, stddev_labels."""
    print("Convert %s and %s to Numpy format" % (inputs_cmvn, labels_cmvn))
    mi, si = cmvn_from_stats(read_binary_file(inputs_cmvn, 0))
    ml, sl = cmvn_from_stats(read_binary_file(labels_cmvn, 0))
    cmvn_name = os.path.join(save_dir, "train_cmvn.npz")
    np.savez(cmvn_name, mean_inputs=mi, stddev_inputs=si, mean_labels=ml, stddev_labels=sl)
    print("Write to %s" % cmvn_name)
    return cmvn_name
