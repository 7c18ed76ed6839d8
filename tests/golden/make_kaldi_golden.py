"""Generates tests/golden/kaldi_*.{ark,npz}: small Kaldi archives (float, double, compressed) and
what the REFERENCE's own reader (io_funcs/kaldi_io.py ArkReader.read_ark, imported from
/root/reference -- it runs under Python 3) returns for them.  Run once in the build container:

    python tests/golden/make_kaldi_golden.py

The .ark files are written with plain struct.pack here (not with the code under test)."""
import importlib.util
import os
import struct

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
spec = importlib.util.spec_from_file_location("ref_kaldi_io", "/root/reference/io_funcs/kaldi_io.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(42)
ark = os.path.join(HERE, "kaldi_small.ark")
entries = []
with open(ark, "wb") as f:
    def put(key, body):
        f.write(key.encode() + b" ")
        entries.append((key, f.tell()))
        f.write(body)
    m = rng.standard_normal((7, 40)).astype(np.float32)
    put("utt_fm", b"\0BFM " + struct.pack("<bi", 4, 7) + struct.pack("<bi", 4, 40) + m.tobytes())
    d = rng.standard_normal((5, 13))
    put("utt_dm", b"\0BDM " + struct.pack("<bi", 4, 5) + struct.pack("<bi", 4, 13) + d.tobytes())
    rows, cols = 11, 6
    hdr = np.sort(rng.integers(0, 65536, size=(cols, 4)), axis=1).astype("<u2")
    data = rng.integers(0, 256, size=(cols, rows)).astype(np.uint8)
    data[0, :4] = [0, 63, 64, 192]
    data[1, :3] = [193, 255, 128]
    put("utt_cm", b"\0BCM " + struct.pack("<ffii", -3.25, 9.5, rows, cols) + hdr.tobytes() + data.tobytes())
    e = np.zeros((1, 1), np.float32)
    put("utt_1x1", b"\0BFM " + struct.pack("<bi", 4, 1) + struct.pack("<bi", 4, 1) + e.tobytes())
with open(os.path.join(HERE, "kaldi_small.scp"), "w") as f:
    for k, pos in entries:
        f.write("%s kaldi_small.ark:%d\n" % (k, pos))
reader = ref.ArkReader()
out = {k: np.asarray(reader.read_ark(ark, pos)) for k, pos in entries}
np.savez(os.path.join(HERE, "kaldi_small_expected.npz"), **out)
# global CMVN stats file as Kaldi's compute-cmvn-stats writes it (double matrix 2 x (D+1))
D, n = 5, 1000.0
x = rng.standard_normal((int(n), D)) * 2 + 1
stats = np.zeros((2, D + 1))
stats[0, :D], stats[0, D], stats[1, :D] = x.sum(0), n, (x * x).sum(0)
with open(os.path.join(HERE, "global.cmvn"), "wb") as f:
    f.write(b"\0BDM " + struct.pack("<bi", 4, 2) + struct.pack("<bi", 4, D + 1) + stats.tobytes())
# the formulas of io_funcs/convert_cmvn_to_numpy.py:29-41 (the script itself only runs under python2)
mean = stats[0, :D] / n
np.savez(os.path.join(HERE, "global_cmvn_expected.npz"), mean=mean, std=np.sqrt(stats[1, :D] / n - mean ** 2))
print("wrote", sorted(os.listdir(HERE)))

# A full-utterance compressed matrix (500 frames x 257 bins): the archive bytes are regenerated from the seed by the
# tests (tests/test_kaldi_io.py::cm_utterance_bytes), only the SHA-256 of the float64 matrix the REFERENCE reader
# returns -- and a few of its rows -- are committed.
import hashlib  # noqa: E402
import sys  # noqa: E402

sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from test_kaldi_io import cm_utterance_bytes  # noqa: E402

body, (rows, cols) = cm_utterance_bytes()
big = os.path.join(HERE, "_cm_utt.ark")
with open(big, "wb") as f:
    f.write(b"utt_big ")
    pos = f.tell()
    f.write(body)
m = np.asarray(ref.ArkReader().read_ark(big, pos))
os.remove(big)
assert m.shape == (rows, cols) and m.dtype == np.float64
np.savez(os.path.join(HERE, "kaldi_cm_utt_expected.npz"), sha256=np.frombuffer(hashlib.sha256(m.tobytes()).digest(), np.uint8),
         rows_0_249_499=m[[0, 249, 499]], shape=np.array(m.shape))
print("reference reader: utterance-sized CM matrix", m.shape, hashlib.sha256(m.tobytes()).hexdigest())
