"""TensorBoard event files for the loss scalars -- the `tf.summary.FileWriter(save_dir/{train,eval})` +
`scalar_summary` surface of models/gan_rnn_placeholder.py:82-86,270-298, written every 100 batches by
scripts/train_gan_rnn_placeholder.py:117-122 (`sess.run(model.summaries)`; `writer.add_summary(summary, counter)`).

File format (TensorFlow's record writer): every record is
    uint64 length | uint32 masked_crc32c(length) | bytes data | uint32 masked_crc32c(data)
and data is a serialized `Event` proto: the first one carries file_version "brain.Event:2", the others
{wall_time, step, summary{value{tag, simple_value}}}.  Checksums and the protobuf encoding are shared with
rsrgan_b200/tf_checkpoint.py; tests read the files back with TensorBoard's own loader.  Histogram summaries of the
reference (weights, logits) are not written.
"""
from __future__ import annotations

import os
import socket
import struct
import time

from .tf_checkpoint import _field, crc32c, mask_crc, put_varint


def _record(data):
    head = struct.pack("<Q", len(data))
    return head + struct.pack("<I", mask_crc(crc32c(head))) + data + struct.pack("<I", mask_crc(crc32c(data)))


def _event(wall_time, step=None, file_version=None, scalars=None):
    out = _field(1, 1, struct.pack("<d", wall_time))
    if step:
        out += _field(2, 0, put_varint(int(step)))
    if file_version is not None:
        v = file_version.encode()
        out += _field(3, 2, put_varint(len(v)) + v)
    if scalars:
        summary = b""
        for tag, value in scalars.items():
            t = tag.encode()
            val = _field(1, 2, put_varint(len(t)) + t) + _field(2, 5, struct.pack("<f", float(value)))
            summary += _field(1, 2, put_varint(len(val)) + val)
        out += _field(5, 2, put_varint(len(summary)) + summary)
    return out


class FileWriter(object):
    """tf.summary.FileWriter(logdir): events.out.tfevents.<time>.<host> in `logdir`."""

    def __init__(self, logdir, graph=None):
        os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir
        self.path = os.path.join(logdir, "events.out.tfevents.%010d.%s" % (int(time.time()), socket.gethostname()))
        self._f = open(self.path, "ab")
        self._f.write(_record(_event(time.time(), file_version="brain.Event:2")))
        self._f.flush()

    def add_scalars(self, scalars, global_step):
        """One Event holding every (tag -> value) of `scalars` -- what add_summary(merged_summary, step) writes."""
        self._f.write(_record(_event(time.time(), step=global_step, scalars=scalars)))

    add_summary = add_scalars

    def flush(self):
        self._f.flush()

    def close(self):
        if not self._f.closed:
            self._f.close()
