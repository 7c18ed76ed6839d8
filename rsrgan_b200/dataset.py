"""Minibatch pipeline with the contract of io_funcs/tfrecords_dataset.py:53-180 (`get_padded_batch`)
and :183-293 (`get_batch`, decode), reading Kaldi archives directly instead of TFRecords.

The reference converts ark -> CMVN -> tf.SequenceExample -> .tfrecords offline (io_funcs/
make_tfrecords.py:60-90) and batches with tf.data.  Here a *list file* (the `--tr_list_file`
argument) names one or more "pair scp" files whose lines are exactly the config lines that
make_tfrecords.py consumes (:62-69):

    <utt_id> <inputs.ark>:<offset> <labels.ark>:<offset>          (train / cv)
    <utt_id> <inputs.ark>:<offset>                                (test)

and the loader applies, per utterance, (x - mean) / stddev in float64 -> float32
(make_tfrecords.py:84-87) -- or, with cmvn_on_device=True, leaves the features RAW and the trainer object
normalises the padded minibatch on the GPU (GAN_RNN.set_cmvn -> rsr_cmvn_apply_padded: the same float64
arithmetic, bit-identical, padded frames exact zeros) --, frame splicing with edge replication
(tfrecords_dataset.py:76-99),
then the batching semantics of get_padded_batch:

  * shuffle buffer of 10 000 utterances (:128,151);
  * repeat `num_epochs` times (:173);
  * bucket key min(num_buckets, (len - 200) // 50) -- floor division, so short utterances get
    negative keys (:176-186); a bucket emits a batch as soon as it holds `batch_size` utterances;
    what is left in the buckets at the end is emitted as ragged batches (tf group_by_window),
    which the trainer skips (scripts/train_gan_rnn_placeholder.py:69-70);
  * zero padding to the longest utterance of the batch (:153-171); lengths returned separately.
"""
from __future__ import annotations

import queue
import random
import threading

import numpy as np

from .kaldi_io import ArkReader


def read_list(path):
    """utils/misc.py:27-35."""
    with open(path) as f:
        return [l.strip() for l in f if l.strip()]


def read_pair_scp(paths):
    out = []
    for p in paths:
        with open(p) as f:
            for line in f:
                parts = line.split()
                if not parts:
                    continue
                utt = parts[0]
                ins = parts[1].rsplit(":", 1)
                lab = parts[2].rsplit(":", 1) if len(parts) > 2 else None
                out.append((utt, ins, lab))
    return out


def splice_feats(feats, left, right):
    """io_funcs/tfrecords_dataset.py:76-99: [row, col] -> [row, col*(left+1+right)], repeating the first /
    last frame at the edges (tf.pad SYMMETRIC applied one row at a time == edge replication)."""
    if left == 0 and right == 0:
        return feats
    row = feats.shape[0]
    idx = np.arange(row)
    cols = [feats[np.clip(idx + o, 0, row - 1)] for o in range(-left, right + 1)]
    return np.concatenate(cols, 1)


class PaddedBatches(object):
    """Iterator of (utt_ids, inputs (B,T,D_in) fp32, labels (B,T,D_out) fp32 | None, lengths (B,) fp32)."""

    def __init__(self, scp_files, batch_size, input_size, output_size, left_context=0, right_context=0,
                 num_threads=4, num_epochs=1, num_buckets=20, cmvn=None, shuffle=True, infer=False, seed=None,
                 buffer_size=10000, cmvn_on_device=False):
        self.items = read_pair_scp(scp_files)
        self.batch_size, self.num_epochs, self.num_buckets = batch_size, num_epochs, num_buckets
        self.left, self.right = left_context, right_context
        self.input_size, self.output_size = input_size, output_size
        # materialise the four arrays once: np.load's NpzFile reads lazily from a zip and is not thread-safe
        self.cmvn = None if cmvn is None else {k: np.asarray(cmvn[k], np.float64) for k in
                                               ("mean_inputs", "stddev_inputs", "mean_labels", "stddev_labels")
                                               if k in cmvn}
        if cmvn_on_device:
            self.cmvn = None              # raw features out; GAN_RNN.set_cmvn(cmvn) normalises the fed batch on the GPU
        self.shuffle, self.infer = shuffle and not infer, infer
        self.buffer_size = buffer_size
        self.rng = random.Random(seed)
        self.reader = ArkReader()
        self.num_threads = max(1, num_threads)

    def _load(self, item):
        utt, ins, lab = item
        x = np.asarray(self.reader.read_ark(ins[0], ins[1]), np.float64)
        y = None if lab is None else np.asarray(self.reader.read_ark(lab[0], lab[1]), np.float64)
        if self.cmvn is not None:
            x = (x - self.cmvn["mean_inputs"]) / self.cmvn["stddev_inputs"]
            if y is not None:
                y = (y - self.cmvn["mean_labels"]) / self.cmvn["stddev_labels"]
        x = splice_feats(x.astype(np.float32), self.left, self.right)
        return utt, x, None if y is None else y.astype(np.float32)

    def _order(self):
        """Utterance order after tf.data's shuffle(buffer_size) over num_epochs passes."""
        epochs = self.num_epochs if self.num_epochs is not None else 1
        for _ in range(epochs):
            if not self.shuffle:
                for it in self.items:
                    yield it
                continue
            buf = []
            for it in self.items:
                buf.append(it)
                if len(buf) >= self.buffer_size:
                    yield buf.pop(self.rng.randrange(len(buf)))
            while buf:
                yield buf.pop(self.rng.randrange(len(buf)))

    def _pad(self, group):
        B = len(group)
        T = max(g[1].shape[0] for g in group)
        x = np.zeros((B, T, group[0][1].shape[1]), np.float32)
        y = None if group[0][2] is None else np.zeros((B, T, group[0][2].shape[1]), np.float32)
        lengths = np.zeros(B, np.float32)
        for i, (_, xi, yi) in enumerate(group):
            x[i, :xi.shape[0]] = xi
            if y is not None:
                y[i, :yi.shape[0]] = yi
            lengths[i] = xi.shape[0]
        return [g[0] for g in group], x, y, lengths

    def __iter__(self):
        # reader threads load + normalise utterances in order; batching happens on the consumer side
        order = list(self._order())
        results = [None] * len(order)
        done = [threading.Event() for _ in order]
        nxt = {"i": 0}
        lock = threading.Lock()
        window = threading.Semaphore(max(4 * self.batch_size, 64))

        def work():
            while True:
                window.acquire()
                with lock:
                    i = nxt["i"]
                    nxt["i"] += 1
                if i >= len(order):
                    return
                try:
                    results[i] = self._load(order[i])
                except BaseException as e:          # handed to the consumer, which re-raises (never a silent hang)
                    results[i] = e
                done[i].set()

        threads = [threading.Thread(target=work, daemon=True) for _ in range(self.num_threads)]
        for t in threads:
            t.start()
        buckets = {}
        for i in range(len(order)):
            done[i].wait()
            item, results[i] = results[i], None
            window.release()
            if isinstance(item, BaseException):
                raise RuntimeError("reading %s failed" % (order[i][0],)) from item
            if self.num_buckets > 1 and not self.infer:
                key = min(self.num_buckets, (item[1].shape[0] - 200) // 50)
            else:
                key = 0
            buckets.setdefault(key, []).append(item)
            if len(buckets[key]) == self.batch_size:
                yield self._pad(buckets.pop(key))
        for key in sorted(buckets):                      # ragged leftovers, emitted at end of data
            if buckets[key]:
                yield self._pad(buckets[key])
        for _ in threads:
            window.release()


def get_padded_batch(filenames, batch_size, input_size, output_size, left_context, right_context,
                     num_threads=4, num_epochs=1, num_buckets=20, cmvn=None, seed=None, cmvn_on_device=False):
    """Same argument list as io_funcs/tfrecords_dataset.py:53-55 (+ cmvn, seed, cmvn_on_device); returns an iterable."""
    return PaddedBatches(filenames, batch_size, input_size, output_size, left_context, right_context,
                         num_threads, num_epochs, num_buckets, cmvn=cmvn, seed=seed, cmvn_on_device=cmvn_on_device)


def get_batch(filenames, batch_size, input_size, output_size, left_context, right_context, num_threads=4,
              num_epochs=1, infer=False, cmvn=None, cmvn_on_device=False):
    """Decode-time reader (io_funcs/tfrecords_dataset.py:183-293 with infer=True): file order, no shuffle,
    no buckets; the trainer uses batch_size=1 (scripts/train_gan_rnn_placeholder.py:214-223)."""
    return PaddedBatches(filenames, batch_size, input_size, output_size, left_context, right_context,
                         num_threads, num_epochs, 1, cmvn=cmvn, shuffle=False, infer=infer, cmvn_on_device=cmvn_on_device)


class Prefetcher(object):
    """Background thread -> Queue(maxsize), the role of run_batch + Queue.Queue(32) in
    scripts/train_gan_rnn_placeholder.py:30-45,463-478 (joined properly, unlike the reference)."""

    def __init__(self, iterable, maxsize=32):
        self.q = queue.Queue(maxsize)
        self._end = object()

        def run():
            try:
                for b in iterable:
                    self.q.put(b)
            except BaseException as e:              # re-raised in the consumer
                self.q.put(e)
            finally:
                self.q.put(self._end)

        self.t = threading.Thread(target=run, daemon=True)
        self.t.start()

    def __iter__(self):
        while True:
            b = self.q.get()
            if b is self._end:
                return
            if isinstance(b, BaseException):
                raise b
            yield b
