#!/usr/bin/env python
"""Convert checkpoints between the two containers GAN_RNN / DNNTrainer / GAN read and write:

    <dir>/<name>-<step>.pt                               torch.save of the state dict ({TF variable name -> array})
    <dir>/<name>-<step>.{index,data-00000-of-00001}      TensorFlow checkpoint-V2 tensor bundle -- what the reference's
                                                         tf.train.Saver writes (models/gan_rnn_placeholder.py:26-60)

    python scripts/convert_checkpoint.py --to tf  exp/gan/GAN_RNN-12.pt      -> exp/gan/GAN_RNN-12.{index,data-...}
    python scripts/convert_checkpoint.py --to pt  exp/ref/GAN_RNN-12 --like exp/gan/GAN_RNN-3.pt
    python scripts/convert_checkpoint.py --list   exp/ref/GAN_RNN-12         (names, dtypes, shapes in the bundle)

`--to pt` needs `--like`: a .pt checkpoint of a model with the SAME architecture (it supplies the state-dict skeleton;
its weights are replaced, optimizer slots / EMA shadows / batch_norm statistics are taken from the bundle when present).
Host-only: no GPU is touched (the CRC-32C routine of the C-ABI library runs on the CPU).
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import tf_checkpoint as T  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("path", help="a .pt file, or the prefix of a TensorFlow bundle (without .index)")
    ap.add_argument("--to", choices=["tf", "pt"])
    ap.add_argument("--like", help="(--to pt) .pt checkpoint of the same architecture")
    ap.add_argument("--out", help="output path / prefix (default: next to the input)")
    ap.add_argument("--list", action="store_true", help="print the variables of a TensorFlow bundle")
    a = ap.parse_args(argv)
    if a.list:
        for name, v in T.read_bundle(a.path).items():
            print("%-90s %-8s %s" % (name, v.dtype, tuple(v.shape)))
        return 0
    if a.to == "tf":
        sd = torch.load(a.path, map_location="cpu", weights_only=False)
        prefix = a.out or a.path[:-3] if a.path.endswith(".pt") else (a.out or a.path + ".tf")
        tensors = T.state_to_tensors(sd)
        T.write_bundle(prefix, tensors)
        # a directory that already has a `checkpoint` index (the trainer's max_to_keep rotation list, in either
        # container) keeps it: the converted bundle is an extra file, not the new "latest"
        d = os.path.dirname(os.path.abspath(prefix))
        if not os.path.exists(os.path.join(d, "checkpoint")):
            T.write_checkpoint_state(d, os.path.basename(prefix), [os.path.basename(prefix)])
        print("wrote %s.index / .data-00000-of-00001 (%d variables)" % (prefix, len(tensors)))
        return 0
    if a.to == "pt":
        if not a.like:
            ap.error("--to pt needs --like <checkpoint of the same architecture>")
        sd = torch.load(a.like, map_location="cpu", weights_only=False)
        missing = T.tensors_to_state(T.read_bundle(a.path), sd)
        out = a.out or a.path + ".pt"
        torch.save(sd, out)
        print("wrote %s (%d optimizer / average variables were not in the bundle)" % (out, len(missing)))
        return 0
    ap.error("nothing to do: give --to tf, --to pt or --list")


if __name__ == "__main__":
    sys.exit(main())
