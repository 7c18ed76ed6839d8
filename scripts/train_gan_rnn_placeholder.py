#!/usr/bin/env python
"""Train / decode the RSRGAN GAN on B200 -- drop-in for the reference's
scripts/train_gan_rnn_placeholder.py (same flags and defaults :587-746, same epoch loop,
LR / noise decay, accept / reject checkpointing :388-584, same decode :204-302, same stdout
loss lines :498-521 that utils/generate_plots.py parses).

Differences that are deliberate:
  * one process per GPU: launch with `torchrun --nproc-per-node N` instead of `--num_gpu N` towers
    inside one process; `--num_gpu` is taken from WORLD_SIZE when run under torchrun;
  * list files name Kaldi pair-scp files instead of TFRecords (rsrgan_b200/dataset.py); CMVN
    (train_cmvn.npz) is applied while loading, so stages 0-1 need no TFRecord conversion;
  * extra flags (--d_type, --g_cell, --g_proj, --g_layers, --dtype) expose what the reference
    hard-codes in its model files; defaults are the reference values.
"""
from __future__ import annotations

import argparse
import datetime
import os
import pprint
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200.dataset import Prefetcher, get_batch, get_padded_batch, read_list  # noqa: E402
from rsrgan_b200.kaldi_io import ArkWriter  # noqa: E402

FLAGS = None


def str2bool(v):
    """utils/misc.py:43-49."""
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def exponential_decay(iteration, num_jobs, num_iters, init_lr, multiply_jobs=True):
    """utils/ops.py:378-391."""
    final_lr = 0.0001 * init_lr
    if iteration + 1 >= num_iters:
        current_lr = final_lr
    else:
        current_lr = init_lr * np.exp(iteration * np.log(final_lr / init_lr) / num_iters)
    return num_jobs * current_lr if multiply_jobs else current_lr


def load_cmvn():
    path = os.path.join(FLAGS.data_dir, "train_cmvn.npz") if FLAGS.data_dir else None
    if path and os.path.isfile(path):
        with np.load(path) as z:
            return {k: z[k] for k in z.files}
    return None


def rank():
    return int(os.environ.get("RANK", "0"))


def log(*a):
    if rank() == 0:
        print(*a)
        sys.stdout.flush()


def world_mean(model, vals):
    """np.mean over towers of the per-tower losses (train...py:85-87,104-107) == mean over ranks."""
    if model.world == 1:
        return vals
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=model.h.device)
    model.dist.all_reduce(t)
    return (t / model.world).tolist()


def _train_and_prefetch(model, cur, nxt):
    """Enqueues the schedule for `cur`, starts the upload of `nxt`, then reads the losses back."""
    d_last, g_last = model.train_batch(cur[1], cur[2], cur[3], sync=False)
    model.prefetch(nxt[1], nxt[2], nxt[3])
    return model.last_update_losses()


def train_one_iteration(model, batches, iteration, tr_num_batch=0):
    """scripts/train_gan_rnn_placeholder.py:48-133."""
    sums = np.zeros(7)
    d_counter = g_counter = batch_counter = 0
    model.d_real, model.d_fake = 1.0, 0.0
    # One call per minibatch runs the whole schedule of :72-101 (disc_updates x D, gen_updates x G on the same
    # batch) on the device; the next minibatch is uploaded meanwhile (GAN_RNN.prefetch).
    full = (b for b in batches if b[1].shape[0] == FLAGS.batch_size)     # ragged tail batches are skipped (:69-70)
    cur = next(full, None)
    if cur is not None:
        model.prefetch(cur[1], cur[2], cur[3])
    while cur is not None:
        nxt = next(full, None)
        _, inputs, labels, lengths = cur
        d_list, g_list = model.train_batch(inputs, labels, lengths, all_updates=True) if nxt is None else \
            _train_and_prefetch(model, cur, nxt)
        for d in d_list:
            d_counter += 1
            sums[0:3] += world_mean(model, [d["d_rl_loss"], d["d_fk_loss"], d["d_loss"]])
        for g in g_list:
            g_counter += 1
            sums[3:7] += world_mean(model, [g["g_adv_loss"], g["g_mse_loss"], g["g_l2_loss"], g["g_loss"]])
        if batch_counter % 100 == 0 and d_list and g_list:
            # `if batch % 100 == 0: ... add_summary(_summaries, iteration*tr_num_batch)` (:116-122): batches 0, 100, ...
            model.write_summaries(dict(d_list[-1], **g_list[-1]), iteration * tr_num_batch)
        if batch_counter % 100 == 99:
            model.check_overflow()       # fp16 operands: skipped updates -> halve the loss scale (no reference counterpart)
        batch_counter += 1
        cur = nxt
    model.check_overflow()
    d_counter, g_counter = max(d_counter, 1), max(g_counter, 1)
    return tuple(sums[0:3] / d_counter) + tuple(sums[3:7] / g_counter)


def eval_one_iteration(model, batches, iteration):
    """scripts/train_gan_rnn_placeholder.py:136-201."""
    sums = np.zeros(7)
    n = 0
    model.d_real, model.d_fake = 1.0, 0.0
    for _, inputs, labels, lengths in batches:
        if inputs.shape[0] != FLAGS.batch_size:
            continue
        e = model.eval_losses(inputs, labels, lengths)
        sums += world_mean(model, [e[k] for k in ("d_rl_loss", "d_fk_loss", "d_loss", "g_adv_loss", "g_mse_loss",
                                                   "g_l2_loss", "g_loss")])
        n += 1
    return tuple(sums / max(n, 1))


def tower_slice(batches):
    """Every rank walks the SAME stream of global batches (batch_size * num_gpu utterances, same seed)
    and keeps its own tower slice -- the reference's per-tower slicing of the fed batch
    (models/gan_rnn_placeholder.py:157-159).  Ragged global batches are dropped on every rank alike
    (train...py:69-70), so all ranks perform the same number of updates (one all-reduce each)."""
    B, r = FLAGS.batch_size, rank()
    for ids, x, y, ln in batches:
        if x.shape[0] != B * FLAGS.num_gpu:
            continue
        sl = slice(r * B, (r + 1) * B)
        yield ids[sl], x[sl], None if y is None else y[sl], ln[sl]


def decode():
    """scripts/train_gan_rnn_placeholder.py:204-302."""
    from rsrgan_b200.gan_rnn import GAN_RNN
    data_list = read_list(FLAGS.test_list_file)
    cmvn = load_cmvn()
    if cmvn is None:
        print("%s not exist, exit now." % os.path.join(str(FLAGS.data_dir), "train_cmvn.npz"))
        sys.exit(1)
    model = GAN_RNN(None, FLAGS, ["/gpu:%d" % int(os.environ.get("LOCAL_RANK", "0"))], cross_validation=True, infer=True)
    if model.load(model.save_dir, moving_average=False):
        print("[*] Load SUCCESS")
    else:
        print("[!] Load failed. Checkpoint not found. Exit now.")
        sys.exit(1)
    out_dir_name = os.path.join(FLAGS.save_dir, "test")
    os.makedirs(out_dir_name, exist_ok=True)
    write_scp_path = os.path.join(out_dir_name, "feats.scp")
    write_ark_path = os.path.join(out_dir_name, "feats.ark")
    if os.path.exists(write_ark_path):
        os.remove(write_ark_path)
    writer = ArkWriter(write_scp_path)
    dev_cmvn = FLAGS.device_cmvn
    if dev_cmvn:
        model.set_cmvn(cmvn)          # (x - mean) / stddev of make_tfrecords.py:84-87 on the GPU, per fed utterance
    batches = list(get_batch(data_list, 1, FLAGS.input_dim, FLAGS.output_dim, FLAGS.left_context,
                             FLAGS.right_context, FLAGS.num_threads, 1, infer=True, cmvn=cmvn, cmvn_on_device=dev_cmvn))
    start = datetime.datetime.now()
    mean, std = cmvn["mean_labels"], cmvn["stddev_labels"]
    for i, (ids, inputs, _, lengths) in enumerate(batches):
        # activations * stddev_labels + mean_labels (:286-287) fused into the device un-staging
        seq = model.generate(inputs, lengths, mean=mean, std=std).cpu().numpy()
        writer.write_next_utt(write_ark_path, ids[0], np.vstack(seq))
        print("[{}/{}] Write inferred {} to {}".format(i + 1, len(batches), ids[0], write_ark_path))
    writer.close()
    print("Decoding time is {}s".format((datetime.datetime.now() - start).total_seconds()))
    sys.stdout.flush()


def count_batches(files, cmvn):
    n = 0
    for _ in get_padded_batch(files, FLAGS.batch_size, FLAGS.input_dim, FLAGS.output_dim, 0, 0,
                              FLAGS.num_threads * 2, 1, cmvn=None, seed=0):
        n += 1
    return n


def main():
    if FLAGS.decode:
        return decode()
    if "WORLD_SIZE" in os.environ and int(os.environ["WORLD_SIZE"]) > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
        FLAGS.num_gpu = dist.get_world_size()
    from rsrgan_b200.gan_rnn import GAN_RNN
    cmvn = load_cmvn()
    tr_files, cv_files = read_list(FLAGS.tr_list_file), read_list(FLAGS.cv_list_file)
    # batch counts are cached in <data_dir>/batch_num_sentence_<B>.txt (:307-323)
    filename = "batch_num_sentence_%s.txt" % FLAGS.batch_size
    batch_file = os.path.join(FLAGS.data_dir, filename)
    if os.path.isfile(batch_file):
        with open(batch_file) as fr:
            cv_num_batch, tr_num_batch = (int(v) for v in fr.readline().split()[:2])
        log("LOG: %s exist, cross validation batches is %d, trian batches is %d." % (filename, cv_num_batch, tr_num_batch))
    else:
        log("Get CV set batch numbers.")
        cv_num_batch = count_batches(cv_files, cmvn)
        log("Get Train set batch numbers.")
        tr_num_batch = count_batches(tr_files, cmvn)
        if rank() == 0:
            with open(batch_file, "w") as fw:
                fw.write("%d %d" % (cv_num_batch, tr_num_batch))
    min_iters, max_iters = int(FLAGS.min_epoches), int(FLAGS.max_epoches)
    log("\nLOG: #train_batch = {}, #valid_batch = {}\nLOG: #min_epoches = {}, #max_epoches = {}\n"
        "LOG: #min_iters = {}, #max_iters = {}\n".format(tr_num_batch, cv_num_batch, FLAGS.min_epoches,
                                                         FLAGS.max_epoches, min_iters, max_iters))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    log("=======================================================")
    log("|                Build Train model                    |")
    log("=======================================================")
    tr_model = GAN_RNN(None, FLAGS, ["/gpu:%d" % local], cross_validation=False)
    log("=======================================================")
    log("|           Build Cross-Validation model              |")
    log("=======================================================")
    cv_model = GAN_RNN(None, FLAGS, ["/gpu:%d" % local], cross_validation=True, share=tr_model)
    if tr_model.load(tr_model.save_dir, moving_average=False):
        log("[*] Load SUCCESS")
    else:
        log("[!] Begin a new model.")
    dev_cmvn = FLAGS.device_cmvn
    if dev_cmvn:                      # the loader hands out raw features; the fed minibatch is normalised on the GPU
        tr_model.set_cmvn(cmvn)
        cv_model.set_cmvn(cmvn)
    g_loss_prev, g_rel_impr, check_interval, windows_g_loss = 10000.0, 1.0, 1, []
    tr_model.g_learning_rate = FLAGS.num_gpu * FLAGS.g_learning_rate       # :458-461
    tr_model.d_learning_rate = FLAGS.num_gpu * FLAGS.d_learning_rate
    iteration = -1
    for iteration in range(max_iters):
        gB = FLAGS.batch_size * FLAGS.num_gpu                     # :395 batch = batch_size * num_gpu
        tr_batches = Prefetcher(tower_slice(get_padded_batch(tr_files, gB, FLAGS.input_dim, FLAGS.output_dim,
                                                             FLAGS.left_context, FLAGS.right_context,
                                                             FLAGS.num_threads, 1, cmvn=cmvn, seed=1000 + iteration,
                                                             cmvn_on_device=dev_cmvn)))
        cv_batches = Prefetcher(tower_slice(get_padded_batch(cv_files, gB, FLAGS.input_dim, FLAGS.output_dim,
                                                             FLAGS.left_context, FLAGS.right_context,
                                                             FLAGS.num_threads, 1, cmvn=cmvn, seed=7,
                                                             cmvn_on_device=dev_cmvn)))
        start = datetime.datetime.now()
        tr = train_one_iteration(tr_model, tr_batches, iteration + 1, tr_num_batch)
        cv = eval_one_iteration(cv_model, cv_batches, iteration + 1)
        end = datetime.datetime.now()
        fmt = ("d_rl_loss = {:.5f}, d_fk_loss = {:.5f}, d_loss = {:.5f}, g_adv_loss = {:.5f}, "
               "g_mse_loss = {:.5f}, g_l2_loss = {:.3e}, g_loss = {:.5f}")
        log("{}/{} (INFO): d_learning_rate = {:.5e}, g_learning_rate = {:.5e}, time = {:.3f} h\n"
            "{}/{} (TRAIN AVG.LOSS): {}\n{}/{} (CROSS AVG.LOSS): {}".format(
                iteration + 1, max_iters, tr_model.d_learning_rate, tr_model.g_learning_rate,
                (end - start).total_seconds() / 3600.0, iteration + 1, max_iters, fmt.format(*tr),
                iteration + 1, max_iters, fmt.format(*cv)))
        cv_g_loss = cv[6]
        # decay (:525-533); only the TRAIN model's noise std decays
        tr_model.g_learning_rate = exponential_decay(iteration + 1, FLAGS.num_gpu, min_iters, FLAGS.g_learning_rate)
        tr_model.d_learning_rate = exponential_decay(iteration + 1, FLAGS.num_gpu, min_iters, FLAGS.d_learning_rate)
        tr_model.disc_noise_std = exponential_decay(iteration + 1, FLAGS.num_gpu, min_iters,
                                                    FLAGS.init_disc_noise_std, multiply_jobs=False)
        windows_g_loss.append(cv_g_loss)
        if (iteration + 1) % check_interval == 0:          # accept / reject (:538-554): only decides whether to save
            g_loss_new = float(np.mean(windows_g_loss))
            g_rel_impr = (g_loss_prev - g_loss_new) / g_loss_prev
            if g_rel_impr > 0.0:
                if rank() == 0:
                    tr_model.save(tr_model.save_dir, iteration + 1)
                log("Iteration {}: Nnet Accepted. Save model SUCCESS. g_loss_prev = {:.5f}, g_loss_new = {:.5f}".format(
                    iteration + 1, g_loss_prev, g_loss_new))
                g_loss_prev = g_loss_new
            else:
                log("Iteration {}: Nnet Rejected. g_loss_prev = {:.5f}, g_loss_new = {:.5f}".format(
                    iteration + 1, g_loss_prev, g_loss_new))
            windows_g_loss = []
        if iteration + 1 > min_iters and (iteration + 1) % check_interval == 0:
            if g_rel_impr < FLAGS.end_improve:
                log("Iteration %d: Finished, too small relative G improvement %g" % (iteration + 1, g_rel_impr))
                break
    if windows_g_loss:
        g_loss_new = float(np.mean(windows_g_loss))
        if (g_loss_prev - g_loss_new) / g_loss_prev > 0.0 and rank() == 0:
            tr_model.save(tr_model.save_dir, iteration + 1)
    log("Training Done.")


def build_parser():
    p = argparse.ArgumentParser()
    p.add_argument("--decode", default=False, action="store_true", help="Flag indicating decoding or training.")
    p.add_argument("--data_dir", type=str, default=None, help="Data directory.")
    p.add_argument("--tr_list_file", type=str, default=None, help="Train set list file (pair-scp files).")
    p.add_argument("--cv_list_file", type=str, default=None, help="Validation set list file.")
    p.add_argument("--test_list_file", type=str, default=None, help="Test set list file.")
    p.add_argument("--input_dim", type=int, default=257, help="The dimension of input.")
    p.add_argument("--output_dim", type=int, default=40, help="The dimension of output.")
    p.add_argument("--left_context", type=int, default=5, help="The number of left context to be added to inputs.")
    p.add_argument("--right_context", type=int, default=5, help="The number of right context to be added to inputs.")
    p.add_argument("--batch_size", type=int, default=32, help="Mini-batch size.")
    p.add_argument("--g_learning_rate", type=float, default=0.0003, help="Initial G learning rate.")
    p.add_argument("--d_learning_rate", type=float, default=0.001, help="Initial D learning rate.")
    p.add_argument("--min_epoches", type=int, default=25, help="Min number of epoches to run trainer without decay.")
    p.add_argument("--max_epoches", type=int, default=30, help="Max number of epoches to run trainer totally.")
    p.add_argument("--end_improve", type=float, default=0.001, help="Stop when relative loss is lower than end_improve.")
    p.add_argument("--num_threads", type=int, default=24, help="The num of threads to read data.")
    p.add_argument("--save_dir", type=str, default="exp/gan_rnn", help="Directory to put the train result.")
    p.add_argument("--init_mse_weight", type=float, default=1.0, help="Init MSE loss lambda.")
    p.add_argument("--g_type", type=str, default="lstm", help="Type of G to use: lstm | res_lstm_l | res_lstm_base")
    p.add_argument("--disc_updates", type=int, default=1, help="Number of D step in a training iteration.")
    p.add_argument("--gen_updates", type=int, default=2, help="Number of G step in a training iteration.")
    p.add_argument("--batch_norm", type=str2bool, nargs="?", default="false", help="Whether use batch normalization.")
    p.add_argument("--keep_prob", type=float, default=1.0, help="The probability that each element is kept for dropout.")
    p.add_argument("--ckpt_format", type=str, default="pt", choices=["pt", "tf"],
                   help="Checkpoint container: torch state dict, or TensorFlow checkpoint-V2 bundles as the reference's "
                        "Saver writes them (either kind is read on resume / decode).")
    p.add_argument("--init_disc_noise_std", type=float, default=0.0, help="Noise std for discriminator.")
    p.add_argument("--l2_scale", type=float, default=0.00001, help="Scale used for L2 regularizer.")
    p.add_argument("--num_gpu", type=int, default=1, help="Number of GPU to use (= WORLD_SIZE under torchrun).")
    # sizes the reference hard-codes in models/*.py (defaults = reference values)
    p.add_argument("--d_type", type=str, default="lstm", help="lstm (discriminator_lstm) | dnn (discriminator_dnn)")
    p.add_argument("--g_cell", type=int, default=None)
    p.add_argument("--g_proj", type=int, default=None)
    p.add_argument("--g_layers", type=int, default=None)
    p.add_argument("--g_units", type=int, default=None, help="hidden width of the dnn generator (models/dnn.py:34)")
    p.add_argument("--d_cell", type=int, default=None)
    p.add_argument("--d_proj", type=int, default=None)
    p.add_argument("--d_layers", type=int, default=None)
    p.add_argument("--d_units", type=int, default=None, help="hidden width of discriminator_dnn (:23)")
    p.add_argument("--dtype", type=str, default="f16", help="tensor-core operand type: f16 | bf16")
    p.add_argument("--device_cmvn", type=str2bool, nargs="?", default="true",
                   help="apply the global CMVN to the fed minibatch on the GPU (bit-identical to the host loader's)")
    p.add_argument("--seed", type=int, default=1234)
    return p


if __name__ == "__main__":
    pp = pprint.PrettyPrinter()
    FLAGS, unparsed = build_parser().parse_known_args()
    if rank() == 0:
        print("*** Parsed arguments ***")
        pp.pprint(FLAGS.__dict__)
        sys.stdout.flush()
    main()
