#!/bin/bash
# Driver of the GAN recipe on B200 -- the surface of the reference's run_gan_rnn_placeholder.sh:117-193 (stage 2: train,
# stage 3: decode) with the same variables, flags and values.  What differs, and why:
#   * stages 0-1 of the reference (convert_cmvn_to_numpy.py + make_tfrecords.py, :19-113) have no counterpart here: the
#     list files name Kaldi pair-scp files and CMVN is applied while loading (rsrgan_b200/dataset.py), so only
#     <train_dir>/train_cmvn.npz (io_funcs/convert_cmvn_to_numpy.py:43-47) has to exist;
#   * `CUDA_VISIBLE_DEVICES="2,3" python ... --num_gpu=2` (in-graph towers, :119-143) becomes one process per GPU:
#     `torchrun --nproc-per-node $num_gpu`; --num_gpu is taken from WORLD_SIZE;
#   * the decode stage reads the save_dir the training stage wrote (the reference decodes another experiment's
#     directory, exp/0124_..., :182).
# Stage 2 is two invocations on ONE save_dir, as in the reference: the first (1 epoch, d_learning_rate 0.001) writes a
# checkpoint that the second (18-20 epochs, d_learning_rate 0.0003) resumes -- every saved variable is restored and only
# the two learning rates are overwritten.
set -euo pipefail

stage=${stage:-2}
num_gpu=${num_gpu:-2}
train_dir=${train_dir:-data/train/train_100h_new}
test_dir=${test_dir:-data/test/test001-3000-real}
save_dir=${save_dir:-exp/0211_gan_res_lstm_l}
tr_list=$train_dir/tr.list
cv_list=$train_dir/cv.list
test_list=$test_dir/test.list
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
port=${MASTER_PORT:-29517}

train() {   # $1 = d_learning_rate, $2 = min_epoches, $3 = max_epoches
  python -m torch.distributed.run --nnodes=1 --nproc-per-node "$num_gpu" --master-addr 127.0.0.1 --master-port "$port" \
    "$here/scripts/train_gan_rnn_placeholder.py" \
      --data_dir=$train_dir \
      --tr_list_file=$tr_list \
      --cv_list_file=$cv_list \
      --g_type="res_lstm_l" \
      --save_dir=$save_dir \
      --batch_size=8 \
      --g_learning_rate=0.00008 \
      --d_learning_rate=$1 \
      --disc_updates=1 \
      --gen_updates=2 \
      --batch_norm=False \
      --l2_scale=0.0 \
      --init_mse_weight=10.0 \
      --input_dim=257 \
      --output_dim=40 \
      --left_context=0 \
      --right_context=0 \
      --min_epoches=$2 \
      --max_epoches=$3 \
      --end_improve=0.001 \
      --num_threads=32 \
      --init_disc_noise_std=0.05 \
      --num_gpu=$num_gpu
}

# Train model
if [ $stage -le 2 ]; then
  echo "$(date): $(hostname)"
  train 0.001 1 1 || exit 1
  train 0.0003 18 20 || exit 1
  echo "Finished training successfully on $(date)"
  echo ""
fi

# Decode
if [ $stage -le 3 ]; then
  echo "Start decoding test data"
  python "$here/scripts/train_gan_rnn_placeholder.py" \
      --decode \
      --data_dir=$train_dir \
      --test_list_file=$test_list \
      --g_type="res_lstm_l" \
      --save_dir=$save_dir \
      --batch_norm=False \
      --input_dim=257 \
      --output_dim=40 \
      --left_context=0 \
      --right_context=0 \
      --batch_size=1 \
      --keep_prob=1.0 \
      --l2_scale=0.0 \
      --num_threads=30 || exit 1
  echo "Decoding done"
fi

exit 0
