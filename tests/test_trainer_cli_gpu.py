"""The driver surface of run_gan_rnn_placeholder.sh stage 2/3 end to end on the GPU: Kaldi scp/ark in,
`TRAIN/CROSS AVG.LOSS` log lines (the format utils/generate_plots.py:126-158 parses), checkpoint + resume,
`--decode` writing <save_dir>/test/feats.{scp,ark} with the inverse CMVN applied
(scripts/train_gan_rnn_placeholder.py:204-302)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCRIPT = os.path.join(ROOT, "scripts", "train_gan_rnn_placeholder.py")


def _make_data(d, n_utt=12, seed=0):
    from rsrgan_b200.kaldi_io import ArkWriter
    rng = np.random.default_rng(seed)
    wi, wl = ArkWriter(os.path.join(d, "inputs.scp")), ArkWriter(os.path.join(d, "labels.scp"))
    lens = []
    for u in range(n_utt):
        T = int(rng.integers(30, 60))
        lens.append(T)
        wi.write_next_utt(os.path.join(d, "inputs.ark"), "utt%02d" % u, rng.standard_normal((T, 257)) * 2 + 1)
        wl.write_next_utt(os.path.join(d, "labels.ark"), "utt%02d" % u, rng.standard_normal((T, 40)) * 3 - 1)
    wi.close(); wl.close()
    ins = [l.split() for l in open(os.path.join(d, "inputs.scp"))]
    labs = [l.split() for l in open(os.path.join(d, "labels.scp"))]
    with open(os.path.join(d, "tr.scp"), "w") as f:
        for a, b in zip(ins[:8], labs[:8]):
            f.write("%s %s %s\n" % (a[0], a[1], b[1]))
    with open(os.path.join(d, "cv.scp"), "w") as f:
        for a, b in zip(ins[8:], labs[8:]):
            f.write("%s %s %s\n" % (a[0], a[1], b[1]))
    with open(os.path.join(d, "test.scp"), "w") as f:
        for a in ins[8:10]:
            f.write("%s %s\n" % (a[0], a[1]))
    for name in ("tr", "cv", "test"):
        with open(os.path.join(d, name + ".list"), "w") as f:
            f.write(os.path.join(d, name + ".scp") + "\n")
    np.savez(os.path.join(d, "train_cmvn.npz"), mean_inputs=np.full(257, 1.0), stddev_inputs=np.full(257, 2.0),
             mean_labels=np.full(40, -1.0), stddev_labels=np.full(40, 3.0))
    return lens


def _run(args):
    r = subprocess.run([sys.executable, SCRIPT] + args, capture_output=True, text=True, cwd=ROOT, timeout=150)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


@pytest.mark.parametrize("extra", [[], ["--batch_norm", "true", "--keep_prob", "0.9", "--ckpt_format", "tf",
                                        "--d_type", "dnn", "--d_units", "128"]],
                         ids=["plain", "batch_norm-dropout-tf_checkpoints"])
def test_train_resume_decode(tmp_path, extra):
    """`extra`: the optional pieces of the same driver surface -- --batch_norm / --keep_prob (first-layer batch_norm and
    DropoutWrapper in the lstm generator, batch_norm + dropout in discriminator_dnn) and TensorFlow checkpoint-V2
    bundles as the container that is saved, resumed from and decoded from."""
    d = str(tmp_path)
    lens = _make_data(d)
    save = os.path.join(d, "exp")
    common = ["--data_dir", d, "--save_dir", save, "--batch_size", "4", "--left_context", "0", "--right_context", "0",
              "--g_type", "lstm", "--d_type", "lstm", "--g_cell", "256", "--g_proj", "64", "--g_layers", "1",
              "--init_mse_weight", "10.0", "--init_disc_noise_std", "0.05", "--num_threads", "2", "--l2_scale", "0"]
    common += extra
    out = _run(common + ["--tr_list_file", os.path.join(d, "tr.list"), "--cv_list_file", os.path.join(d, "cv.list"),
                         "--min_epoches", "2", "--max_epoches", "2"])
    pat = re.compile(r"(\d+)/(\d+) \((TRAIN|CROSS) AVG\.LOSS\): d_rl_loss = ([-\d.e+]+), d_fk_loss = ([-\d.e+]+), "
                     r"d_loss = ([-\d.e+]+), g_adv_loss = ([-\d.e+]+), g_mse_loss = ([-\d.e+]+), "
                     r"g_l2_loss = ([-\d.e+]+), g_loss = ([-\d.e+]+)")
    lines = pat.findall(out)
    assert [(l[0], l[2]) for l in lines] == [("1", "TRAIN"), ("1", "CROSS"), ("2", "TRAIN"), ("2", "CROSS")], out[-2000:]
    for l in lines:
        vals = [float(v) for v in l[3:]]
        assert all(np.isfinite(vals)) and vals[2] == pytest.approx(vals[0] + vals[1], rel=1e-3, abs=1e-4)
    assert os.path.exists(os.path.join(save, "checkpoint"))
    # TensorBoard scalars of the first batch of each iteration (train...py:116-122) in <save_dir>/train
    ev = [f for f in os.listdir(os.path.join(save, "train")) if f.startswith("events.out.tfevents.")]
    assert len(ev) == 1 and os.path.getsize(os.path.join(save, "train", ev[0])) > 200
    if "tf" in extra:
        from rsrgan_b200 import tf_checkpoint
        latest, _ = tf_checkpoint.read_checkpoint_state(save)
        names = tf_checkpoint.read_bundle(os.path.join(save, latest))
        assert "g_model/fully_connected/BatchNorm/gamma" in names and "d_model/fully_connected/BatchNorm/moving_mean" in names
    # second invocation on the same save_dir resumes (run_gan_rnn_placeholder.sh runs the script twice, :117-168)
    out2 = _run(common + ["--tr_list_file", os.path.join(d, "tr.list"), "--cv_list_file", os.path.join(d, "cv.list"),
                          "--min_epoches", "1", "--max_epoches", "1", "--d_learning_rate", "0.0003"])
    assert "[*] Load SUCCESS" in out2
    # decode: G(x) * stddev_labels + mean_labels written as Kaldi ark, one matrix per test utterance
    _run(common + ["--decode", "--test_list_file", os.path.join(d, "test.list")])
    from rsrgan_b200.kaldi_io import ArkReader
    scp = [l.split() for l in open(os.path.join(save, "test", "feats.scp"))]
    assert [s[0] for s in scp] == ["utt08", "utt09"]
    for (utt, loc), T in zip(scp, lens[8:10]):
        path, off = loc.rsplit(":", 1)
        mat = ArkReader().read_ark(path, int(off))
        assert mat.shape == (T, 40) and np.isfinite(mat).all()
