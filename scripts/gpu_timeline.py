"""Concurrent timeline of ONE eager batch schedule (both streams live): start / end of every C-ABI call relative to
the start of the schedule, from CUDA events.  Shows what is on the critical path and what hides under what.
    python scripts/gpu_timeline.py [cfg2] > gpurun_out/timeline.txt"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["RSR_NO_GRAPH"] = "1"
import bench  # noqa: E402
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B, T = cfg["B"], cfg["T"]
args = Namespace(g_type=cfg["g_type"], d_type=cfg["d_type"], batch_size=B, g_cell=cfg["g_cell"], g_proj=cfg["g_proj"],
                 g_layers=cfg["g_layers"], init_mse_weight=10.0, init_disc_noise_std=0.05, l2_scale=0.0,
                 dtype=os.environ.get("RSR_DTYPE", "f16"), seed=1234, g_learning_rate=8e-5, d_learning_rate=1e-3)
m = GAN_RNN(None, args, ["/gpu:0"])
rng = np.random.default_rng(0)
x = torch.tensor(rng.standard_normal((B, T, 257), dtype=np.float32)).cuda()
y = torch.tensor(rng.standard_normal((B, T, 40), dtype=np.float32)).cuda()
ln = torch.full((B,), T, dtype=torch.int32).cuda()
for _ in range(3):
    m.train_batch(x, y, ln, sync=False)
torch.cuda.synchronize()
m.h.timeline = []
base = torch.cuda.Event(enable_timing=True)
base.record()
m.train_batch(x, y, ln, sync=False)
end = torch.cuda.Event(enable_timing=True)
end.record()
torch.cuda.synchronize()
tl = m.h.timeline
m.h.timeline = None
print("schedule %.1f us, %d calls" % (base.elapsed_time(end) * 1e3, len(tl)))
print("%9s %9s %8s  %-5s %s" % ("start", "end", "dur", "strm", "call"))
prev_end = {"main": 0.0, "side": 0.0}
for name, strm, e0, e1, work in tl:
    s, e = base.elapsed_time(e0) * 1e3, base.elapsed_time(e1) * 1e3
    gap = s - prev_end[strm]
    prev_end[strm] = e
    print("%9.1f %9.1f %8.1f  %-5s %-24s gap %7.1f %s" % (s, e, e - s, strm, name, gap,
                                                         "%.0f TF" % (work / (e - s) / 1e6) if work else ""))
