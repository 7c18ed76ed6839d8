"""Kernel timeline of the CUDA-GRAPH replay of one batch schedule (CUPTI via torch.profiler): start, duration and
stream of every kernel, relative to the first kernel of the schedule.
    python scripts/gpu_timeline_graph.py [cfg2] > gpurun_out/timeline_graph.txt"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
for _ in range(6):
    m.train_batch(x, y, ln, sync=False)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(2):
        m.train_batch(x, y, ln, sync=False)
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
if not ev:
    print("no CUDA events captured")
    sys.exit(0)
# keep the second replay
n = len(ev) // 2
ev = ev[n:]
t0 = ev[0].time_range.start
print("%d kernels, span %.1f us" % (len(ev), ev[-1].time_range.end - t0))
streams = {}
for e in ev:
    sid = getattr(e, "stream", None)
    if sid is None:
        sid = getattr(e, "device_resource_id", 0)
    tag = streams.setdefault(sid, "s%d" % len(streams))
    print("%9.1f %9.1f %8.1f  %-3s %s" % (e.time_range.start - t0, e.time_range.end - t0,
                                         e.time_range.end - e.time_range.start, tag, e.name[:60]))
