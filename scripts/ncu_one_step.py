"""Runs a few batch schedules of a bench config eagerly -- the target of `ncu -k regex:<kernel>` captures.
    python scripts/ncu_one_step.py [cfg2|cfgP|cfgR] [steps]"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from rsrgan_b200.gan_rnn import GAN_RNN  # noqa: E402

cfg = bench.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
B, T = cfg["B"], cfg["T"]
args = Namespace(g_type=cfg["g_type"], d_type=cfg["d_type"], batch_size=B, g_cell=cfg["g_cell"], g_proj=cfg["g_proj"],
                 g_layers=cfg["g_layers"], init_mse_weight=10.0, init_disc_noise_std=0.05, l2_scale=0.0,
                 dtype=os.environ.get("RSR_DTYPE", "f16"), seed=1234, g_learning_rate=8e-5, d_learning_rate=1e-3)
m = GAN_RNN(None, args, ["/gpu:0"])
rng = np.random.default_rng(0)
x = torch.tensor(rng.standard_normal((B, T, 257), dtype=np.float32)).cuda()
y = torch.tensor(rng.standard_normal((B, T, 40), dtype=np.float32)).cuda()
ln = torch.full((B,), T, dtype=torch.int32).cuda()
for _ in range(steps):
    m.train_batch(x, y, ln, sync=False)
torch.cuda.synchronize()
print("done", m.h.launches, "kernels")
