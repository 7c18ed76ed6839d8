"""One DNNTrainer step of the RCED generator with batch_norm on the convolutions (cfg-4 frame count) -- the target of
the ncu DRAM table of the lines kernels (scripts/r2_prof2.sh)."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200.dnn_trainer import DNNTrainer  # noqa: E402

N = 256
args = Namespace(g_type="rced", batch_size=N, input_dim=257, output_dim=40, batch_norm=True, g_learning_rate=1e-4, seed=1,
                 dtype="f16")
m = DNNTrainer(None, args, ["/gpu:0"])
rng = np.random.default_rng(0)
x = torch.tensor(rng.standard_normal((N, 257), dtype=np.float32)).cuda()
y = torch.tensor(rng.standard_normal((N, 40), dtype=np.float32)).cuda()
for _ in range(2):
    out = m.train_step(x, y)
torch.cuda.synchronize()
print("done", out)
