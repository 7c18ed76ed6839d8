"""Probe: rsr_affine_act_drop mask vs the oracle across ticks and sizes (debug aid for tests/test_batchnorm_gpu.py)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import rsr_oracle as O          # noqa: E402  (checker only)
from rsrgan_b200 import ops                 # noqa: E402

h = ops.Handle(0, "f16")
dev = h.device
for rows, N in ((300, 280), (4096, 1024), (12800, 1024)):
    rng = np.random.default_rng(0)
    z = (rng.standard_normal((rows, N)) + 3.0).astype(np.float32)       # all positive: kept <=> nonzero
    zt = torch.tensor(z, device=dev)
    one, zero = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    buf = torch.tensor([99, 5], dtype=torch.int64, device=dev)
    for step in range(3):
        out = torch.zeros(rows, N + 8, dtype=torch.float16, device=dev)
        h.affine_act_drop(zt, rows, N, one, zero, 0, 0.8, buf, 513, out)
        torch.cuda.synchronize()
        got = out[:, :N].float().cpu().numpy() != 0
        st = buf.tolist()
        res = []
        for t in range(4, 9):
            m = O.dropout_mask(99, t, 513, rows, N, 0.8)
            res.append((t, int((m != got).sum())))
        print(rows, N, "buf", st, "kept frac %.4f" % got.mean(), "mismatches by tick", res, flush=True)
        h.rng_tick(buf)
