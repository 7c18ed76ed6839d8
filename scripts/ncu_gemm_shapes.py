"""A few launches of selected cfg-2 GEMM shapes -- the target of `ncu --set full -k regex:gemm` captures.
    python scripts/ncu_gemm_shapes.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops  # noqa: E402

h = ops.Handle(0, "f16")
dev = h.device
M = 12800
x1024 = (torch.randn(M, 1024, device=dev) * 0.5).to(h.h16)
x40 = (torch.randn(M, 40, device=dev) * 0.5).to(h.h16)
W = (torch.randn(1024, 1024, device=dev) * 0.03).to(h.h16)
W0 = (torch.randn(40, 1024, device=dev) * 0.1).to(h.h16)
b = torch.randn(1024, device=dev) * 0.1
y = torch.zeros(M, 1024, dtype=h.h16, device=dev)
dx = torch.zeros(M, 1024, dtype=h.h16, device=dev)
dW = torch.zeros(1024, 1024, device=dev)
for _ in range(3):
    h.gemm(x1024, W, M, 1024, 1024, b_mn=True, bias=b, act=ops.ACT_RELU, out16=y)                  # D fc   (fwd)
    h.gemm(x1024, W, M, 1024, 1024, dact_src=y, dact=ops.ACT_RELU, out16=dx)                       # D dx   (bwd data)
    h.gemm(x1024, y, 1024, 1024, M, a_mn=True, b_mn=True, beta=1.0, out32=dW)                      # D dW
    h.gemm(x40, W0, M, 1024, 40, b_mn=True, bias=b, act=ops.ACT_RELU, out16=y)                     # D fc0  (K = 40)
torch.cuda.synchronize()
print("done")
