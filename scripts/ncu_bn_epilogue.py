"""fully_connected + batch_norm statistics at the cfg-2 discriminator size, two-pass (rsr_gemm, rsr_bn_train_stats) and
with the statistics in the GEMM epilogue (rsr_gemm(stats), rsr_bn_train_finish), K = 1024 and K = 40 -- the target of the
ncu duration / DRAM table profiles/r2_bn_epilogue_ncu.csv (launch order: two-pass then epilogue form, per K)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops                 # noqa: E402

rows, N = 12800, 1024
h = ops.Handle(0, "f16")
dev = h.device
g = torch.Generator(device=dev).manual_seed(0)
gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
state = torch.zeros(6, N, device=dev)
state[1].fill_(1.0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for K in (1024, 40):
    x16 = (0.5 * torch.randn(rows, K, device=dev, generator=g)).half()
    w16 = (0.05 * torch.randn(K, N, device=dev, generator=g)).half()
    z, coef, scratch = torch.zeros(rows, N, device=dev), torch.zeros(8, N, device=dev), torch.zeros(768, N, device=dev)
    for rep in range(2):
        flush.zero_()
        h.gemm(x16, w16, rows, N, K, b_mn=True, out32=z)
        h.bn_train_stats(z, rows, N, gamma, beta, state, coef, scratch, update_state=True)
        flush.zero_()
        h.gemm(x16, w16, rows, N, K, b_mn=True, out32=z, stats=scratch)
        h.bn_train_finish((rows + 127) // 128, rows, N, gamma, beta, state, coef, scratch, update_state=True)
torch.cuda.synchronize()
print("done")
