"""Probe: one big GEMM (12800 x 1024 x 1024, 16-bit out) with forced tile widths, repeated back to back so the
launch/cold-start cost is amortised: separates the steady-state rate from the per-launch overhead."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops
h = ops.Handle(0, "f16")
dev = h.device
M, N, K = 12800, 1024, 1024
A = (torch.randn(M, K, device=dev) * 0.1).to(h.h16)
B = (torch.randn(K, N, device=dev) * 0.1).to(h.h16)
out = torch.zeros(M, N, dtype=h.h16, device=dev)
for tn in (128, 256, 0):      # 0 = automatic choice (two-CTA kernel for this shape)
    for reps in (1, 8):
        ts = []
        for it in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                h.gemm(A, B, M, N, K, b_mn=True, out16=out, tile_n=tn)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / reps)
        us = sorted(ts)[2]
        print("tile_n %3d reps %d: %6.1f us/gemm  %6.1f TFLOP/s" % (tn, reps, us, 2.0 * M * N * K / us / 1e6), flush=True)
Am = A; Bm = B
for reps in (1, 8):
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            torch.matmul(Am, Bm)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / reps)
    print("cublas reps %d: %6.1f us/gemm" % (reps, sorted(ts)[2]))
