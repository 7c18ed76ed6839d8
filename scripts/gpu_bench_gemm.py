"""Times rsr_gemm on the GEMM shapes of a bench config (CUDA events, L2 flushed between calls) and prints
TFLOP/s and the HBM bytes each call must move, next to torch.matmul (cuBLAS) on the same operands as a
calibration line.  Diagnostic tool for `gpurun`.

    python scripts/gpu_bench_gemm.py [f16|bf16] [substring of the case label]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f16"
only = sys.argv[2] if len(sys.argv) > 2 else ""
h = ops.Handle(0, dt)
dev = h.device
R = 12800
# (label, M, N, K, a_mn, b_mn, out32, out16, beta)
CASES = [
    ("G fc in   y=xW      ", R, 256, 264, 0, 1, 0, 1, 0.0),
    ("G zx      y=xKx     ", R, 2048, 256, 0, 1, 1, 0, 0.0),
    ("G zx16    y=xKx     ", R, 2048, 256, 0, 1, 0, 1, 0.0),
    ("G proj    y=mtWp    ", R, 256, 512, 0, 1, 0, 1, 0.0),
    ("G fc out  y=xW      ", R, 40, 256, 0, 1, 1, 0, 0.0),
    ("D fc0     y=xW      ", R, 1024, 40, 0, 1, 0, 1, 0.0),
    ("D fc      y=xW      ", R, 1024, 1024, 0, 1, 0, 1, 0.0),
    ("D fc out  y=xW      ", R, 8, 1024, 0, 1, 1, 0, 0.0),
    ("D dx      dx=dyW^T  ", R, 1024, 1024, 0, 0, 0, 1, 0.0),
    ("D dW      dW=x^Tdy  ", 1024, 1024, R, 1, 1, 1, 0, 1.0),
    ("D dW0     dW=x^Tdy  ", 40, 1024, R, 1, 1, 1, 0, 1.0),
    ("G dmt     =doutWp^T ", R, 512, 256, 0, 0, 1, 0, 0.0),
    ("G dKx     dW=x^Tdz  ", 256, 2048, R, 1, 1, 1, 0, 1.0),
    ("G dx      dx=dzKx^T ", R, 256, 2048, 0, 0, 0, 1, 0.0),
    ("G dmtot   =dzKh^T   ", R, 256, 2048, 0, 0, 0, 1, 0.0),
    ("G dWp     =mt^Tdmtot", 512, 256, R, 1, 1, 1, 0, 1.0),
    ("G Wc      =WpKh     ", 512, 2048, 256, 0, 1, 0, 1, 0.0),
]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("%-22s %6s %5s %6s  %8s %8s %7s | cublas us" % ("case", "M", "N", "K", "us", "TFLOP/s", "GB/s"))
for (lab, M, N, K, a_mn, b_mn, o32, o16, beta) in CASES:
    if only and only not in lab:
        continue
    A = (torch.randn((K, M) if a_mn else (M, K), device=dev) * 0.1).to(h.h16)
    B = (torch.randn((K, N) if b_mn else (N, K), device=dev) * 0.1).to(h.h16)
    out32 = torch.zeros(M, N, device=dev) if o32 else None
    out16 = torch.zeros(M, N, dtype=h.h16, device=dev) if o16 else None
    ts = []
    for it in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.gemm(A, B, M, N, K, a_mn=bool(a_mn), b_mn=bool(b_mn), beta=beta, out32=out32, out16=out16)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    us = sorted(ts[2:])[len(ts[2:]) // 2]
    Am = A.t() if a_mn else A
    Bm = B if b_mn else B.t()
    tc = []
    for it in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(Am, Bm)
        e1.record()
        torch.cuda.synchronize()
        tc.append(e0.elapsed_time(e1) * 1e3)
    byts = 2 * (M * K + K * N) + M * N * (4 * o32 * (2 if beta else 1) + 2 * o16)
    print("%-22s %6d %5d %6d  %8.1f %8.1f %7.0f | %8.1f" % (lab, M, N, K, us, 2.0 * M * N * K / us / 1e6, byts / us / 1e3,
                                                          sorted(tc[1:])[len(tc[1:]) // 2]), flush=True)
