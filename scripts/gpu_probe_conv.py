"""Probe: 1-D SAME convolution as rsr_gemm over an overlapped strided view (no im2col) vs torch conv1d.
    python scripts/gpu_probe_conv.py"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops  # noqa: E402

h = ops.Handle(0, "f16")
dev = h.device
torch.manual_seed(0)
L, S, G = 257, 264, 8
for (N, w, cin, cout) in [(5, 13, 1, 12), (64, 11, 12, 16), (256, 7, 24, 32), (256, 9, 24, 20)]:
    cip, cop = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    x = torch.randn(N, L, cin, device=dev)
    W = torch.randn(w, cin, cout, device=dev) / (w * cin) ** 0.5
    b = torch.randn(cout, device=dev)
    xb = torch.zeros(N * S + 2 * G, cip, dtype=torch.float16, device=dev)
    xb[G:G + N * S].view(N, S, cip)[:, :L, :cin] = x.half()
    Wp = torch.zeros(w, cip, cop, dtype=torch.float16, device=dev)
    Wp[:, :cin, :cout] = W.half()
    bp = torch.zeros(cop, device=dev)
    bp[:cout] = b
    yb = torch.zeros(N * S + 2 * G, cop, dtype=torch.float16, device=dev)
    A = xb.as_strided((N * S, w * cip), (cip, 1), (G - w // 2) * cip)
    h.gemm(A, Wp.view(w * cip, cop), N * S, cop, w * cip, b_mn=True, bias=bp, act=ops.ACT_RELU, out16=yb[G:])
    h.conv_mask_rows(yb[G:], N, S, L, cop)
    torch.cuda.synchronize()
    y = yb[G:G + N * S].view(N, S, cop)
    ref = F.relu(F.conv1d(x.half().float().permute(0, 2, 1), W.half().float().permute(2, 1, 0), b, padding=w // 2)).permute(0, 2, 1)
    err = (y[:, :L, :cout].float() - ref).abs().max().item()
    pad = y[:, L:].abs().sum().item() + y[:, :, cout:].abs().sum().item()
    # weight gradient: dW = A^T dY
    dY = torch.zeros_like(yb)
    dY[G:G + N * S].view(N, S, cop)[:, :L, :cout] = torch.randn(N, L, cout, device=dev).half()
    dW = torch.zeros(w * cip, cop, device=dev)
    h.gemm(A, dY[G:], w * cip, cop, N * S, a_mn=True, b_mn=True, beta=1.0, out32=dW)
    torch.cuda.synchronize()
    dWref = A.float().t() @ dY[G:G + N * S].float()
    errw = ((dW - dWref).abs().max() / dWref.abs().max()).item()
    # data gradient with flipped taps
    Wf = torch.zeros(w, cop, cip, dtype=torch.float16, device=dev)
    h.conv_w_flip(Wp, w, cip, cop, Wf)
    dX = torch.zeros_like(xb)
    Ad = dY.as_strided((N * S, w * cop), (cop, 1), (G - w // 2) * cop)
    h.gemm(Ad, Wf.view(w * cop, cip), N * S, cip, w * cop, b_mn=True, out16=dX[G:])
    torch.cuda.synchronize()
    xr = x.half().float().permute(0, 2, 1).requires_grad_(True)
    yr = F.conv1d(xr, W.half().float().permute(2, 1, 0), None, padding=w // 2)
    yr.backward(dY[G:G + N * S].view(N, S, cop)[:, :L, :cout].float().permute(0, 2, 1))
    errx = (dX[G:G + N * S].view(N, S, cip)[:, :L, :cin].float() - xr.grad.permute(0, 2, 1)).abs().max().item()
    print("N=%d w=%d %d->%d  fwd max err %.3e  pad residue %.1e  dW rel err %.3e  dX max err %.3e" %
          (N, w, cin, cout, err, pad, errw, errx))
