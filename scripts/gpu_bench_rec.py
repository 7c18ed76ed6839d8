"""Times the persistent LSTMP recurrence kernels (rsr_lstmp_rec_fwd / _bwd) alone: microseconds per time
step for several (B, Cp).  Diagnostic tool for `gpurun`.

    python scripts/gpu_bench_rec.py [f16|bf16]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops  # noqa: E402

dt = sys.argv[1] if len(sys.argv) > 1 else "f16"
h = ops.Handle(0, dt)
dev = h.device


def run(B, T, Cp):
    rows = T * B
    zx = torch.randn(rows, 4 * Cp, device=dev) * 0.5
    wcT = (torch.randn(4 * Cp, Cp, device=dev) * 0.03).to(h.h16)
    wc = wcT.t().contiguous()
    w = [torch.randn(Cp, device=dev) * 0.1 for _ in range(3)]
    ln = torch.full((B,), T, dtype=torch.int32, device=dev)
    mt = torch.zeros(rows + B, Cp, dtype=h.h16, device=dev)
    save = torch.zeros(rows, 5 * Cp, device=dev)
    dmt = torch.randn(rows, Cp, device=dev) * 0.01
    dz = torch.zeros(rows + B, 4 * Cp, dtype=h.h16, device=dev)
    db = torch.zeros(4 * Cp, device=dev)
    dw = [torch.zeros(Cp, device=dev) for _ in range(3)]
    I = 256
    x16 = (torch.randn(rows, I, device=dev) * 0.5).to(h.h16)
    kxT = (torch.randn(4 * Cp, I, device=dev) * 0.03).to(h.h16)
    bias = torch.randn(4 * Cp, device=dev) * 0.1
    res = []
    for which in ("fwd", "fused", "bwd"):
        ts = []
        for it in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if which == "fwd":
                h.lstmp_rec_fwd(B, T, Cp, zx, wcT, w[0], w[1], w[2], ln, mt, save)
            elif which == "fused":
                if not h.lstmp_fused_fwd(B, T, I, Cp, x16, kxT, bias, wcT, w[0], w[1], w[2], ln, mt, save):
                    break
            else:
                h.lstmp_rec_bwd(B, T, Cp, dmt, wc, w[0], w[1], w[2], ln, save, dz, db, dw[0], dw[1], dw[2])
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(sorted(ts[1:])[len(ts[1:]) // 2] if len(ts) > 1 else float("nan"))
    return res


print("%5s %5s %5s | %10s %10s | %10s %10s | %10s %10s" % ("B", "T", "Cp", "fwd us", "us/step", "fused us", "us/step", "bwd us", "us/step"))
shapes = os.environ.get("RSR_REC_SHAPES")          # "B,T,Cp;B,T,Cp;..." overrides the default sweep
for (B, T, Cp) in [tuple(int(v) for v in s.split(",")) for s in shapes.split(";")] if shapes else [(16, 100, 512), (32, 100, 512), (48, 100, 512), (64, 100, 512), (96, 100, 512), (112, 100, 512),
                   (128, 100, 512), (128, 200, 512), (8, 100, 768), (64, 100, 1024), (8, 100, 256), (32, 100, 256),
                   (128, 100, 256)]:
    try:
        f, fu, b = run(B, T, Cp)
        print("%5d %5d %5d | %10.1f %10.2f | %10.1f %10.2f | %10.1f %10.2f" % (B, T, Cp, f, f / T, fu, fu / T, b, b / T), flush=True)
    except Exception as e:  # noqa: BLE001
        print(B, T, Cp, "failed:", e, flush=True)
