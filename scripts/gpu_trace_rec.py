"""Phase timing of the cluster LSTMP forward kernel (library built with RSR_EXTRA_NVCC_FLAGS=-DRSR_TRACE).
    python scripts/gpu_trace_rec.py [B] [Cp]"""
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
Cp = int(sys.argv[2]) if len(sys.argv) > 2 else 512
which = sys.argv[3] if len(sys.argv) > 3 else "fwd"
T = 40
h = ops.Handle(0, "f16")
dev = h.device
rows = T * B
zx = torch.randn(rows, 4 * Cp, device=dev) * 0.5
wcT = (torch.randn(4 * Cp, Cp, device=dev) * 0.03).to(h.h16)
w = [torch.randn(Cp, device=dev) * 0.1 for _ in range(3)]
ln = torch.full((B,), T, dtype=torch.int32, device=dev)
mt = torch.zeros(rows + B, Cp, dtype=h.h16, device=dev)
save = torch.zeros(rows, 5 * Cp, device=dev)
h.lstmp_rec_fwd(B, T, Cp, zx, wcT, w[0], w[1], w[2], ln, mt, save)
if which == "pfwd":      # fused forward, CTA-pair kernel (lstmp_pair_sm100.cu)
    I = 256
    x16 = (torch.randn(rows, I, device=dev) * 0.5).to(h.h16)
    kxT = (torch.randn(4 * Cp, I, device=dev) * 0.03).to(h.h16)
    bias = torch.randn(4 * Cp, device=dev) * 0.1
    for _ in range(3):
        assert h.lstmp_fused_fwd(B, T, I, Cp, x16, kxT, bias, wcT, w[0], w[1], w[2], ln, mt, save)
elif which == "wave":    # two stacked layers as one wavefront launch; the trace is layer 1's first cluster
    I = 256
    x16 = (torch.randn(rows, I, device=dev) * 0.5).to(h.h16)
    mk = lambda: ((torch.randn(4 * Cp, I, device=dev) * 0.03).to(h.h16), torch.randn(4 * Cp, device=dev) * 0.1, wcT, w[0], w[1], w[2])
    wpT = (torch.randn(I, Cp, device=dev) * 0.03).to(h.h16)
    mt2, out1, save2 = torch.zeros_like(mt), torch.zeros(rows + B, I, dtype=h.h16, device=dev), torch.zeros_like(save)
    l1, l2 = mk(), mk()
    for _ in range(3):
        assert h.lstmp_wave_fwd(B, T, Cp, I, I, ln, x16, l1, mt, save, wpT, out1, l2, mt2, save2)
elif which == "wavebwd":  # backward wavefront; the trace is layer 2's first cluster
    I = 256
    wc = wcT.t().contiguous()
    fT = (torch.randn(Cp, 4 * Cp, device=dev) * 0.03).to(h.h16)
    dmt = torch.randn(rows, Cp, device=dev) * 0.01
    dz1, dz2 = (torch.zeros(rows + B, 4 * Cp, dtype=h.h16, device=dev) for _ in range(2))
    part = torch.zeros(T * (B + 48), Cp, device=dev)
    g = lambda: (torch.zeros(4 * Cp, device=dev), torch.zeros(Cp, device=dev), torch.zeros(Cp, device=dev), torch.zeros(Cp, device=dev))
    for _ in range(3):
        assert h.lstmp_wave_bwd(B, T, Cp, ln, dmt, (wc, w[0], w[1], w[2]), save, dz2, g(), fT, part, (wc, w[0], w[1], w[2]), save, dz1, g())
elif which == "bwd":
    wc = wcT.t().contiguous()
    dmt = torch.randn(rows, Cp, device=dev) * 0.01
    dz = torch.zeros(rows + B, 4 * Cp, dtype=h.h16, device=dev)
    db = torch.zeros(4 * Cp, device=dev)
    dw = [torch.zeros(Cp, device=dev) for _ in range(3)]
    for _ in range(3):
        h.lstmp_rec_bwd(B, T, Cp, dmt, wc, w[0], w[1], w[2], ln, save, dz, db, dw[0], dw[1], dw[2])
else:
    for _ in range(3):
        h.lstmp_rec_fwd(B, T, Cp, zx, wcT, w[0], w[1], w[2], ln, mt, save)
torch.cuda.synchronize()
buf = (C.c_ulonglong * (64 * 8))()
fn = h.lib.rsr_debug_trace_pair if (which.startswith("p") or which.startswith("wave") or (which == "bwd" and B > 16 and not os.environ.get("RSR_NO_PAIR"))) else h.lib.rsr_debug_trace
fn.argtypes = [C.c_void_p, C.c_int]
rc = fn(buf, 64 * 8)
tr = np.array(buf[:], dtype=np.int64).reshape(64, 8)[:T]
names = (["top", "full-wait done", "mma issued", "mma done", "xchg+sync", "gate math", "st.async sends", "global stores"] if which in ("fwd", "pfwd", "wave") else
         ["top", "partials landed", "summed", "gate math", "dz stores", "bar+mma issued", "mma done", "ld+send"])
print(which, "B %d Cp %d rc %d; cycles between trace points (median over steps 5..%d)" % (B, Cp, rc, T - 2))
nt = len(names)
d = np.diff(tr[5:T - 1, :nt], axis=1)
for i in range(nt - 1):
    print("  %-16s -> %-16s %7.0f" % (names[i], names[i + 1], np.median(d[:, i])))
print("  step period %7.0f cycles" % np.median(np.diff(tr[5:T - 1, 0])))
