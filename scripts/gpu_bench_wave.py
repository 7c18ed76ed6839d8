"""Times rsr_lstmp_wave_fwd (two stacked LSTMP layers as one wavefront launch) against the same two layers run one
after the other (fused forward, projection GEMM, fused forward).  usage: gpu_bench_wave.py [f16|bf16]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops, packing  # noqa: E402

h = ops.Handle(0, sys.argv[1] if len(sys.argv) > 1 else "f16")
dev = h.device


def layer(rng, I, Cp, P):
    Ik, Pp = packing.round_up(I, 16), packing.round_up(P, 8)
    r = lambda *s: torch.tensor(rng.standard_normal(s).astype(np.float32) * 0.05, device=dev)
    return (r(4 * Cp, Ik).to(h.h16), r(4 * Cp), r(4 * Cp, Cp).to(h.h16), r(Cp), r(Cp), r(Cp)), r(Pp, Cp).to(h.h16)


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


print("%5s %5s %5s %5s | %10s %8s | %10s %8s | %s" % ("B", "T", "Cp", "nbp", "serial us", "us/step", "wave us", "us/step", "ok"))
for (B, T, I, Cp, P, nbp) in [(128, 100, 256, 512, 256, 48), (96, 100, 256, 512, 256, 32), (96, 100, 256, 512, 256, 48),
                              (64, 100, 256, 512, 256, 32), (128, 200, 256, 512, 256, 48), (64, 100, 40, 256, 40, 0)]:
    if nbp:
        os.environ["RSR_WAVE_NBP"] = str(nbp)
    else:
        os.environ.pop("RSR_WAVE_NBP", None)
    rng = np.random.default_rng(1)
    Ip, Pp = packing.round_up(I, 8), packing.round_up(P, 8)
    (d1, wpT1), (d2, _) = layer(rng, I, Cp, P), layer(rng, P, Cp, P)
    x16 = torch.tensor(rng.standard_normal((T * B, Ip)).astype(np.float32), device=dev).to(h.h16)
    d_len = torch.full((B,), T, dtype=torch.int32, device=dev)
    mk = lambda cols: torch.zeros((T + 1) * B, cols, dtype=h.h16, device=dev)
    mt1, mt2, out1 = mk(Cp), mk(Cp), mk(Pp)
    sv1, sv2 = (torch.zeros(T * B, 5 * Cp, dtype=torch.float32, device=dev) for _ in range(2))

    def serial():
        h.lstmp_fused_fwd(B, T, I, Cp, x16, *d1, d_len, mt1, sv1)
        h.gemm(mt1[B:], wpT1, T * B, Pp, Cp, out16=out1[B:])
        h.lstmp_fused_fwd(B, T, P, Cp, out1[B:], *d2, d_len, mt2, sv2)

    ok = [True]

    def wave():
        ok[0] = h.lstmp_wave_fwd(B, T, Cp, I, P, d_len, x16, d1, mt1, sv1, wpT1, out1, d2, mt2, sv2) and ok[0]

    ts = timeit(serial)
    tw = timeit(wave)
    print("%5d %5d %5d %5d | %10.1f %8.2f | %10.1f %8.2f | %s" % (B, T, Cp, nbp, ts, ts / (2 * T), tw, tw / T, ok[0]))
    # backward
    wc1, wc2 = d1[2].t().contiguous(), d2[2].t().contiguous()
    kx2 = d2[0].t().contiguous()[:Pp]
    wp1 = wpT1.t().contiguous()
    fT = (wp1.float() @ kx2.float()).to(h.h16).contiguous()
    dmt2 = torch.tensor(rng.standard_normal((T * B, Cp)).astype(np.float32) * 0.01, device=dev)
    dmt1 = torch.zeros(T * B, Cp, dtype=torch.float32, device=dev)
    dz1, dz2 = (torch.zeros((T + 1) * B, 4 * Cp, dtype=h.h16, device=dev) for _ in range(2))
    dx2 = torch.zeros(T * B, Pp, dtype=h.h16, device=dev)
    part = torch.zeros(T * (B + 48), Cp, dtype=torch.float32, device=dev)
    z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
    g1, g2 = (z(4 * Cp), z(Cp), z(Cp), z(Cp)), (z(4 * Cp), z(Cp), z(Cp), z(Cp))

    def serial_b():
        h.lstmp_rec_bwd(B, T, Cp, dmt2, wc2, *d2[3:6], d_len, sv2, dz2, *g2)
        h.gemm(dz2, kx2, T * B, Pp, 4 * Cp, out16=dx2)
        h.gemm(dx2, wp1, T * B, Cp, Pp, out32=dmt1)
        h.lstmp_rec_bwd(B, T, Cp, dmt1, wc1, *d1[3:6], d_len, sv1, dz1, *g1)

    def wave_b():
        ok[0] = h.lstmp_wave_bwd(B, T, Cp, d_len, dmt2, (wc2,) + tuple(d2[3:6]), sv2, dz2, g2, fT, part,
                                 (wc1,) + tuple(d1[3:6]), sv1, dz1, g1) and ok[0]

    ts = timeit(serial_b)
    tw = timeit(wave_b)
    print("%5s %5s %5s %5s | %10.1f %8.2f | %10.1f %8.2f | %s  (backward)" % ("", "", "", "", ts, ts / (2 * T), tw, tw / T, ok[0]))
