"""HBM roofline of the batch_norm / dropout stream kernels (csrc/batchnorm.cu) at the cfg-2 discriminator size
(12800 frames x 1024 units): CUDA-event time per C-ABI call over rotating buffer sets larger than L2, algorithmic
bytes per call from DESIGN.md section 4, against the measured copy bandwidth in MEASURED_PEAKS.json.
Prints one JSON line per entry point."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops                 # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
peak = 6551.0
try:
    mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    for k in ("hbm_gbs", "hbm_gbps"):
        if k in mp:
            peak = float(mp[k])
except Exception:
    pass

rows, N, SETS, ITERS = 12800, 1024, 6, int(os.environ.get("RSR_BN_ITERS", "30"))
h = ops.Handle(0, "f16")
dev = h.device
g = torch.Generator(device=dev).manual_seed(0)
sets = []
for _ in range(SETS):
    sets.append(dict(z=torch.randn(rows, N, device=dev, generator=g),
                     da=(0.1 * torch.randn(rows, N, device=dev, generator=g)).half(),
                     out=torch.zeros(rows, N, dtype=torch.float16, device=dev),
                     dz=torch.zeros(rows, N, dtype=torch.float16, device=dev),
                     coef=torch.zeros(8, N, device=dev), scratch=torch.zeros(768, N, device=dev),
                     state=torch.zeros(6, N, device=dev)))
gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
dgam, dbet = torch.zeros(N, device=dev), torch.zeros(N, device=dev)
rng = torch.tensor([1, 0], dtype=torch.int64, device=dev)
for s in sets:
    s["state"][1].fill_(1.0)


def stats(s):
    h.bn_train_stats(s["z"], rows, N, gamma, beta, s["state"], s["coef"], s["scratch"], update_state=True)


def norm(keep):
    return lambda s: h.affine_act_drop(s["z"], rows, N, s["coef"][0], s["coef"][1], 1, keep, rng, 3, s["out"])


def bwd(keep):
    return lambda s: h.bn_bwd(s["da"], s["z"], rows, N, 1, keep, rng, 3, True, s["coef"], None, dgam, dbet, s["dz"],
                              s["scratch"])


E = rows * N
cases = [("rsr_bn_train_stats", stats, 4 * E), ("rsr_affine_act_drop", norm(1.0), 6 * E),
         ("rsr_affine_act_drop(keep=0.8)", norm(0.8), 6 * E), ("rsr_bn_bwd", bwd(1.0), 14 * E),
         ("rsr_bn_bwd(keep=0.8)", bwd(0.8), 14 * E)]
for s in sets:
    stats(s)
for name, fn, nbytes in cases:
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(ITERS):
        fn(sets[i % SETS])
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / ITERS
    gbs = nbytes / ms / 1e6
    print(json.dumps({"call": name, "rows": rows, "N": N, "us": round(ms * 1e3, 2), "algorithmic_bytes": nbytes,
                      "achieved_GBps": round(gbs, 1), "peak_GBps": peak, "frac": round(gbs / peak, 3),
                      "buffers": "%d rotating sets (%.0f MB) > 126 MB L2" % (SETS, SETS * E * 10 / 1e6)}), flush=True)

# batch_norm statistics in the GEMM epilogue (rsr_gemm_args.stats + rsr_bn_train_finish) against the two-pass form
# (GEMM writes the pre-activation, rsr_bn_train_stats reads it back): the cfg-2 discriminator's two hidden layers
for K in (1024, 40):
    x16 = [(0.5 * torch.randn(rows, K, device=dev, generator=g)).half() for _ in range(SETS)]
    w16 = (0.05 * torch.randn(K, N, device=dev, generator=g)).half()

    def two_pass(i):
        s = sets[i]
        h.gemm(x16[i], w16, rows, N, K, b_mn=True, out32=s["z"])
        h.bn_train_stats(s["z"], rows, N, gamma, beta, s["state"], s["coef"], s["scratch"], update_state=True)

    def fused(i):
        s = sets[i]
        h.gemm(x16[i], w16, rows, N, K, b_mn=True, out32=s["z"], stats=s["scratch"])
        h.bn_train_finish((rows + 127) // 128, rows, N, gamma, beta, s["state"], s["coef"], s["scratch"], update_state=True)

    def gemm_only(i):
        h.gemm(x16[i], w16, rows, N, K, b_mn=True, out32=sets[i]["z"])

    for name, fn in (("gemm", gemm_only), ("gemm + rsr_bn_train_stats", two_pass),
                     ("gemm(stats) + rsr_bn_train_finish", fused)):
        for i in range(SETS):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(ITERS):
            fn(i % SETS)
        e1.record()
        torch.cuda.synchronize()
        print(json.dumps({"call": name, "rows": rows, "N": N, "K": K, "us": round(e0.elapsed_time(e1) / ITERS * 1e3, 2)}),
              flush=True)
