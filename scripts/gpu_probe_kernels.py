"""GPU bring-up probe: runs each kernel of librsrgan_sm100.so on small seeded inputs and prints the
error against a float64 CPU computation (numpy / the oracle).  Diagnostic tool for `gpurun`; the
asserting versions of these checks live in tests/test_kernels_gpu.py.

    python scripts/gpu_probe_kernels.py [gemm] [rec] [elt]
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rsrgan_b200 import ops, packing  # noqa: E402
from oracle import rsr_oracle as O  # noqa: E402  (checker only)


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max()), float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def probe_gemm(h):
    dev = h.device
    rng = np.random.default_rng(0)
    cases = [(128, 128, 64), (256, 128, 128), (200, 40, 257), (333, 280, 40), (128, 16, 64), (1000, 1, 1024),
             (512, 1024, 1024), (257, 3040, 560)]
    for (M, N, K) in cases:
        for a_mn in (0, 1):
            for b_mn in (0, 1):
                A = rng.standard_normal((M, K)).astype(np.float32)
                Bm = rng.standard_normal((K, N)).astype(np.float32)
                ldk, ldm, ldn = packing.round_up(K, 8), packing.round_up(M, 8), packing.round_up(N, 8)
                if a_mn:
                    At = torch.zeros(K, ldm, dtype=h.h16, device=dev); At[:, :M] = torch.tensor(A.T)
                else:
                    At = torch.zeros(M, ldk, dtype=h.h16, device=dev); At[:, :K] = torch.tensor(A)
                if b_mn:
                    Bt = torch.zeros(K, ldn, dtype=h.h16, device=dev); Bt[:, :N] = torch.tensor(Bm)
                else:
                    Bt = torch.zeros(N, ldk, dtype=h.h16, device=dev); Bt[:, :K] = torch.tensor(Bm.T)
                ld32 = packing.round_up(N, 4)
                out = torch.full((M, ld32), 7.0, dtype=torch.float32, device=dev)
                h.gemm(At, Bt, M, N, K, a_mn=a_mn, b_mn=b_mn, out32=out)
                torch.cuda.synchronize()
                A16 = (At[:, :M].T if a_mn else At[:, :K]).double().cpu().numpy()
                B16 = (Bt[:, :N] if b_mn else Bt[:, :K].T).double().cpu().numpy()
                ref = A16 @ B16
                got = out[:, :N].cpu().numpy()
                e = rel(got, ref)
                pad_ok = bool((out[:, N:] == 7.0).all().item()) if ld32 > N else True
                print("gemm M%d N%d K%d a_mn%d b_mn%d  maxabs %.3e relrms %.3e pad_untouched %s %s"
                      % (M, N, K, a_mn, b_mn, e[0], e[1], pad_ok, "OK" if e[1] < 1e-5 else "BAD"), flush=True)
    # epilogue: bias + resid + lrelu + dact + beta + out16
    M, N, K = 300, 280, 257
    A = rng.standard_normal((M, K)).astype(np.float32) * 0.3
    W = rng.standard_normal((K, N)).astype(np.float32) * 0.1
    bias = rng.standard_normal(N).astype(np.float32)
    resid = rng.standard_normal((M, N)).astype(np.float32)
    dsrc = rng.standard_normal((M, N)).astype(np.float32)
    old = rng.standard_normal((M, N)).astype(np.float32)
    At = torch.zeros(M, packing.round_up(K, 8), dtype=h.h16, device=dev); At[:, :K] = torch.tensor(A)
    Wt = torch.zeros(K, N, dtype=h.h16, device=dev); Wt[:] = torch.tensor(W)
    ds = torch.tensor(dsrc, device=dev).to(h.h16)
    o32 = torch.tensor(old, device=dev)
    o16 = torch.zeros(M, N, dtype=h.h16, device=dev)
    h.gemm(At, Wt, M, N, K, b_mn=True, alpha=0.5, beta=2.0, bias=torch.tensor(bias, device=dev),
           resid=torch.tensor(resid, device=dev), act=ops.ACT_LRELU, dact_src=ds, dact=ops.ACT_RELU,
           out32=o32, out16=o16)
    torch.cuda.synchronize()
    v = 0.5 * (At[:, :K].double().cpu().numpy() @ Wt.double().cpu().numpy()) + bias + resid
    v = np.maximum(v, 0.3 * v) * (ds.double().cpu().numpy() > 0)
    print("gemm epilogue fp32 (beta): maxabs %.3e relrms %.3e" % rel(o32.cpu().numpy(), v + 2.0 * old))
    print("gemm epilogue h16: maxabs %.3e relrms %.3e" % rel(o16.float().cpu().numpy(), v))


def probe_rec(h, B=8, T=12, I=40, C=256, P=40, ragged=True, seed=1):
    dev = h.device
    rng = np.random.default_rng(seed)
    Cp = packing.cell_pad(C)
    x = rng.standard_normal((B, T, I))
    K = O.xavier(rng, (I + P, 4 * C)) * 2.0
    b = rng.standard_normal(4 * C) * 0.1
    wi, wf, wo = (O.xavier(rng, (C,)) for _ in range(3))
    Wp = O.xavier(rng, (C, P)) * 2.0
    lengths = rng.integers(T // 2, T + 1, size=B) if ragged else np.full(B, T)
    out_ref, cache = O.lstmp_fwd(x, lengths, K, b, wi, wf, wo, Wp)
    steps = cache[-1]
    # device inputs: Zx packed [T*B, 4Cp] time-major
    zx = np.einsum("bti,ig->tbg", x, K[:I]) + b                       # (T,B,4C)
    zx_p = packing.pack_cols(zx.reshape(T * B, 4 * C), C).astype(np.float32)
    Wc = Wp @ K[I:]                                                    # (C,4C)
    Wc_p = packing.pad_first(packing.pack_cols(Wc, C), Cp)             # (Cp,4Cp)
    wc16 = torch.tensor(Wc_p, device=dev).to(h.h16).contiguous()
    wcT16 = wc16.t().contiguous()
    pk = lambda v: torch.tensor(packing.pad_last(v, Cp).astype(np.float32), device=dev)
    d_wi, d_wf, d_wo = pk(wi), pk(wf), pk(wo)
    d_len = torch.tensor(lengths.astype(np.int32), device=dev)
    mt_seq = torch.zeros((T + 1) * B, Cp, dtype=h.h16, device=dev)
    save = torch.zeros(T * B, 5, Cp, dtype=torch.float32, device=dev)
    d_zx = torch.tensor(zx_p, device=dev)
    t0 = time.time()
    h.lstmp_rec_fwd(B, T, Cp, d_zx, wcT16, d_wi, d_wf, d_wo, d_len, mt_seq, save)
    torch.cuda.synchronize()
    print("rec fwd B%d T%d C%d(Cp%d) P%d done in %.1f ms" % (B, T, C, Cp, P, 1e3 * (time.time() - t0)), flush=True)
    mt_ref = np.zeros((T, B, C))
    for t in range(T):
        act = steps[t][9]
        mt_ref[t] = np.where(act, steps[t][8], 0.0)
    got = mt_seq[B:].float().cpu().numpy().reshape(T, B, Cp)
    print("  mt vs oracle: maxabs %.3e relrms %.3e ; pad cells max %.3e" %
          (rel(got[:, :, :C], mt_ref) + (float(np.abs(got[:, :, C:]).max()) if Cp > C else 0.0,)))
    out_dev = got[:, :, :C] @ Wp
    print("  out (=mt Wp) vs oracle: maxabs %.3e relrms %.3e" % rel(out_dev.transpose(1, 0, 2), out_ref))
    # ---- backward
    dout = rng.standard_normal((B, T, P)) * 0.1
    dx_ref, g_ref = O.lstmp_bwd(dout, cache)
    dmt = np.einsum("btp,cp->tbc", dout, Wp).reshape(T * B, C)
    d_dmt = torch.tensor(packing.pad_last(dmt, Cp).astype(np.float32), device=dev)
    dz16 = torch.zeros(T * B, 4 * Cp, dtype=h.h16, device=dev)
    dbias = torch.zeros(4 * Cp, dtype=torch.float32, device=dev)
    dwi, dwf, dwo = (torch.zeros(Cp, dtype=torch.float32, device=dev) for _ in range(3))
    t0 = time.time()
    h.lstmp_rec_bwd(B, T, Cp, d_dmt, wc16, d_wi, d_wf, d_wo, d_len, save, dz16, dbias, dwi, dwf, dwo)
    torch.cuda.synchronize()
    print("rec bwd done in %.1f ms" % (1e3 * (time.time() - t0)), flush=True)
    db_got = packing.unpack_cols(dbias.cpu().numpy(), C)
    print("  dbias: maxabs %.3e relrms %.3e" % rel(db_got, g_ref["bias"]))
    print("  dw_i : maxabs %.3e relrms %.3e" % rel(dwi.cpu().numpy()[:C], g_ref["w_i_diag"]))
    print("  dw_f : maxabs %.3e relrms %.3e" % rel(dwf.cpu().numpy()[:C], g_ref["w_f_diag"]))
    print("  dw_o : maxabs %.3e relrms %.3e" % rel(dwo.cpu().numpy()[:C], g_ref["w_o_diag"]))
    dz = packing.unpack_cols(dz16.float().cpu().numpy(), C).reshape(T, B, 4 * C)
    dx_got = np.einsum("tbg,ig->bti", dz, K[:I])
    print("  dx (=dz Kx^T): maxabs %.3e relrms %.3e" % rel(dx_got, dx_ref))
    xin = np.concatenate([x.transpose(1, 0, 2), np.concatenate([np.zeros((1, B, P)), out_ref.transpose(1, 0, 2)[:-1]], 0)], 2)
    dK_got = np.einsum("tbi,tbg->ig", xin, dz)
    print("  dK (=xin^T dz): maxabs %.3e relrms %.3e" % rel(dK_got, g_ref["kernel"]))


def probe_elt(h):
    dev = h.device
    rng = np.random.default_rng(2)
    B, T, D = 5, 7, 40
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    mean = rng.standard_normal(D).astype(np.float32)
    std = (rng.random(D) + 0.5).astype(np.float32)
    noise = rng.standard_normal((B, D)).astype(np.float32)
    o16 = torch.zeros(T * B, 40, dtype=h.h16, device=dev)
    o32 = torch.zeros(T * B, 40, dtype=torch.float32, device=dev)
    tt = lambda a: torch.tensor(a, device=dev)
    h.stage_input(tt(x), B, T, D, out16=o16, out32=o32, mean=tt(mean), istd=tt(1.0 / std), noise=tt(noise))
    ref = ((x - mean) * (1.0 / std) + noise[:, None, :]).transpose(1, 0, 2).reshape(T * B, D)
    print("stage_input fp32: maxabs %.3e relrms %.3e" % rel(o32.cpu().numpy(), ref))
    back = torch.zeros(B, T, D, dtype=torch.float32, device=dev)
    h.unstage_output(o32, B, T, D, back)
    print("unstage: maxabs %.3e" % rel(back.cpu().numpy(), ref.reshape(T, B, D).transpose(1, 0, 2))[0])
    # losses
    n = B * T
    rl = rng.standard_normal(n).astype(np.float32); fk = rng.standard_normal(n).astype(np.float32)
    g = rng.standard_normal((n, D)).astype(np.float32); y = rng.standard_normal((n, D)).astype(np.float32)
    losses = torch.zeros(8, dtype=torch.float32, device=dev)
    rl4 = torch.zeros(n, 4, device=dev); rl4[:, 0] = tt(rl)
    fk4 = torch.zeros(n, 4, device=dev); fk4[:, 0] = tt(fk)
    g1 = torch.zeros(n, 8, dtype=h.h16, device=dev); g2 = torch.zeros_like(g1); g3 = torch.zeros_like(g1)
    dg = torch.zeros(n, D, device=dev)
    h.lsgan_mse_losses(losses, rl=rl4, fk=fk4, ld_logit=4, n_logit=n, clip=True, g=tt(g), y=tt(y), n_frames=n,
                       d_out=D, lam=10.0, gscale=64.0, d_rl_grad=g1, d_fk_grad=g2, g_adv_grad=g3, ld_grad=8, dg_mse=dg)
    rc, fc = np.clip(rl, -0.5, 1.5), np.clip(fk, -0.5, 1.5)
    L = O.lsgan_mse_losses(rc, fc, g.astype(np.float64), y.astype(np.float64))
    got = losses.cpu().numpy()
    print("losses got", got[:4], "ref", [L["d_rl_loss"], L["d_fk_loss"], L["g_adv_loss"], L["g_mse_loss"]])
    in_rl = ((rl >= -0.5) & (rl <= 1.5)); in_fk = ((fk >= -0.5) & (fk <= 1.5))
    print("d_rl_grad: %.3e" % rel(g1[:, 0].float().cpu().numpy(), 64 * 2 * (rc - 1) / n * in_rl)[1])
    print("g_adv_grad: %.3e" % rel(g3[:, 0].float().cpu().numpy(), 64 * 2 * (fc - 1) / n * in_fk)[1])
    print("dg_mse: %.3e" % rel(dg.cpu().numpy(), 64 * 10.0 * (g - y) / n)[1])
    # colsum
    X = rng.standard_normal((1000, 77)).astype(np.float32)
    Xd = torch.zeros(1000, 80, dtype=h.h16, device=dev); Xd[:, :77] = tt(X)
    cs = torch.zeros(77, device=dev)
    h.colsum16(Xd, 1000, 77, cs)
    print("colsum16: relrms %.3e" % rel(cs.cpu().numpy(), Xd[:, :77].double().sum(0).cpu().numpy())[1])
    # update sweep: 3 segments
    sizes = [1024 * 3, 1024, 2048]
    n_el = sum(sizes)
    seg_id = np.concatenate([np.full(s // 1024, i) for i, s in enumerate(sizes)]).astype(np.int32)
    theta = rng.standard_normal(n_el).astype(np.float32); grad = rng.standard_normal(n_el).astype(np.float32)
    grad[:3072] *= 5.0     # first segment exceeds the clip norm
    d_theta, d_grad, d_seg = tt(theta), tt(grad * 8.0), tt(seg_id)
    d_ema = tt(theta); d_m = torch.zeros(n_el, device=dev); d_v = torch.zeros(n_el, device=dev)
    sumsq = torch.zeros(3, device=dev); th16 = torch.zeros(n_el, dtype=h.h16, device=dev)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999, 0, 0], dtype=torch.float32, device=dev)
    p = {"a": theta[:3072].astype(np.float64), "b": theta[3072:4096].astype(np.float64), "c": theta[4096:].astype(np.float64)}
    gd = {"a": grad[:3072].astype(np.float64), "b": grad[3072:4096].astype(np.float64), "c": grad[4096:].astype(np.float64)}
    m = {k: np.zeros_like(v) for k, v in p.items()}; v_ = {k: np.zeros_like(v) for k, v in p.items()}
    ema = {k: v.copy() for k, v in p.items()}
    tstep = 0
    for it in range(3):
        h.seg_sumsq(d_grad, 1.0 / 8.0, d_seg, 3, sumsq)
        h.clip_adam_ema(d_grad, 1.0 / 8.0, d_seg, sumsq, 15.0, hyper, 0.9999, d_theta, d_m, d_v, d_ema, th16)
        cl = {k: O.clip_by_norm(gd[k], 15.0) for k in gd}
        p, m, v_, tstep = O.adam_update_tf(p, cl, m, v_, tstep, 1e-3)
        ema = O.ema_update(ema, p)
    ref_t = np.concatenate([p["a"], p["b"], p["c"]]); ref_e = np.concatenate([ema["a"], ema["b"], ema["c"]])
    print("adam theta: maxabs %.3e ; ema maxabs %.3e ; theta16 maxabs %.3e" %
          (rel(d_theta.cpu().numpy(), ref_t)[0], rel(d_ema.cpu().numpy(), ref_e)[0],
           rel(th16.float().cpu().numpy(), ref_t)[0]))
    d_theta2 = tt(theta); d_ema2 = tt(theta)
    hy2 = torch.tensor([0.05, 0, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device=dev)
    h.seg_sumsq(d_grad, 1.0 / 8.0, d_seg, 3, sumsq)
    h.clip_sgd_ema(d_grad, 1.0 / 8.0, d_seg, sumsq, 15.0, hy2, 0.9999, d_theta2, d_ema2, None)
    ref = np.concatenate([theta[a:b] - 0.05 * O.clip_by_norm(grad[a:b].astype(np.float64)) for a, b in ((0, 3072), (3072, 4096), (4096, 6144))])
    print("sgd theta: maxabs %.3e" % rel(d_theta2.cpu().numpy(), ref)[0])


if __name__ == "__main__":
    which = sys.argv[1:] or ["elt", "gemm", "rec"]
    dtype = os.environ.get("RSR_DTYPE", "f16")
    h = ops.Handle(0, dtype)
    print("device", torch.cuda.get_device_name(0), "sms", h.num_sms, "dtype", dtype, flush=True)
    if "elt" in which:
        probe_elt(h)
    if "gemm" in which:
        probe_gemm(h)
    if "rec" in which:
        probe_rec(h)                                    # D-LSTM sizes
        probe_rec(h, B=8, T=20, I=280, C=760, P=280)    # ref-native G layer
        probe_rec(h, B=40, T=10, I=256, C=512, P=256, ragged=False)   # multi-group
    torch.cuda.synchronize()
    print("probe done")
