"""Parity of every sm_100a kernel with float64 CPU math, through the C ABI (ctypes -> librsrgan_sm100.so)."""
import numpy as np
import pytest
import torch

from oracle import rsr_oracle as O
from rsrgan_b200 import packing

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["f16", "bf16"])
def h(request):
    from rsrgan_b200 import ops
    hd = ops.Handle(0, request.param)
    yield hd
    hd.close()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def tol(h, f16, bf16):
    return f16 if h.dtype_id == 0 else bf16


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 40, 257), (333, 280, 40), (128, 16, 64), (1000, 1, 1024),
                                   (512, 1024, 1024), (257, 3040, 560), (1, 8, 8), (4100, 264, 3072)])
@pytest.mark.parametrize("a_mn,b_mn", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_all_operand_layouts(h, M, N, K, a_mn, b_mn):
    """tcgen05 GEMM == exact product of the 16-bit operands (fp32 accumulate): rel RMS 1e-5."""
    dev, rng = h.device, np.random.default_rng(M + N + K)
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
    ld32 = packing.round_up(N, 4) + 4
    out = torch.full((M, ld32), 7.0, dtype=torch.float32, device=dev)
    h.gemm(At, Bt, M, N, K, a_mn=a_mn, b_mn=b_mn, out32=out)
    torch.cuda.synchronize()
    A16 = (At[:, :M].T if a_mn else At[:, :K]).double().cpu().numpy()
    B16 = (Bt[:, :N] if b_mn else Bt[:, :K].T).double().cpu().numpy()
    assert rel(out[:, :N].cpu().numpy(), A16 @ B16) < 1e-5
    assert bool((out[:, N:] == 7.0).all())               # nothing written outside [M, N]


def test_gemm_fused_epilogue(h):
    dev, rng = h.device, np.random.default_rng(0)
    M, N, K = 300, 280, 257
    A = rng.standard_normal((M, K)).astype(np.float32) * 0.3
    W = rng.standard_normal((K, N)).astype(np.float32) * 0.1
    bias, resid = rng.standard_normal(N).astype(np.float32), rng.standard_normal((M, N)).astype(np.float32)
    dsrc, old = rng.standard_normal((M, N)).astype(np.float32), rng.standard_normal((M, N)).astype(np.float32)
    At = torch.zeros(M, packing.round_up(K, 8), dtype=h.h16, device=dev); At[:, :K] = torch.tensor(A)
    Wt = torch.tensor(W, device=dev).to(h.h16)
    ds = torch.tensor(dsrc, device=dev).to(h.h16)
    o32, o16 = torch.tensor(old, device=dev), torch.zeros(M, N, dtype=h.h16, device=dev)
    from rsrgan_b200 import ops
    h.gemm(At, Wt, M, N, K, b_mn=True, alpha=0.5, beta=2.0, bias=torch.tensor(bias, device=dev),
           resid=torch.tensor(resid, device=dev), act=ops.ACT_LRELU, dact_src=ds, dact=ops.ACT_RELU, out32=o32, out16=o16)
    torch.cuda.synchronize()
    v = 0.5 * (At[:, :K].double().cpu().numpy() @ Wt.double().cpu().numpy()) + bias + resid
    v = np.maximum(v, 0.3 * v) * (ds.double().cpu().numpy() > 0)
    assert rel(o32.cpu().numpy(), v + 2.0 * old) < 1e-5
    assert rel(o16.float().cpu().numpy(), v) < tol(h, 5e-4, 4e-3)


def _rec_case(h, B, T, I, C, P, ragged, seed, lengths=None):
    dev, rng = h.device, np.random.default_rng(seed)
    Cp = packing.cell_pad(C)
    x = rng.standard_normal((B, T, I))
    K = O.xavier(rng, (I + P, 4 * C)) * 2.0
    b = rng.standard_normal(4 * C) * 0.1
    wi, wf, wo = (O.xavier(rng, (C,)) for _ in range(3))
    Wp = O.xavier(rng, (C, P)) * 2.0
    drawn = rng.integers(max(T // 2, 1), T + 1, size=B) if ragged else np.full(B, T)
    lengths = drawn if lengths is None else np.asarray(lengths)
    out_ref, cache = O.lstmp_fwd(x, lengths, K, b, wi, wf, wo, Wp)
    steps = cache[-1]
    zx = np.einsum("bti,ig->tbg", x, K[:I]) + b
    zx_p = packing.pack_cols(zx.reshape(T * B, 4 * C), C).astype(np.float32)
    Wc_p = packing.pad_first(packing.pack_cols(Wp @ K[I:], C), Cp)
    wc16 = torch.tensor(Wc_p, device=dev).to(h.h16).contiguous()
    wcT16 = wc16.t().contiguous()
    pk = lambda v: torch.tensor(packing.pad_last(v, Cp).astype(np.float32), device=dev)
    d_wi, d_wf, d_wo = pk(wi), pk(wf), pk(wo)
    d_len = torch.tensor(lengths.astype(np.int32), device=dev)
    mt_seq = torch.zeros((T + 1) * B, Cp, dtype=h.h16, device=dev)
    save = torch.zeros(T * B, 5, Cp, dtype=torch.float32, device=dev)
    h.lstmp_rec_fwd(B, T, Cp, torch.tensor(zx_p, device=dev), wcT16, d_wi, d_wf, d_wo, d_len, mt_seq, save)
    torch.cuda.synchronize()
    mt_ref = np.stack([np.where(steps[t][9], steps[t][8], 0.0) for t in range(T)])
    got = mt_seq[B:].float().cpu().numpy().reshape(T, B, Cp)
    res = dict(mt=rel(got[:, :, :C], mt_ref), pad=float(np.abs(got[:, :, C:]).max()) if Cp > C else 0.0,
               out=rel((got[:, :, :C] @ Wp).transpose(1, 0, 2), out_ref))
    dout = rng.standard_normal((B, T, P)) * 0.1
    dx_ref, g_ref = O.lstmp_bwd(dout, cache)
    dmt = np.einsum("btp,cp->tbc", dout, Wp).reshape(T * B, C)
    d_dmt = torch.tensor(packing.pad_last(dmt, Cp).astype(np.float32), device=dev)
    dz16 = torch.zeros(T * B, 4 * Cp, dtype=h.h16, device=dev)
    dbias = torch.zeros(4 * Cp, dtype=torch.float32, device=dev)
    dwi, dwf, dwo = (torch.zeros(Cp, dtype=torch.float32, device=dev) for _ in range(3))
    h.lstmp_rec_bwd(B, T, Cp, d_dmt, wc16, d_wi, d_wf, d_wo, d_len, save, dz16, dbias, dwi, dwf, dwo)
    torch.cuda.synchronize()
    dz = packing.unpack_cols(dz16.float().cpu().numpy(), C).reshape(T, B, 4 * C)
    xin = np.concatenate([x.transpose(1, 0, 2), np.concatenate([np.zeros((1, B, P)), out_ref.transpose(1, 0, 2)[:-1]], 0)], 2)
    res.update(dbias=rel(packing.unpack_cols(dbias.cpu().numpy(), C), g_ref["bias"]),
               dwi=rel(dwi.cpu().numpy()[:C], g_ref["w_i_diag"]), dwf=rel(dwf.cpu().numpy()[:C], g_ref["w_f_diag"]),
               dwo=rel(dwo.cpu().numpy()[:C], g_ref["w_o_diag"]),
               dx=rel(np.einsum("tbg,ig->bti", dz, K[:I]), dx_ref), dK=rel(np.einsum("tbi,tbg->ig", xin, dz), g_ref["kernel"]))
    return res


@pytest.mark.parametrize("B,T,I,C,P,ragged", [
    (8, 12, 40, 256, 40, True),          # discriminator_lstm layer (models/discriminator_lstm.py:26-28)
    (8, 20, 280, 760, 280, True),        # lstm generator layer (models/lstm.py:43-45), C padded 760 -> 768
    (40, 10, 256, 512, 256, False),      # BASELINE cfg-2 layer, several utterance groups
    (1, 9, 257, 760, 257, True),         # decode: one utterance (train...py batch_size=1), res_lstm_l layer
    (3, 1, 40, 256, 40, False),          # a single frame
    (20, 7, 257, 1024, 257, True),       # BASELINE configs[4] layer (res_lstm_l, C = 1024): weight slab half in TMEM
    (64, 5, 257, 1000, 257, False),      # same family at the cfg-5 batch (two groups of 32 in the backward), C padded
    (50, 6, 257, 1024, 257, True),       # ... ragged, second group partial (256-thread backward blocks)
])
def test_lstmp_recurrence_fwd_bwd(h, B, T, I, C, P, ragged):
    r = _rec_case(h, B, T, I, C, P, ragged, seed=B + T)
    t = tol(h, 1.5e-3, 1e-2)
    assert r["pad"] == 0.0
    for k, v in r.items():
        if k != "pad":
            assert v < t, (k, v, r)


@pytest.mark.parametrize("B,T,I,C,P,ragged", [
    (2, 1000, 40, 256, 40, True),        # discriminator_lstm layer, two long ragged utterances
    (1, 1200, 257, 760, 257, False),     # decode of ONE whole utterance (batch_size = 1) through a res_lstm_l layer
    (8, 800, 256, 512, 256, True),       # cfg-2 layer at the reference driver's batch of 8
])
def test_lstmp_recurrence_long_utterances(h, B, T, I, C, P, ragged):
    """Utterance-scale lengths (the reference trains on and decodes whole utterances of up to ~1000 frames,
    scripts/train_gan_rnn_placeholder.py:262-300): forward states and every gradient against the float64 oracle after
    800-1200 dependent steps.  Bars = 2x the measured deviation (profiles/r2_long_T_v1.jsonl: fp16 <= 9.7e-4, bf16 <= 7.4e-3,
    relative RMS)."""
    r = _rec_case(h, B, T, I, C, P, ragged, seed=B + T)
    t = tol(h, 2e-3, 1.5e-2)
    assert r["pad"] == 0.0
    for k, v in r.items():
        if k != "pad":
            assert v < t, (k, v, r)


@pytest.mark.parametrize("B,T,I,C,P,lengths", [
    (4, 6, 40, 256, 40, [0, 6, 3, 0]),                     # cluster / pair kernels: first and last utterance empty
    (3, 5, 257, 760, 257, [5, 0, 1]),                      # L2-exchange kernels (Cp = 768)
    (40, 8, 256, 512, 256, [0] * 32 + [8, 0, 3, 8, 0, 1, 2, 8]),   # a whole utterance group of empty rows
])
def test_lstmp_recurrence_empty_utterances(h, B, T, I, C, P, lengths):
    """sequence_length = 0 (tf.nn.dynamic_rnn copies zero state through and emits zeros, models/lstm.py:104-112): the rows
    of an empty utterance stay exactly zero, it contributes nothing to any gradient, and its neighbours are unaffected."""
    r = _rec_case(h, B, T, I, C, P, True, seed=B + T, lengths=lengths)
    t = tol(h, 1.5e-3, 1e-2)
    assert r["pad"] == 0.0
    for k, v in r.items():
        if k != "pad":
            assert v < t, (k, v, r)


@pytest.mark.parametrize("B,T,I,C,P", [(40, 10, 256, 512, 256), (8, 12, 40, 256, 40)])
def test_lstmp_recurrence_l2_exchange_variant(h, monkeypatch, B, T, I, C, P):
    """Cp <= 512 normally runs the cluster/DSMEM kernels; RSR_NO_CLUSTER forces the L2-exchange kernels
    (the only variant for Cp > 512) on the same shapes."""
    monkeypatch.setenv("RSR_NO_CLUSTER", "1")
    r = _rec_case(h, B, T, I, C, P, True, seed=7)
    t = tol(h, 1.5e-3, 1e-2)
    assert r["pad"] == 0.0
    for k, v in r.items():
        if k != "pad":
            assert v < t, (k, v, r)


def test_lstmp_forward_l2_multicast_exchange(h, monkeypatch):
    """RSR_FWD_XCHG=l2mc: mt_t travels through its global copy + one multicast TMA load per CTA instead of st.async."""
    monkeypatch.setenv("RSR_FWD_XCHG", "l2mc")
    for (B, T, I, C, P, ragged) in ((40, 10, 256, 512, 256, True), (8, 12, 40, 256, 40, True)):
        r = _rec_case(h, B, T, I, C, P, ragged, seed=11)
        assert r["pad"] == 0.0
        for k, v in r.items():
            if k != "pad":
                assert v < tol(h, 1.5e-3, 1e-2), (k, v, r)


@pytest.mark.parametrize("B,T,I,C,P,ragged", [
    (40, 10, 256, 512, 256, True),       # BASELINE cfg-2 layer: three groups, the last one partial
    (8, 12, 40, 256, 40, True),          # discriminator_lstm layer: Ik = 48 > I (zero padded K)
    (20, 6, 257, 256, 257, False),       # 257-wide input (res_lstm_l): ldx = 264, Ik = 272, five 64-wide x sub-tiles
])
def test_lstmp_fused_forward(h, B, T, I, C, P, ragged):
    """rsr_lstmp_fused_fwd (x_t K_x + b + mt_{t-1} Wc + gates in one kernel) == oracle LSTMP forward."""
    dev, rng = h.device, np.random.default_rng(B + I)
    Cp, Ip, Ik = packing.cell_pad(C), packing.round_up(I, 8), packing.round_up(I, 16)
    x = rng.standard_normal((B, T, I))
    K = O.xavier(rng, (I + P, 4 * C)) * 2.0
    b = rng.standard_normal(4 * C) * 0.1
    wi, wf, wo = (O.xavier(rng, (C,)) for _ in range(3))
    Wp = O.xavier(rng, (C, P)) * 2.0
    lengths = rng.integers(max(T // 2, 1), T + 1, size=B) if ragged else np.full(B, T)
    out_ref, cache = O.lstmp_fwd(x, lengths, K, b, wi, wf, wo, Wp)
    steps = cache[-1]
    x16 = torch.zeros(T * B, Ip, dtype=h.h16, device=dev)
    x16[:, :I] = torch.tensor(x.transpose(1, 0, 2).reshape(T * B, I), device=dev).to(h.h16)
    kx_p = packing.pack_cols(K[:I], C)                                   # [I, 4Cp]
    kxT = torch.zeros(4 * Cp, Ik, dtype=h.h16, device=dev)
    kxT[:, :I] = torch.tensor(kx_p.T.copy(), device=dev).to(h.h16)
    Wc_p = packing.pad_first(packing.pack_cols(Wp @ K[I:], C), Cp)
    wcT16 = torch.tensor(Wc_p, device=dev).to(h.h16).t().contiguous()
    pk = lambda v: torch.tensor(packing.pad_last(v, Cp).astype(np.float32), device=dev)
    bias_p = torch.tensor(packing.pack_cols(b[None], C)[0].astype(np.float32), device=dev)
    mt_seq = torch.zeros((T + 1) * B, Cp, dtype=h.h16, device=dev)
    save = torch.zeros(T * B, 5, Cp, dtype=torch.float32, device=dev)
    ok = h.lstmp_fused_fwd(B, T, I, Cp, x16, kxT, bias_p, wcT16, pk(wi), pk(wf), pk(wo),
                           torch.tensor(lengths.astype(np.int32), device=dev), mt_seq, save)
    torch.cuda.synchronize()
    assert ok
    mt_ref = np.stack([np.where(steps[t][9], steps[t][8], 0.0) for t in range(T)])
    got = mt_seq[B:].float().cpu().numpy().reshape(T, B, Cp)
    t16 = tol(h, 1.5e-3, 1e-2)
    assert rel(got[:, :, :C], mt_ref) < t16
    assert Cp == C or float(np.abs(got[:, :, C:]).max()) == 0.0
    assert rel((got[:, :, :C] @ Wp).transpose(1, 0, 2), out_ref) < t16
    # transposition helper used to keep K_x^T in step with the weights
    src = torch.tensor(rng.standard_normal((70, 130)).astype(np.float32), device=dev).to(h.h16)
    dst = torch.zeros(136, 72, dtype=h.h16, device=dev)
    h.transpose16(src, 70, 130, dst)
    assert torch.equal(dst[:130, :70], src.t()) and float(dst[130:].abs().max()) == 0 and float(dst[:, 70:].abs().max()) == 0


def _fused_operands(h, rng, I, C, P, Cp):
    """Random LSTMP layer: oracle weights + the device operands of the fused forward kernels."""
    dev = h.device
    Ik = packing.round_up(I, 16)
    K = O.xavier(rng, (I + P, 4 * C)) * 2.0
    b = rng.standard_normal(4 * C) * 0.1
    wi, wf, wo = (O.xavier(rng, (C,)) for _ in range(3))
    Wp = O.xavier(rng, (C, P)) * 2.0
    kxT = torch.zeros(4 * Cp, Ik, dtype=h.h16, device=dev)
    kxT[:, :I] = torch.tensor(packing.pack_cols(K[:I], C).T.copy(), device=dev).to(h.h16)
    wcT16 = torch.tensor(packing.pad_first(packing.pack_cols(Wp @ K[I:], C), Cp), device=dev).to(h.h16).t().contiguous()
    pk = lambda v: torch.tensor(packing.pad_last(v, Cp).astype(np.float32), device=dev)
    bias_p = torch.tensor(packing.pack_cols(b[None], C)[0].astype(np.float32), device=dev)
    Pp = packing.round_up(P, 8)
    wpT = torch.zeros(Pp, Cp, dtype=h.h16, device=dev)
    wpT[:P, :C] = torch.tensor(Wp.T.copy(), device=dev).to(h.h16)
    return (K, b, wi, wf, wo, Wp), (kxT, bias_p, wcT16, pk(wi), pk(wf), pk(wo)), wpT


@pytest.mark.parametrize("B,T,I,C,P,ragged,nbp", [
    (40, 10, 256, 512, 256, True, 0),       # BASELINE cfg-2 stack: 32 utterances per cluster, 2 + 2 + 1 clusters
    (128, 12, 256, 512, 256, True, 0),      # ... at the benchmarked batch: 48 per cluster, 3 + 3 + 1 clusters (all that fit)
    (100, 7, 256, 512, 256, False, 48),     # last group partial (4 of 48)
    (8, 12, 40, 256, 40, True, 0),          # discriminator_lstm stack: 8-CTA clusters, one 40-wide feature tile
    (70, 9, 257, 256, 257, True, 0),        # 257-wide projection: three feature tiles, ldo = 264, Ik = 272
    (3, 1, 40, 256, 40, False, 32),         # a single frame
])
def test_lstmp_wave_forward(h, monkeypatch, B, T, I, C, P, ragged, nbp):
    """rsr_lstmp_wave_fwd (two stacked LSTMP layers, layer 2 a few steps behind layer 1 in the same launch) == the oracle's
    two dynamic_rnn layers, and == the one-after-the-other kernels bit for bit."""
    if nbp:
        monkeypatch.setenv("RSR_WAVE_NBP", str(nbp))
    dev, rng = h.device, np.random.default_rng(B + I + T)
    Cp, Ip, Pp = packing.cell_pad(C), packing.round_up(I, 8), packing.round_up(P, 8)
    x = rng.standard_normal((B, T, I))
    lengths = rng.integers(max(T // 2, 1), T + 1, size=B) if ragged else np.full(B, T)
    (w1, d1, wpT1), (w2, d2, _) = _fused_operands(h, rng, I, C, P, Cp), _fused_operands(h, rng, P, C, P, Cp)
    out1_ref, c1 = O.lstmp_fwd(x, lengths, *w1)
    out2_ref, c2 = O.lstmp_fwd(out1_ref, lengths, *w2)
    x16 = torch.zeros(T * B, Ip, dtype=h.h16, device=dev)
    x16[:, :I] = torch.tensor(x.transpose(1, 0, 2).reshape(T * B, I), device=dev).to(h.h16)
    d_len = torch.tensor(lengths.astype(np.int32), device=dev)
    mk = lambda cols, dt: torch.zeros((T + 1) * B, cols, dtype=dt, device=dev)
    mt1, mt2, out1 = mk(Cp, h.h16), mk(Cp, h.h16), mk(Pp, h.h16)
    sv1, sv2 = (torch.zeros(T * B, 5 * Cp, dtype=torch.float32, device=dev) for _ in range(2))
    ok = h.lstmp_wave_fwd(B, T, Cp, I, P, d_len, x16, d1, mt1, sv1, wpT1, out1, d2, mt2, sv2)
    torch.cuda.synchronize()
    assert ok
    t16 = tol(h, 1.5e-3, 1e-2)
    mref = lambda c: np.stack([np.where(c[-1][t][9], c[-1][t][8], 0.0) for t in range(T)])
    got1 = mt1[B:].float().cpu().numpy().reshape(T, B, Cp)
    got2 = mt2[B:].float().cpu().numpy().reshape(T, B, Cp)
    o1 = out1[B:].float().cpu().numpy().reshape(T, B, Pp)
    assert rel(got1[:, :, :C], mref(c1)) < t16
    assert rel(o1[:, :, :P].transpose(1, 0, 2), out1_ref) < t16
    assert rel(got2[:, :, :C], mref(c2)) < 2 * t16
    assert float(out1[:B].abs().max()) == 0.0 and (Pp == P or float(np.abs(o1[:, :, P:]).max()) == 0.0)
    # the same two layers one after the other (fused forward, projection GEMM, fused forward)
    s_mt1, s_mt2, s_out1 = mk(Cp, h.h16), mk(Cp, h.h16), mk(Pp, h.h16)
    s_sv1, s_sv2 = torch.zeros_like(sv1), torch.zeros_like(sv2)
    assert h.lstmp_fused_fwd(B, T, I, Cp, x16, *d1, d_len, s_mt1, s_sv1)
    h.gemm(s_mt1[B:], wpT1, T * B, Pp, Cp, out16=s_out1[B:])
    assert h.lstmp_fused_fwd(B, T, P, Cp, s_out1[B:], *d2, d_len, s_mt2, s_sv2)
    torch.cuda.synchronize()
    assert torch.equal(mt1, s_mt1) and torch.equal(sv1, s_sv1)
    assert float((out1.float() - s_out1.float()).abs().max()) <= 2e-3 * float(s_out1.float().abs().max())
    assert rel(got2, s_mt2[B:].float().cpu().numpy().reshape(T, B, Cp)) < 2e-3


@pytest.mark.parametrize("B,T,I,C,P,ragged,nbp", [
    (40, 10, 256, 512, 256, True, 0),       # BASELINE cfg-2 stack: 32 utterances per cluster
    (128, 12, 256, 512, 256, True, 0),      # ... at the benchmarked batch: 48 per cluster, all 7 placeable clusters
    (100, 7, 256, 512, 256, False, 48),     # last group partial
    (8, 12, 40, 256, 40, True, 0),          # discriminator_lstm stack: 8-CTA clusters
    (3, 1, 40, 256, 40, False, 32),         # a single frame
])
def test_lstmp_wave_backward(h, monkeypatch, B, T, I, C, P, ragged, nbp):
    """rsr_lstmp_wave_bwd (reversed recurrences of two stacked layers in one launch, dmt1_t = dz2_t (W_p1 K_x2)^T formed
    between them) == the oracle's gradients of the two dynamic_rnn layers, and ~= the one-after-the-other kernels."""
    if nbp:
        monkeypatch.setenv("RSR_WAVE_NBP", str(nbp))
    dev, rng = h.device, np.random.default_rng(B + I + T)
    Cp, Ip, Pp = packing.cell_pad(C), packing.round_up(I, 8), packing.round_up(P, 8)
    x = rng.standard_normal((B, T, I))
    lengths = rng.integers(max(T // 2, 1), T + 1, size=B) if ragged else np.full(B, T)
    (w1, d1, wpT1), (w2, d2, wpT2) = _fused_operands(h, rng, I, C, P, Cp), _fused_operands(h, rng, P, C, P, Cp)
    out1_ref, c1 = O.lstmp_fwd(x, lengths, *w1)
    out2_ref, c2 = O.lstmp_fwd(out1_ref, lengths, *w2)
    dout2 = rng.standard_normal((B, T, P)) * 0.1
    dx2_ref, g2 = O.lstmp_bwd(dout2, c2)
    dx1_ref, g1 = O.lstmp_bwd(dx2_ref, c1)
    x16 = torch.zeros(T * B, Ip, dtype=h.h16, device=dev)
    x16[:, :I] = torch.tensor(x.transpose(1, 0, 2).reshape(T * B, I), device=dev).to(h.h16)
    d_len = torch.tensor(lengths.astype(np.int32), device=dev)
    mk = lambda cols, dt: torch.zeros((T + 1) * B, cols, dtype=dt, device=dev)
    mt1, mt2, out1 = mk(Cp, h.h16), mk(Cp, h.h16), mk(Pp, h.h16)
    sv1, sv2 = (torch.zeros(T * B, 5 * Cp, dtype=torch.float32, device=dev) for _ in range(2))
    assert h.lstmp_wave_fwd(B, T, Cp, I, P, d_len, x16, d1, mt1, sv1, wpT1, out1, d2, mt2, sv2)
    # operands of the backward: Wc (untransposed), dmt2 = dOut2 W_p2^T, F = W_p1 K_x2 (16-bit operands, rounded once)
    wc1, wc2 = d1[2].t().contiguous(), d2[2].t().contiguous()
    dmt2 = torch.tensor(packing.pad_last(np.einsum("btp,cp->tbc", dout2, w2[5]).reshape(T * B, C), Cp).astype(np.float32), device=dev)
    kx2 = torch.zeros(Pp, 4 * Cp, dtype=h.h16, device=dev)
    kx2[:P] = torch.tensor(packing.pack_cols(w2[0][:P], C), device=dev).to(h.h16)
    fT = (wpT1.float().t() @ kx2.float()).to(h.h16).contiguous()
    z = lambda n: torch.zeros(n, dtype=torch.float32, device=dev)
    grads = lambda: (z(4 * Cp), z(Cp), z(Cp), z(Cp))
    dz1, dz2 = mk(4 * Cp, h.h16), mk(4 * Cp, h.h16)
    gw1, gw2 = grads(), grads()
    part = torch.zeros(T * (B + 48), Cp, dtype=torch.float32, device=dev)
    ok = h.lstmp_wave_bwd(B, T, Cp, d_len, dmt2, (wc2,) + tuple(d2[3:6]), sv2, dz2, gw2, fT, part,
                          (wc1,) + tuple(d1[3:6]), sv1, dz1, gw1)
    torch.cuda.synchronize()
    assert ok
    t16 = tol(h, 3e-3, 2e-2)

    def check(dz, gw, w, xin, dx_ref, g_ref, Iw, t):
        dzu = packing.unpack_cols(dz[:T * B].float().cpu().numpy(), C).reshape(T, B, 4 * C)
        assert rel(packing.unpack_cols(gw[0].cpu().numpy(), C), g_ref["bias"]) < t
        for k, n in ((1, "w_i_diag"), (2, "w_f_diag"), (3, "w_o_diag")):
            assert rel(gw[k].cpu().numpy()[:C], g_ref[n]) < t, n
        assert rel(np.einsum("tbg,ig->bti", dzu, w[0][:Iw]), dx_ref) < t
        assert rel(np.einsum("tbi,tbg->ig", xin, dzu), g_ref["kernel"]) < t

    prev = lambda o: np.concatenate([np.zeros((1, B, P)), o.transpose(1, 0, 2)[:-1]], 0)
    check(dz2, gw2, w2, np.concatenate([out1_ref.transpose(1, 0, 2), prev(out2_ref)], 2), dx2_ref, g2, P, t16)
    check(dz1, gw1, w1, np.concatenate([x.transpose(1, 0, 2), prev(out1_ref)], 2), dx1_ref, g1, I, 2 * t16)
    # the same two layers one after the other
    s_dz1, s_dz2 = mk(4 * Cp, h.h16), mk(4 * Cp, h.h16)
    s_gw1, s_gw2 = grads(), grads()
    h.lstmp_rec_bwd(B, T, Cp, dmt2, wc2, *d2[3:6], d_len, sv2, s_dz2, *s_gw2)
    dx2_16 = torch.zeros(T * B, Pp, dtype=h.h16, device=dev)
    h.gemm(s_dz2, kx2, T * B, Pp, 4 * Cp, out16=dx2_16)
    dmt1 = torch.zeros(T * B, Cp, dtype=torch.float32, device=dev)
    h.gemm(dx2_16, wpT1.t().contiguous(), T * B, Cp, Pp, out32=dmt1)
    h.lstmp_rec_bwd(B, T, Cp, dmt1, wc1, *d1[3:6], d_len, sv1, s_dz1, *s_gw1)
    torch.cuda.synchronize()
    if B > 16:        # (B <= 16: rsr_lstmp_rec_bwd runs the single-CTA cluster kernel, another summation order)
        assert torch.equal(dz2, s_dz2)
    assert rel(dz2.float().cpu().numpy(), s_dz2.float().cpu().numpy()) < tol(h, 1e-3, 8e-3)
    assert rel(dz1.float().cpu().numpy(), s_dz1.float().cpu().numpy()) < tol(h, 4e-3, 3e-2)
    assert float(part.abs().max()) == 0.0        # the scratch is handed back zeroed


def test_gauss_noise_stream(h):
    """rsr_gauss_noise == the oracle's restatement of the counter-based stream; draws differ by tick and salt."""
    dev = h.device
    rng = torch.tensor([4242, 7], dtype=torch.int64, device=dev)
    out = torch.zeros(64, 40, dtype=torch.float32, device=dev)
    h.gauss_noise(rng, 0x4e01, out, 0.05)
    ref = O.gauss_noise(4242, 7, 0x4e01, 64 * 40, 0.05).reshape(64, 40)
    got = out.cpu().numpy()
    assert np.abs(got - ref).max() < 2e-6
    h.rng_tick(rng)
    out2 = torch.zeros_like(out)
    h.gauss_noise(rng, 0x4e01, out2, 0.05)
    torch.cuda.synchronize()
    assert int(rng[1]) == 8 and np.abs(out2.cpu().numpy() - O.gauss_noise(4242, 8, 0x4e01, 64 * 40, 0.05).reshape(64, 40)).max() < 2e-6
    assert np.abs(out2.cpu().numpy() - got).max() > 0.01


def test_lstmp_shape_errors(h):
    z = torch.zeros(8, device=h.device)
    from rsrgan_b200 import _lib
    with pytest.raises(_lib.RsrError):
        h.lstmp_rec_fwd(8, 4, 100, z, z, z, z, z, z.int(), z, None)        # Cp not a multiple of 256
    with pytest.raises(_lib.RsrError):
        h.lstmp_rec_fwd(4096, 4, 768, z, z, z, z, z, z.int(), z, None)     # L2-exchange groups would not be co-resident


def test_staging_losses_update(h):
    dev, rng = h.device, np.random.default_rng(2)
    tt = lambda a: torch.tensor(a, device=dev)
    B, T, D = 5, 7, 40
    x = rng.standard_normal((B, T, D)).astype(np.float32)
    mean, std = rng.standard_normal(D).astype(np.float32), (rng.random(D) + 0.5).astype(np.float32)
    noise = rng.standard_normal((B, D)).astype(np.float32)
    o16 = torch.zeros(T * B, 40, dtype=h.h16, device=dev)
    o32 = torch.zeros(T * B, 40, dtype=torch.float32, device=dev)
    h.stage_input(tt(x), B, T, D, out16=o16, out32=o32, mean=tt(mean), istd=tt(1.0 / std), noise=tt(noise))
    ref = ((x - mean) * (1.0 / std) + noise[:, None, :]).transpose(1, 0, 2).reshape(T * B, D)
    assert np.allclose(o32.cpu().numpy(), ref, atol=1e-6)
    assert rel(o16.float().cpu().numpy(), ref) < tol(h, 5e-4, 4e-3)
    back = torch.zeros(B, T, D, dtype=torch.float32, device=dev)
    h.unstage_output(o32, B, T, D, back)
    assert np.array_equal(back.cpu().numpy(), o32.cpu().numpy().reshape(T, B, D).transpose(1, 0, 2))
    # CMVN apply / invert (io_funcs/make_tfrecords.py:84-87; train...py:286-287): fp32 within 1 ulp-ish of float64 math
    xm = rng.standard_normal((1000, 257)).astype(np.float32) * 3 + 1
    m2, s2 = rng.standard_normal(257).astype(np.float32), (rng.random(257) + 0.5).astype(np.float32)
    out = torch.empty(1000, 257, device=dev)
    h.cmvn_apply(tt(xm), tt(m2), tt(s2), out)
    assert np.allclose(out.cpu().numpy(), O.cmvn_apply(xm, m2.astype(np.float64), s2.astype(np.float64)), rtol=2e-6, atol=2e-6)
    inv = torch.empty_like(out)
    h.cmvn_invert(out, tt(m2), tt(s2), inv)
    assert np.allclose(inv.cpu().numpy(), xm, rtol=1e-5, atol=1e-5)
    # losses + gradients, with the discriminator_dnn clip
    n = B * T
    rl, fk = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    g, y = rng.standard_normal((n, D)).astype(np.float32), rng.standard_normal((n, D)).astype(np.float32)
    losses = torch.zeros(8, dtype=torch.float32, device=dev)
    rl4 = torch.zeros(n, 4, device=dev); rl4[:, 0] = tt(rl)
    fk4 = torch.zeros(n, 4, device=dev); fk4[:, 0] = tt(fk)
    g1, g2, g3 = (torch.zeros(n, 8, dtype=h.h16, device=dev) for _ in range(3))
    dg = torch.zeros(n, D, device=dev)
    h.lsgan_mse_losses(losses, rl=rl4, fk=fk4, ld_logit=4, n_logit=n, clip=True, g=tt(g), y=tt(y), n_frames=n,
                       d_out=D, lam=10.0, gscale=64.0, d_rl_grad=g1, d_fk_grad=g2, g_adv_grad=g3, ld_grad=8, dg_mse=dg)
    rc, fc = np.clip(rl, -0.5, 1.5), np.clip(fk, -0.5, 1.5)
    L = O.lsgan_mse_losses(rc.astype(np.float64), fc.astype(np.float64), g.astype(np.float64), y.astype(np.float64))
    got = losses.cpu().numpy()
    for i, k in enumerate(("d_rl_loss", "d_fk_loss", "g_adv_loss", "g_mse_loss")):
        assert got[i] == pytest.approx(L[k], rel=1e-5)
    in_rl, in_fk = (rl >= -0.5) & (rl <= 1.5), (fk >= -0.5) & (fk <= 1.5)
    t16 = tol(h, 5e-4, 4e-3)
    assert rel(g1[:, 0].float().cpu().numpy(), 64 * 2 * (rc - 1) / n * in_rl) < t16
    assert rel(g2[:, 0].float().cpu().numpy(), 64 * 2 * (fc - 0) / n * in_fk) < t16
    assert rel(g3[:, 0].float().cpu().numpy(), 64 * 2 * (fc - 1) / n * in_fk) < t16
    assert rel(dg.cpu().numpy(), 64 * 10.0 * (g - y) / n) < 1e-6
    # column sums
    X = rng.standard_normal((1000, 77)).astype(np.float32)
    Xd = torch.zeros(1000, 80, dtype=h.h16, device=dev); Xd[:, :77] = tt(X)
    cs = torch.zeros(77, device=dev)
    h.colsum16(Xd, 1000, 77, cs)
    assert rel(cs.cpu().numpy(), Xd[:, :77].double().sum(0).cpu().numpy()) < 1e-5
    # 16-byte vectorised variant (N, ld multiples of 8), accumulating into a non-zero vector
    Xv = tt(rng.standard_normal((3001, 520)).astype(np.float32)).to(h.h16)
    cv = torch.ones(520, device=dev)
    h.colsum16(Xv, 3001, 520, cv, accumulate=True)
    assert rel(cv.cpu().numpy(), 1.0 + Xv.double().sum(0).cpu().numpy()) < 1e-5


def test_fused_clip_update_sweep(h):
    """rsr_seg_sumsq + rsr_clip_{adam,sgd}_ema == clip_by_norm per tensor, TF Adam, EMA (gan_rnn_placeholder.py:144-189)."""
    dev, rng = h.device, np.random.default_rng(3)
    tt = lambda a: torch.tensor(a, device=dev)
    sizes = [3072, 1024, 2048]
    n_el = sum(sizes)
    seg_id = np.concatenate([np.full(s // 1024, i) for i, s in enumerate(sizes)]).astype(np.int32)
    theta, grad = rng.standard_normal(n_el).astype(np.float32), rng.standard_normal(n_el).astype(np.float32) * 0.1
    grad[:3072] *= 50.0                                      # first tensor exceeds the clip norm, the others do not
    bounds = [(0, 3072), (3072, 4096), (4096, 6144)]
    d_theta, d_grad, d_seg, d_ema = tt(theta), tt(grad * 8.0), tt(seg_id), tt(theta)
    d_m, d_v = torch.zeros(n_el, device=dev), torch.zeros(n_el, device=dev)
    sumsq, th16 = torch.zeros(3, device=dev), torch.zeros(n_el, dtype=h.h16, device=dev)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999, 0, 0], dtype=torch.float32, device=dev)
    p = {i: theta[a:b].astype(np.float64) for i, (a, b) in enumerate(bounds)}
    gd = {i: grad[a:b].astype(np.float64) for i, (a, b) in enumerate(bounds)}
    m, v_ = {k: np.zeros_like(x) for k, x in p.items()}, {k: np.zeros_like(x) for k, x in p.items()}
    ema, tstep = {k: x.copy() for k, x in p.items()}, 0
    for _ in range(3):
        h.seg_sumsq(d_grad, 1.0 / 8.0, d_seg, 3, sumsq)
        h.clip_adam_ema(d_grad, 1.0 / 8.0, d_seg, sumsq, 15.0, hyper, 0.9999, d_theta, d_m, d_v, d_ema, th16)
        cl = {k: O.clip_by_norm(gd[k], 15.0) for k in gd}
        p, m, v_, tstep = O.adam_update_tf(p, cl, m, v_, tstep, 1e-3)
        ema = O.ema_update(ema, p)
    cat = lambda d: np.concatenate([d[0], d[1], d[2]])
    assert np.abs(d_theta.cpu().numpy() - cat(p)).max() < 1e-6
    assert np.abs(d_ema.cpu().numpy() - cat(ema)).max() < 1e-6
    assert rel(th16.float().cpu().numpy(), cat(p)) < tol(h, 5e-4, 4e-3)
    assert hyper[4].item() == pytest.approx(0.9 ** 4, rel=1e-6) and hyper[5].item() == pytest.approx(0.999 ** 4, rel=1e-6)
    d_theta2, d_ema2 = tt(theta), tt(theta)
    hy2 = torch.tensor([0.05, 0, 0, 0, 0, 0, 0, 0], dtype=torch.float32, device=dev)
    h.seg_sumsq(d_grad, 1.0 / 8.0, d_seg, 3, sumsq)
    h.clip_sgd_ema(d_grad, 1.0 / 8.0, d_seg, sumsq, 15.0, hy2, 0.9999, d_theta2, d_ema2, None)
    ref = np.concatenate([theta[a:b] - 0.05 * O.clip_by_norm(grad[a:b].astype(np.float64)) for a, b in bounds])
    assert np.abs(d_theta2.cpu().numpy() - ref).max() < 1e-6


@pytest.mark.parametrize("B,T,D", [(3, 7, 257), (5, 11, 40), (2, 9, 771)])
def test_cmvn_kernels(h, B, T, D):
    """rsr_cmvn_apply / _invert (fp32 vector stream over the flat matrix; D = 257 rows are not 16-byte aligned, totals not
    a multiple of 4) against numpy, and rsr_cmvn_apply_padded bit for bit against the loader's float64 statement of
    io_funcs/make_tfrecords.py:84-87 with zero padding after the normalisation (tfrecords_dataset.py:149-152)."""
    dev, rng = h.device, np.random.default_rng(B * T + D)
    x = (rng.standard_normal((B, T, D)) * 3 + 1).astype(np.float32)
    mean, std = rng.standard_normal(D), rng.uniform(0.5, 2.0, D)
    xd = torch.tensor(x.reshape(B * T, D), device=dev)
    out = torch.empty_like(xd)
    m32, s32 = torch.tensor(mean.astype(np.float32), device=dev), torch.tensor(std.astype(np.float32), device=dev)
    h.cmvn_apply(xd, m32, s32, out)
    ref = (x.reshape(B * T, D) - mean.astype(np.float32)) / std.astype(np.float32)
    assert np.allclose(out.cpu().numpy(), ref, rtol=2e-6, atol=1e-6)
    back = torch.empty_like(xd)
    h.cmvn_invert(out, m32, s32, back)
    assert np.allclose(back.cpu().numpy(), x.reshape(B * T, D), rtol=1e-5, atol=1e-5)
    lengths = rng.integers(1, T + 1, size=B).astype(np.int32)
    lengths[0] = T
    x3 = torch.tensor(x, device=dev)
    o3 = torch.empty_like(x3)
    h.cmvn_apply_padded(x3, torch.tensor(lengths, device=dev), torch.tensor(mean, device=dev), torch.tensor(std, device=dev), o3)
    want = ((x.astype(np.float64) - mean) / std).astype(np.float32)
    for b in range(B):
        want[b, lengths[b]:] = 0.0
    assert np.array_equal(o3.cpu().numpy(), want)
    h.cmvn_apply_padded(x3, torch.tensor(lengths, device=dev), torch.tensor(mean, device=dev), torch.tensor(std, device=dev), x3)
    assert np.array_equal(x3.cpu().numpy(), want)                  # in place


def test_update_sweep_overflow_guard(h):
    """n_seg > 0: a non-finite per-tensor norm (an fp16 overflow somewhere in the backward pass) leaves weights, Adam
    slots, shadows, the 16-bit copy and the beta powers untouched and counts the skipped update in hyper[7]."""
    dev = h.device
    n = 4 * 1024
    seg = torch.tensor([0, 0, 1, 2], dtype=torch.int32, device=dev)
    grad = torch.randn(n, device=dev)
    theta, m, v, ema = (torch.randn(n, device=dev) for _ in range(4))
    v.abs_()
    th16 = theta.to(h.h16)
    hyper = torch.tensor([1e-3, 0.9, 0.999, 1e-8, 0.9, 0.999, 0.0, 0.0], device=dev)
    sumsq = torch.zeros(3, device=dev)
    keep = [t.clone() for t in (theta, m, v, ema, th16, hyper)]
    grad[2048 + 7] = float("inf")
    h.seg_sumsq(grad, 1.0, seg, 3, sumsq)
    h.clip_adam_ema(grad, 1.0, seg, sumsq, 15.0, hyper, 0.9999, theta, m, v, ema, th16, n_seg=3)
    h.clip_sgd_ema(grad, 1.0, seg, sumsq, 15.0, hyper, 0.9999, theta, ema, th16, n_seg=3)
    torch.cuda.synchronize()
    assert not bool(torch.isfinite(sumsq).all())
    for t, k in zip((theta, m, v, ema, th16), keep):
        assert torch.equal(t, k)
    assert torch.equal(hyper[:7], keep[5][:7]) and float(hyper[7]) == 2.0
    grad[2048 + 7] = 0.5                                 # finite again: the update goes through
    h.seg_sumsq(grad, 1.0, seg, 3, sumsq)
    h.clip_adam_ema(grad, 1.0, seg, sumsq, 15.0, hyper, 0.9999, theta, m, v, ema, th16, n_seg=3)
    torch.cuda.synchronize()
    assert not torch.equal(theta, keep[0]) and float(hyper[7]) == 2.0 and float(hyper[4]) == pytest.approx(0.81)


@pytest.mark.parametrize("N,w,cin,cout", [(5, 13, 1, 12), (64, 11, 12, 16), (37, 9, 16, 20), (256, 7, 24, 32),
                                          (256, 9, 24, 20), (3, 13, 16, 12)])
def test_conv1d_same_overlapped_view_gemm(h, N, w, cin, cout):
    """models/rced.py:94-101 (splice = 1): forward, weight gradient and data gradient of the SAME convolution as
    rsr_gemm over the overlapped frame view + rsr_conv_mask_rows / rsr_conv_w_flip, against the oracle's
    conv1d_same_fwd / _bwd on the same 16-bit operands.  Tolerance: 16-bit output rounding (fp16 2^-11, bf16 2^-8
    relative) on top of the exact fp32-accumulated product."""
    from rsrgan_b200.nets import ConvFrames
    dev, rng = h.device, np.random.default_rng(N + w + cin)
    L = 257
    fl = ConvFrames(L, 13)
    S, G = fl.S, fl.GUARD
    cip, cop = packing.round_up(cin, 8), packing.round_up(cout, 8)
    x16 = torch.tensor(rng.standard_normal((N, L, cin)).astype(np.float32)).to(h.h16)
    W16 = torch.tensor((rng.standard_normal((w, cin, cout)) / np.sqrt(w * cin)).astype(np.float32)).to(h.h16)
    b = rng.standard_normal(cout).astype(np.float32) * 0.1
    xb = torch.zeros(N * S + 2 * G, cip, dtype=h.h16, device=dev)
    xb[G:G + N * S].view(N, S, cip)[:, :L, :cin] = x16.to(dev)
    Wp = torch.zeros(w, cip, cop, dtype=h.h16, device=dev)
    Wp[:, :cin, :cout] = W16.to(dev)
    bp = torch.zeros(cop, dtype=torch.float32, device=dev)
    bp[:cout] = torch.tensor(b)
    yb = torch.full((N * S + 2 * G, cop), 3.0, dtype=h.h16, device=dev)
    yb[:G] = 0; yb[G + N * S:] = 0
    h.gemm(fl.window(xb, N, cip, w), Wp.view(w * cip, cop), N * S, cop, w * cip, b_mn=True, bias=bp,
           act=O.ACT_RELU, out16=yb[G:])
    h.conv_mask_rows(yb[G:], N, S, L, cop)
    torch.cuda.synchronize()
    xo, Wo = x16.double().numpy(), W16.double().numpy()[None]
    y_ref, cache = O.conv1d_same_fwd(xo, Wo, b.astype(np.float64))
    y = yb[G:G + N * S].view(N, S, cop).float().cpu().numpy()
    assert rel(y[:, :L, :cout], y_ref) < tol(h, 5e-4, 4e-3)
    assert not y[:, L:].any() and not y[:, :, cout:].any()          # SAME padding rows and padded channels stay zero
    assert not yb[:G].any() and not yb[G + N * S:].any()            # guard rows untouched
    # backward: dY = pre-activation gradient (already masked by relu'), zero in the padding rows
    dy16 = torch.tensor((rng.standard_normal((N, L, cout)) * (y_ref > 0)).astype(np.float32)).to(h.h16)
    dyb = torch.zeros(N * S + 2 * G, cop, dtype=h.h16, device=dev)
    dyb[G:G + N * S].view(N, S, cop)[:, :L, :cout] = dy16.to(dev)
    dW = torch.zeros(w * cip, cop, dtype=torch.float32, device=dev)
    h.gemm(fl.window(xb, N, cip, w), dyb[G:G + N * S], w * cip, cop, N * S, a_mn=True, b_mn=True, beta=1.0, out32=dW)
    db = torch.zeros(cop, dtype=torch.float32, device=dev)
    h.colsum16(dyb[G:G + N * S], N * S, cop, db)
    Wf = torch.zeros(w * cop, cip, dtype=h.h16, device=dev)
    h.conv_w_flip(Wp, w, cip, cop, Wf)
    dxb = torch.zeros(N * S + 2 * G, cip, dtype=h.h16, device=dev)
    h.gemm(fl.window(dyb, N, cop, w), Wf, N * S, cip, w * cop, b_mn=True, out16=dxb[G:])
    torch.cuda.synchronize()
    # oracle backward with act = none on the already-masked dy (cache's u only matters through act_bwd)
    xp, Wc, u, _ = cache
    dx_ref, dW_ref, db_ref = O.conv1d_same_bwd(dy16.double().numpy(), (xp, Wc, u, O.ACT_NONE))
    assert rel(dW.view(w, cip, cop)[:, :cin, :cout].cpu().numpy(), dW_ref[0]) < 1e-5
    assert not dW.view(w, cip, cop)[:, cin:].any() and not dW.view(w, cip, cop)[:, :, cout:].any()
    assert rel(db[:cout].cpu().numpy(), db_ref) < 1e-5
    dx = dxb[G:G + N * S].view(N, S, cip).float().cpu().numpy()
    assert rel(dx[:, :L, :cin], dx_ref) < tol(h, 5e-4, 4e-3)


def test_conv_stage_frames(h):
    """(B, T, 257) fp32 frames -> channels-last padded rows, time-major frame order, CMVN fused (models/rced.py:46-57)."""
    dev, rng = h.device, np.random.default_rng(3)
    B, T, L, S, Cp = 3, 4, 257, 264, 8
    x = rng.standard_normal((B, T, L)).astype(np.float32)
    mean, std = rng.standard_normal(L).astype(np.float32), (rng.random(L) + 0.5).astype(np.float32)
    out = torch.full((B * T * S, Cp), 5.0, dtype=h.h16, device=dev)
    t = lambda a: torch.tensor(a, device=dev)
    h.conv_stage_frames(t(x), B, T, L, S, Cp, out, mean=t(mean), istd=t(1.0 / std))
    torch.cuda.synchronize()
    o = out.view(T, B, S, Cp).float().cpu().numpy()
    ref = ((x - mean) * (1.0 / std)).transpose(1, 0, 2)
    assert rel(o[:, :, :L, 0], ref) < tol(h, 5e-4, 4e-3)
    assert not o[:, :, L:].any() and not o[:, :, :, 1:].any()


@pytest.mark.parametrize("rows,K", [(12800, 1024), (37, 64), (1000, 264)])
def test_fc1_forward_and_data_gradient(h, rows, K):
    """One-output fully_connected (discriminator heads): dot product per row and the relu'-masked outer product, against
    the exact product of the same 16-bit operands."""
    dev, rng = h.device, np.random.default_rng(rows + K)
    x16 = torch.tensor(rng.standard_normal((rows, K)).astype(np.float32), device=dev).to(h.h16)
    w16 = torch.zeros(K, 8, dtype=h.h16, device=dev)
    w16[:, 0] = torch.tensor((rng.standard_normal(K) * 0.05).astype(np.float32), device=dev).to(h.h16)
    bias = torch.tensor([0.25] + [0.0] * 7, device=dev)
    out = torch.full((rows, 8), 9.0, device=dev)
    h.fc1_fwd(x16, rows, K, w16, bias, out)
    torch.cuda.synchronize()
    ref = x16.double().cpu().numpy() @ w16[:, 0].double().cpu().numpy() + 0.25
    assert rel(out[:, 0].cpu().numpy(), ref) < 1e-5
    assert bool((out[:, 1:] == 9.0).all())
    y16 = torch.tensor(np.maximum(rng.standard_normal((rows, K)), 0).astype(np.float32), device=dev).to(h.h16)
    dy16 = torch.zeros(rows, 8, dtype=h.h16, device=dev)
    dy16[:, 0] = torch.tensor(rng.standard_normal(rows).astype(np.float32), device=dev).to(h.h16)
    for src, act in ((y16, O.ACT_RELU), (y16, O.ACT_LRELU), (None, O.ACT_NONE)):
        dx = torch.zeros(rows, K, dtype=h.h16, device=dev)
        h.fc1_bwd_dx(dy16, rows, K, w16, dx, dact_src=src, dact=act)
        torch.cuda.synchronize()
        r = np.outer(dy16[:, 0].double().cpu().numpy(), w16[:, 0].double().cpu().numpy())
        if src is not None:
            r = r * np.where(y16.double().cpu().numpy() > 0, 1.0, 0.3 if act == O.ACT_LRELU else 0.0)
        assert rel(dx.float().cpu().numpy(), r) < tol(h, 5e-4, 4e-3)


@pytest.mark.parametrize("rows,K,which,clip", [(12800, 1024, 1, True), (12800, 1024, 0, True), (333, 64, 1, False), (1000, 264, 0, True)])
def test_fc1_head_equals_the_three_kernels_it_replaces(h, rows, K, which, clip):
    """rsr_fc1_head == rsr_fc1_fwd -> rsr_lsgan_mse_losses (logit terms) -> rsr_fc1_bwd_dx: same logits, same 16-bit
    d loss / d logit, same head data gradient (bit for bit), same loss sums; and the losses against the float64 formula."""
    from rsrgan_b200 import ops
    dev, rng = h.device, np.random.default_rng(rows + K + which)
    x16 = torch.tensor(np.maximum(rng.standard_normal((rows, K)), 0).astype(np.float32), device=dev).to(h.h16)
    w16 = torch.zeros(K, 8, dtype=h.h16, device=dev)
    w16[:, 0] = torch.tensor((rng.standard_normal(K) * 0.08).astype(np.float32), device=dev).to(h.h16)
    bias = torch.tensor([0.4] + [0.0] * 7, device=dev)
    d_real, d_fake, gs = 1.0, 0.0, 4096.0
    target = d_real if which == 0 else d_fake
    # the three kernels
    lg = torch.zeros(rows, 8, device=dev)
    g16 = torch.zeros(rows, 8, dtype=h.h16, device=dev)
    dx = torch.zeros(rows, K, dtype=h.h16, device=dev)
    losses = torch.zeros(8, device=dev)
    h.fc1_fwd(x16, rows, K, w16, bias, lg)
    kw = dict(ld_logit=8, n_logit=rows, clip=clip, d_real=d_real, d_fake=d_fake, gscale=gs, ld_grad=8)
    if which == 0:
        h.lsgan_mse_losses(losses, rl=lg, d_rl_grad=g16, **kw)
    else:
        h.lsgan_mse_losses(losses, fk=lg, d_fk_grad=g16, **kw)
    h.fc1_bwd_dx(g16, rows, K, w16, dx, dact_src=x16, dact=ops.ACT_RELU)
    # the fused head
    lg2 = torch.zeros(rows, 8, device=dev)
    g2 = torch.zeros(rows, 8, dtype=h.h16, device=dev)
    dx2 = torch.zeros(rows, K, dtype=h.h16, device=dev)
    losses2 = torch.zeros(8, device=dev)
    h.fc1_head(x16, rows, K, w16, bias, which, clip, d_real, d_fake, target, gs, losses2, lg2, dlogit16=g2, dact=ops.ACT_RELU,
               dx16=dx2)
    torch.cuda.synchronize()
    assert rel(lg2[:, 0].cpu().numpy(), lg[:, 0].cpu().numpy()) < 1e-6
    # (the two dot products sum in a different order: a logit may differ in its last bits, and with it -- rarely -- the
    #  rounding of its 16-bit gradient)
    gd = (g2[:, 0].float() - g16[:, 0].float()).abs().cpu().numpy()
    assert rel(g2[:, 0].float().cpu().numpy(), g16[:, 0].float().cpu().numpy()) < 1e-3 and (gd > 0).mean() < 0.02
    assert rel(dx2.float().cpu().numpy(), dx.float().cpu().numpy()) < 2e-3
    u = x16.double().cpu().numpy() @ w16[:, 0].double().cpu().numpy() + 0.4
    l = np.clip(u, -0.5, 1.5) if clip else u
    ref = [((l - d_real) ** 2).mean(), 0.0, 0.0] if which == 0 else [0.0, ((l - d_fake) ** 2).mean(), ((l - d_real) ** 2).mean()]
    got, got3 = losses2.cpu().numpy(), losses.cpu().numpy()
    for k in range(3):
        assert got[k] == pytest.approx(ref[k], rel=1e-4, abs=1e-7) and got[k] == pytest.approx(got3[k], rel=1e-5, abs=1e-7)


def test_ark_decompress_bit_exact(h):
    """rsr_ark_decompress == the reference's compressed-matrix reader, bit for bit: (i) the `CM` entry of the golden
    archive, whose expected float64 matrix was produced by the REFERENCE's io_funcs/kaldi_io.py (tests/golden/
    make_kaldi_golden.py); (ii) random headers / bytes at ragged and utterance sizes against our host reader (itself
    pinned to the reference on that fixture), with and without the float64 CMVN of make_tfrecords.py:84-87."""
    import io
    import os
    from rsrgan_b200 import kaldi_io
    gold = os.path.join(os.path.dirname(__file__), "golden")
    dev = h.device
    exp = np.load(os.path.join(gold, "kaldi_small_expected.npz"))["utt_cm"]
    r = kaldi_io.ArkReader()
    cwd = os.getcwd()
    os.chdir(gold)
    try:
        r(os.path.join(gold, "kaldi_small.scp"))
        path, off = r.scp_data[r.utt_ids.index("utt_cm")]
        mn, rg, rows, cols, hdr, data = kaldi_io.ArkReader.read_compressed_raw(path, off)
        got32 = r.read_ark_device(h, path, off).cpu().numpy()
    finally:
        os.chdir(cwd)
    out64 = torch.zeros(rows, cols, dtype=torch.float64, device=dev)
    h.ark_decompress(torch.from_numpy(hdr.view(np.int16).copy()).to(dev), torch.from_numpy(data.copy()).to(dev), mn, rg,
                     rows, cols, out64=out64)
    assert np.array_equal(out64.cpu().numpy(), exp)                      # float64, identical bits
    assert np.array_equal(got32, exp.astype(np.float32))
    rng = np.random.default_rng(0)
    for rows, cols in ((1, 1), (9, 40), (131, 257), (1000, 257), (129, 33)):
        hdr = np.sort(rng.integers(0, 65536, (cols, 4)), axis=1).astype("<u2")
        data = rng.integers(0, 256, (cols, rows)).astype(np.uint8)
        mn, rg = np.float32(-12.5), np.float32(31.25 + rows)
        want = kaldi_io.ArkReader().read_compress(mn, rg, rows, cols, io.BytesIO(hdr.tobytes() + data.tobytes()))
        mean, std = rng.standard_normal(cols), 0.5 + rng.random(cols)
        ld = cols + 3
        o64 = torch.full((rows, ld), 7.0, dtype=torch.float64, device=dev)
        o32 = torch.full((rows, ld), 7.0, dtype=torch.float32, device=dev)
        h.ark_decompress(torch.from_numpy(hdr.view(np.int16).copy()).to(dev), torch.from_numpy(data).to(dev), mn, rg,
                         rows, cols, out64=o64, out32=o32, mean=torch.from_numpy(mean).to(dev),
                         std=torch.from_numpy(std).to(dev))
        assert np.array_equal(o64[:, :cols].cpu().numpy(), want), (rows, cols)
        assert np.array_equal(o32[:, :cols].cpu().numpy(), ((want - mean) / std).astype(np.float32)), (rows, cols)
        assert bool((o64[:, cols:] == 7.0).all()) and bool((o32[:, cols:] == 7.0).all())


def test_ark_decompress_utterance_sha256_of_reference_reader(h):
    """A 500 x 257 compressed utterance decoded on the device hashes to the SHA-256 of what the REFERENCE's reader
    returned for the same bytes (tests/golden/kaldi_cm_utt_expected.npz, computed by importing io_funcs/kaldi_io.py);
    the float64-CMVN'd float32 output equals the host expression bit for bit."""
    import hashlib
    import os
    import struct
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_kaldi_io import cm_utterance_bytes
    exp = np.load(os.path.join(os.path.dirname(__file__), "golden", "kaldi_cm_utt_expected.npz"))
    body, (rows, cols) = cm_utterance_bytes()
    mn, rg, r2, c2 = struct.unpack("<ffii", body[5:21])
    assert (r2, c2) == (rows, cols)
    hdr = np.frombuffer(body[21:21 + 8 * cols], dtype="<u2").reshape(cols, 4)
    data = np.frombuffer(body[21 + 8 * cols:], dtype=np.uint8).reshape(cols, rows)
    dev = h.device
    out64 = torch.zeros(rows, cols, dtype=torch.float64, device=dev)
    out32 = torch.zeros(rows, cols, dtype=torch.float32, device=dev)
    mean, std = np.linspace(-3, 3, cols), np.linspace(0.5, 4, cols)
    h.ark_decompress(torch.from_numpy(hdr.view(np.int16).copy()).to(dev), torch.from_numpy(data.copy()).to(dev), mn, rg,
                     rows, cols, out64=out64, out32=out32, mean=torch.from_numpy(mean).to(dev),
                     std=torch.from_numpy(std).to(dev))
    m = out64.cpu().numpy()
    assert np.array_equal(m[[0, 249, 499]], exp["rows_0_249_499"])
    assert hashlib.sha256(m.tobytes()).digest() == exp["sha256"].tobytes()
    assert np.array_equal(out32.cpu().numpy(), ((m - mean) / std).astype(np.float32))
