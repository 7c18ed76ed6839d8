"""The strided convolution family of utils/ops.py (downconv, conv1d, deconv) and virtual batch norm (utils/bnorm.py) on
the GPU (rsrgan_b200/conv_ops.py: rsr_gemm over strided overlapped views + the batch-norm stream kernels) against the
float64 oracle, which tests/test_oracle.py pins to torch.nn.functional.conv1d / conv_transpose1d and autograd.
Tolerance: 16-bit operands, fp32 accumulate -> relative RMS 3e-3 (fp16) / 2e-2 (bf16) per tensor."""
import numpy as np
import pytest
import torch

from oracle import rsr_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["f16", "bf16"])
def h(request):
    from rsrgan_b200 import ops
    return ops.Handle(0, request.param)


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).mean()) / (np.sqrt((b ** 2).mean()) + 1e-30))


def tol(h):
    return 3e-3 if h.h16 == torch.float16 else 2e-2


def r16(h, a):
    """the 16-bit rounding the device applies to operands"""
    return torch.tensor(np.asarray(a, np.float32)).to(h.h16).double().numpy()


@pytest.mark.parametrize("B,L,ci,co,k", [(3, 64, 3, 5, 31), (2, 128, 16, 32, 31), (4, 32, 8, 8, 5), (1, 256, 1, 16, 31)])
def test_downconv_forward_and_gradients(h, B, L, ci, co, k):
    from rsrgan_b200.conv_ops import DownConv1d, Seq
    rng = np.random.default_rng(L + ci)
    x, W, b = rng.standard_normal((B, L, ci)), rng.standard_normal((k, ci, co)) * 0.2, rng.standard_normal(co)
    dy = rng.standard_normal((B, L // 2, co))
    xq, Wq, dyq = r16(h, x), r16(h, W), r16(h, dy)
    y_ref, cache = O.downconv_fwd(xq, Wq, b)
    dx_ref, dW_ref, db_ref = O.downconv_bwd(dyq, cache)
    seq = Seq(B, L, L + 32)
    lay = DownConv1d(h, ci, co, k)
    lay.load(W, b)
    xw = seq.stage(h, x, lay.ap)
    yw, z32 = lay.fwd(seq, xw, want32=True)
    lo = seq.half()
    y = lo.unstage(yw[lo.GUARD:lo.GUARD + lo.rows], co).cpu().numpy()
    assert rel(y, y_ref) < tol(h)
    assert rel(z32.view(B, lo.S, -1)[:, :lo.L, :co].cpu().numpy(), y_ref) < 1e-5 + (0 if h.h16 == torch.float16 else 1e-4)
    body = yw[lo.GUARD:lo.GUARD + lo.rows].view(B, lo.S, -1)
    assert float(body[:, lo.L:].abs().max()) == 0.0                          # the SAME padding rows stay zero
    assert lay.bp == co or float(body[:, :, co:].abs().max()) == 0.0         # ... and the padded channels
    dyw = lo.stage(h, dy, lay.bp)
    dxw = lay.bwd(seq, xw, dyw)
    torch.cuda.synchronize()
    assert rel(seq.unstage(dxw[seq.GUARD:seq.GUARD + seq.rows], ci).cpu().numpy(), dx_ref) < tol(h)
    gw = lay.gw.view(k, lay.ap, lay.bp)[:, :ci, :co].cpu().numpy()
    assert rel(gw, dW_ref) < 1e-4 and rel(lay.gb[:co].cpu().numpy(), db_ref) < 1e-4
    assert lay.ap == ci or float(lay.gw.view(k, lay.ap, lay.bp)[:, ci:].abs().max()) == 0.0


@pytest.mark.parametrize("B,L,ci,co,k", [(3, 32, 5, 3, 31), (2, 64, 32, 16, 31), (4, 16, 8, 8, 5)])
def test_deconv_forward_and_gradients(h, B, L, ci, co, k):
    from rsrgan_b200.conv_ops import Deconv1d, Seq
    rng = np.random.default_rng(L + co)
    x, W, b = rng.standard_normal((B, L, ci)), rng.standard_normal((k, co, ci)) * 0.2, rng.standard_normal(co)
    dy = rng.standard_normal((B, 2 * L, co))
    xq, Wq, dyq = r16(h, x), r16(h, W), r16(h, dy)
    y_ref, cache = O.deconv_fwd(xq, Wq, b)
    dx_ref, dW_ref, db_ref = O.deconv_bwd(dyq, cache)
    seq = Seq(B, L, L + 16)
    lay = Deconv1d(h, ci, co, k)
    lay.load(W, b)
    xw = seq.stage(h, x, lay.bp)
    yw = lay.fwd(seq, xw)
    hi = seq.double()
    assert rel(hi.unstage(yw[hi.GUARD:hi.GUARD + hi.rows], co).cpu().numpy(), y_ref) < tol(h)
    dyw = hi.stage(h, dy, lay.ap)
    dxw = lay.bwd(seq, xw, dyw)
    torch.cuda.synchronize()
    assert rel(seq.unstage(dxw[seq.GUARD:seq.GUARD + seq.rows], ci).cpu().numpy(), dx_ref) < tol(h)
    assert rel(lay.gw.view(k, lay.ap, lay.bp)[:, :co, :ci].cpu().numpy(), dW_ref) < 1e-4
    assert rel(lay.gb[:co].cpu().numpy(), db_ref) < 1e-4


def test_conv1d_k31_and_virtual_batch_norm(h):
    """One discriminator block of models/discriminator.py:38-66 -- downconv(31, stride 2) -> VBN -> leaky ReLU -- on a
    reference batch and on a live batch, then the logits_conv (conv1d, kwidth 31, one kernel): forward and the gradients
    through everything, against the oracle."""
    from rsrgan_b200.conv_ops import Conv1d, DownConv1d, Seq, VBN
    from rsrgan_b200.ops import ACT_LRELU, ACT_NONE
    rng = np.random.default_rng(3)
    B, L, ci, co, k = 4, 128, 1, 16, 31
    W, Wl = rng.standard_normal((k, ci, co)) * 0.3, rng.standard_normal((k, co, 1)) * 0.2
    gamma, beta = 1.0 + 0.1 * rng.standard_normal(co), 0.1 * rng.standard_normal(co)
    x_ref, x_live = rng.standard_normal((B, L, ci)), rng.standard_normal((B, L, ci)) * 1.5 + 0.3
    seq = Seq(B, L, L + 32)
    lo = seq.half()
    down, vbn, logit = DownConv1d(h, ci, co, k), VBN(h, co), Conv1d(h, co, 1, k)
    down.load(W)
    logit.load(Wl)
    vbn.gamma[:co] = torch.tensor(gamma, dtype=torch.float32)
    vbn.beta[:co] = torch.tensor(beta, dtype=torch.float32)
    Wq, Wlq = r16(h, W), r16(h, Wl)
    ref_stats = None
    for x in (x_ref, x_live):
        xw = seq.stage(h, x, down.ap)
        _, z32 = down.fwd(seq, xw, want32=True, bias=False)
        aw = vbn.fwd(lo, z32, act=ACT_LRELU)
        lw, l32 = logit.fwd(lo, aw, want32=True)
        z_o, c_down = O.downconv_fwd(r16(h, x), Wq)
        y_o, c_vbn = O.vbn_fwd(z_o, gamma, beta, ref=ref_stats)
        if ref_stats is None:
            ref_stats = O.vbn_reference(z_o)
        a_o = np.maximum(y_o, 0.3 * y_o)
        l_o, c_log = O.conv1d_same_fwd(r16(h, a_o)[:, :, :], Wlq[None], np.zeros(1), act=O.ACT_NONE)
        got_a = lo.unstage(aw[lo.GUARD:lo.GUARD + lo.rows], co).cpu().numpy()
        assert rel(got_a, a_o) < tol(h)
        got_l = l32.view(B, lo.S, -1)[:, :lo.L, :1].cpu().numpy()
        assert rel(got_l, l_o) < tol(h)
    # gradients of the live pass: d(sum logits * g) through conv1d, leaky ReLU + VBN (batch weight 1/(B+1)), downconv
    g = rng.standard_normal((B, lo.L, 1))
    dlw = lo.stage(h, g, logit.bp)
    daw = logit.bwd(lo, aw, dlw)
    dzw = vbn.bwd(lo, daw)
    down.bwd(seq, xw, dzw, want_dx=False)
    torch.cuda.synchronize()
    da_o, dWl_o, _ = O.conv1d_same_bwd(r16(h, g), c_log)
    dy_o = np.where(y_o > 0, 1.0, 0.3) * da_o
    dz_o, dgamma_o, dbeta_o = O.vbn_bwd(dy_o, c_vbn)
    _, dW_o, _ = O.downconv_bwd(dz_o, c_down)
    t = 3 * tol(h)                                       # three 16-bit roundings in a row
    assert rel(logit.gw.view(k, logit.ap, logit.bp)[:, :co, :1].cpu().numpy(), dWl_o[0]) < t
    assert rel(vbn.ggamma[:co].cpu().numpy(), dgamma_o) < t and rel(vbn.gbeta[:co].cpu().numpy(), dbeta_o) < t
    assert rel(down.gw.view(k, down.ap, down.bp)[:, :ci, :co].cpu().numpy(), dW_o) < t
