"""ctypes binding of the C-ABI library `librsrgan_sm100.so` (include/rsrgan_b200.h).

There is no fallback: if the shared library is missing (not built) importing the
compute entry points raises, and `Handle()` raises when no sm_100 device exists.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RSR_LIB: another build of the same library (e.g. the -DRSR_TRACE phase-timing build used by scripts/gpu_trace_rec.py)
LIB_PATH = os.environ.get("RSR_LIB") or os.path.join(_HERE, "librsrgan_sm100.so")

RSR_DTYPE_F16, RSR_DTYPE_BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_CLIP = 0, 1, 2, 3
RSR_E_RESIDENT = -4
ERRORS = {-1: "RSR_E_ARG", -2: "RSR_E_SHAPE", -3: "RSR_E_NODEV", -4: "RSR_E_RESIDENT"}

vp, ci, cf, cll = C.c_void_p, C.c_int, C.c_float, C.c_longlong


class GemmArgs(C.Structure):
    _fields_ = [("M", ci), ("N", ci), ("K", ci),
                ("A", vp), ("lda", ci), ("a_mn", ci),
                ("B", vp), ("ldb", ci), ("b_mn", ci),
                ("alpha", cf), ("beta", cf),
                ("bias", vp),
                ("resid", vp), ("ldr", ci),
                ("act", ci),
                ("dact_src", vp), ("ldd", ci), ("dact", ci),
                ("out32", vp), ("ldc32", ci),
                ("out16", vp), ("ldc16", ci),
                ("tile_n", ci), ("split_k", ci),
                ("stats", vp)]


class WaveArgs(C.Structure):
    """rsr_wave_args (include/rsrgan_b200.h)."""
    _fields_ = [("B", ci), ("T", ci), ("Cp", ci), ("I1", ci), ("P1", ci), ("forget_bias", cf),
                ("lengths", vp),
                ("x16", vp), ("ldx", ci),
                ("kxT1", vp), ("bias1", vp), ("wcT1", vp), ("w_i1", vp), ("w_f1", vp), ("w_o1", vp),
                ("mt1", vp), ("save1", vp),
                ("wpT1", vp),
                ("out1", vp), ("ldo1", ci),
                ("kxT2", vp), ("bias2", vp), ("wcT2", vp), ("w_i2", vp), ("w_f2", vp), ("w_o2", vp),
                ("mt2", vp), ("save2", vp)]


class WaveBwdArgs(C.Structure):
    """rsr_wave_bwd_args (include/rsrgan_b200.h)."""
    _fields_ = [("B", ci), ("T", ci), ("Cp", ci), ("max_nbp", ci),
                ("lengths", vp),
                ("dmt2", vp), ("wc2", vp), ("w_i2", vp), ("w_f2", vp), ("w_o2", vp), ("save2", vp),
                ("dz2", vp), ("dbias2", vp), ("dw_i2", vp), ("dw_f2", vp), ("dw_o2", vp),
                ("fT", vp), ("part", vp),
                ("wc1", vp), ("w_i1", vp), ("w_f1", vp), ("w_o1", vp), ("save1", vp),
                ("dz1", vp), ("dbias1", vp), ("dw_i1", vp), ("dw_f1", vp), ("dw_o1", vp)]


# name -> argtypes (every symbol include/rsrgan_b200.h declares)
SIGNATURES = {
    "rsr_version": [],
    "rsr_create": [C.POINTER(vp), ci, ci],
    "rsr_destroy": [vp],
    "rsr_num_sms": [vp],
    "rsr_gemm": [vp, vp, C.POINTER(GemmArgs)],
    "rsr_stage_input": [vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp, ci],
    "rsr_unstage_output": [vp, vp, vp, ci, ci, ci, ci, vp, vp, vp],
    "rsr_cmvn_apply": [vp, vp, vp, vp, vp, cll, ci, vp],
    "rsr_cmvn_invert": [vp, vp, vp, vp, vp, cll, ci, vp],
    "rsr_cmvn_apply_padded": [vp, vp, vp, vp, vp, vp, ci, ci, ci, vp],
    "rsr_lstmp_rec_fwd": [vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, cf, vp, vp, vp],
    "rsr_lstmp_fused_fwd": [vp, vp, ci, ci, ci, ci, vp, ci, vp, vp, vp, vp, vp, vp, cf, vp, vp, vp],
    "rsr_lstmp_wave_fwd": [vp, vp, C.POINTER(WaveArgs)],
    "rsr_lstmp_wave_bwd": [vp, vp, C.POINTER(WaveBwdArgs)],
    "rsr_transpose16": [vp, vp, vp, ci, ci, ci, vp, ci],
    "rsr_peer_alloc": [vp, cll, C.POINTER(vp), vp],
    "rsr_peer_open": [vp, vp, C.POINTER(vp)],
    "rsr_peer_close": [vp, vp],
    "rsr_peer_free": [vp, vp],
    "rsr_peer_error": [vp, vp, C.POINTER(ci)],
    "rsr_peer_allreduce": [vp, vp, C.POINTER(vp), ci, ci, cll, cll, ci],
    "rsr_lstmp_rec_bwd": [vp, vp, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp],
    "rsr_lsgan_mse_losses": [vp, vp, vp, vp, ci, cll, ci, vp, ci, vp, ci, cll, ci, cf, cf, cf, cf,
                             vp, vp, vp, vp, ci, vp, ci],
    "rsr_colsum16": [vp, vp, vp, ci, cll, ci, vp, ci],
    "rsr_colsum32": [vp, vp, vp, ci, cll, ci, vp, ci],
    "rsr_seg_sumsq": [vp, vp, vp, cf, vp, cll, ci, vp],
    "rsr_clip_sgd_ema": [vp, vp, vp, cf, vp, vp, ci, cf, vp, cf, cll, vp, vp, vp],
    "rsr_clip_adam_ema": [vp, vp, vp, cf, vp, vp, ci, cf, vp, cf, cll, vp, vp, vp, vp, vp],
    "rsr_l2_grad": [vp, vp, vp, vp, vp, vp, cf, cll],
    "rsr_add_cast": [vp, vp, vp, vp, cll, vp, vp],
    "rsr_cast16": [vp, vp, vp, cll, vp],
    "rsr_fill32": [vp, vp, vp, cll, cf],
    "rsr_fc1_fwd": [vp, vp, vp, ci, cll, ci, vp, ci, vp, vp, ci],
    "rsr_fc1_bwd_dx": [vp, vp, vp, ci, cll, ci, vp, ci, vp, ci, ci, vp, ci],
    "rsr_fc1_head": [vp, vp, vp, ci, cll, ci, vp, ci, vp, ci, ci, cf, cf, cf, cf, vp, vp, ci, vp, ci, ci, vp, ci],
    "rsr_conv_stage_frames": [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp],
    "rsr_conv_mask_rows": [vp, vp, vp, cll, ci, ci, ci],
    "rsr_conv_w_flip": [vp, vp, vp, ci, ci, ci, vp],
    "rsr_conv_w_phase": [vp, vp, vp, ci, ci, ci, ci, ci, vp],
    "rsr_conv_toeplitz_expand": [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp],
    "rsr_conv_toeplitz_fold": [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp],
    "rsr_vec_tile": [vp, vp, vp, ci, ci, ci, vp],
    "rsr_vec_fold": [vp, vp, vp, ci, ci, vp],
    "rsr_conv_stage_lines": [vp, vp, vp, ci, ci, ci, ci, ci, ci, ci, ci, vp, vp, vp],
    "rsr_vbn_stats": [vp, vp, vp, ci, cll, ci, vp, vp, cf, cf, vp, vp, vp, vp],
    "rsr_vbn_bwd": [vp, vp, vp, ci, vp, ci, cll, ci, ci, cf, vp, vp, vp, vp, ci, vp, ci, vp],
    "rsr_bn_train_stats": [vp, vp, vp, ci, cll, ci, vp, vp, cf, vp, cf, cf, ci, vp, vp],
    "rsr_bn_train_finish": [vp, vp, ci, cll, ci, vp, vp, cf, vp, cf, cf, ci, vp, vp],
    "rsr_bn_eval_coef": [vp, vp, ci, vp, vp, cf, vp, vp],
    "rsr_bn_train_stats_lines": [vp, vp, vp, ci, cll, ci, ci, ci, ci, ci, vp, vp, cf, vp, ci, cf, cf, ci, vp, vp],
    "rsr_bn_eval_coef_lines": [vp, vp, ci, ci, ci, vp, vp, cf, vp, ci, vp],
    "rsr_affine_act_lines": [vp, vp, vp, ci, cll, ci, ci, ci, vp, vp, ci, vp, ci],
    "rsr_bn_bwd_lines": [vp, vp, vp, ci, vp, ci, cll, ci, ci, ci, ci, ci, ci, vp, vp, vp, vp, ci, vp],
    "rsr_affine_act_drop": [vp, vp, vp, ci, cll, ci, vp, vp, ci, cf, vp, C.c_uint, vp, ci, vp, ci],
    "rsr_bn_bwd": [vp, vp, vp, ci, vp, ci, cll, ci, ci, cf, vp, C.c_uint, ci, vp, vp, vp, vp, vp, ci, vp, ci, vp],
    "rsr_rng_tick": [vp, vp, vp],
    "rsr_gauss_noise": [vp, vp, vp, C.c_uint, vp, cll, cf],
    "rsr_ark_decompress": [vp, vp, vp, vp, cf, cf, ci, ci, vp, ci, vp, vp, vp, ci],
    "rsr_crc32c_host": [vp, C.c_ulonglong, C.c_uint],
}

RESTYPES = {"rsr_crc32c_host": C.c_uint}       # everything else returns an int status

_lib = None


def load():
    """Loads the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "rsrgan_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = RESTYPES.get(name, ci)
        _lib = lib
    return _lib


class RsrError(RuntimeError):
    pass


def check(rc, what):
    if rc != 0:
        if rc < 0:
            raise RsrError("%s failed: %s" % (what, ERRORS.get(rc, rc)))
        raise RsrError("%s failed: cudaError %d" % (what, rc))
