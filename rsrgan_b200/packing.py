"""Layout helpers shared by the host code and the tests.

Packed gate order (see include/rsrgan_b200.h, rsr_lstmp_rec_fwd): the 4C gate
columns of TF's LSTMCell kernel/bias are [i | j | f | o] blocks of C
(models/BNLSTMCell.py:176-179).  The recurrent kernels want the four gates of
32 consecutive cells in one 128-wide block:

    packed_col(cell, gate) = (cell // 32) * 128 + gate * 32 + cell % 32

with C padded to Cp (multiple of 256) by zero columns.
"""
from __future__ import annotations

import numpy as np


def round_up(x, m):
    return (x + m - 1) // m * m


def cell_pad(C):
    return round_up(C, 256)


def packed_index(C):
    """int64 array of length 4*Cp: TF column index for each packed column, -1 for padding."""
    Cp = cell_pad(C)
    idx = np.full(4 * Cp, -1, dtype=np.int64)
    cell = np.arange(C)
    for g in range(4):
        idx[(cell // 32) * 128 + g * 32 + cell % 32] = g * C + cell
    return idx


def pack_cols(a, C):
    """a: (..., 4C) numpy in TF gate order -> (..., 4Cp) packed, zeros in padding."""
    idx = packed_index(C)
    out = np.zeros(a.shape[:-1] + (idx.shape[0],), dtype=a.dtype)
    valid = idx >= 0
    out[..., valid] = a[..., idx[valid]]
    return out


def unpack_cols(a, C):
    """inverse of pack_cols: (..., 4Cp) packed -> (..., 4C) TF order."""
    idx = packed_index(C)
    out = np.zeros(a.shape[:-1] + (4 * C,), dtype=a.dtype)
    valid = idx >= 0
    out[..., idx[valid]] = a[..., valid]
    return out


def pad_last(a, n):
    """zero-pad the last axis of a numpy array to length n."""
    if a.shape[-1] == n:
        return a
    out = np.zeros(a.shape[:-1] + (n,), dtype=a.dtype)
    out[..., :a.shape[-1]] = a
    return out


def pad_first(a, n):
    if a.shape[0] == n:
        return a
    out = np.zeros((n,) + a.shape[1:], dtype=a.dtype)
    out[:a.shape[0]] = a
    return out
