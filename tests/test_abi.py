"""The C-ABI library loads and exports every symbol include/rsrgan_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rsrgan_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\bint\s+(rsr_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    from rsrgan_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        ge.build()
    return _lib.load()


def test_header_and_binding_agree(lib):
    from rsrgan_b200 import _lib
    syms = declared_symbols()
    assert len(syms) >= 20
    assert sorted(_lib.SIGNATURES) == syms
    for s in syms:
        assert isinstance(getattr(lib, s), ctypes._CFuncPtr)


def test_version_and_argument_errors_without_gpu(lib):
    assert lib.rsr_version() >= 100
    assert lib.rsr_create(None, 0, 0) == -1                     # RSR_E_ARG: null out pointer
    h = ctypes.c_void_p()
    rc = lib.rsr_create(ctypes.byref(h), 0, 7)
    assert rc == -1                                             # bad dtype
    import torch
    if not torch.cuda.is_available():
        assert lib.rsr_create(ctypes.byref(h), 0, 0) == -3      # RSR_E_NODEV: the library has no CPU path
        from rsrgan_b200 import ops
        with pytest.raises(Exception):
            ops.Handle(0, "f16")


def test_sass_contains_blackwell_tensor_and_tma_ops():
    """cuobjdump -sass shows UTC*MMA (tcgen05.mma), LDTM (tcgen05.ld) and UTMALDG (TMA)."""
    import shutil
    import subprocess
    from rsrgan_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(exe):
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([exe, "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "LDTM" in sass and "UTMALDG" in sass
