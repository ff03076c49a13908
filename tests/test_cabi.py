"""The C-ABI shared library builds, loads and exports every symbol include/qinfer_b200.h declares.
No compute call is made here (no GPU on the CPU test box)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "qinfer_b200.h")


def _declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qb_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path_entry_points():
    syms = _declared_symbols()
    for needed in ("qb_fused_update", "qb_moments", "qb_cdf", "qb_draw", "qb_lw_move", "qb_lw_retry",
                   "qb_compact_invalid", "qb_tomo_canonicalize", "qb_likelihood", "qb_are_models_valid"):
        assert needed in syms


def test_library_exports_every_declared_symbol():
    from qinfer_b200 import _lib
    lib = ctypes.CDLL(_lib.library_path())
    for name in _declared_symbols():
        assert hasattr(lib, name), "libqinfer_b200.so does not export %s" % name


def test_binding_covers_every_declared_symbol_and_loads():
    from qinfer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared_symbols()
    lib = _lib.load()
    assert lib.qb_abi_version() == _lib.QB_ABI_VERSION == 4


def test_struct_layouts_match_the_header():
    from qinfer_b200 import _lib
    assert ctypes.sizeof(_lib.QbModel) == 48          # + d_extra, extra_rule, fast_math, reserved (ABI 3)
    assert ctypes.sizeof(_lib.QbExpparams) == 8 + 8 + 8 + 4 + 4 + 8 + 8 * _lib.QB_MAX_D
    assert _lib.QbExpparams.meas.offset == 40
    sizes = (ctypes.c_int32 * 3)()
    _lib.load().qb_struct_sizes(sizes)                       # what the C compiler actually laid out
    assert list(sizes) == [ctypes.sizeof(_lib.QbModel), ctypes.sizeof(_lib.QbExpparams),
                           ctypes.sizeof(_lib.QbUpdateCtl)]
    assert ctypes.sizeof(_lib.QbUpdateCtl) == 56 + 8 * _lib.QB_MAX_RANKS + 8 + 8     # + h_shard_norms (ABI 4)


def test_argument_validation_without_a_gpu():
    """Pure host-side checks of the C ABI (they return before any CUDA call)."""
    from qinfer_b200 import _lib
    lib = _lib.load()
    m = _lib.QbModel(kind=99, d=1, binomial=0, interleaved=0, min_freq=0.0)
    ep = _lib.QbExpparams()
    rc = lib.qb_fused_update(ctypes.byref(m), ctypes.byref(ep), 0, None, 10, None, None, None, None, None, None, 0, None)
    assert rc == -2 and b"no CPU fallback" in lib.qb_last_error()
    m = _lib.QbModel(kind=_lib.QB_MODEL_RB, d=2, binomial=0, interleaved=0, min_freq=0.0)
    rc = lib.qb_fused_update(ctypes.byref(m), ctypes.byref(ep), 0, None, 10, None, None, None, None, None, None, 0, None)
    assert rc == -2
    m = _lib.QbModel(kind=_lib.QB_MODEL_PRECESSION, d=1, binomial=0, interleaved=0, min_freq=0.0)
    rc = lib.qb_fused_update(ctypes.byref(m), ctypes.byref(ep), 0, None, 10, None, None, None, None, None, None, 0, None)
    assert rc == -1 and b"NULL" in lib.qb_last_error()
    assert lib.qb_update_workspace_bytes(1000, 1) > 0
    assert lib.qb_moments_workspace_bytes(1000, 16) >= 153 * 8
    assert lib.qb_cdf_workspace_bytes(10 ** 7) >= (10 ** 7 // 2048) * 8
