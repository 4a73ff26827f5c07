"""CPU tier: the C-ABI shared library builds, loads and exports every symbol that include/npgp.h declares (no compute
call is made: there is no GPU here), and the product refuses to run without CUDA instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "npgp.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(npgp_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from nonstationary_precip_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "run `python -m nonstationary_precip_b200.build` (or __graft_entry__.build())"
    handle = ctypes.CDLL(_lib.LIB_PATH)
    syms = declared_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(handle, s)]
    assert not missing, missing
    # every declared symbol has a ctypes signature in the binding, and vice versa
    assert set(syms) == set(_lib.exported_symbols())
    assert handle.npgp_version() == 100


def test_no_cpu_fallback():
    from nonstationary_precip_b200 import ops
    from nonstationary_precip_b200._lib import NpgpError
    x = torch.rand(4, 2, dtype=torch.float64)
    ell = torch.ones(2, 4, dtype=torch.float64)
    with pytest.raises((NpgpError, RuntimeError)):
        ops.gibbs_diag_fwd(x, ell, x, ell)


def test_argument_errors_are_reported_without_a_gpu():
    from nonstationary_precip_b200 import _lib
    lib = _lib.lib()
    assert lib.npgp_gibbs_diag_fwd(2, -1, 4, None, None, None, None, None, None, 4, None, None, None) == -1
    assert lib.npgp_potrf_workspace_bytes(1024) > 8 * 1024 * 1024
    assert lib.npgp_set_gemm_config(99) == -1
