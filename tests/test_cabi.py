"""Host-side checks that need no GPU: the C-ABI library loads, exports every symbol the header
declares, and the Python mirror fails loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest
import torch

import fthmc_b200
from fthmc_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "fthmc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(fthmc_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    h = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 19
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/fthmc_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature in fthmc_b200/_lib.py"
    assert set(_lib.SIGNATURES) == set(syms)
    assert fthmc_b200.lib().fthmc_version() >= 100


def test_api_mirrors_reference_names():
    for name in ["action", "force", "leapfrog", "hmc", "topocharge", "topo_charge", "regularize", "ft_flow",
                 "ft_flow_inv", "ft_action", "ft_force", "ft_leapfrog", "ft_hmc", "Param"]:
        assert callable(getattr(fthmc_b200, name))
    p = fthmc_b200.Param(beta=4.0, lat=(32, 32), tau=1.0, nstep=10)
    assert p.dt == 0.1 and p.volume == 1024 and p.nd == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="no CPU path"):
        fthmc_b200.action(fthmc_b200.Param(beta=2.0, lat=(8, 8)), torch.zeros(2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU path"):
        fthmc_b200.ft_force(fthmc_b200.Param(beta=2.0, lat=(8, 8)), None, torch.zeros(1, 2, 8, 8))


def test_argument_errors_without_gpu():
    L = fthmc_b200.lib()
    # null pointers / bad sizes are rejected before any CUDA call
    assert L.fthmc_action(None, 1, 8, 8, 1.0, 0, None, 0, None) == -1
    assert b"null" in L.fthmc_last_error_string()
    h = ctypes.c_void_p()
    import numpy as np
    raw = np.zeros((2, 955)); mu = np.zeros(2, dtype=np.int32); off = np.zeros(2, dtype=np.int32)
    rc = L.fthmc_flow_pack(raw.ctypes.data, 2, mu.ctypes.data, off.ctypes.data, 16, 8, 2, 3, 0, 0, 1e-6, 1000, ctypes.byref(h))
    assert rc == -5
    rc = L.fthmc_flow_pack(raw.ctypes.data, 2, mu.ctypes.data, off.ctypes.data, 8, 8, 2, 3, 7, 0, 1e-6, 1000, ctypes.byref(h))
    assert rc == -1
