"""CPU tests of the C-ABI boundary: the shared library loads, exports every symbol that
include/stereo_b200.h declares, validates arguments like the reference gateway does, and
fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

import stereo_b200 as sb
from stereo_b200 import _lib, synth
from stereo_b200._lib import SB_EINVAL, SB_ENODEV, SB_ENOTGRID, SB_EUNSUP, SbError

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for fn in os.listdir(os.path.join(ROOT, "include")):
        if fn.endswith(".h"):
            src = open(os.path.join(ROOT, "include", fn)).read()
            names |= set(re.findall(r"SB_API\s+[\w\s\*]+?\b(sb_\w+)\s*\(", src))
    return sorted(names)


def test_library_loads_and_reports_version():
    L = _lib.lib()
    assert b"sm_100a" in L.sb_version()


def test_every_declared_symbol_is_exported():
    L = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 9
    for s in syms:
        assert hasattr(L, s), f"{s} declared in include/ but not exported"


def test_no_oracle_in_product():
    """The product must never route through oracle/: nothing under stereo_b200/ mentions it."""
    for dirpath, _, files in os.walk(os.path.join(ROOT, "stereo_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".inc", ".m")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower() or f in ("_lib.py",) and "oracle" not in txt, (dirpath, f)


def test_trws_argument_validation():
    pr = synth.trws_problem(5, 6, 4, seed=0)
    with pytest.raises(SbError) as ei:  # trws_mex.cpp:162 "Unsupported kernel"
        sb.trws(3, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"], {})
    assert ei.value.code == SB_EINVAL and "Unsupported kernel" in str(ei.value)
    if _lib.lib().sb_device_count() > 0:  # the NaN scan of trws.m:9-15 runs on the device
        q = pr["q"].copy()
        q[1, 2] = np.nan
        with pytest.raises(ValueError, match="q contains NaN"):
            sb.trws(1, pr["unary"], pr["connectivity"], q, pr["qprim"], pr["alphas"], pr["tol"], {})
    conn = pr["connectivity"].copy()
    conn[:, [0, 1]] = conn[:, [1, 0]]
    with pytest.raises(SbError) as ei:
        sb.trws(1, pr["unary"], conn, pr["q"], pr["qprim"], pr["alphas"], pr["tol"], {})
    assert ei.value.code == SB_ENOTGRID
    with pytest.raises(AssertionError):  # trws.m:5 connectivity must be 1-based
        sb.trws(1, pr["unary"], pr["connectivity"] - 1, pr["q"], pr["qprim"], pr["alphas"], pr["tol"], {})


def test_too_many_labels():
    H, W, L = 4, 4, 257
    i1, i2 = sb.construct_neighborhood(H, W)
    E = i1.size
    with pytest.raises(SbError) as ei:
        sb.trws(1, np.zeros((L, H * W)), np.stack([i1, i2]), np.zeros((L, E)), np.zeros((L, E)), np.ones(E), 1.0, {})
    assert ei.value.code in (SB_EUNSUP, SB_ENODEV)


def test_no_device_fails_loudly():
    """Without a GPU the compute entry points return SB_ENODEV -- never a CPU result."""
    if _lib.lib().sb_device_count() > 0:
        pytest.skip("a CUDA device is present")
    pr = synth.trws_problem(5, 6, 4, seed=0)
    with pytest.raises(SbError) as ei:
        sb.trws(1, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"], {})
    assert ei.value.code == SB_ENODEV
