"""The product called the way MATLAB would call it: stereo_b200/matlab/{trws,rd}_mex.cpp compiled
against the mex.h stand-in, fed fabricated mxArrays by the same driver that feeds the reference
gateways (oracle/ref_driver.cpp).  Same inputs, same ABI, results compared with the reference."""
import ctypes

import numpy as np
import pytest

from stereo_b200 import synth
from util import golden

pytestmark = pytest.mark.gpu
libc = ctypes.CDLL(None)


def test_trws_mex_gateway_matches_golden():
    from oracle import oracle
    T = golden("trws_solve.npz")
    for i in (1, 3, 6, 8):
        H, W, L, k, seed, it, gap = T[f"case{i}_params"]
        pr = synth.trws_problem(int(H), int(W), int(L), seed=int(seed), kernel=int(k))
        lab, e, lb, n = oracle.trws_solve(int(k), pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T,
                                          pr["alphas"], pr["tol"], int(it), float(gap), kind="gateway")
        ge, glb, gn = T[f"case{i}_scalars"]
        assert abs(e - ge) <= 1e-4 * abs(ge) and abs(lb - glb) <= 1e-4 * abs(glb)
        if abs(ge - glb) > 1e-9 * abs(ge):   # the count is rounding-dependent once the gap has closed
            assert n == gn
        assert np.mean(lab == T[f"case{i}_labels"]) >= 0.995


def test_rd_mex_gateway_matches_golden():
    from oracle import oracle
    R = golden("rd_solve.npz")
    for i in (2, 4, 6, 12):
        H, W, seed, kernel, mode, improve = R[f"case{i}_params"]
        pr = synth.rd_problem(int(H), int(W), seed=int(seed), kernel=int(kernel), mode="frustrated" if mode else "stereo")
        libc.srand(1)
        lab, e, lb, nu = oracle.rd_solve(pr["U0"], pr["U1"], pr["E00"], pr["E01"], pr["E10"], pr["E11"],
                                         (pr["connectivity"] - 1).T, improve=bool(improve), kind="gateway")
        assert np.array_equal(lab, R[f"case{i}_labels"].astype(np.float64))
        ge, glb, gnu = R[f"case{i}_scalars"]
        assert abs(e - ge) <= 1e-9 * abs(ge) and nu == gnu


def test_gateway_error_paths():
    """Unsupported kernel -> mexErrMsgTxt("Unsupported kernel") like trws_mex.cpp:162."""
    from oracle import oracle
    pr = synth.trws_problem(5, 6, 4, seed=0)
    with pytest.raises(RuntimeError) as ei:
        oracle.trws_solve(3, pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T, pr["alphas"],
                          pr["tol"], 5, 0.0, kind="gateway")
    assert "Unsupported kernel" in str(ei.value)
