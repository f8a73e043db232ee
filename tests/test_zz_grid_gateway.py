"""stereo_b200/matlab/sb_grid_mex.cpp, the grid-native MATLAB gateway (what dispmap_super.simultaneous_fusion calls when the
L x E arrays of trws(...) do not fit), behind the mex.h stand-in (oracle/gw_builders_driver.cpp, -DSB_GW_GRID).

The argument checks and the error path run on the CPU.  The solve itself was written after this round's GPU budget was
spent: it has NOT been executed on a device yet, so the GPU case is a non-strict xfail (a pass shows up as XPASS and
cannot break the suite); it is the last file of the collection for the same reason."""
import numpy as np
import pytest

from oracle import oracle
from stereo_b200 import synth


def _problem():
    return synth.trws_problem(12, 17, 15, seed=3, kernel=1)


def test_grid_gateway_argument_checks():
    pr = _problem()
    with pytest.raises(RuntimeError, match="Unsupported kernel"):                    # trws_mex.cpp:162
        oracle.grid_gateway_solve(3, 12, 17, pr["planes"], pr["unary"], pr["alphas"], pr["tol"])


def test_grid_gateway_without_device_reports_no_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    pr = _problem()
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        oracle.grid_gateway_solve(1, 12, 17, pr["planes"], pr["unary"], pr["alphas"], pr["tol"], maxiter=3)


@pytest.mark.gpu
@pytest.mark.xfail(strict=False, reason="sb_grid_mex.cpp was written after the round's GPU budget was spent: first execution on a device")
def test_grid_gateway_matches_grid_entry():
    from stereo_b200.gridsolver import trws_grid
    pr = _problem()
    lab, e, lb, it = oracle.grid_gateway_solve(1, 12, 17, pr["planes"], pr["unary"], pr["alphas"], pr["tol"], maxiter=6)
    sol, e2, lb2, it2 = trws_grid(1, pr["unary"], pr["planes"], pr["alphas"], pr["tol"], 12, 17, dict(maxiter=6, max_relgap=0.0))
    assert it == it2 == 6
    assert e == e2 and lb == lb2 and np.array_equal(lab, sol)
