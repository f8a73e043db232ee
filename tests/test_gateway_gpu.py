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


def test_builders_gateway_ops():
    """stereo_b200/matlab/sb_builders_mex.cpp through the mex.h stand-in (oracle/gw_builders_driver.cpp): the operations
    the dispmap_* classes call, with MATLAB-shaped arguments, against the same entry points called directly."""
    import stereo_b200 as sb
    from stereo_b200 import builders
    from oracle import oracle
    H, W = 20, 26
    N = H * W
    E = 2 * ((H - 1) * W + H * (W - 1))
    rng = np.random.default_rng(3)
    props = synth.random_plane_proposals(H, W, 3, rng)
    w = rng.random(E) * 5
    un = rng.random(N)
    sz = np.array([[float(H), float(W)]])
    # energy / pairwise tables (dispmap_super.update_energy, all_pairwise_costs)
    e, = oracle.builders_gateway("energy", sz, 1.0, un.reshape(N, 1), props[0], w.reshape(1, E), 0.5, 0.0, 1.0)
    assert e.shape == (1, 1) and e[0, 0] == builders.energy(H, W, 1, un, props[0], w, 0.5, 0.0, 1.0)
    t = oracle.builders_gateway("pairwise_tables", sz, 2.0, props[0], props[1], w.reshape(1, E), 0.3, 0.0, 1.0, nlhs=4)
    for a, b in zip(t, builders.pairwise_tables(H, W, 2, props[0], props[1], w, 0.3, 0.0, 1.0)):
        assert a.shape == (1, E) and np.array_equal(a.reshape(-1), b)
    with pytest.raises(RuntimeError):           # a proposal is given: four outputs or none at all
        oracle.builders_gateway("pairwise_tables", sz, 2.0, props[0], props[1], w.reshape(1, E), 0.3, 0.0, 1.0, nlhs=2)
    with pytest.raises(RuntimeError):
        oracle.builders_gateway("no_such_operation", sz)
    # smoothness weights from a segment image
    seg = rng.integers(1, 4, size=(H, W)).astype(np.uint32)
    sw, = oracle.builders_gateway("smooth_weights", seg, 5.0, 0.5, 2.0)
    assert sw.shape == (1, E) and np.array_equal(sw.reshape(-1), builders.smooth_weights(seg, 5.0, 0.5, 2.0))
    # plane proposal around a point
    cc, rr = np.meshgrid(np.arange(1, W + 1), np.arange(1, H + 1))
    disp = 3.0 + 0.1 * cc - 0.05 * rr + rng.normal(0, 0.2, size=(H, W))
    p, prop = oracle.builders_gateway("plane_from_disparity", disp, 12.0, 9.0, 5.0, 2.0, nlhs=2)
    ref_p, ref_prop = builders.plane_from_disparity(disp, 12.0, 9.0, 5.0, 2, return_proposal=True)
    assert p.shape == (4, 1) and prop.shape == (4, N)
    assert np.array_equal(p.reshape(-1), ref_p) and np.array_equal(prop, ref_prop)
    # the segpln window-matching volume
    im0, im1, _ = synth.stereo_pair(H, W, 4, seed=2)
    P = np.zeros((3, 4, 2))
    P[:, :3, 0] = np.eye(3)
    P[:, :3, 1] = np.eye(3)
    P[0, 3, 1] = -0.25
    disps = np.arange(16.0, -1.0, -1.0)
    corr, score = oracle.builders_gateway("segpln_wta", np.stack([im0, im1], axis=3), P, disps.reshape(1, -1), 1.0, 30.0, 0.07, nlhs=2)
    rc, rs = builders.segpln_wta([im0, im1], P, disps, 1, 30.0, return_score=True)
    assert np.array_equal(corr, rc) and np.array_equal(score, rs)
    # the fusion scheduler
    unaries = rng.random((3, N))
    cur = props[2]
    ucur = rng.random(N)
    ids = np.array([1, 2, 3, 1, 3, 2, 1, 2, 3], dtype=np.int32)
    fused, Es, ufused = oracle.builders_gateway("fuse_until_convergence", sz, 1.0, np.stack(list(props), axis=2), unaries.T.copy(), cur,
                                                ucur.reshape(N, 1), w.reshape(1, E), 0.5, 0.0, 1.0, 0.0, 6.0, ids.reshape(1, -1), nlhs=3)
    a, u, Eref, _ = sb.binary_fuse_until_convergence_grid(H, W, 1, list(props), unaries, cur, ucur, w, 0.5, 6, ids)
    assert Es.shape == (1, Eref.size) and np.array_equal(Es.reshape(-1), Eref)
    assert np.array_equal(fused, a) and np.array_equal(ufused.reshape(-1), u)
    assert all(Eref[i + 1] <= Eref[i] * (1 + 1e-12) for i in range(Eref.size - 1))
