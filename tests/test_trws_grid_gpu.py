"""GPU parity tests of the grid-native TRW-S entry (sb_trws_grid_*) against the
oracle (the compiled reference where oracle/_ref exists) and against the MATLAB-layout entry.

The reference solver needs q / qprim.  Two ways to get them:
  * from the exact double planes (synth.trws_problem builds them like dispmap_super.m:180-183):
    the grid entry rounds own disparity and slopes separately, so positions agree to rounding
    only -> fp64 instantiation 1e-9, fp32 1e-4 (BASELINE north_star);
  * from the planes AS STORED on the device (get_label -> positions_from_labels): the oracle then
    solves the bit-identical problem -> fp64 labels must be exactly equal."""
import numpy as np
import pytest

import stereo_b200 as sb
from stereo_b200 import synth
from stereo_b200.gridsolver import TrwsGrid, positions_from_labels, trws_grid
from util import trws_oracle

pytestmark = pytest.mark.gpu


def _grid_solve(pr, H, W, maxiter, relgap=0.0, **kw):
    return trws_grid(pr["kernel"], pr["unary"], pr["planes"], pr["alphas"], pr["tol"], H, W,
                     dict(maxiter=maxiter, max_relgap=relgap, **kw))


def _stored_problem(g, pr, H, W, dtype):
    """The problem the device actually holds, in trws() shapes."""
    L = pr["unary"].shape[0]
    lab = [g.get_label(l) for l in range(L)]
    unary = np.stack([x[0] for x in lab])
    own, gx, gy = (np.stack([x[i] for x in lab]) for i in (1, 2, 3))
    q, qprim = positions_from_labels(H, W, own, gx, gy, dtype=dtype)
    out = dict(pr)
    out.update(unary=unary, q=q, qprim=qprim, alphas=g.get_weights())
    return out


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("H,W,L,it", [(4, 4, 5, 6), (5, 7, 9, 6), (12, 10, 16, 8), (21, 34, 24, 7), (33, 29, 40, 5),
                                      (16, 18, 70, 4), (11, 13, 130, 3), (9, 12, 200, 3), (8, 9, 256, 3)])
def test_f64_bit_identical_problem(H, W, L, it, kernel):
    pr = synth.trws_problem(H, W, L, seed=H * 100 + W, kernel=kernel)
    g = TrwsGrid(kernel, H, W, L, pr["tol"], dict(precision="f64"))
    g.set_labels(0, pr["planes"], pr["unary"])
    g.set_weights(pr["alphas"])
    g.finalize()
    e, lb, n = g.minimize(it, 0.0)
    sol = g.labels()
    r = trws_oracle(_stored_problem(g, pr, H, W, np.float64), it)
    g.close()
    assert abs(e - r[1]) <= 1e-9 * abs(r[1]) and abs(lb - r[2]) <= 1e-9 * abs(r[2])
    assert np.array_equal(sol, r[0])
    if abs(r[1] - r[2]) > 1e-9 * abs(r[1]):
        assert n == r[3]   # fixed iteration count unless the gap closed to rounding level


@pytest.mark.parametrize("precision,rtol,min_equal", [("f32", 1e-4, 0.99), ("f64", 1e-9, 0.999)])
@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("H,W,L,it", [(24, 31, 16, 8), (48, 64, 8, 20), (30, 40, 64, 6), (17, 23, 129, 4)])
def test_against_exact_plane_problem(H, W, L, it, kernel, precision, rtol, min_equal):
    pr = synth.trws_problem(H, W, L, seed=7 + L, kernel=kernel)
    sol, e, lb, n = _grid_solve(pr, H, W, it, precision=precision)
    r = trws_oracle(pr, it)
    assert abs(e - r[1]) <= rtol * abs(r[1]) and abs(lb - r[2]) <= rtol * abs(r[2])
    assert np.mean(sol == r[0]) >= min_equal
    if abs(r[1] - r[2]) > 1e-6 * abs(r[1]):
        assert n == r[3]


@pytest.mark.parametrize("fuse", [True, False])
def test_fused_rounding_and_relgap_stop(fuse):
    H, W, L = 30, 30, 24
    pr = synth.trws_problem(H, W, L, seed=9, kernel=1)
    pr["alphas"][::7] = 0.0   # typeStereoLinear.h:390-395 shortcut (both terms of a pair: stride 7 hits singles too)
    sol, e, lb, n = _grid_solve(pr, H, W, 100, 1e-2, fuse_rounding=fuse)
    r = trws_oracle(pr, 100, 1e-2)
    assert abs(n - r[3]) <= 1
    assert abs(e - r[1]) <= 1e-4 * abs(r[1]) and abs(lb - r[2]) <= 1e-4 * abs(r[2])


def test_equals_matlab_layout_entry():
    """Same kernels' arithmetic on the same fp32 positions -> same energies as sb_trws_solve."""
    H, W, L = 40, 56, 48
    pr = synth.trws_problem(H, W, L, seed=3, kernel=1)
    g = TrwsGrid(1, H, W, L, pr["tol"], dict(precision="f32"))
    g.set_labels(0, pr["planes"], pr["unary"])
    g.set_weights(pr["alphas"])
    g.finalize()
    e, lb, n = g.minimize(10, 0.0)
    sol = g.labels()
    st = _stored_problem(g, pr, H, W, np.float32)
    g.close()
    sol2, e2, lb2, n2 = sb.trws(1, st["unary"], st["connectivity"], st["q"], st["qprim"], st["alphas"], st["tol"],
                                dict(maxiter=10, max_relgap=0))
    assert abs(e - e2) <= 1e-6 * abs(e2) and abs(lb - lb2) <= 1e-6 * abs(lb2)
    assert np.mean(sol == sol2) >= 0.999


def test_device_synth_problem_against_oracle():
    """The on-device generator (bench.py's big configs) produces a problem the reference solves to the
    same answer."""
    H, W, L = 40, 52, 32
    for kernel in (1, 2):
        g = TrwsGrid(kernel, H, W, L, 0.02 if kernel == 1 else 0.02 ** 2, dict(precision="f32"))
        g.synth(0xB200 + 4)
        g.finalize()
        e, lb, n = g.minimize(6, 0.0)
        sol = g.labels()
        ind1, ind2 = sb.construct_neighborhood(H, W)
        pr = dict(kernel=kernel, connectivity=np.stack([ind1, ind2]), tol=0.02 if kernel == 1 else 0.02 ** 2,
                  unary=np.zeros((L, H * W)))
        st = _stored_problem(g, pr, H, W, np.float32)
        g.close()
        r = trws_oracle(st, 6)
        assert abs(e - r[1]) <= 1e-4 * abs(r[1]) and abs(lb - r[2]) <= 1e-4 * abs(r[2])
        assert np.mean(sol == r[0]) >= 0.99
        assert len(np.unique(st["alphas"])) >= 2 and st["unary"].max() <= np.log(2.0) + 1e-6


def test_continue_and_reset():
    H, W, L = 20, 25, 12
    pr = synth.trws_problem(H, W, L, seed=11, kernel=1)
    g = TrwsGrid(1, H, W, L, pr["tol"], dict(precision="f64", fuse_rounding=False))
    g.set_labels(0, pr["planes"][:5], pr["unary"][:5])
    g.set_labels(5, pr["planes"][5:], pr["unary"][5:])     # labels may arrive in groups
    g.set_weights(pr["alphas"])
    g.finalize()
    a = g.minimize(3, 0.0)
    b = g.minimize(3, 0.0)     # continues from the current messages (6 iterations in all)
    g.reset()
    c = g.minimize(6, 0.0)
    g.close()
    assert b[1] >= a[1] - 1e-9 * abs(a[1])
    assert abs(b[0] - c[0]) <= 1e-9 * abs(c[0]) and abs(b[1] - c[1]) <= 1e-9 * abs(c[1])


def test_errors():
    from stereo_b200._lib import SbError
    with pytest.raises(SbError, match="Unsupported kernel"):
        TrwsGrid(3, 8, 8, 4, 0.1)
    with pytest.raises(SbError):
        TrwsGrid(1, 3, 9, 4, 0.1)            # degenerate grids go through sb_trws_solve
    g = TrwsGrid(1, 6, 6, 3, 0.1)
    pl = np.zeros((3, 4, 36))
    with pytest.raises(SbError, match="Infinite disparity"):
        g.set_labels(0, pl, np.zeros((3, 36)))    # c == 0 (dispmap_super.m:321-323)
    with pytest.raises(SbError, match="finalize"):
        g.minimize(1, 0.0)
    g.close()


def test_memory_footprint():
    """45 bytes per label and node (fp32): what makes BASELINE configs 4 / 5 exist."""
    g = TrwsGrid(1, 64, 96, 64, 0.02)
    info = g.info()
    g.close()
    per = info["hbm_bytes"] / (64 * 96 * 64)
    assert 45.0 <= per <= 46.5, per


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("H,W,L,it,world,blocks", [(12, 17, 8, 5, 2, 1), (24, 31, 16, 6, 3, 1), (9, 40, 40, 4, 4, 1), (40, 64, 70, 3, 4, 2),
                                                   (31, 33, 130, 3, 2, 4), (16, 64, 24, 4, 8, 2), (20, 130, 12, 4, 2, 3),
                                                   (14, 200, 20, 3, 4, 0), (10, 37, 9, 5, 3, 2)])
def test_column_bands_one_gpu(H, W, L, it, world, blocks, kernel):
    """The multi-GPU sweep (column bands, boundary messages written into the neighbour's arrays, self-validating
    words, passes launched without barriers) with all ranks on THIS device: same labels / energy / bound as the
    single-rank sweep of the same problem."""
    from stereo_b200.gridsolver import TrwsGridLocalGroup
    pr = synth.trws_problem(H, W, L, seed=7 * H + W + world, kernel=kernel)
    ref = TrwsGrid(kernel, H, W, L, pr["tol"])
    ref.set_labels(0, pr["planes"], pr["unary"])
    ref.set_weights(pr["alphas"])
    ref.finalize()
    e1, lb1, _ = ref.minimize(it, 0.0)
    lab1 = ref.labels()
    ref.close()
    grp = TrwsGridLocalGroup(kernel, H, W, L, pr["tol"], world, dict(col_blocks=blocks))
    try:
        grp.each(lambda g: g.set_labels(0, pr["planes"], pr["unary"]))
        grp.each(lambda g: g.set_weights(pr["alphas"]))
        grp.finalize()
        e, lb, _ = grp.minimize(it)
        lab = grp.labels()
        # every node is swept by exactly one rank (labels are 1-based; a node no rank owns would stay 0)
        assert lab.min() >= 1 and lab.max() <= L
    finally:
        grp.close()
    # the bands sum their energy / bound contributions in a different order: summation-order differences only
    assert abs(e - e1) <= 1e-5 * abs(e1) and abs(lb - lb1) <= 1e-5 * abs(lb1)
    assert np.mean(lab == lab1) >= 0.999


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("H,W,L,it,world", [(24, 31, 16, 5, 1), (18, 44, 64, 4, 1), (33, 40, 100, 3, 1), (21, 60, 192, 3, 1), (14, 48, 250, 2, 1),
                                            (18, 44, 64, 4, 2), (40, 64, 100, 3, 4), (21, 60, 192, 3, 3), (14, 48, 250, 2, 2)])
def test_latency_build(H, W, L, it, world, kernel):
    """sb_trws_options.latency_mode: the sweep kernel built for at most two strip walkers per SM (no spills, operands taken
    ahead of the dependent chain), which small grids and banded runs pick on their own, against the throughput build
    (latency_mode = -1).  Same DAG, same arithmetic: identical labels / energy / bound on a single rank, and the banded
    sweep on the latency build agrees with the single-rank sweep on the throughput build."""
    from stereo_b200.gridsolver import TrwsGridLocalGroup
    pr = synth.trws_problem(H, W, L, seed=3 * H + W + world, kernel=kernel)
    ref = TrwsGrid(kernel, H, W, L, pr["tol"], dict(latency_mode=-1))
    assert ref.info()["latency_build"] == 0
    ref.set_labels(0, pr["planes"], pr["unary"])
    ref.set_weights(pr["alphas"])
    ref.finalize()
    e1, lb1, _ = ref.minimize(it, 0.0)
    lab1 = ref.labels()
    ref.close()
    if world == 1:
        g = TrwsGrid(kernel, H, W, L, pr["tol"], dict(latency_mode=1))
        assert g.info()["latency_build"] == 1
        g.set_labels(0, pr["planes"], pr["unary"])
        g.set_weights(pr["alphas"])
        g.finalize()
        e, lb, _ = g.minimize(it, 0.0)
        lab = g.labels()
        g.close()
        assert e == e1 and lb == lb1 and np.array_equal(lab, lab1)
        return
    grp = TrwsGridLocalGroup(kernel, H, W, L, pr["tol"], world, dict(latency_mode=1))
    try:
        assert all(g.info()["latency_build"] == 1 for g in grp.ranks)
        grp.each(lambda g: g.set_labels(0, pr["planes"], pr["unary"]))
        grp.each(lambda g: g.set_weights(pr["alphas"]))
        grp.finalize()
        e, lb, _ = grp.minimize(it)
        lab = grp.labels()
    finally:
        grp.close()
    assert abs(e - e1) <= 1e-5 * abs(e1) and abs(lb - lb1) <= 1e-5 * abs(lb1)
    assert np.mean(lab == lab1) >= 0.999


def test_column_bands_one_gpu_f64_exact():
    from stereo_b200.gridsolver import TrwsGridLocalGroup
    H, W, L, it, world = 21, 34, 24, 6, 3
    pr = synth.trws_problem(H, W, L, seed=11, kernel=1)
    ref = TrwsGrid(1, H, W, L, pr["tol"], dict(precision="f64"))
    ref.set_labels(0, pr["planes"], pr["unary"])
    ref.set_weights(pr["alphas"])
    ref.finalize()
    e1, lb1, _ = ref.minimize(it, 0.0)
    lab1 = ref.labels()
    ref.close()
    grp = TrwsGridLocalGroup(1, H, W, L, pr["tol"], world, dict(precision="f64"))
    try:
        grp.each(lambda g: g.set_labels(0, pr["planes"], pr["unary"]))
        grp.each(lambda g: g.set_weights(pr["alphas"]))
        grp.finalize()
        e, lb, _ = grp.minimize(it)
        lab = grp.labels()
    finally:
        grp.close()
    assert abs(e - e1) <= 1e-12 * abs(e1) and abs(lb - lb1) <= 1e-12 * abs(lb1)
    assert np.array_equal(lab, lab1)


def test_alternating_schedule_on_device():
    """BASELINE configs[4]'s schedule at a small size (scripts/cfg5_alternating.py): binary fusions from device-resident
    fields (sb_binary_fusion_grid, on_device) alternating with the grid TRW-S whose last label is re-set from a device
    pointer.  A QPBO fusion move never increases the energy; the first TRW-S energy can be checked independently."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import cfg5_alternating
    out = cfg5_alternating.run(40, 56, 12, rounds=2, fusions=4, iters=8, verbose=False, check=True)
    for rec in out["log"]:
        # the solver's (fp32) energy is the energy of the assignment it hands back, evaluated independently in fp64
        assert abs(rec["trws"]["energy"] - rec["trws"]["energy_of_assignment"]) <= 1e-4 * abs(rec["trws"]["energy_of_assignment"])
        e = [f["energy"] for f in rec["fusions"]]
        assert all(e[i + 1] <= e[i] * (1 + 1e-12) for i in range(len(e) - 1))
        assert rec["trws"]["lower_bound"] <= rec["trws"]["energy"] * (1 + 1e-6)
    # the simultaneous fusion starts from the fused assignment as one of its labels: it should not end far above it
    assert out["log"][0]["trws"]["energy"] <= out["log"][0]["fusions"][-1]["energy"] * 1.05


def test_set_labels_from_device_pointer():
    import torch
    H, W, L = 16, 20, 6
    pr = synth.trws_problem(H, W, L, seed=5, kernel=1)
    res = []
    for on_dev in (False, True):
        g = TrwsGrid(1, H, W, L, pr["tol"])
        if on_dev:
            pl = torch.from_numpy(np.ascontiguousarray(pr["planes"].transpose(0, 2, 1))).cuda()
            un = torch.from_numpy(np.ascontiguousarray(pr["unary"])).cuda()
            torch.cuda.synchronize()
            g.set_labels_ptr(0, L, pl.data_ptr(), un.data_ptr())
        else:
            g.set_labels(0, pr["planes"], pr["unary"])
        g.set_weights(pr["alphas"])
        g.finalize()
        res.append(g.minimize(5, 0.0) + (g.labels(),))
        g.close()
    assert res[0][:3] == res[1][:3] and np.array_equal(res[0][3], res[1][3])
