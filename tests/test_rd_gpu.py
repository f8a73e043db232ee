"""GPU parity tests of the QPBO path through the C ABI (sb_rd_solve) against the reference's own
rd_mex.cpp + QPBO 1.3 (oracle/_ref, or the golden fixtures generated from it).

Bar (BASELINE.json north_star): integer label assignments bit-exact; energy / bound within 1e-9
relative (fp64 on both sides, different summation order)."""
import ctypes

import numpy as np
import pytest

import stereo_b200 as sb
from stereo_b200 import synth
from util import golden

pytestmark = pytest.mark.gpu

RD = golden("rd_solve.npz")
NCASES = len([k for k in RD.files if k.endswith("_params")])
libc = ctypes.CDLL(None)


def _solve(pr, improve=False):
    libc.srand(1)  # QPBO::Improve draws from the process-global libc stream (QPBO_extra.cpp:23)
    return sb.rd(pr["U0"], pr["U1"], pr["E00"], pr["E01"], pr["E10"], pr["E11"], pr["connectivity"],
                 dict(improve=improve))


def _ref(pr, improve=False):
    from oracle import oracle
    libc.srand(1)
    return oracle.rd_solve(pr["U0"], pr["U1"], pr["E00"], pr["E01"], pr["E10"], pr["E11"],
                           (pr["connectivity"] - 1).T, improve=improve)


def _problem(params):
    H, W, seed, kernel, mode, improve = params
    return synth.rd_problem(int(H), int(W), seed=int(seed), kernel=int(kernel),
                            mode="frustrated" if int(mode) else "stereo"), bool(improve)


@pytest.mark.parametrize("i", range(NCASES))
def test_golden_cases(i):
    pr, improve = _problem(RD[f"case{i}_params"])
    lab, e, lb, nu = _solve(pr, improve)
    glab = RD[f"case{i}_labels"].astype(np.float64)
    ge, glb, gnu = RD[f"case{i}_scalars"]
    assert np.array_equal(lab, glab), f"{int((lab != glab).sum())} labels differ"
    assert abs(e - ge) <= 1e-9 * abs(ge)
    assert nu == gnu
    if not (improve and gnu > 0):  # after Improve the reference's bound depends on its particular flow
        assert abs(lb - glb) <= 1e-9 * abs(glb)


def test_reference_when_available():
    from oracle import oracle
    if not oracle.have_ref("rd"):
        pytest.skip("oracle/_ref/libref_rd.so not present")
    for seed in range(6):
        for mode in ("stereo", "frustrated"):
            pr = synth.rd_problem(23 + seed, 31 - seed, seed=100 + seed, kernel=1 + seed % 2, mode=mode)
            a, r = _solve(pr), _ref(pr)
            assert np.array_equal(a[0], r[0]), (seed, mode, int((a[0] != r[0]).sum()))
            assert abs(a[1] - r[1]) <= 1e-9 * abs(r[1]) and abs(a[2] - r[2]) <= 1e-9 * abs(r[2]) and a[3] == r[3]


def test_brute_force_tiny_grids():
    """Known answer: on grids small enough to enumerate, the strong labels agree with every
    global minimum and the bound does not exceed the minimum energy."""
    for seed in range(8):
        H, W = 3, 4
        pr = synth.rd_problem(H, W, seed=seed, mode="frustrated" if seed % 2 else "stereo")
        lab, e, lb, nu = _solve(pr)
        N = H * W
        i1, i2 = pr["connectivity"] - 1
        T = np.stack([pr["E00"], pr["E01"], pr["E10"], pr["E11"]])
        best, mins = np.inf, []
        for code in range(1 << N):
            x = (code >> np.arange(N)) & 1
            en = np.where(x == 1, pr["U1"], pr["U0"]).sum() + T[2 * x[i1] + x[i2], np.arange(i1.size)].sum()
            if en < best - 1e-12:
                best, mins = en, [x]
            elif abs(en - best) <= 1e-12:
                mins.append(x)
        assert lb <= best + 1e-9
        for x in mins:
            fixed = lab >= 0
            strong_ok = True  # weak persistencies may pick one of several optima; check consistency with at least one
            if not np.array_equal(x[fixed], lab[fixed].astype(int)):
                strong_ok = False
            if strong_ok:
                break
        else:
            pytest.fail("labels agree with no global minimum")


def test_full_size_properties():
    """BASELINE config 3 shape (1080 x 1920): too slow for per-label oracle comparison in the test
    budget here, so check size-independent properties: lb <= e, labels in {-1, 0, 1}, energy equals
    an independent evaluation, fusing never increases the energy over keeping (x = 0)."""
    pr = synth.rd_problem(1080, 1920, seed=5, mode="stereo")
    lab, e, lb, nu = _solve(pr)
    assert set(np.unique(lab)).issubset({-1.0, 0.0, 1.0})
    assert lb <= e * (1 + 1e-12)
    x = (lab == 1).astype(int)
    i1, i2 = pr["connectivity"] - 1
    T = np.stack([pr["E00"], pr["E01"], pr["E10"], pr["E11"]])
    en = np.where(x == 1, pr["U1"], pr["U0"]).sum() + T[2 * x[i1] + x[i2], np.arange(i1.size)].sum()
    assert abs(en - e) <= 1e-9 * abs(en)
    e0 = pr["U0"].sum() + pr["E00"].sum()
    assert e <= e0 * (1 + 1e-12)
    from oracle import oracle
    if oracle.have_ref("rd"):
        r = _ref(pr)
        assert np.array_equal(lab, r[0])
        assert abs(e - r[1]) <= 1e-9 * abs(r[1]) and abs(lb - r[2]) <= 1e-9 * abs(r[2])


def test_argument_validation():
    pr = synth.rd_problem(5, 6, seed=0)
    conn = pr["connectivity"].copy()
    conn[:, [0, 1]] = conn[:, [1, 0]]
    with pytest.raises(sb._lib.SbError) as ei:
        sb.rd(pr["U0"], pr["U1"], pr["E00"], pr["E01"], pr["E10"], pr["E11"], conn, {})
    assert ei.value.code == sb._lib.SB_ENOTGRID
