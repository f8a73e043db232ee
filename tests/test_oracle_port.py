"""CPU tests: the C port of the reference algorithms (oracle/*.c) is pinned against
(a) golden vectors generated from the UNMODIFIED reference (tests/golden/make_golden.py) and
(b) the compiled reference itself where oracle/_ref is present."""
import numpy as np
import pytest

from oracle import oracle
from stereo_b200 import synth
from util import golden

TRWS = golden("trws_solve.npz")
NCASES = len([k for k in TRWS.files if k.endswith("_params")])


@pytest.mark.parametrize("i", range(NCASES))
def test_port_trws_matches_reference_golden(i):
    H, W, L, k, seed, it, gap = TRWS[f"case{i}_params"]
    H, W, L, k, seed, it = int(H), int(W), int(L), int(k), int(seed), int(it)
    pr = synth.trws_problem(H, W, L, seed=seed, kernel=k)
    sums = np.array([pr["unary"].sum(), pr["q"].sum(), pr["qprim"].sum(), pr["alphas"].sum()])
    np.testing.assert_allclose(sums, TRWS[f"case{i}_inputsum"], rtol=1e-12)  # generator did not drift
    lab, e, lb, n = oracle.trws_solve(k, pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T,
                                      pr["alphas"], pr["tol"], it, gap, kind="port")
    ge, glb, gn = TRWS[f"case{i}_scalars"]
    assert n == gn
    np.testing.assert_allclose([e, lb], [ge, glb], rtol=1e-12)
    assert np.array_equal(lab, TRWS[f"case{i}_labels"].astype(np.int64))


def test_port_ordering_matches_reference_golden():
    g = golden("ordering.npz")
    for key in g.files:
        H, W = map(int, key.split("x"))
        assert np.array_equal(oracle.trws_ordering(H, W, kind="port"), g[key]), key


def test_port_update_message_matches_reference_golden():
    g = golden("update_message.npz")
    n = len([k for k in g.files if k.endswith("_par")])
    assert n == 32
    for i in range(n):
        Di, msg, s0, s1 = g[f"u{i}_in"]
        k, alpha, lam, gamma, dir_, sw = g[f"u{i}_par"]
        m2, vmin = oracle.trws_update_message(int(k), Di, msg, s0, s1, alpha, lam, gamma, int(dir_), int(sw), kind="port")
        out = g[f"u{i}_out"]
        np.testing.assert_allclose(m2, out[:-1], rtol=0, atol=1e-13)
        assert abs(vmin - out[-1]) <= 1e-13


def test_update_message_is_bruteforce_minplus():
    """the envelope algorithms equal the O(L^2) definition
    msg[j] = min_i (gamma*Di[i] - msg[i] + alpha*min(|x_dst[j]-x_src[i]|^k, lambda)) - min."""
    g = golden("update_message.npz")
    n = len([k for k in g.files if k.endswith("_par")])
    for i in range(n):
        Di, msg, s0, s1 = g[f"u{i}_in"]
        k, alpha, lam, gamma, dir_, sw = g[f"u{i}_par"]
        src, dst = (s1, s0) if int(dir_) == int(sw) else (s0, s1)  # typeStereoLinear.h:343-357
        H = gamma * Di - msg
        d = np.abs(dst[:, None] - src[None, :]) ** int(k)
        bf = (H[None, :] + alpha * np.minimum(d, lam)).min(axis=1)
        out = g[f"u{i}_out"]
        np.testing.assert_allclose(bf - bf.min(), out[:-1], rtol=0, atol=1e-12)
        assert abs(bf.min() - out[-1]) <= 1e-12


def test_port_interp2_matches_reference_golden():
    g = golden("interp2.npz")
    B = oracle.interp2_linear(g["A"], g["X"], g["Y"], -1000.0, kind="port")
    assert np.array_equal(B, g["B"])


@pytest.mark.skipif(not oracle.have_ref("trws"), reason="oracle/_ref not built (no /root/reference)")
def test_port_equals_compiled_reference_live():
    for (H, W, L, k, it) in [(9, 11, 6, 1, 6), (9, 11, 6, 2, 6), (15, 14, 40, 1, 5)]:
        pr = synth.trws_problem(H, W, L, seed=11, kernel=k)
        args = (k, pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T, pr["alphas"], pr["tol"], it, 0)
        a = oracle.trws_solve(*args, kind="reference")
        b = oracle.trws_solve(*args, kind="port")
        assert np.array_equal(a[0], b[0]) and a[1] == b[1] and a[2] == b[2] and a[3] == b[3]


def test_teddy_fixture_integrity():
    """tests/golden/teddy.npz (made by tests/golden/make_teddy.py from the reference's data/teddy) is the pair the
    examples run on: shape and checksums as recorded when it was generated."""
    import numpy as np
    from util import golden
    g = golden("teddy.npz")
    assert g["im2"].shape == g["im6"].shape == (375, 450, 3) and g["im2"].dtype == np.uint8
    assert int(g["im2"].astype(np.int64).sum()) == 60448401 and int(g["im6"].astype(np.int64).sum()) == 60462544
