"""CPU tests of the NumPy restatements behind the SURVEY 8(f) rows (oracle/stereo_np.py): the vectorised forms against
direct per-pixel definitions, so the checker the GPU tests compare with is itself checked.  (No MATLAB exists here: these
restatements are unpinned against the reference)"""
import numpy as np

from oracle import oracle, stereo_np
from stereo_b200 import synth


def _interp(A, X, Y, oobv):
    return oracle.interp2_linear(A, X, Y, oobv, kind="reference" if oracle.have_ref("interp2") else "port")


def test_segpln_wta_restatement_against_per_pixel_definition():
    """dispmap_globalstereo.m:83-117, one output pixel at a time: mean photo cost over the (2 w + 1)^2 window around the
    pixel, normalised, first maximum over the levels, 0.07 threshold, symmetric padding."""
    H, W, w = 14, 18, 1
    im0, im1, _ = synth.stereo_pair(H, W, 3, seed=4)
    P = np.zeros((3, 4, 2))
    P[:, :3, 0] = np.eye(3)
    P[:, :3, 1] = np.eye(3)
    P[0, 3, 1] = -0.5
    disps = np.arange(8.0, -1.0, -1.0)
    col_thresh = 30.0
    got, vol = stereo_np.segpln_wta([im0, im1], P, disps, w, col_thresh, _interp)
    # direct definition
    C = 3
    cost = np.zeros((H, W, disps.size))
    for b, dv in enumerate(disps):
        for r in range(H):
            for c in range(W):
                tot = 0.0
                for a, im in enumerate((im0, im1)):
                    x, y = c + 1.0, r + 1.0
                    X = P[:, :3, a] @ np.array([x, y, 1.0]) + dv * P[:, 3, a]
                    s = _interp(im, np.array([X[0] / X[2]]), np.array([X[1] / X[2]]), -1000.0)[0]
                    f = s - im0[r, c]
                    tot += np.log(2.0) - np.log(np.exp((f ** 2).sum() * (-1.0 / (col_thresh * C))) + 1.0)
                cost[r, c, b] = tot
    x1 = 2 * (np.log(2.0) - np.log(np.exp(((-1000.0 - im0[0, 0]) ** 2).sum() * (-1.0 / (col_thresh * C))) + 1.0))
    ref_vol = np.zeros((H - 2 * w, W - 2 * w, disps.size))
    for r in range(H - 2 * w):
        for c in range(W - 2 * w):
            ref_vol[r, c] = (x1 - cost[r:r + 2 * w + 1, c:c + 2 * w + 1].mean(axis=(0, 1))) / x1
    assert np.allclose(vol, ref_vol, rtol=1e-12, atol=1e-14)
    idx = ref_vol.argmax(axis=2)
    inner = disps[idx]
    inner[ref_vol.max(axis=2) < 0.07] = 0
    top2 = np.sort(ref_vol, axis=2)[:, :, -2:]
    clear = (top2[:, :, 1] - top2[:, :, 0]) > 1e-10
    assert np.array_equal(got[w:H - w, w:W - w][clear], inner[clear])
    # symmetric padding (padarray 'symmetric'): the border repeats the first interior ring, mirrored
    assert np.array_equal(got[0, w:W - w], got[w, w:W - w]) and np.array_equal(got[w:H - w, W - 1], got[w:H - w, W - 1 - w])


def test_smooth_weights_restatement():
    """dispmap_globalstereo.m:396-400 term by term, in construct_neighborhood order."""
    H, W = 5, 7
    rng = np.random.default_rng(0)
    seg = rng.integers(0, 3, size=(H, W))
    got = stereo_np.smooth_weights(H, W, seg, 4.0, 0.25, 2.0)
    ind1, ind2 = stereo_np.construct_neighborhood(H, W)
    flat = seg.reshape(-1, order="F")
    for p in range(ind1.size):
        assert got[p] == (4.0 if flat[ind1[p] - 1] == flat[ind2[p] - 1] else 0.25) * 2.0
    assert got.size == 2 * ((H - 1) * W + H * (W - 1))


def test_segpln_host_glue_ransac():
    """The host glue of the Python mirror's segpln: dispmap_globalstereo.rplane / nsamples (dispmap_globalstereo.m:417-466)
    restated over a NumPy generator -- on points of one plane plus gross outliers the inlier set is the plane."""
    from stereo_b200.dispmap import _nsamples, _rplane
    rng = np.random.default_rng(0)
    N = np.array([0.01, -0.02, -0.5])
    P = rng.random((200, 3)) * np.array([50, 40, 1]) + np.array([0, 0, 1.5])
    P[:, 2] = (-1 - P[:, 0] * N[0] - P[:, 1] * N[1]) / N[2]
    P[:40, 2] += rng.normal(0, 2, 40)
    inl = _rplane(P, 0.1, rng)
    assert inl[40:].all() and inl[:40].sum() <= 10
    # nsamples: log(1 - conf) / log(1 - q) with q = prod((ni-2:ni) ./ (n-2:n)), at least 1
    q = (148 / 198) * (149 / 199) * (150 / 200)
    assert abs(_nsamples(150, 200, 3, 0.95) - np.log(0.05) / np.log(1 - q)) < 1e-12
    assert _nsamples(200, 200, 3, 0.95) == 1.0
