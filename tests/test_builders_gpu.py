"""GPU parity tests of the cost-volume / unary / pairwise builders through the C ABI against
the NumPy restatement of the MATLAB code (oracle/stereo_np.py) and, for the vgg_interp2 gather,
against the compiled reference / its golden fixture.  Tolerance 1e-4 relative (north star) for the
fp32-accumulated NCC volume on well-conditioned windows; 1e-12 for the fp64 elementwise paths."""
import numpy as np
import pytest

import stereo_b200 as sb
from stereo_b200 import builders, synth
from util import golden

pytestmark = pytest.mark.gpu


def _np():
    from oracle import stereo_np
    return stereo_np


@pytest.mark.parametrize("H,W,D,p,frac", [(37, 53, 8, 2, False), (64, 48, 6, 4, False), (30, 41, 5, 2, True),
                                          (7, 9, 3, 2, False), (40, 33, 4, 0, False)])
def test_ncc_volume(H, W, D, p, frac):
    im0, im1, _ = synth.stereo_pair(H, W, D, seed=H + W)
    disps = np.arange(D, dtype=np.float64) * (1.5 if frac else 1.0)
    got = builders.ncc_volume(im0, im1, disps, p)
    ref = _np().compute_ncc(im0, im1, disps, p)
    cond = _np().ncc_conditioning(im0, im1, disps, p)
    ok = cond > 1e-6
    assert ok.mean() > 0.9
    assert np.all(np.abs(got - ref)[ok] <= 1e-4 * np.maximum(np.abs(ref[ok]), 1e-2))
    # masked columns and the zero padding region are exact zeros in both
    assert np.array_equal(got == 0, ref == 0) or np.mean((got == 0) == (ref == 0)) > 0.999


@pytest.mark.parametrize("H,W,disps,p", [(37, 53, list(range(8)), 2), (101, 300, list(range(0, 120, 3)), 4), (64, 200, list(range(0, 64, 4)), 2),
                                         (7, 9, [0, 1, 2], 1), (33, 70, [5, 0, 17, 2], 8), (130, 96, list(range(20)), 3),
                                         (50, 400, list(range(130)), 4)])
def test_ncc_one_pass_kernel_equals_general(H, W, disps, p, monkeypatch):
    """The one-pass volume kernel (exact integer window sums, TMA-staged ring, all levels per CTA) against the general
    per-level kernel on the same 8-bit pair, and the device-resident handle against the host-array entry points."""
    im0, im1, _ = synth.stereo_pair(H, W, max(disps) + 1, seed=H * 7 + W)
    d = np.asarray(disps, dtype=np.float64)
    vol = builders.NccVolume(im0, im1, d, p)
    assert vol.info()["one_pass"]
    fast = vol.get()
    monkeypatch.setenv("SB_NCC_GENERAL", "1")
    gen = builders.ncc_volume(im0, im1, d, p)
    monkeypatch.delenv("SB_NCC_GENERAL")
    cond = _np().ncc_conditioning(im0, im1, d, p)
    ok = cond > 1e-6
    # the general kernel combines fp32 sums in doubles with cancellation; the one-pass kernel has none
    assert np.all(np.abs(fast - gen)[ok] <= 2e-5)
    assert np.all(np.abs(fast) <= 1.0 + 1e-5)
    ref = _np().compute_ncc(im0, im1, d, p)
    assert np.all(np.abs(fast - ref)[ok] <= 1e-4 * np.maximum(np.abs(ref[ok]), 1e-2))
    # handle methods == host-array entry points on the same volume
    assert np.array_equal(vol.best_disp(), builders.ncc_best_disp(fast, d))
    x = np.random.default_rng(1).random((H, W)) * (d.max() + 2) - 1
    assert np.array_equal(vol.sample(x, 40.0, True), builders.ncc_sample(fast, d, x.reshape(-1, order="F"), 40.0, True))
    vol.close()


def test_ncc_sampling_and_wta():
    H, W, D = 33, 47, 9
    im0, im1, _ = synth.stereo_pair(H, W, D, seed=3)
    disps = np.arange(D, dtype=np.float64) * 2.0
    ncc = _np().compute_ncc(im0, im1, disps, 2)
    best = builders.ncc_best_disp(ncc, disps)
    rbest = _np().best_disp_from_ncc(ncc, disps)
    m = np.isfinite(rbest)
    assert np.allclose(best[m], rbest[m], rtol=1e-12, atol=1e-12)
    rng = np.random.default_rng(0)
    x = rng.random((H, W)) * (disps.max() + 4) - 2      # includes out-of-range and exact levels
    x[0, :D] = disps
    x[1, : D - 1] = (disps[:-1] + disps[1:]) / 2         # exact ties between two levels (:232)
    got = builders.ncc_sample(ncc, disps, x.reshape(-1, order="F"))
    ref = _np().sample_ncc_from_disp(ncc, disps, x)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-12)
    u = builders.ncc_sample(ncc, disps, x.reshape(-1, order="F"), 40.0, True)
    assert np.allclose(u, 40.0 * (1 - ref), rtol=1e-12, atol=1e-9)


def test_interp2_golden():
    g = golden("interp2.npz")
    got = builders.interp2_linear(g["A"], g["X"], g["Y"], -1000.0)
    assert np.array_equal(got, g["B"])       # same operation order as vgg_interp2.cxx:262-266 -> bit-exact


def test_photo_unary_and_plane_disparity():
    H, W = 29, 37
    im0, im1, _ = synth.stereo_pair(H, W, 12, seed=9)
    rng = np.random.default_rng(1)
    planes = np.zeros((4, H * W))
    planes[0] = (rng.random(H * W) - 0.5) * 0.01
    planes[1] = (rng.random(H * W) - 0.5) * 0.01
    planes[2] = 1.0
    planes[3] = -rng.random(H * W) * 50
    pts = sb.get_points(H, W)
    d_min, d_step = 0.0, 48.0
    nd = builders.plane_disparity(planes, pts, d_min, d_step)
    assert np.allclose(nd, _np().disparity_from_assignment(planes, pts, d_min, d_step), rtol=1e-14, atol=1e-14)
    P2 = np.array([[1, 0, 0, -0.25], [0, 1, 0, 0], [0, 0, 1, 0]], dtype=np.float64).T   # example_global.m:17-18
    got = builders.photo_unary(im0, im1, P2, planes, d_min, d_step, 30.0)
    from oracle import oracle

    def interp(A, X, Y, oobv):
        kind = "reference" if oracle.have_ref("interp2") else "port"
        return oracle.interp2_linear(A, X, Y, oobv, kind=kind)
    ref = _np().photo_unary_cost([im0, im1], P2, planes, d_min, d_step, 30.0, interp)
    assert np.allclose(got, ref, rtol=1e-12, atol=1e-13)
    bad = planes.copy()
    bad[2, 5] = 0
    with pytest.raises(sb._lib.SbError) as ei:          # dispmap_super.m:323-325
        builders.plane_disparity(bad, pts)
    assert "Infinite disparity" in str(ei.value)


@pytest.mark.parametrize("kernel", [1, 2])
def test_pairwise_tables_positions_energy(kernel):
    H, W, L = 17, 23, 5
    rng = np.random.Generator(np.random.PCG64(4))
    props = synth.random_plane_proposals(H, W, L, rng)
    E = 2 * ((H - 1) * W + H * (W - 1))
    w = rng.random(E) * 10
    tol = 0.02 if kernel == 1 else 0.02 ** 2
    got = builders.pairwise_tables(H, W, kernel, props[1], props[2], w, tol, 0.1, 2.0)
    ref = _np().all_pairwise_costs(H, W, kernel, w, tol, props[1], props[2], 0.1, 2.0)
    for a, b in zip(got, ref):
        assert np.allclose(a, b, rtol=1e-12, atol=1e-15)
    q, qp = builders.fusion_positions(H, W, list(props), 0.1, 2.0)
    rq, rqp = _np().fusion_positions(H, W, list(props), 0.1, 2.0)
    assert np.allclose(q, rq, rtol=1e-13, atol=1e-14) and np.allclose(qp, rqp, rtol=1e-13, atol=1e-14)
    un = rng.random(H * W)
    e = builders.energy(H, W, kernel, un, props[1], w, tol, 0.1, 2.0)
    assert abs(e - _np().energy(H, W, kernel, w, tol, un, props[1], 0.1, 2.0)) <= 1e-11 * abs(e)


def test_dispmap_ncc_class_end_to_end():
    """example_ncc.m plumbing at a small size: constructor (NCC volume + WTA init), one binary
    fusion (QPBO) and one simultaneous fusion (TRW-S); energies must not increase."""
    H, W = 32, 40
    im0, im1, _ = synth.stereo_pair(H, W, 10, seed=2)
    dm = sb.dispmap_ncc([im0, im1], np.arange(0, 11, dtype=np.float64), 1, 40.0, 8.0)
    e0 = dm.energy()
    prop = np.zeros((4, H * W))
    prop[2] = 1
    prop[3] = -5.0
    dm.binary_fusion(prop)
    e1 = dm.energy()
    assert e1 <= e0 + 1e-9 * abs(e0)
    dm.maxiter = 30
    props = []
    for d in (0.0, 3.0, 6.0, 9.0):
        p = np.zeros((4, H * W))
        p[2] = 1
        p[3] = -d
        props.append(p)
    dm.simultaneous_fusion(props)
    assert dm.energy() <= e1 + 1e-6 * abs(e1)
    assert dm.current_dispmap().shape == (H, W)


@pytest.mark.parametrize("n_images,window", [(2, 2), (3, 1), (2, 0)])
def test_segpln_wta_volume(n_images, window):
    """dispmap_globalstereo.segpln, the window-matching WTA disparity (dispmap_globalstereo.m:83-117), against the NumPy
    restatement (the gather through the compiled vgg_interp2 where it is built): same disparity wherever the best level
    wins by more than rounding, same scores; and the disparity found is the true one of the synthetic pair."""
    from oracle import oracle
    H, W = 36, 52
    im0, im1, dtrue = synth.stereo_pair(H, W, 8, seed=11)
    images = [im0, im1] + ([im1[:, ::-1].copy()] if n_images == 3 else [])
    P = np.zeros((3, 4, n_images))
    for a in range(n_images):
        P[:, :3, a] = np.eye(3)
    P[0, 3, 1] = -0.25                                   # example_global.m:17-18: 4 disparity units per pixel
    if n_images == 3:
        P[0, 3, 2] = 0.125
    disps = np.arange(40.0, -1.0, -1.0)                  # descending, like self.disps (:48)
    corr, score = builders.segpln_wta(images, P, disps, window, 30.0, return_score=True)

    def interp(A, X, Y, oobv):
        kind = "reference" if oracle.have_ref("interp2") else "port"
        return oracle.interp2_linear(A, X, Y, oobv, kind=kind)
    ref, vol = _np().segpln_wta(images, P, disps, window, 30.0, interp)
    assert corr.shape == (H, W) and score.shape == vol.shape[:2]
    assert np.allclose(score, vol.max(axis=2), rtol=1e-12, atol=1e-13)
    top2 = np.sort(vol, axis=2)[:, :, -2:]
    clear = np.pad((top2[:, :, 1] - top2[:, :, 0]) > 1e-9, window, mode="symmetric")
    near_thresh = np.pad(np.abs(vol.max(axis=2) - 0.07) < 1e-9, window, mode="symmetric")
    ok = clear & ~near_thresh
    assert ok.mean() > 0.6
    assert np.array_equal(corr[ok], ref[ok])
    if n_images == 2 and window == 2:
        # the pair is im1(r, c - d) = im0(r, c): with 4 units per pixel the winning level is about 4 d away from the borders
        inner = (slice(6, H - 6), slice(14, W - 6))
        found = corr[inner] > 0
        assert found.mean() > 0.4
        assert np.median(np.abs(corr[inner][found] / 4.0 - dtrue[inner][found])) <= 1.0


@pytest.mark.parametrize("kernel", [1, 2])
@pytest.mark.parametrize("x,y,r", [(20.0, 15.0, 6.0), (3.0, 2.0, 5.5), (47.5, 30.2, 9.0), (25.0, 18.0, 2.0)])
def test_plane_from_disparity(x, y, r, kernel):
    """dispmap_ncc.generate_new_plane_RANSAC / fit_plane_to_points (dispmap_ncc.m:48-91) against the literal SVD / IRLS
    restatement: same plane to rounding (eigenvector of the scatter matrix == right singular vector, the sign cancels)."""
    H, W = 36, 52
    rng = np.random.default_rng(5)
    cc, rr = np.meshgrid(np.arange(1, W + 1), np.arange(1, H + 1))
    disp = 4.0 + 0.11 * cc - 0.07 * rr + rng.normal(0, 0.3, size=(H, W))
    disp[rng.random((H, W)) < 0.05] += 6.0                      # outliers: what the IRLS rounds are for
    plane, prop = builders.plane_from_disparity(disp, x, y, r, kernel, return_proposal=True)
    ref, rprop = _np().generate_new_plane(disp, x, y, r, kernel)
    assert plane[2] == 1.0
    assert np.allclose(plane, ref, rtol=2e-7, atol=1e-9), (plane, ref)
    assert prop.shape == (4, H * W) and np.array_equal(prop, np.repeat(plane[:, None], H * W, axis=1))
    if r >= 5 and kernel == 2:
        # the least-squares plane is close to the generating one, d = -(a x + b y + d0), despite the outliers
        assert abs(-plane[0] - 0.11) < 0.08 and abs(-plane[1] + 0.07) < 0.08
    with pytest.raises(sb._lib.SbError):
        builders.plane_from_disparity(disp, 10.0, 10.0, 0.5, kernel)     # one point within the radius


def test_smooth_weights_from_segments():
    """dispmap_globalstereo.preprocess (:396-401): lambda_h inside a segment, lambda_l across, scaled by the image count."""
    H, W = 13, 17
    rng = np.random.default_rng(2)
    seg = (rng.integers(0, 3, size=(H, W)) + 7).astype(np.uint32)
    got = builders.smooth_weights(seg, 5.0, 0.5, 2.0)
    assert np.array_equal(got, _np().smooth_weights(H, W, seg, 5.0, 0.5, 2.0))
    opts = dict(smoothness_kernel=1, disp_thresh=0.02, lambda_h=5.0, lambda_l=0.5, col_thresh=30.0, improve=0, window=1)
    im0, im1, _ = synth.stereo_pair(H, W, 3, seed=1)
    P = np.zeros((3, 4, 2))
    P[:, :3, 0] = np.eye(3)
    P[:, :3, 1] = np.eye(3)
    P[0, 3, 1] = -0.25
    dm = sb.dispmap_globalstereo([im0, im1], P, [0, 3], 4, opts, segment=seg, rng=np.random.default_rng(0))
    assert np.array_equal(dm.smooth_weights, got)
    w = dm.segpln_wta()
    assert w.shape == (H, W) and set(np.unique(w)).issubset(set(dm.disps) | {0.0})


def test_segpln_proposals_with_injected_segments():
    """dispmap_globalstereo.segpln (dispmap_globalstereo.m:60-201) in the Python mirror: WTA volume on the GPU, per-segment
    RANSAC + least-squares planes as host glue (like the reference), segmentations injected: one plane per segment, finite
    (NaN / Inf -> 1e-100 like :197-200), and the proposal cell feeds binary_fuse_until_convergence."""
    H, W = 40, 60
    im0, im1, _ = synth.stereo_pair(H, W, 6, seed=9)
    opts = dict(smoothness_kernel=1, disp_thresh=0.02, lambda_h=5.0, lambda_l=0.5, col_thresh=30.0, improve=0, window=2)
    P = np.zeros((3, 4, 2))
    P[:, :3, 0] = np.eye(3)
    P[:, :3, 1] = np.eye(3)
    P[0, 3, 1] = -0.25
    dm = sb.dispmap_globalstereo([im0, im1], P, [0, 8], 4, opts, rng=np.random.default_rng(0))
    seg_a = (1 + (np.arange(H)[:, None] // 20) * 3 + (np.arange(W)[None, :] // 20)).astype(np.uint32)     # 2 x 3 blocks
    seg_b = np.ones((H, W), dtype=np.uint32)
    props = dm.segpln([seg_a, seg_b], rng=np.random.default_rng(1))
    assert len(props) == 2 and all(p.shape == (4, H * W) and np.isfinite(p).all() for p in props)
    for seg, p in zip((seg_a, seg_b), props):
        flat = seg.reshape(-1, order="F")
        for a in np.unique(flat):
            assert np.ptp(p[:, flat == a], axis=1).max() == 0          # one plane per segment
    # the proposals are usable: fusing them never raises the energy
    e0 = dm.energy()
    dm.maxiter = 4
    dm.binary_fuse_until_convergence(props, rng=np.random.default_rng(2))
    assert dm.energy() <= e0 * (1 + 1e-12)


@pytest.mark.parametrize("kernel", [1, 2])
def test_binary_fuse_until_convergence_device_loop(kernel):
    """dispmap_super.binary_fuse_until_convergence (dispmap_super.m:85-152): the one-call loop over device-resident
    fields (sb_binary_fuse_until_convergence_grid) follows the per-fusion loop of the class -- same visiting order, same
    energies after every move, same final assignment -- including the reference's bookkeeping (starts at ids(2), stops
    when every proposal has been tried without a change of E)."""
    H, W = 30, 44
    im0, im1, _ = synth.stereo_pair(H, W, 9, seed=5)
    props = []
    for i, d in enumerate((0.0, 2.0, 4.0, 6.0, 8.0)):
        p = np.zeros((4, H * W))
        p[0] = 0.01 * (i - 2)          # slanted planes: a x + b y + c d + d0 = 0
        p[2] = 1
        p[3] = -d - p[0] * W / 2
        props.append(p)
    runs = []
    for device_loop in (False, True):
        dm = sb.dispmap_ncc([im0, im1], np.arange(0, 10, dtype=np.float64), kernel, 40.0, 8.0 if kernel == 1 else 3.0)
        dm.maxiter = 12
        dm.device_loop = device_loop
        n = dm.binary_fuse_until_convergence(props, rng=np.random.default_rng(3))
        assert n == len(dm.fusion_energies)
        runs.append((n, np.array(dm.fusion_energies), dm.assignment.copy(), dm.energy()))
    (n0, e0, a0, f0), (n1, e1, a1, f1) = runs
    assert n0 == n1 and n0 >= 3
    assert np.allclose(e0, e1, rtol=1e-12, atol=0)
    assert np.array_equal(a0, a1) and abs(f0 - f1) <= 1e-12 * abs(f0)
    assert all(e1[i + 1] <= e1[i] * (1 + 1e-12) for i in range(len(e1) - 1)) and e1[-1] < e1[0]


def test_binary_fuse_until_convergence_stops_when_all_visited():
    """Proposals equal to the current assignment change nothing: every proposal is marked after one try and the loop
    ends (dispmap_super.m:136-150) -- ids(1) is skipped by the reference's `iter = iter + 1`, so with n proposals the
    1:n prefix leaves proposal 1 for the random tail."""
    H, W = 12, 15
    N = H * W
    rng = np.random.default_rng(0)
    cur = np.zeros((4, N))
    cur[2] = 1
    cur[3] = -3.0
    un = rng.random(N)
    w = np.ones(2 * ((H - 1) * W + H * (W - 1)))
    ids = np.array([1, 2, 3, 2, 1, 3, 1], dtype=np.int32)
    a, u, E, st = sb.binary_fuse_until_convergence_grid(H, W, 1, [cur, cur, cur], [un, un, un], cur, un, w, 0.5, 20, ids)
    # visits ids(2)=2, ids(3)=3, ids(4)=2 (already marked: skipped), ids(5)=1 -> all marked
    assert st["fusions"] == 3
    assert len(E) == 4 and np.all(E == E[0])
    assert np.array_equal(a, cur) and np.array_equal(u, un)
    with pytest.raises(Exception):
        sb.binary_fuse_until_convergence_grid(H, W, 1, [cur], [un], cur, un, w, 0.5, 5, np.array([1, 2], dtype=np.int32))


def _teddy():
    g = golden("teddy.npz")
    return g["im2"].astype(np.float64), g["im6"].astype(np.float64)


def test_teddy_ncc_pipeline():
    """example_ncc.m:9-46 plumbing on the real Middlebury pair (BASELINE configs[0] shape, 375 x 450): the NCC volume
    (one-pass kernel == general kernel == NumPy restatement on real data), the WTA initial solution, and fronto-parallel
    fusion moves.  There is no MATLAB output to compare with; the anchors are photo-consistency of the WTA
    disparity and monotone fusion energies."""
    import os
    im2, im6 = _teddy()
    H, W, _ = im2.shape
    disps = np.arange(0, 51, dtype=np.float64)               # example_ncc.m:12
    dm = sb.dispmap_ncc([im2, im6], disps, 1, 40.0, 8.0)     # :13-19
    assert dm._vol.info()["one_pass"]
    fast = dm.ncc
    os.environ["SB_NCC_GENERAL"] = "1"
    try:
        gen = builders.ncc_volume(im2, im6, disps[::10], 2)
    finally:
        del os.environ["SB_NCC_GENERAL"]
    cond = _np().ncc_conditioning(im2, im6, disps[::10], 2) > 1e-6
    assert cond.mean() > 0.8
    assert np.all(np.abs(fast[:, :, ::10] - gen)[cond] <= 5e-5)
    ref = _np().compute_ncc(im2, im6, disps[:3], 2)
    c3 = _np().ncc_conditioning(im2, im6, disps[:3], 2) > 1e-6
    assert np.all(np.abs(fast[:, :, :3] - ref)[c3] <= 1e-4 * np.maximum(np.abs(ref[c3]), 1e-2))
    # WTA disparity: warping the right image by it must explain the left image far better than no disparity
    best = dm.best_disp_from_ncc()
    assert np.isfinite(best).all() and best.min() >= -1 and best.max() <= 51
    cols = np.clip(np.round(np.arange(W)[None, :] - best).astype(int), 0, W - 1)
    warped = im6[np.arange(H)[:, None], cols]
    inner = (slice(8, H - 8), slice(60, W - 8))
    mad_w = np.abs(warped - im2)[inner].mean()
    mad_0 = np.abs(im6 - im2)[inner].mean()
    assert mad_w < 0.5 * mad_0, (mad_w, mad_0)
    # fusion moves with fronto-parallel proposals (example_ncc.m:35-46): energy never increases
    e = [dm.energy()]
    for d in range(0, 51, 10):
        prop = np.zeros((4, H * W))
        prop[2] = 1
        prop[3] = -d
        dm.binary_fusion(prop)
        e.append(dm.energy())
    assert all(e[i + 1] <= e[i] * (1 + 1e-12) for i in range(len(e) - 1)), e
    assert e[-1] < e[0]
