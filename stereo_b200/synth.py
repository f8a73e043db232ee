"""Seeded synthetic inputs for the hot path (tests and bench.py).

Everything derives from ``numpy.random.Generator(PCG64(seed))``.  The TRW-S
problems are built the way dispmap_super.simultaneous_fusion builds them
(dispmap_super.m:158-188): L plane proposals per pixel -> unary L x N,
q / qprim L x E through disparitymap_from_assignment (dispmap_super.m:318-328).
"""
from __future__ import annotations

import numpy as np

from .grid import construct_neighborhood, get_points


def disparity_from_planes(planes, points):
    """dispmap_super.disparitymap_from_assignment (dispmap_super.m:318-328):
    planes 4 x M ([a; b; c; d0]), points 2 x M ([x; y]) -> -(a x + b y + d0) / c."""
    return -(planes[0] * points[0] + planes[1] * points[1] + planes[3]) / planes[2]


def random_plane_proposals(H, W, L, rng, slope=0.02, segments=6):
    """L proposals, each a 4 x N plane field.  Proposals are piecewise planar over a random
    rectangular segmentation (like SegPln output, dispmap_globalstereo.m:157-201) with a few
    fronto-parallel ones; disparities stay within roughly [0, 1] (normalised units)."""
    N = H * W
    pts = get_points(H, W)
    props = np.zeros((L, 4, N))
    props[:, 2, :] = 1.0
    for l in range(L):
        if l % 4 == 0:
            props[l, 3, :] = -(l + 0.5) / L  # fronto-parallel level
            continue
        nseg = int(rng.integers(1, segments + 1))
        # random rectangular segmentation
        rcut = np.sort(rng.integers(0, H, size=nseg))
        ccut = np.sort(rng.integers(0, W, size=nseg))
        seg = (np.searchsorted(rcut, pts[1] - 1, side="right") * (nseg + 1)
               + np.searchsorted(ccut, pts[0] - 1, side="right"))
        ids = np.unique(seg)
        a = slope * (rng.random(ids.size) - 0.5) / max(W, 1) * 8
        b = slope * (rng.random(ids.size) - 0.5) / max(H, 1) * 8
        d = rng.random(ids.size)
        lut = {s: i for i, s in enumerate(ids)}
        idx = np.vectorize(lut.get)(seg)
        cx, cy = (W + 1) / 2.0, (H + 1) / 2.0
        props[l, 0, :] = a[idx]
        props[l, 1, :] = b[idx]
        props[l, 3, :] = -(d[idx] - a[idx] * cx - b[idx] * cy) * 1.0 - 0.0
        # disparity at centre = d
        props[l, 3, :] = -(d[idx] + a[idx] * cx + b[idx] * cy)
    return props


def trws_problem(H, W, L, seed=0, kernel=1, tol=None, duplicate_last=True):
    """A complete trws() argument set in MATLAB shapes.

    Returns dict(kernel, unary LxN, connectivity 2xE (1-based), q LxE, qprim LxE, alphas E, tol).
    alphas take the two values of dispmap_globalstereo.m:400-403 (lambda_l, lambda_h times 2)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = H * W
    ind1, ind2 = construct_neighborhood(H, W)
    E = ind1.size
    pts = get_points(H, W)
    props = random_plane_proposals(H, W, L, rng)
    if duplicate_last and L > 2:
        # the last label is the "current assignment": a per-pixel mix of other proposals
        pick = rng.integers(0, L - 1, size=N)
        props[L - 1] = props[pick, :, np.arange(N)].T
    unary = rng.random((L, N)) * np.log(2.0)
    q = np.empty((L, E))
    qprim = np.empty((L, E))
    p2 = pts[:, ind2 - 1]
    for l in range(L):
        q[l] = disparity_from_planes(props[l][:, ind2 - 1], p2)
        qprim[l] = disparity_from_planes(props[l][:, ind1 - 1], p2)
    same = rng.random(E // 2) < 0.8
    w = np.where(same, 108.0, 9.0) * 2.0
    # both directions of a neighbour pair share the segmentation weight
    nV = (H - 1) * W
    nH = H * (W - 1)
    wv = w[:nV]
    wh = w[nV:nV + nH]
    alphas = np.concatenate([wv, wv, wh, wh])
    if tol is None:
        tol = 0.02 if kernel == 1 else 0.02 ** 2
    if kernel == 2:
        alphas = alphas / 0.02
    return dict(kernel=kernel, unary=unary, connectivity=np.stack([ind1, ind2]), q=q, qprim=qprim,
                alphas=alphas, tol=float(tol), planes=props)
