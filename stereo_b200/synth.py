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


def stereo_pair(H, W, max_disp, seed=0, noise=2.0):
    """Synthetic rectified pair: image 1 = band-limited random RGB texture
    quantised to 8 bits, ground-truth disparity = a few random planes over a rectangular
    segmentation in [0, max_disp], image 2 = image 1 warped by -disparity (linear) + noise.
    Returns (im0, im1) float64 H x W x 3 with integer values 0..255, and the disparity."""
    rng = np.random.Generator(np.random.PCG64(seed))
    base = rng.random((H // 4 + 3, W // 4 + 3, 3))
    yy = np.linspace(0, base.shape[0] - 1.001, H)
    xx = np.linspace(0, base.shape[1] - 1.001, W)
    y0, x0 = np.floor(yy).astype(int), np.floor(xx).astype(int)
    fy, fx = (yy - y0)[:, None, None], (xx - x0)[None, :, None]
    tex = (base[y0][:, x0] * (1 - fy) * (1 - fx) + base[y0 + 1][:, x0] * fy * (1 - fx)
           + base[y0][:, x0 + 1] * (1 - fy) * fx + base[y0 + 1][:, x0 + 1] * fy * fx)
    tex = 0.6 * tex + 0.4 * rng.random((H, W, 3))
    im0 = np.floor(tex * 255.999)
    disp = np.zeros((H, W))
    rows = np.arange(H)[:, None]
    cols = np.arange(W)[None, :]
    for _ in range(5):
        r0, r1 = np.sort(rng.integers(0, H, 2))
        c0, c1 = np.sort(rng.integers(0, W, 2))
        a, b = (rng.random(2) - 0.5) * 0.1
        d0 = rng.random() * max_disp
        m = (rows >= r0) & (rows <= r1) & (cols >= c0) & (cols <= c1)
        disp = np.where(m, np.clip(d0 + a * (cols - c0) + b * (rows - r0), 0, max_disp), disp)
    # image 2 (x - d) = image 1 (x)  =>  im1[:, c] = im0 sampled at c + disp (approximately)
    src = np.clip(cols + disp, 0, W - 1)
    s0 = np.floor(src).astype(int)
    f = (src - s0)[:, :, None]
    s1 = np.minimum(s0 + 1, W - 1)
    im1 = im0[rows, s0] * (1 - f) + im0[rows, s1] * f + rng.normal(0, noise, (H, W, 3))
    im1 = np.clip(np.round(im1), 0, 255)
    return im0, im1, disp


def rd_problem(H, W, seed=0, kernel=1, mode="stereo", tol=0.02):
    """A complete rd() argument set in MATLAB shapes (dispmap_super.binary_fusion,
    dispmap_super.m:61-84): U0, U1 (N), E00..E11 (E), connectivity 2 x E (1-based).

    mode "stereo": tables from two plane fields through the truncated pairwise cost
    (mostly submodular, exact ties where both fields agree); "frustrated": random tables
    with many non-submodular terms (hard part 3)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    N = H * W
    ind1, ind2 = construct_neighborhood(H, W)
    E = ind1.size
    if mode == "frustrated":
        U0, U1 = rng.random(N), rng.random(N)
        T = rng.random((4, E)) * 2
        return dict(U0=U0, U1=U1, E00=T[0], E01=T[1], E10=T[2], E11=T[3], connectivity=np.stack([ind1, ind2]))
    pts = get_points(H, W)
    props = random_plane_proposals(H, W, 3, rng)
    cur, new = props[1].copy(), props[2].copy()
    same = rng.random(N) < 0.15               # pixels where the proposal equals the current plane
    new[:, same] = cur[:, same]
    U0 = rng.random(N) * np.log(2.0)
    U1 = rng.random(N) * np.log(2.0)
    U1[same] = U0[same]
    p2 = pts[:, ind2 - 1]
    q, qp = disparity_from_planes(cur[:, ind2 - 1], p2), disparity_from_planes(cur[:, ind1 - 1], p2)
    nq, nqp = disparity_from_planes(new[:, ind2 - 1], p2), disparity_from_planes(new[:, ind1 - 1], p2)
    nV, nH = (H - 1) * W, H * (W - 1)
    w = np.where(rng.random(nV + nH) < 0.8, 108.0, 9.0) * 2.0
    w = np.concatenate([w[:nV], w[:nV], w[nV:], w[nV:]])
    if kernel == 2:
        w, tol = w / tol, tol ** 2
    pc = (lambda a, b: w * np.minimum(np.abs(a - b), tol)) if kernel == 1 else (lambda a, b: w * np.minimum((a - b) ** 2, tol))
    return dict(U0=U0, U1=U1, E00=pc(q, qp), E01=pc(nq, qp), E10=pc(q, nqp), E11=pc(nq, nqp),
                connectivity=np.stack([ind1, ind2]), cur=cur, new=new, weights=w, tol=tol)


# ---------------------------------------------------------------------------
# numpy mirror of the ON-DEVICE synthetic generator (csrc/gtrws_solve.cu: gsynth_kernel), used
# where no GPU code may run (bench.py --impl reference).  Same hashes, same float32 formulas; the
# device may contract a*b+c into an FMA, so own disparities can differ in the last bit.
def _hash32(x):
    x = np.asarray(x, dtype=np.uint64) & 0xFFFFFFFF
    x ^= x >> 16
    x = (x * 0x7FEB352D) & 0xFFFFFFFF
    x ^= x >> 15
    x = (x * 0x846CA68B) & 0xFFFFFFFF
    x ^= x >> 16
    return x


def _hash3(seed, a, b, c):
    m = 0xFFFFFFFF
    a, b, c = (np.asarray(v, dtype=np.uint64) for v in (a, b, c))
    h = _hash32((np.uint64(seed) ^ ((a * 0x9E3779B9) & m)) & m)
    h = _hash32((h ^ ((b * 0x85EBCA6B) & m)) & m)
    return _hash32((h ^ ((c * 0xC2B2AE35) & m)) & m)


def _u01(h):
    return (np.asarray(h, dtype=np.uint64) >> 8).astype(np.float32) * np.float32(1.0 / 16777216.0)


def grid_synth_np(seed, H, W, L, kernel=1, scene=None, offset=(0, 0)):
    """(unary, own, gx, gy) as L x N float32-valued float64 arrays (MATLAB node order) and alphas (E) of the
    h x w window at `offset` of the scene; see TrwsGrid.synth."""
    Hs, Ws = scene if scene is not None else (H, W)
    seed = int(seed)
    seed32 = (seed ^ (seed >> 32)) & 0xFFFFFFFF
    rr, cc = np.meshgrid(np.arange(H) + offset[0], np.arange(W) + offset[1], indexing="ij")
    r = rr.T.reshape(-1).astype(np.int64)      # MATLAB node order u = r + H c
    c = cc.T.reshape(-1).astype(np.int64)
    f32 = np.float32

    def plane_of(lab):
        lab = np.broadcast_to(np.asarray(lab, dtype=np.int64), r.shape)
        hs = _hash3(seed32, 0x51, lab, 0)
        cw = 16 + (hs & 127).astype(np.int64)
        ch = 16 + ((hs >> 8) & 127).astype(np.int64)
        cx, cy = c // cw, r // ch
        hc = _hash3(seed32, (0x52 + lab) & 0xFFFFFFFF, cx, cy)
        gx = (_u01(_hash32(hc ^ 1)) - f32(0.5)) * f32(0.16) / f32(Ws)
        gy = (_u01(_hash32(hc ^ 2)) - f32(0.5)) * f32(0.16) / f32(Hs)
        d = _u01(_hash32(hc ^ 3))
        dx = c.astype(f32) - (cx.astype(f32) + f32(0.5)) * cw.astype(f32)
        dy = r.astype(f32) - (cy.astype(f32) + f32(0.5)) * ch.astype(f32)
        own = (d + gx * dx + gy * dy).astype(f32)
        fp = (lab % 4) == 0
        own = np.where(fp, ((lab.astype(f32) + f32(0.5)) / f32(L)).astype(f32), own)
        return own, np.where(fp, f32(0), gx).astype(f32), np.where(fp, f32(0), gy).astype(f32)

    N = H * W
    unary = np.empty((L, N))
    own = np.empty((L, N))
    gx = np.empty((L, N))
    gy = np.empty((L, N))
    for l in range(L):
        if l == L - 1 and L > 2:
            pick = (_hash3(seed32, 0x53, r, c) % np.uint64(L - 1)).astype(np.int64)
            o, a, b = plane_of(pick)
        else:
            o, a, b = plane_of(l)
        own[l], gx[l], gy[l] = o, a, b
        unary[l] = _u01(_hash3(seed32, (0x54 + l) & 0xFFFFFFFF, r, c)) * f32(0.69314718)
    # weights: pair (node, down) -> l = 0 hash, (node, right) -> l = 1 hash; reference term order
    def wt(which, rr_, cc_):
        w = np.where(_u01(_hash3(seed32, 0x55 + which, rr_, cc_)) < f32(0.8), f32(108.0), f32(9.0)) * f32(2.0)
        if kernel == 2:
            w = w / f32(0.02)
        return w.astype(np.float64)
    R, C = np.meshgrid(np.arange(H - 1) + offset[0], np.arange(W) + offset[1], indexing="ij")
    wv = wt(0, R.T.reshape(-1), C.T.reshape(-1))
    R, C = np.meshgrid(np.arange(H) + offset[0], np.arange(W - 1) + offset[1], indexing="ij")
    wh = wt(1, R.T.reshape(-1), C.T.reshape(-1))
    alphas = np.concatenate([wv, wv, wh, wh])
    return unary, own, gx, gy, alphas
