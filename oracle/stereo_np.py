"""NumPy restatement of the MATLAB-side array builders of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may import this module.

PARITY UNPINNED for everything in this file except ``interp2_linear``: the functions below
restate MATLAB code (dispmap_ncc.m, dispmap_super.m, dispmap_globalstereo.m) and there is no
MATLAB / Octave in the build container and no golden vector anywhere in the reference tree
(SURVEY.md 8(c)), so they could only be checked against the MATLAB documentation of conv2 /
interp2 / max / round and against brute-force definitions (tests/test_oracle_np.py).  The
vgg_interp2 gather used by the photo-consistency unary IS pinned: oracle.interp2_linear runs the
compiled reference mex source.

All arrays use MATLAB conventions: images (H, W, 3) float64 with values 0..255, node index
u = r + H*c, planes 4 x M ([a; b; c; d0]), points 2 x M ([x; y] = [column; row], 1-based).
"""
from __future__ import annotations

import numpy as np


# ----------------------------------------------------------------------------- helpers
def _box_same(a, p):
    """conv2(a, ones(2p+1), 'same'): zero-padded box sum (dispmap_ncc.m:128-137)."""
    H, W = a.shape
    pad = np.zeros((H + 2 * p, W + 2 * p), dtype=a.dtype)
    pad[p:p + H, p:p + W] = a
    c = np.cumsum(np.cumsum(pad, axis=0), axis=1)
    c = np.pad(c, ((1, 0), (1, 0)))
    k = 2 * p + 1
    return c[k:, k:] - c[:-k, k:] - c[k:, :-k] + c[:-k, :-k]


def _matlab_round(x):
    return np.floor(np.abs(x) + 0.5) * np.sign(x)


def get_points(H, W):
    """dispmap_super.m:275-278."""
    xx, yy = np.meshgrid(np.arange(1, W + 1, dtype=np.float64), np.arange(1, H + 1, dtype=np.float64))
    return np.stack([xx.T.reshape(-1), yy.T.reshape(-1)])


def construct_neighborhood(H, W):
    """dispmap_super.m:279-302: ind1, ind2 (1-based)."""
    nodenr = np.arange(1, H * W + 1, dtype=np.int64).reshape(W, H).T
    vs, vf = nodenr[:-1, :].T.reshape(-1), nodenr[1:, :].T.reshape(-1)
    hs, hf = nodenr[:, :-1].T.reshape(-1), nodenr[:, 1:].T.reshape(-1)
    return np.concatenate([vs, vf, hs, hf]), np.concatenate([vf, vs, hf, hs])


# ----------------------------------------------------------------------------- a5
def disparity_from_assignment(assignment, points, d_min=0.0, d_step=1.0):
    """dispmap_super.m:318-328 and the override dispmap_globalstereo.m:336-345."""
    a = np.asarray(assignment, dtype=np.float64)
    if np.any(a[2] == 0):
        raise ValueError("Infinite disparity")
    d = -((a[0:2] * points).sum(axis=0) + a[3]) / a[2]
    return (d - d_min) / d_step


# ----------------------------------------------------------------------------- a1
def shifted_image(im1, d):
    """dispmap_ncc.m:145-154: image 2 moved right by d through interp2 on a linspace grid."""
    H, W, C = im1.shape
    imtr = np.zeros((H, W, C))
    c0 = int(np.ceil(d + 1))            # first column of y_span (1-based)
    n = W - c0 + 1
    if n <= 0:
        return imtr
    X = np.linspace(1.0, W - d, n)
    x0 = np.floor(X).astype(np.int64)
    f = X - x0
    x0 = np.clip(x0, 1, W)
    x1 = np.clip(x0 + 1, 1, W)
    for ch in range(C):
        A = im1[:, :, ch]
        imtr[:, c0 - 1:, ch] = A[:, x0 - 1] * (1 - f) + A[:, x1 - 1] * f
    return imtr


def compute_ncc(im0, im1, disparities, patchsize=2):
    """dispmap_ncc.compute_ncc (dispmap_ncc.m:116-198) -> (H, W, D) float64."""
    im0 = np.asarray(im0, dtype=np.float64)
    im1 = np.asarray(im1, dtype=np.float64)
    H, W, C = im0.shape
    assert C == 3
    p = patchsize
    n3 = float((2 * p + 1) ** 2 * 3)
    d = np.asarray(disparities, dtype=np.float64).reshape(-1)
    sR = sum(_box_same(im0[:, :, c], p) for c in range(3))
    sRR = sum(_box_same(im0[:, :, c] ** 2, p) for c in range(3))
    mean_right = sR / n3
    norm_right = np.sqrt((sRR - 2 * mean_right * sR + n3 * mean_right ** 2).astype(np.complex128))
    ncc = np.zeros((H, W, d.size))
    cols = np.arange(1, W + 1)
    for i, di in enumerate(d):
        bnd = cols >= _matlab_round(di + 1)
        tr = shifted_image(im1, di)
        sT = sum(_box_same(tr[:, :, c], p) for c in range(3))
        sTT = sum(_box_same(tr[:, :, c] ** 2, p) for c in range(3))
        sRT = sum(_box_same(im0[:, :, c] * tr[:, :, c], p) for c in range(3))
        mean_tr = sT / n3
        norm_tr = np.sqrt((sTT - 2 * mean_tr * sT + n3 * mean_tr ** 2).astype(np.complex128))
        num = sRT - mean_right * sT - mean_tr * sR + n3 * mean_tr * mean_right
        with np.errstate(divide="ignore", invalid="ignore"):
            v = num / norm_right / norm_tr
        v[~np.isfinite(v)] = 0
        v[:, ~bnd] = 0
        ncc[:, :, i] = v.real
    return ncc


def ncc_conditioning(im0, im1, disparities, patchsize=2):
    """min(var_R, var_T) / max(sum of squares) per output: windows where this is tiny are
    numerically degenerate in the reference itself (0/0 of rounding noise)."""
    im0 = np.asarray(im0, dtype=np.float64)
    im1 = np.asarray(im1, dtype=np.float64)
    p = patchsize
    n3 = float((2 * p + 1) ** 2 * 3)
    d = np.asarray(disparities, dtype=np.float64).reshape(-1)
    sR = sum(_box_same(im0[:, :, c], p) for c in range(3))
    sRR = sum(_box_same(im0[:, :, c] ** 2, p) for c in range(3))
    vR = sRR - sR ** 2 / n3
    out = np.zeros(im0.shape[:2] + (d.size,))
    for i, di in enumerate(d):
        tr = shifted_image(im1, di)
        sT = sum(_box_same(tr[:, :, c], p) for c in range(3))
        sTT = sum(_box_same(tr[:, :, c] ** 2, p) for c in range(3))
        vT = sTT - sT ** 2 / n3
        out[:, :, i] = np.minimum(vR / np.maximum(sRR, 1), vT / np.maximum(sTT, 1))
    return out


# ----------------------------------------------------------------------------- a2
def interpolate_ncc(ncc, disparities, t2, y2, okdepth):
    """dispmap_ncc.m:246-275 (t2 1-based)."""
    d = np.asarray(disparities, dtype=np.float64).reshape(-1)
    H, W, _ = ncc.shape
    d2 = d[t2 - 1]
    t1 = np.where(okdepth, t2 - 1, t2)
    t3 = np.where(okdepth, t2 + 1, t2)
    d1, d3 = d[t1 - 1], d[t3 - 1]
    rr, cc = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    y1 = ncc[rr, cc, t1 - 1]
    y3 = ncc[rr, cc, t3 - 1]
    with np.errstate(divide="ignore", invalid="ignore"):
        a = y1 / (d1 - d2) / (d1 - d3)
        b = y2 / (d2 - d1) / (d2 - d3)
        c = y3 / (d3 - d1) / (d3 - d2)
        r = a + b + c
        p = -(a * (d2 + d3) + b * (d1 + d3) + c * (d1 + d2))
        q = a * d2 * d3 + b * d1 * d3 + c * d1 * d2
    return r, p, q, d2


def best_disp_from_ncc(ncc, disparities):
    """dispmap_ncc.m:208-221: WTA level + parabola refinement."""
    D = ncc.shape[2]
    t2 = np.argmax(ncc, axis=2) + 1          # MATLAB max: first occurrence
    y2 = np.max(ncc, axis=2)
    okdepth = (t2 < D) & (t2 > 1)
    r, p, q, d2 = interpolate_ncc(ncc, disparities, t2, y2, okdepth)
    with np.errstate(divide="ignore", invalid="ignore"):
        best = -p / r / 2
    best[~okdepth] = d2[~okdepth]
    return best


def sample_ncc_from_disp(ncc, disparities, disps):
    """dispmap_ncc.m:222-245."""
    d = np.asarray(disparities, dtype=np.float64).reshape(-1)
    H, W, D = ncc.shape
    x = np.asarray(disps, dtype=np.float64).reshape(W, H).T if np.ndim(disps) == 1 else np.asarray(disps, dtype=np.float64)
    t2 = np.ones((H, W), dtype=np.int64)
    smallest = np.abs(x - d[0])
    y2 = np.ones((H, W))
    for i in range(D):
        nd = np.abs(x - d[i])
        m = nd <= smallest
        t2[m] = i + 1
        y2[m] = ncc[:, :, i][m]
        smallest[m] = nd[m]
    okdepth = (t2 < D) & (t2 > 1)
    good = (x <= d.max()) & (x >= d.min())
    r, p, q, _ = interpolate_ncc(ncc, d, t2, y2, okdepth)
    with np.errstate(invalid="ignore"):
        nccs = r * x ** 2 + p * x + q
    nccs[t2 == 1] = ncc[:, :, 0][t2 == 1]
    nccs[t2 == D] = ncc[:, :, -1][t2 == D]
    nccs[~good] = -1e6
    return nccs


def ncc_unary_cost(ncc, disparities, assignment, unary_weight, H, W):
    """dispmap_ncc.unary_cost (dispmap_ncc.m:107-115) -> N vector."""
    disps = disparity_from_assignment(assignment, get_points(H, W))
    nccs = sample_ncc_from_disp(ncc, disparities, disps)
    return unary_weight * (1 - nccs.T.reshape(-1))


# ----------------------------------------------------------------------------- a3
def photo_unary_cost(images, P2, assignment, d_min, d_step, col_thresh, interp2):
    """dispmap_globalstereo.unary_cost (dispmap_globalstereo.m:355-375) with ephoto (:405).
    ``P2`` is self.P(:,:,2), i.e. the TRANSPOSE (4 x 3) of the user's 3 x 4 camera matrix
    (dispmap_globalstereo.m:42); ``interp2(A, X, Y, oobv)`` is the vgg_interp2 linear gather."""
    im0 = np.asarray(images[0], dtype=np.float64)
    H, W = im0.shape[:2]
    colors = im0.shape[2] if im0.ndim == 3 else 1
    nd = disparity_from_assignment(assignment, get_points(H, W), d_min, d_step)
    disp = d_step * (nd + d_min)                    # literal :356
    pts = get_points(H, W)
    WC = np.stack([pts[0], pts[1], np.ones(H * W), disp], axis=1)
    T = WC @ np.asarray(P2, dtype=np.float64)
    Nn = 1.0 / T[:, 2]
    X, Y = T[:, 0] * Nn, T[:, 1] * Nn
    R = im0.reshape(H, W, -1).transpose(1, 0, 2).reshape(H * W, -1)
    M = interp2(np.asarray(images[1], dtype=np.float64), X, Y, -1000.0) - R
    return np.log(2.0) - np.log(np.exp((M ** 2).sum(axis=1) * (-1.0 / (col_thresh * colors))) + 1.0)


# ----------------------------------------------------------------------------- f2 / f4
def segpln_wta(images, P, disps, window, col_thresh, interp2, min_corr=0.07):
    """The window-matching volume of dispmap_globalstereo.segpln (dispmap_globalstereo.m:83-117), literally: per image
    and disparity the photo cost at the projected point, conv2(filt, filt', ., 'valid') with the 1 x (2 w + 1) average
    filter, normalisation by X(1), first maximum over the levels, score < 0.07 -> 0, symmetric padding.
    ``P``: 3 x 4 x n camera matrices; returns (corr H x W, volume (H-2w) x (W-2w) x D of normalised scores)."""
    ims = [np.asarray(im, dtype=np.float64) for im in images]
    ims = [im[:, :, None] if im.ndim == 2 else im for im in ims]
    H, W, C = ims[0].shape
    P = np.asarray(P, dtype=np.float64).reshape(3, 4, -1)
    disps = np.asarray(disps, dtype=np.float64).reshape(-1)
    w = int(window)
    Rvec = ims[0].transpose(1, 0, 2).reshape(H * W, C)                     # reshape(double(R), [], sz(3)) (:74)
    pts = get_points(H, W)
    WC = np.stack([pts[0], pts[1], np.ones(H * W)], axis=1)                # :75-78

    def ephoto(F):
        return np.log(2.0) - np.log(np.exp((F ** 2).sum(axis=1) * (-1.0 / (col_thresh * C))) + 1.0)

    f = 1.0 / (2 * w + 1)
    corr = np.zeros((H - 2 * w, W - 2 * w, disps.size))
    for a in range(len(ims)):
        X = WC @ P[:, :3, a].T                                             # :89
        P_ = P[:, 3, a]
        for b, dv in enumerate(disps):
            d = dv * P_                                                    # :95
            Z = 1.0 / (X[:, 2] + d[2])
            Y = interp2(ims[a], (X[:, 0] + d[0]) * Z, (X[:, 1] + d[1]) * Z, -1000.0)
            Y = ephoto(Y - Rvec).reshape(W, H).T                           # reshape(Y, sz(1:2)) (:104)
            hb = sum(Y[:, k:k + W - 2 * w] * f for k in range(2 * w + 1))
            vb = sum(hb[k:k + H - 2 * w, :] * f for k in range(2 * w + 1))
            corr[:, :, b] += vb
    X1 = ephoto((-1000.0 - Rvec)[:1])[0] * len(ims)                        # :110
    vol = (X1 - corr) / X1
    idx = vol.argmax(axis=2)                                               # first maximum, like max(., [], 3)
    best = np.take_along_axis(vol, idx[:, :, None], axis=2)[:, :, 0]
    out = disps[idx]
    out[best < min_corr] = 0
    return np.pad(out, w, mode="symmetric"), vol


def fit_plane_to_points(points, kernel):
    """dispmap_ncc.fit_plane_to_points (dispmap_ncc.m:67-91), literally (SVD, 20 IRLS rounds for kernel 1)."""
    points = np.asarray(points, dtype=np.float64)
    c = points.mean(axis=1, keepdims=True)
    cost = -(points - c).T
    p = np.zeros(4)
    if kernel == 1:
        w = np.ones(cost.shape[0])
        for _ in range(20):
            V = np.linalg.svd(w[:, None] * cost, full_matrices=False)[2].T
            p[:3] = V[:, -1]
            w = np.sqrt(np.abs(cost @ V[:, -1]))
    else:
        V = np.linalg.svd(cost, full_matrices=False)[2].T
        p[:3] = V[:, -1]
    p[3] = -(p[:3] @ points[:3].mean(axis=1))
    return p / p[2]


def generate_new_plane(best_disp, x, y, r, kernel):
    """dispmap_ncc.generate_new_plane_RANSAC (dispmap_ncc.m:48-66) -> (plane 4, proposal 4 x N)."""
    H, W = best_disp.shape
    pts = get_points(H, W)
    ids = np.sqrt((pts[0] - x) ** 2 + (pts[1] - y) ** 2) < r
    p = fit_plane_to_points(np.vstack([pts[:, ids], best_disp.reshape(-1, order="F")[ids][None]]), kernel)
    return p, np.repeat(p[:, None], H * W, axis=1)


def smooth_weights(H, W, segment, lambda_h, lambda_l, scale):
    """dispmap_globalstereo.m:396-400."""
    ind1, ind2 = construct_neighborhood(H, W)
    seg = np.asarray(segment).reshape(-1, order="F")
    same = seg[ind1 - 1] == seg[ind2 - 1]
    return (same * lambda_h + (~same) * lambda_l) * scale


# ----------------------------------------------------------------------------- a6 / a7
def pairwise_cost(kernel, weights, tol, p, q):
    """dispmap_super.m:226-235."""
    if kernel == 1:
        return weights * np.minimum(np.abs(p - q), tol)
    if kernel == 2:
        return weights * np.minimum((p - q) ** 2, tol)
    raise ValueError("Unkown kernel type")


def all_pairwise_costs(H, W, kernel, weights, tol, assignment, proposal, d_min=0.0, d_step=1.0):
    """dispmap_super.m:236-262 -> E00, E01, E10, E11 (each E)."""
    ind1, ind2 = construct_neighborhood(H, W)
    pts = get_points(H, W)[:, ind2 - 1]
    dd = lambda a, ind: disparity_from_assignment(a[:, ind - 1], pts, d_min, d_step)  # noqa: E731
    q, qprim = dd(assignment, ind2), dd(assignment, ind1)
    nq, nqprim = dd(proposal, ind2), dd(proposal, ind1)
    E00 = pairwise_cost(kernel, weights, tol, q, qprim)
    E11 = pairwise_cost(kernel, weights, tol, nq, nqprim)
    E10 = pairwise_cost(kernel, weights, tol, q, nqprim)
    E01 = pairwise_cost(kernel, weights, tol, nq, qprim)
    return E00, E01, E10, E11


def fusion_positions(H, W, proposals, d_min=0.0, d_step=1.0):
    """dispmap_super.m:170-183: q, qprim (L x E) of a list of 4 x N proposals."""
    ind1, ind2 = construct_neighborhood(H, W)
    pts = get_points(H, W)[:, ind2 - 1]
    q = np.stack([disparity_from_assignment(p[:, ind2 - 1], pts, d_min, d_step) for p in proposals])
    qprim = np.stack([disparity_from_assignment(p[:, ind1 - 1], pts, d_min, d_step) for p in proposals])
    return q, qprim


def energy(H, W, kernel, weights, tol, unary, assignment, d_min=0.0, d_step=1.0):
    """dispmap_super.update_energy (dispmap_super.m:263-274)."""
    ind1, ind2 = construct_neighborhood(H, W)
    pts = get_points(H, W)[:, ind2 - 1]
    q = disparity_from_assignment(assignment[:, ind2 - 1], pts, d_min, d_step)
    qprim = disparity_from_assignment(assignment[:, ind1 - 1], pts, d_min, d_step)
    return float(np.sum(unary) + np.sum(pairwise_cost(kernel, weights, tol, q, qprim)))
