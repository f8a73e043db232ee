"""ctypes wrappers of the cost-volume / unary / pairwise builders of libstereo_b200.so
(include/stereo_b200.h).  Arrays use MATLAB shapes; everything is computed on the GPU."""
from __future__ import annotations

import ctypes

import numpy as np

from ._lib import _dp, check, lib


def _f(a):
    return np.asfortranarray(a, dtype=np.float64)


def _p(a):
    return a.ctypes.data_as(_dp)


def ncc_volume(im0, im1, disparities, patchsize=2):
    """dispmap_ncc.compute_ncc (dispmap_ncc.m:116-198) -> (H, W, D)."""
    im0, im1 = _f(im0), _f(im1)
    H, W, C = im0.shape
    assert im1.shape == im0.shape
    d = _f(np.asarray(disparities).reshape(-1))
    out = np.zeros((H, W, d.size), dtype=np.float64, order="F")
    check(lib().sb_ncc_volume(H, W, C, _p(im0), _p(im1), d.size, _p(d), int(patchsize), _p(out)))
    return out


class NccVolume:
    """The NCC volume of dispmap_ncc.compute_ncc kept ON the device (``sb_ncc_vol_*``): ``best_disp`` and ``sample``
    are dispmap_ncc.best_disp_from_ncc / sample_ncc_from_disp (dispmap_ncc.m:208-245) on the resident volume;
    ``get()`` copies it out as (H, W, D) doubles only when asked."""

    def __init__(self, im0, im1, disparities, patchsize=2):
        import ctypes
        im0, im1 = _f(im0), _f(im1)
        self.H, self.W, C = im0.shape
        assert im1.shape == im0.shape
        self.disparities = _f(np.asarray(disparities).reshape(-1))
        self.D = self.disparities.size
        self._h = ctypes.c_void_p()
        check(lib().sb_ncc_vol_create(self.H, self.W, C, _p(im0), _p(im1), self.D, _p(self.disparities), int(patchsize),
                                      ctypes.byref(self._h)))

    def info(self):
        import ctypes
        out = (ctypes.c_double * 3)()
        check(lib().sb_ncc_vol_info(self._h, out))
        return dict(kernel_ms=out[0], one_pass=bool(out[1]), bytes=int(out[2]))

    def get(self):
        out = np.zeros((self.H, self.W, self.D), dtype=np.float64, order="F")
        check(lib().sb_ncc_vol_get(self._h, _p(out)))
        return out

    def best_disp(self):
        out = np.zeros((self.H, self.W), dtype=np.float64, order="F")
        check(lib().sb_ncc_vol_best_disp(self._h, _p(out)))
        return out

    def sample(self, disps, unary_weight=1.0, as_unary=False):
        x = _f(np.asarray(disps).reshape(-1, order="F"))
        assert x.size == self.H * self.W
        out = np.zeros((self.H, self.W), dtype=np.float64, order="F")
        check(lib().sb_ncc_vol_sample(self._h, _p(x), float(unary_weight), int(bool(as_unary)), _p(out)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_ncc_vol_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def ncc_best_disp(ncc, disparities):
    """dispmap_ncc.best_disp_from_ncc (dispmap_ncc.m:208-221) -> (H, W)."""
    ncc = _f(ncc)
    H, W, D = ncc.shape
    d = _f(np.asarray(disparities).reshape(-1))
    out = np.zeros((H, W), dtype=np.float64, order="F")
    check(lib().sb_ncc_best_disp(H, W, D, _p(ncc), _p(d), _p(out)))
    return out


def ncc_sample(ncc, disparities, disps, unary_weight=1.0, as_unary=False):
    """dispmap_ncc.sample_ncc_from_disp (dispmap_ncc.m:222-245); as_unary: dispmap_ncc.unary_cost."""
    ncc = _f(ncc)
    H, W, D = ncc.shape
    d = _f(np.asarray(disparities).reshape(-1))
    x = _f(np.asarray(disps).reshape(-1, order="F"))
    out = np.zeros((H, W), dtype=np.float64, order="F")
    check(lib().sb_ncc_sample(H, W, D, _p(ncc), _p(d), _p(x), float(unary_weight), int(bool(as_unary)), _p(out)))
    return out


def plane_disparity(planes, points, d_min=0.0, d_step=1.0):
    """disparitymap_from_assignment (dispmap_super.m:318-328 / dispmap_globalstereo.m:336-345)."""
    planes, points = _f(planes), _f(points)
    M = planes.shape[1]
    assert planes.shape[0] == 4 and points.shape == (2, M)
    out = np.zeros(M, dtype=np.float64)
    check(lib().sb_plane_disparity(M, _p(planes), _p(points), float(d_min), float(d_step), _p(out)))
    return out


def interp2_linear(A, X, Y, oobv):
    """vgg_interp2(A, X, Y, 'linear', oobv) (vgg_interp2.cxx:246-322) -> (n, col)."""
    A = _f(A)
    if A.ndim == 2:
        A = A[:, :, None]
    h, w, col = A.shape
    X, Y = _f(np.asarray(X).reshape(-1)), _f(np.asarray(Y).reshape(-1))
    out = np.zeros((X.size, col), dtype=np.float64, order="F")
    check(lib().sb_interp2_linear(_p(A), h, w, col, _p(X), _p(Y), X.size, float(oobv), _p(out)))
    return out


def photo_unary(im0, im1, P2, planes, d_min, d_step, col_thresh):
    """dispmap_globalstereo.unary_cost (dispmap_globalstereo.m:355-375,405) -> N."""
    im0, im1 = _f(im0), _f(im1)
    if im0.ndim == 2:
        im0, im1 = im0[:, :, None], im1[:, :, None]
    H, W, C = im0.shape
    P2 = _f(P2)
    assert P2.shape == (4, 3)
    planes = _f(planes)
    out = np.zeros(H * W, dtype=np.float64)
    check(lib().sb_photo_unary(H, W, C, _p(im0), _p(im1), _p(P2), _p(planes), float(d_min), float(d_step),
                               float(col_thresh), _p(out)))
    return out


def segpln_wta(images, P, disps, window, col_thresh, min_corr=0.07, return_score=False):
    """The window-matching volume of dispmap_globalstereo.segpln (dispmap_globalstereo.m:83-117): winner-takes-all
    disparity per pixel (H x W; 0 where the best match scores below ``min_corr``).  images: list of H x W x C arrays,
    images[0] the reference; P: 3 x 4 x n camera matrices (the constructor's argument); disps: self.disps."""
    ims = [_f(np.asarray(im, dtype=np.float64)) for im in images]
    if ims[0].ndim == 2:
        ims = [im[:, :, None] for im in ims]
    H, W, C = ims[0].shape
    n = len(ims)
    stack = np.ascontiguousarray(np.stack([im.reshape(-1, order="F") for im in ims]))       # n x (H W C), MATLAB order
    P = np.asarray(P, dtype=np.float64)
    if P.ndim == 2:
        P = P[:, :, None]
    assert P.shape == (3, 4, n)
    Pm = np.ascontiguousarray(np.stack([P[:, :, a].reshape(-1, order="F") for a in range(n)]))
    disps = np.ascontiguousarray(np.asarray(disps, dtype=np.float64).reshape(-1))
    corr = np.zeros((H, W), dtype=np.float64, order="F")
    score = np.zeros((H - 2 * window, W - 2 * window), dtype=np.float64, order="F")
    check(lib().sb_segpln_wta(H, W, C, n, _p(stack), _p(Pm), disps.size, _p(disps), int(window), float(col_thresh),
                              float(min_corr), _p(corr), _p(score)))
    return (corr, score) if return_score else corr


def plane_from_disparity(disp, x, y, r, kernel, return_proposal=False):
    """dispmap_ncc.generate_new_plane_RANSAC + fit_plane_to_points (dispmap_ncc.m:48-91): the plane [a; b; 1; d0] through
    the points (col, row, disp) within radius r of (x, y); with ``return_proposal`` also the 4 x N field repmat(p, [1 N])."""
    d = np.asfortranarray(np.asarray(disp, dtype=np.float64))
    H, W = d.shape
    plane = np.zeros(4, dtype=np.float64)
    npts = ctypes.c_double(0)
    prop = np.zeros((4, H * W), dtype=np.float64, order="F") if return_proposal else None
    check(lib().sb_plane_from_disparity(H, W, ctypes.c_void_p(d.ctypes.data), float(x), float(y), float(r), int(kernel), 0, _p(plane),
                                        ctypes.c_void_p(prop.ctypes.data) if return_proposal else None, ctypes.byref(npts)))
    return (plane, prop) if return_proposal else plane


def smooth_weights(segment, lambda_h, lambda_l, scale=1.0):
    """The smoothness weights of dispmap_globalstereo.preprocess (dispmap_globalstereo.m:396-401) from a segment label
    image (H x W integers): scale * lambda_h inside a segment, scale * lambda_l across a boundary, E values in
    construct_neighborhood order."""
    seg = np.asfortranarray(np.asarray(segment).astype(np.uint32))
    H, W = seg.shape
    E = 2 * ((H - 1) * W + H * (W - 1))
    out = np.zeros(E, dtype=np.float64)
    if E:
        check(lib().sb_smooth_weights(H, W, seg.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), float(lambda_h), float(lambda_l),
                                      float(scale), _p(out)))
    return out


def pairwise_tables(H, W, kernel, assignment, proposal, weights, tol, d_min=0.0, d_step=1.0):
    """dispmap_super.all_pairwise_costs (dispmap_super.m:236-262) -> E00, E01, E10, E11
    (E00 only when proposal is None)."""
    E = 2 * ((H - 1) * W + H * (W - 1))
    a = _f(assignment)
    w = _f(np.asarray(weights).reshape(-1))
    assert a.shape == (4, H * W) and w.size == E
    E00 = np.zeros(E)
    if proposal is None:
        check(lib().sb_pairwise_tables(H, W, int(kernel), _p(a), None, _p(w), float(tol), float(d_min), float(d_step),
                                       _p(E00), None, None, None))
        return E00
    b = _f(proposal)
    E01, E10, E11 = np.zeros(E), np.zeros(E), np.zeros(E)
    check(lib().sb_pairwise_tables(H, W, int(kernel), _p(a), _p(b), _p(w), float(tol), float(d_min), float(d_step),
                                   _p(E00), _p(E01), _p(E10), _p(E11)))
    return E00, E01, E10, E11


def fusion_positions(H, W, proposals, d_min=0.0, d_step=1.0):
    """q, qprim (L x E) of dispmap_super.simultaneous_fusion (dispmap_super.m:170-183)."""
    L = len(proposals)
    E = 2 * ((H - 1) * W + H * (W - 1))
    stack = np.empty((L, 4 * H * W), dtype=np.float64)
    for l, pr in enumerate(proposals):
        stack[l] = _f(pr).reshape(-1, order="F")
    q = np.zeros((L, E), dtype=np.float64, order="F")
    qp = np.zeros((L, E), dtype=np.float64, order="F")
    check(lib().sb_fusion_positions(H, W, L, _p(stack), float(d_min), float(d_step), _p(q), _p(qp)))
    return q, qp


def energy(H, W, kernel, unary, assignment, weights, tol, d_min=0.0, d_step=1.0):
    """dispmap_super.update_energy (dispmap_super.m:263-274)."""
    u = _f(np.asarray(unary).reshape(-1))
    a = _f(assignment)
    w = _f(np.asarray(weights).reshape(-1))
    e = ctypes.c_double()
    check(lib().sb_energy(H, W, int(kernel), _p(u), _p(a), _p(w), float(tol), float(d_min), float(d_step),
                          ctypes.byref(e)))
    return e.value
