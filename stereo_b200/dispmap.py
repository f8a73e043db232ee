"""Host-side mirror of the reference's model classes dispmap_super / dispmap_ncc /
dispmap_globalstereo (same property and method names, argument order and error behaviour), over
the C ABI of libstereo_b200.so.  Arrays keep MATLAB shapes: ``assignment`` is 4 x N (one plane
[a; b; c; d0] per pixel, node u = r + H*c), proposals are lists ("cell arrays") of 4 x N arrays.
Nothing is computed here: every array-producing method forwards to a CUDA entry point.
"""
from __future__ import annotations

import numpy as np

from . import builders
from .grid import construct_neighborhood, get_points
from .solvers import binary_fuse_until_convergence_grid, binary_fusion_grid, rd, trws


class dispmap_super:
    """dispmap_super.m:3-329."""

    def __init__(self, images, kernel):
        # dispmap_super.m:24-36
        self.images = [np.asarray(im, dtype=np.float64) for im in images]
        self.sz = tuple(self.images[0].shape[:2])
        self._kernel = kernel
        self.maxiter = 1000              # dispmap_super.m:9
        self._max_relgap = 1e-4          # :10
        self._improve = False            # :13
        self.grid_native = True          # binary_fusion through sb_binary_fusion_grid (tables built on the device)
        self.device_loop = True          # binary_fuse_until_convergence as one call (fields stay on the device)
        self.fusion_energies = None
        self.last_fusion_stats = None
        self._assignment = None
        self.stored_energy = np.inf
        ind1, ind2 = construct_neighborhood(*self.sz)
        self.neighborhood = dict(ind1=ind1, ind2=ind2)
        self.smooth_weights = np.ones(ind1.size)
        self.d_min, self.d_step = 0.0, 1.0   # disparity normalisation (identity here)

    # ---- properties with the reference's validating setters (dispmap_super.m:39-56)
    @property
    def max_relgap(self):
        return self._max_relgap

    @max_relgap.setter
    def max_relgap(self, v):
        if v < 0:
            raise ValueError("Maximum relative gap must be non-negative")
        self._max_relgap = v

    @property
    def improve(self):
        return self._improve

    @improve.setter
    def improve(self, v):
        self._improve = bool(v)

    @property
    def assignment(self):
        return self._assignment

    @assignment.setter
    def assignment(self, a):
        self._assignment = np.asarray(a, dtype=np.float64)
        self.update_energy()

    @property
    def smoothness_kernel(self):
        return self._kernel

    @smoothness_kernel.setter
    def smoothness_kernel(self, k):
        self._kernel = k
        self.update_energy()

    def energy(self):
        return self.stored_energy

    # ---- fusion moves
    def binary_fusion(self, proposal):
        """dispmap_super.m:61-84: one QPBO fusion move."""
        proposal = np.asarray(proposal, dtype=np.float64)
        if proposal.shape != self._assignment.shape:
            raise ValueError("Binary fusion: Proposals is of wrong size")
        U0 = self.unary_cost(self._assignment)
        U1 = self.unary_cost(proposal)
        if self.grid_native:
            # the tables of all_pairwise_costs are built on the device and feed the solver directly
            labelling, e, lb, num_unlabelled, self.last_fusion_stats = binary_fusion_grid(
                self.sz[0], self.sz[1], self.smoothness_kernel, self._assignment, proposal, U0, U1, self.smooth_weights,
                self.tol, self.d_min, self.d_step, dict(improve=self.improve))
        else:
            E00, E01, E10, E11 = self.all_pairwise_costs(self._assignment, proposal)
            connectivity = np.stack([self.neighborhood["ind1"], self.neighborhood["ind2"]]).astype(np.uint32)
            labelling, e, lb, num_unlabelled = rd(U0, U1, E00, E01, E10, E11, connectivity, dict(improve=self.improve))
        a = self._assignment.copy()
        take = labelling == 1
        a[:, take] = proposal[:, take]
        self.assignment = a
        return e, lb, num_unlabelled

    def binary_fuse_until_convergence(self, proposal_cell, show_steps=False, rng=None):
        """dispmap_super.m:85-152 (including its ``iter = iter + 1`` skip of ids(1))."""
        if not isinstance(proposal_cell, (list, tuple)):
            raise TypeError("Input proposals should be given in cell array.")
        rng = rng or np.random.default_rng()
        n = len(proposal_cell)
        nrand = self.maxiter * 5
        ids = np.concatenate([np.arange(1, n + 1), rng.integers(1, n + 1, size=nrand)])
        keep = np.ones(ids.size, dtype=bool)
        keep[:-1][np.diff(ids) == 0] = False   # ids([diff(ids) == 0]) = 0; removed
        ids = ids[keep]
        if self.grid_native and self.device_loop:
            # the whole loop as one call on device-resident fields (sb_binary_fuse_until_convergence_grid); the visiting
            # order stays the one drawn above
            props = [np.asarray(p, dtype=np.float64) for p in proposal_cell]
            for p in props:
                if p.shape != self._assignment.shape:
                    raise ValueError("Binary fusion: Proposals is of wrong size")
            a, _, E, self.last_fusion_stats = binary_fuse_until_convergence_grid(
                self.sz[0], self.sz[1], self.smoothness_kernel, props, [self.unary_cost(p) for p in props], self._assignment,
                self.unary_cost(self._assignment), self.smooth_weights, self.tol, self.maxiter, ids, self.d_min, self.d_step,
                dict(improve=self.improve))
            self.assignment = a
            self.fusion_energies = list(E)
            return len(E)
        E = [self.energy()]
        visited = np.zeros(n, dtype=bool)
        for it in range(1, self.maxiter + 1):
            if it > nrand:
                ids = np.concatenate([ids, ids])
            it1 = it + 1                        # :116
            if it1 > ids.size:
                break
            if visited[ids[it1 - 1] - 1]:
                continue
            self.binary_fusion(proposal_cell[ids[it1 - 1] - 1])
            E.append(self.energy())
            if E[-2] != E[-1]:
                visited[:] = False
            else:
                visited[ids[it1 - 1] - 1] = True
            if visited.all():
                break
        self.fusion_energies = list(E)
        return len(E)

    def simultaneous_fusion(self, proposal_cell):
        """dispmap_super.m:153-198: TRW-S over the proposals plus the current assignment."""
        if not isinstance(proposal_cell, (list, tuple)):
            raise TypeError("Input proposals should be given in cell array.")
        proposal_cell = list(proposal_cell) + [self._assignment]
        unary = np.stack([self.unary_cost(p) for p in proposal_cell])
        connectivity = np.stack([self.neighborhood["ind1"], self.neighborhood["ind2"]]).astype(np.uint32)
        q, qprim = builders.fusion_positions(self.sz[0], self.sz[1], proposal_cell, self.d_min, self.d_step)
        options = dict(maxiter=self.maxiter, max_relgap=self.max_relgap)
        L, e, lb, iterations = trws(np.int32(self.smoothness_kernel), unary, connectivity, q, qprim,
                                    self.smooth_weights.reshape(-1), self.tol, options)
        assignments = np.zeros_like(proposal_cell[0])
        for i, p in enumerate(proposal_cell):
            m = L == i + 1
            assignments[:, m] = p[:, m]
        self.assignment = assignments
        return e, lb, iterations

    def current_dispmap(self):
        return self.disparitymap_from_assignment(self._assignment).reshape(self.sz[1], self.sz[0]).T

    # ---- protected methods of the reference
    def unary_cost(self, assignment):
        raise NotImplementedError("Overload unary_cost")

    def all_pairwise_costs(self, assignment, proposals=None):
        """dispmap_super.m:236-262."""
        return builders.pairwise_tables(self.sz[0], self.sz[1], self.smoothness_kernel, assignment, proposals,
                                        self.smooth_weights, self.tol, self.d_min, self.d_step)

    def update_energy(self):
        """dispmap_super.m:263-274."""
        if self._assignment is None:
            self.stored_energy = np.inf
            return
        U = self.unary_cost(self._assignment)
        self.stored_energy = builders.energy(self.sz[0], self.sz[1], self.smoothness_kernel, U, self._assignment,
                                             self.smooth_weights, self.tol, self.d_min, self.d_step)

    def get_points(self):
        return get_points(*self.sz)

    def disparitymap_from_assignment(self, assignment, points=None):
        """dispmap_super.m:318-328 (dispmap_globalstereo.m:336-345 through d_min / d_step)."""
        if points is None:
            points = self.get_points()
        return builders.plane_disparity(assignment, points, self.d_min, self.d_step)


class dispmap_ncc(dispmap_super):
    """dispmap_ncc.m:5-277.  ``patchsize`` is the one surface extension (reference: fixed 2,
    dispmap_ncc.m:24) needed for BASELINE config 3's 9x9 window."""

    def __init__(self, images, disparities, kernel, unary_weight, tol, patchsize=2):
        super().__init__(images, kernel)
        self.disparities = np.asarray(disparities, dtype=np.float64).reshape(-1)
        if unary_weight < 0:
            raise ValueError("Unary weight must be positive")
        if tol < 0:
            raise ValueError("Tolerance weight must be positive")
        self._unary_weight = unary_weight
        self._tol = tol
        # :24 -- the volume stays on the device; `ncc` (the reference's public property) copies it out on first use
        self._vol = builders.NccVolume(self.images[0], self.images[1], self.disparities, patchsize)
        self._ncc_host = None
        self.init_solution()                                                                        # :27

    @property
    def ncc(self):
        if self._ncc_host is None:
            self._ncc_host = self._vol.get()
        return self._ncc_host

    @property
    def tol(self):
        return self._tol

    @tol.setter
    def tol(self, v):
        if v < 0:
            raise ValueError("Tolerance weight must be positive")
        self._tol = v
        self.update_energy()

    @property
    def unary_weight(self):
        return self._unary_weight

    @unary_weight.setter
    def unary_weight(self, v):
        if v < 0:
            raise ValueError("Unary weight must be positive")
        self._unary_weight = v
        self.update_energy()

    def restart(self):
        self.init_solution()

    def unary_cost(self, assignment):
        """dispmap_ncc.m:107-115."""
        disps = self.disparitymap_from_assignment(assignment)
        return self._vol.sample(disps, self.unary_weight, True).reshape(-1, order="F")

    def best_disp_from_ncc(self):
        return self._vol.best_disp()

    def generate_new_plane_RANSAC(self, x, y, r):
        """dispmap_ncc.m:48-66: the plane fitted to the WTA disparities within radius r of (x, y), as a 4 x N proposal."""
        _, prop = builders.plane_from_disparity(self.best_disp_from_ncc(), x, y, r, self.smoothness_kernel, return_proposal=True)
        return prop

    def init_solution(self):
        """dispmap_ncc.m:199-207."""
        best = self.best_disp_from_ncc()
        a = np.zeros((4, self.sz[0] * self.sz[1]))
        a[2] = 1
        a[3] = -best.reshape(-1, order="F")
        self.assignment = a


class dispmap_globalstereo(dispmap_super):
    """dispmap_globalstereo.m:10-480, hot-path part: the photo-consistency unary, the disparity
    normalisation and the segmentation-weighted smoothness.  The mean-shift segmentation of the
    reference (vgg_segment_ms) is preprocessing outside the hot path: pass
    its label image as ``segment`` (H x W integers); without it all edges get lambda_l."""

    def __init__(self, images, P, disp_range, disparity_factor, options, segment=None, start_disparity=None,
                 rng=None):
        kernel = options["smoothness_kernel"]
        super().__init__(images, kernel)
        self._tol = options["disp_thresh"]
        P = np.asarray(P, dtype=np.float64)
        if P.ndim == 2:
            P = P[:, :, None]
        if np.max(np.abs(P.reshape(-1, order="F")[[0, 1, 2, 3, 4, 5, 8]] - [1, 0, 0, 0, 1, 0, 1])) > 1e-12:
            raise ValueError("First image must be reference image")
        self.P = np.transpose(P, (1, 0, 2))                       # :42
        self.disp_range = disp_range
        self.disparity_factor = disparity_factor
        disps = np.sort(np.arange(disp_range[0] * disparity_factor, disp_range[1] * disparity_factor + 1))[::-1]
        self.disps = disps
        self.d_min = float(disps[-1])
        self.d_step = float(disps[0] - self.d_min)
        self.options = dict(options)
        self._preprocess(segment)
        if start_disparity is None:
            rng = rng or np.random.default_rng()
            start_disparity = rng.random(self.sz) * self.d_step + self.d_min   # :56
        self.start_disparity = np.asarray(start_disparity, dtype=np.float64)
        self.init_solution()

    @property
    def tol(self):
        return self._tol

    @tol.setter
    def tol(self, v):
        self._tol = v

    def _preprocess(self, segment):
        """dispmap_globalstereo.m:377-414 without the segmentation call itself."""
        o = self.options
        self.improve = o.get("improve", 0) > 0
        num_in = len(self.images)
        scale = num_in / ((o.get("connect", 4) == 8) + 1)
        if segment is None:
            self.smooth_weights = np.full(self.neighborhood["ind1"].size, o["lambda_l"] * scale, dtype=np.float64)
        else:
            # :396-400 on the device (sb_smooth_weights)
            self.smooth_weights = builders.smooth_weights(np.asarray(segment).reshape(self.sz), o["lambda_h"],
                                                          o["lambda_l"], scale)
        if self._kernel == 2:
            self.smooth_weights = self.smooth_weights / self._tol
            self._tol = self._tol ** 2

    def segpln_wta(self):
        """The winner-takes-all window-matching disparity map segpln starts from (dispmap_globalstereo.m:83-117): the
        input of its per-segment plane fits."""
        P = np.transpose(self.P, (1, 0, 2))        # :66 (back to 3 x 4 x n)
        return builders.segpln_wta(self.images, P, self.disps, self.options["window"], self.options["col_thresh"])

    def segpln(self, segments, rng=None, corr=None):
        """dispmap_globalstereo.segpln (dispmap_globalstereo.m:60-201) with the segmentations injected: ``segments`` is a
        list of H x W label images (the reference computes 14 of them with vgg_segment_ms / vgg_segment_gb, :118-135).
        The window-matching WTA disparity comes from the GPU (sb_segpln_wta); the per-segment plane fits stay host glue
        exactly as in the reference -- RANSAC over random point triples (rplane, :417-453; the random stream is ``rng``
        here, MATLAB's randperm there), least squares on the inliers, the plane written to every pixel of the segment.
        Returns the proposal cell: one 4 x N plane field per segmentation."""
        rng = rng or np.random.default_rng()
        H, W = self.sz
        if corr is None:
            corr = self.segpln_wta()
        with np.errstate(divide="ignore", invalid="ignore"):
            z = 1.0 / np.asarray(corr, dtype=np.float64).reshape(-1, order="F")             # :142
            pts = self.get_points()
            WC = np.stack([z * pts[0], z * pts[1], z], axis=1)                               # :143-144
        proposals = []
        for seg in segments:
            seg = np.asarray(seg).reshape(-1, order="F")
            prop = np.zeros((4, H * W))
            prop[2] = 1.0                                                                   # :162-164 (0 disparity)
            for a in range(1, int(seg.max()) + 1):                                          # :171
                M = seg == a
                Np = WC[M]
                Np = Np[Np[:, 2] != 0]                                                      # :175
                local = Np
                if Np.shape[0] > 3:
                    local = Np[_rplane(Np, 0.1, rng)]                                       # :178-182
                if local.shape[0] > 2:
                    with np.errstate(all="ignore"):
                        try:
                            n_ = np.linalg.lstsq(local, -np.ones(local.shape[0]), rcond=None)[0]   # :186
                        except np.linalg.LinAlgError:
                            n_ = np.full(3, np.nan)
                    prop[:, M] = np.array([n_[0], n_[1], 1.0, n_[2]])[:, None]              # :191-192
            prop[~np.isfinite(prop)] = 1e-100                                               # :197-200
            proposals.append(prop)
        return proposals

    def init_solution(self):
        a = np.zeros((4, self.sz[0] * self.sz[1]))
        a[2] = 1
        a[3] = -self.start_disparity.reshape(-1, order="F")
        self.assignment = a

    def unary_cost(self, assignment):
        """dispmap_globalstereo.m:355-375."""
        return builders.photo_unary(self.images[0], self.images[1], self.P[:, :, 1], assignment, self.d_min,
                                    self.d_step, self.options["col_thresh"])


def _nsamples(ni, pt_num, pf, conf):
    """dispmap_globalstereo.nsamples (dispmap_globalstereo.m:454-466)."""
    q = np.prod(np.arange(ni - pf + 1, ni + 1) / np.arange(pt_num - pf + 1, pt_num + 1))
    cnt = 1.0 if (1 - q) < np.finfo(float).eps else np.log(1 - 0.95 if conf is None else 1 - conf) / np.log(1 - q)
    return max(cnt, 1.0)


def _rplane(pts, th, rng):
    """dispmap_globalstereo.rplane (dispmap_globalstereo.m:417-453): RANSAC inliers of the plane pts * N = -1."""
    n = pts.shape[0]
    max_i, max_sam, no_sam = 3, 500.0, 0
    inls = np.zeros(n, dtype=bool)
    with np.errstate(all="ignore"):
        while no_sam < max_sam:
            no_sam += 1
            sam = rng.permutation(n)[:3]
            try:
                N = np.linalg.solve(pts[sam], -np.ones(3))
            except np.linalg.LinAlgError:
                continue
            v = np.abs(pts @ N + 1) < th
            no_i = int(v.sum())
            if max_i < no_i:
                N = np.linalg.lstsq(pts[v], -np.ones(no_i), rcond=None)[0]
                v = np.abs(pts @ N + 1) < th
                if v.sum() > inls.sum():
                    inls = v
                    max_i = no_i
                    max_sam = min(max_sam, _nsamples(int(inls.sum()), n, 3, 0.95))
    return inls
