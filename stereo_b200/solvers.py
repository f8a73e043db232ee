"""Host-side mirrors of the reference's solver wrappers rd.m and trws.m.

Same names, argument order, argument meaning and error behaviour as the MATLAB
wrappers; arrays use MATLAB shapes (e.g. ``unary`` is L x N, ``connectivity`` is
2 x E and 1-BASED exactly as dispmap_super builds it) so code reads like the
reference's.  Everything is forwarded to the C ABI of libstereo_b200.so, which
runs on the GPU; nothing is computed here.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double, c_int

import numpy as np

from . import _lib
from ._lib import SB_F32, SB_F64, TrwsOptions, TrwsTiming, _dp, _ip, _up, check, lib


def _f(a):
    """float64, column-major (MATLAB) memory."""
    return np.asfortranarray(a, dtype=np.float64)


def _conn0(connectivity):
    """connectivity - 1 as 2 x E uint32 in MATLAB memory order (one pass; the callers have already
    asserted min > 0)."""
    c = np.asfortranarray(connectivity, dtype=np.uint32)
    return np.subtract(c, np.uint32(1), order="F")


def _opt(options, key, default):
    if options is None:
        return default
    if isinstance(options, dict):
        return options.get(key, default)
    return getattr(options, key, default)


last_timing: dict = {}


def _trws_args(kernel, unary, connectivity, q, qprim, alphas, tol):
    """Argument checks of trws.m:5-15 and trws_mex.cpp:42-52; returns ctypes-ready arrays."""
    unary = _f(unary)
    q = _f(q)
    qprim = _f(qprim)
    alphas = _f(np.asarray(alphas).reshape(-1))
    connectivity = np.asarray(connectivity)
    # trws.m:5-6
    assert connectivity.size == 0 or connectivity.min() > 0
    assert connectivity.size == 0 or connectivity.max() <= unary.size
    kernel = int(np.int32(kernel))
    # trws.m:9-15 (the NaN scan itself runs on the GPU while the arrays are converted; the library
    # reports "q contains NaN" / "qprim contains NaN" and _raise_nan turns that into the same error)
    L, N = unary.shape
    # trws_mex.cpp:42-52
    assert connectivity.shape[0] == 2
    E = connectivity.shape[1]
    assert q.shape[1] == qprim.shape[1] == E
    assert q.shape[0] == L and qprim.shape[0] == L
    assert alphas.shape[0] == E
    assert np.size(tol) == 1
    conn0 = _conn0(connectivity)  # trws.m:33
    return kernel, L, N, E, unary, conn0, q, qprim, alphas, float(np.asarray(tol).reshape(-1)[0])


def _raise_nan(rc):
    if rc == _lib.SB_EINVAL:
        msg = lib().sb_last_error().decode("utf-8", "replace")
        if msg in ("q contains NaN", "qprim contains NaN"):
            raise ValueError(msg)


def _trws_options(options):
    opt = TrwsOptions()
    lib().sb_trws_default_options(ctypes.byref(opt))
    opt.maxiter = float(_opt(options, "maxiter", 1000))
    opt.max_relgap = float(_opt(options, "max_relgap", 0))
    prec = _opt(options, "precision", "f32")
    opt.precision = SB_F64 if prec in ("f64", SB_F64, "double") and prec != 0 else SB_F32
    opt.fuse_rounding = int(bool(_opt(options, "fuse_rounding", True)))
    opt.col_blocks = int(_opt(options, "col_blocks", 0))      # grid-native multi-GPU: column blocks per rank (0 = default)
    opt.latency_mode = int(_opt(options, "latency_mode", 0))  # grid-native: 0 auto, 1 / -1 force the latency build on / off
    return opt


def _timing_dict(tm):
    return dict(setup_ms=tm.setup_ms, solve_ms=tm.solve_ms, sweep_ms_avg=tm.sweep_ms_avg,
                download_ms=tm.download_ms, kernel_launches=tm.kernel_launches,
                sweep_kernel_ms=tm.sweep_kernel_ms, sweep_kernel_launches=tm.sweep_kernel_launches)


def trws(kernel, unary, connectivity, q, qprim, alphas, tol, options=None):
    """[solution, energy, lower_bound, iterations] = trws(kernel, unary, connectivity, q, qprim,
    alphas, tol, options)  -- trws.m:2-33.

    kernel: 1 (truncated linear) or 2 (truncated quadratic).  unary: L x N.
    connectivity: 2 x E, 1-based.  q, qprim: L x E.  alphas: E.  tol: scalar.
    options: dict/object with ``maxiter`` (default 1000) and ``max_relgap`` (default 0)
    (trws_mex.cpp:38-40); extra keys ``precision`` ("f32"|"f64") and ``fuse_rounding``.
    Returns (solution N float64 1-based, energy, lower_bound, iterations).
    """
    kernel, L, N, E, unary, conn0, q, qprim, alphas, tol = _trws_args(kernel, unary, connectivity, q, qprim,
                                                                      alphas, tol)
    opt = _trws_options(options)
    solution = np.zeros(N, dtype=np.float64)
    e = c_double()
    lb = c_double()
    it = c_double()
    tm = TrwsTiming()
    rc = lib().sb_trws_solve(kernel, L, N, E, unary.ctypes.data_as(_dp), conn0.ctypes.data_as(_up),
                             q.ctypes.data_as(_dp), qprim.ctypes.data_as(_dp), alphas.ctypes.data_as(_dp),
                             tol, ctypes.byref(opt), solution.ctypes.data_as(_dp),
                             ctypes.byref(e), ctypes.byref(lb), ctypes.byref(it), ctypes.byref(tm))
    _raise_nan(rc)
    check(rc)
    last_timing.clear()
    last_timing.update(_timing_dict(tm))
    return solution, e.value, lb.value, it.value


class TrwsSolver:
    """Resident form of ``trws``: the MRFEnergy object of trws_mex.cpp:58-146 kept alive in HBM.
    ``minimize`` continues from the current messages (Minimize_TRW_S, minimize.cpp:7-116),
    ``reset`` is ZeroMessages (MRFEnergy.cpp:115-131)."""

    def __init__(self, kernel, unary, connectivity, q, qprim, alphas, tol, options=None):
        kernel, L, N, E, unary, conn0, q, qprim, alphas, tol = _trws_args(kernel, unary, connectivity, q, qprim,
                                                                          alphas, tol)
        self.N, self.L, self.E = N, L, E
        opt = _trws_options(options)
        self._h = ctypes.c_void_p()
        rc = lib().sb_trws_create(kernel, L, N, E, unary.ctypes.data_as(_dp), conn0.ctypes.data_as(_up),
                                  q.ctypes.data_as(_dp), qprim.ctypes.data_as(_dp), alphas.ctypes.data_as(_dp),
                                  tol, ctypes.byref(opt), ctypes.byref(self._h))
        _raise_nan(rc)
        check(rc)
        self.timing = {}

    def reset(self):
        check(lib().sb_trws_reset(self._h))

    def minimize(self, maxiter=1000, max_relgap=0.0):
        e = c_double()
        lb = c_double()
        it = c_double()
        tm = TrwsTiming()
        check(lib().sb_trws_minimize(self._h, float(maxiter), float(max_relgap), ctypes.byref(e), ctypes.byref(lb),
                                     ctypes.byref(it), ctypes.byref(tm)))
        self.timing = _timing_dict(tm)
        return e.value, lb.value, it.value

    def labels(self):
        out = np.zeros(self.N, dtype=np.float64)
        check(lib().sb_trws_get_labels(self._h, out.ctypes.data_as(_dp)))
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_trws_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def rd(U0, U1, E00, E01, E10, E11, connectivity, options=None):
    """[solution, energy, lower_bound, num_unlabelled] = rd(U0, U1, E00, E01, E10, E11,
    connectivity, options)  -- rd.m:3-21.

    U0, U1: N.  E00..E11: E.  connectivity: 2 x E, 1-based.  options: ``improve`` (bool,
    default False; rd_mex.cpp:34).  Returns (solution N float64 in {0, 1, negative}, energy,
    lower_bound, num_unlabelled)."""
    U0, U1 = _f(np.asarray(U0).reshape(-1)), _f(np.asarray(U1).reshape(-1))
    E00, E01, E10, E11 = (_f(np.asarray(x).reshape(-1)) for x in (E00, E01, E10, E11))
    connectivity = np.asarray(connectivity)
    # rd.m:5-6
    assert connectivity.size == 0 or connectivity.min() > 0
    assert connectivity.size == 0 or connectivity.max() <= U0.size
    # rd_mex.cpp:36-49
    assert U0.shape == U1.shape
    assert E00.shape == E01.shape == E10.shape == E11.shape
    assert connectivity.shape == (2, E00.size)
    improve = bool(_opt(options, "improve", False))
    conn0 = _conn0(connectivity)  # rd.m:21
    N, E = U0.size, E00.size
    solution = np.zeros(N, dtype=np.float64)
    e, lb, nu = c_double(), c_double(), c_double()
    rc = lib().sb_rd_solve(N, E, U0.ctypes.data_as(_dp), U1.ctypes.data_as(_dp), E00.ctypes.data_as(_dp),
                           E01.ctypes.data_as(_dp), E10.ctypes.data_as(_dp), E11.ctypes.data_as(_dp),
                           conn0.ctypes.data_as(_up), int(improve), solution.ctypes.data_as(_dp), ctypes.byref(e),
                           ctypes.byref(lb), ctypes.byref(nu))
    check(rc)
    return solution, e.value, lb.value, nu.value


def binary_fusion_grid(H, W, kernel, assignment, proposal, U0, U1, weights, tol, d_min=0.0, d_step=1.0, options=None,
                       device_ptrs=None):
    """dispmap_super.binary_fusion (dispmap_super.m:61-84) as one grid-native call (``sb_binary_fusion_grid``):
    the tables of all_pairwise_costs are built on the device.  assignment / proposal: 4 x N; U0 / U1: N; weights: E
    (the reference's term order).  Returns (labelling N float64 in {0, 1, negative}, energy, lower_bound,
    num_unlabelled, stats) with stats = dict(rounds, relabels, bfs_sweeps, solve_ms).

    ``device_ptrs``: dict of integer device addresses (assignment, proposal, U0, U1, weights, labels) -- the
    arrays stay where they are; the labelling is then written to the ``labels`` address and None is returned
    in its place."""
    N = int(H) * int(W)
    improve = bool(_opt(options, "improve", False))
    e, lb, nu = c_double(), c_double(), c_double()
    stats = (c_double * 4)()
    if device_ptrs is not None:
        args = [ctypes.c_void_p(int(device_ptrs[k])) for k in ("assignment", "proposal", "U0", "U1", "weights")]
        out = ctypes.c_void_p(int(device_ptrs["labels"]))
        solution = None
        on_device = 1
    else:
        def planes(a):
            # 4 x N in MATLAB (column-major) memory order; an F-contiguous (4, N) view is passed through untouched
            a = np.asarray(a, dtype=np.float64)
            assert a.shape == (4, N)
            return np.asfortranarray(a)
        arrs = [planes(assignment), planes(proposal),
                _f(np.asarray(U0).reshape(-1)), _f(np.asarray(U1).reshape(-1)), _f(np.asarray(weights).reshape(-1))]
        assert arrs[2].size == N and arrs[3].size == N
        args = [ctypes.c_void_p(a.ctypes.data) for a in arrs]
        solution = np.zeros(N, dtype=np.float64)
        out = ctypes.c_void_p(solution.ctypes.data)
        on_device = 0
    check(lib().sb_binary_fusion_grid(int(H), int(W), int(kernel), *args, float(tol), float(d_min), float(d_step), int(improve),
                                      on_device, out, ctypes.byref(e), ctypes.byref(lb), ctypes.byref(nu), stats))
    return solution, e.value, lb.value, nu.value, dict(rounds=stats[0], relabels=stats[1], bfs_sweeps=stats[2], solve_ms=stats[3])


def binary_fuse_until_convergence_grid(H, W, kernel, proposals, unaries, assignment, unary, weights, tol, maxiter, ids,
                                       d_min=0.0, d_step=1.0, options=None):
    """dispmap_super.binary_fuse_until_convergence (dispmap_super.m:85-152) as one call
    (``sb_binary_fuse_until_convergence_grid``): proposals n x (4 x N), unaries n x N (unary_cost of each proposal),
    assignment 4 x N and unary N (the current field and its unary cost), ids = the 1-based visiting order of :96-101.
    Returns (assignment 4 x N, unary N, energies E, stats)."""
    N = int(H) * int(W)
    n = len(proposals)
    props = np.ascontiguousarray(np.stack([np.asfortranarray(np.asarray(p, dtype=np.float64)).T for p in proposals]))  # n x N x 4
    assert props.shape == (n, N, 4)
    uns = np.ascontiguousarray(np.asarray(unaries, dtype=np.float64).reshape(n, N))
    cur = np.ascontiguousarray(np.asarray(assignment, dtype=np.float64).T)        # N x 4 == MATLAB 4 x N memory
    ucur = np.ascontiguousarray(np.asarray(unary, dtype=np.float64).reshape(N))
    wt = _f(np.asarray(weights).reshape(-1))
    ids = np.ascontiguousarray(np.asarray(ids, dtype=np.int32))
    energies = np.zeros(int(maxiter) + 1, dtype=np.float64)
    ne = ctypes.c_int(0)
    stats = (c_double * 4)()
    vp = ctypes.c_void_p
    check(lib().sb_binary_fuse_until_convergence_grid(
        int(H), int(W), int(kernel), n, vp(props.ctypes.data), vp(uns.ctypes.data), vp(cur.ctypes.data), vp(ucur.ctypes.data),
        vp(wt.ctypes.data), float(tol), float(d_min), float(d_step), int(bool(_opt(options, "improve", False))), int(maxiter),
        ids.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), int(ids.size), 0, energies.ctypes.data_as(_dp), ctypes.byref(ne), stats))
    return cur.T.copy(), ucur, energies[:ne.value].copy(), dict(fusions=stats[0], rounds=stats[1], solve_ms=stats[2], adopted=stats[3])


def trws_grid_ordering(H, W):
    """m_ordering of SetAutomaticOrdering (ordering.cpp:7-157) on the H x W grid, as (H, W) int32."""
    out = np.zeros(H * W, dtype=np.int32)
    check(lib().sb_trws_grid_ordering(int(H), int(W), out.ctypes.data_as(_ip)))
    return out.reshape(W, H).T.copy()


def grid_from_connectivity(connectivity0, N):
    """(H, W) of a 0-based 2 x E connectivity list, or raises SbError(SB_ENOTGRID)."""
    conn0 = np.asfortranarray(connectivity0, dtype=np.uint32)
    H = c_int()
    W = c_int()
    check(lib().sb_grid_from_connectivity(int(N), conn0.shape[1], conn0.ctypes.data_as(_up), ctypes.byref(H),
                                          ctypes.byref(W)))
    return H.value, W.value
