"""Grid-native TRW-S (``sb_trws_grid_*``, include/stereo_b200.h).

``TrwsGrid`` is what ``dispmap_super.simultaneous_fusion`` (dispmap_super.m:153-198) hands the
solver when it does NOT first expand its proposals into the L x E arrays ``q`` / ``qprim``: L
plane proposals (4 x N each), an L x N unary slab and the E smooth weights.  The positions
q(:,p) = d(plane@ind2, pt ind2), qprim(:,p) = d(plane@ind1, pt ind2) (dispmap_super.m:180-183) are
recomputed on the device.  With ``world > 1`` every rank holds and sweeps one COLUMN band
(torch.distributed plumbing as in multigpu.py; the halo is peer stores over NVLink): the strips of
the sweep run along the image rows, so the ranks work as the stages of a pipeline.
``TrwsGridLocalGroup`` runs the same banded sweep with all ranks in ONE process on one device.

``positions_from_labels`` rebuilds q / qprim in numpy from the planes AS STORED on the device
(``get_label``), so a parity test can hand the reference solver bit-identical inputs.
"""
from __future__ import annotations

import ctypes
from ctypes import c_double

import numpy as np

from ._lib import TrwsTiming, _dp, check, lib
from .grid import construct_neighborhood
from .solvers import _f, _timing_dict, _trws_options

MODE_SEND, MODE_ROUND = 1, 2


class TrwsGrid:
    def __init__(self, kernel, H, W, L, tol, options=None, group=None, rank=0, world=1, local=False):
        self.H, self.W, self.L = int(H), int(W), int(L)
        self.N = self.H * self.W
        self.E = 2 * ((self.H - 1) * self.W + self.H * (self.W - 1))
        self.group = group
        self.dist = None
        if (group is not None or world > 1) and not local:
            import torch
            import torch.distributed as dist
            self.dist, self.torch = dist, torch
            rank, world = dist.get_rank(group), dist.get_world_size(group)
        self.rank, self.world = rank, world
        opt = _trws_options(options)
        self.fuse = bool(opt.fuse_rounding)
        self.precision = opt.precision
        self._h = ctypes.c_void_p()
        check(lib().sb_trws_grid_create(int(np.int32(kernel)), self.H, self.W, self.L, float(tol), ctypes.byref(opt),
                                        self.rank, self.world, ctypes.byref(self._h)))
        self.timing = {}
        self._attached = False

    # ---- problem
    def set_labels(self, l0, planes, unary, d_min=0.0, d_step=1.0):
        """planes: nl x 4 x N (or 4 x N), unary: nl x N (or N)."""
        planes = np.asarray(planes, dtype=np.float64)
        unary = np.asarray(unary, dtype=np.float64)
        if planes.ndim == 2:
            planes, unary = planes[None], unary[None]
        nl = planes.shape[0]
        assert planes.shape == (nl, 4, self.N) and unary.shape == (nl, self.N)
        # proposal-major, each proposal a MATLAB 4 x N array (column-major: [a b c d0] per pixel contiguous)
        pl = np.ascontiguousarray(planes.transpose(0, 2, 1))
        un = np.ascontiguousarray(unary)
        check(lib().sb_trws_grid_set_labels(self._h, int(l0), nl, pl.ctypes.data_as(_dp), un.ctypes.data_as(_dp),
                                            float(d_min), float(d_step)))

    def set_labels_ptr(self, l0, nl, planes_ptr, unary_ptr, d_min=0.0, d_step=1.0):
        """The same from raw addresses (host or device): nl proposals of 4 x N doubles (MATLAB layout) / nl x N unaries."""
        check(lib().sb_trws_grid_set_labels(self._h, int(l0), int(nl), ctypes.cast(ctypes.c_void_p(int(planes_ptr)), _dp),
                                            ctypes.cast(ctypes.c_void_p(int(unary_ptr)), _dp), float(d_min), float(d_step)))

    def set_weights(self, alphas):
        a = _f(np.asarray(alphas).reshape(-1))
        assert a.size == self.E
        check(lib().sb_trws_grid_set_weights(self._h, a.ctypes.data_as(_dp)))

    def synth(self, seed, scene=None, offset=(0, 0)):
        """Seeded synthetic problem generated on the device; with ``scene=(Hs, Ws)`` this solver's grid is the
        window at ``offset`` of that larger scene."""
        if scene is None:
            check(lib().sb_trws_grid_synth(self._h, ctypes.c_uint64(int(seed))))
        else:
            check(lib().sb_trws_grid_synth_window(self._h, ctypes.c_uint64(int(seed)), int(scene[0]), int(scene[1]),
                                                  int(offset[0]), int(offset[1])))

    def finalize(self):
        check(lib().sb_trws_grid_finalize(self._h))
        if self.world > 1 and not self._attached and self.dist is not None:
            mine = ctypes.create_string_buffer(128)
            check(lib().sb_trws_grid_ipc_export(self._h, mine))
            gathered = [None] * self.world
            self.dist.all_gather_object(gathered, bytes(mine.raw), group=self.group)
            # the column blocks are dealt round robin, so the neighbours wrap around the ranks
            up = gathered[(self.rank - 1) % self.world]
            down = gathered[(self.rank + 1) % self.world]
            check(lib().sb_trws_grid_ipc_attach(self._h, up, down))
            self.dist.barrier(group=self.group)
            self._attached = True

    def get_label(self, l):
        """(unary, own, gx, gy) of proposal l as stored (N doubles each, MATLAB node order)."""
        out = np.zeros((4, self.N), dtype=np.float64)
        check(lib().sb_trws_grid_get_label(self._h, int(l), out.ctypes.data_as(_dp)))
        return out[0], out[1], out[2], out[3]

    def get_weights(self):
        out = np.zeros(self.E, dtype=np.float64)
        check(lib().sb_trws_grid_get_weights(self._h, out.ctypes.data_as(_dp)))
        return out

    def info(self):
        out = (ctypes.c_int64 * 8)()
        check(lib().sb_trws_grid_info(self._h, out))
        keys = ("hbm_bytes", "nodes_stored", "col_lo", "col_hi", "ctas_fwd", "ctas_bwd", "smem_per_cta", "LP")
        d = dict(zip(keys, [int(v) for v in out]))
        on = ctypes.c_int(0)
        check(lib().sb_trws_grid_latency_mode(self._h, ctypes.byref(on)))
        d["latency_build"] = int(on.value)
        return d

    def counters(self):
        """Cumulative (sweep kernel ms, sweep launches, set-up ms) since creation."""
        out = (ctypes.c_double * 3)()
        check(lib().sb_trws_grid_counters(self._h, out))
        return float(out[0]), int(out[1]), float(out[2])

    # ---- solve
    def reset(self):
        check(lib().sb_trws_grid_reset(self._h))
        if self.world > 1 and self.dist is not None:
            self.dist.barrier(group=self.group)

    # asynchronous passes: launch only enqueues; wait returns [(energy, bound contribution)] per launched pass
    def launch(self, which, mode):
        check(lib().sb_trws_grid_launch_pass(self._h, which, mode))

    def wait(self, max_passes=256):
        acc = (ctypes.c_double * (2 * max_passes))()
        n = ctypes.c_int()
        check(lib().sb_trws_grid_wait(self._h, acc, max_passes, ctypes.byref(n)))
        return [(acc[2 * i], acc[2 * i + 1]) for i in range(n.value)]

    @staticmethod
    def pass_list(fuse, iters):
        """(pass, mode) sequence of `iters` iterations under max_relgap = 0 (minimize.cpp:31-113; with fused
        rounding the rounding of iteration t rides in the forward sweep of t + 1)."""
        seq = []
        for it in range(1, iters + 1):
            seq.append((0, MODE_SEND | (MODE_ROUND if (fuse and it > 1) else 0)))
            seq.append((1, 0))
            if (not fuse) or it >= iters:
                seq.append((0, MODE_ROUND))
        return seq

    def _minimize_async(self, iters):
        """Fixed iteration count on several ranks: every rank launches all its passes at once (the message words
        validate themselves, so no barrier separates the passes) and the sums are all-reduced once."""
        seq = self.pass_list(self.fuse, iters)
        accs = []
        for i0 in range(0, len(seq), 200):
            for which, mode in seq[i0:i0 + 200]:
                self.launch(which, mode)
            accs += self.wait()
        t = self.torch.tensor(accs, dtype=self.torch.float64)
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = t.to(dev)
        self.dist.all_reduce(t, group=self.group)
        t = t.cpu()
        energy = lb = 0.0
        for (which, mode), row in zip(seq, t.tolist()):
            if which == 0 and (mode & MODE_ROUND):
                energy = row[0]
            if which == 1:
                lb = row[1]
        return energy, lb, float(iters)

    def _pass(self, which, mode):
        acc = (ctypes.c_double * 2)()
        check(lib().sb_trws_grid_pass(self._h, which, mode, acc))
        if self.world == 1:
            return acc[0], acc[1]
        t = self.torch.tensor([acc[0], acc[1]], dtype=self.torch.float64)
        dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
        t = t.to(dev)
        self.dist.all_reduce(t, group=self.group)   # the stop rule needs the sums; also separates the passes
        t = t.cpu()
        return float(t[0]), float(t[1])

    def minimize(self, maxiter=1000, max_relgap=0.0):
        """Minimize_TRW_S (minimize.cpp:7-116); returns (energy, lower_bound, iterations)."""
        if self.world == 1:
            e, lb, it = c_double(), c_double(), c_double()
            tm = TrwsTiming()
            check(lib().sb_trws_grid_minimize(self._h, float(maxiter), float(max_relgap), ctypes.byref(e),
                                              ctypes.byref(lb), ctypes.byref(it), ctypes.byref(tm)))
            self.timing = _timing_dict(tm)
            return e.value, lb.value, it.value
        iter_max = int(maxiter)
        if max_relgap <= 0.0 and iter_max >= 1:
            return self._minimize_async(iter_max)
        energy = lb = 0.0
        it = 1
        while True:
            e, _ = self._pass(0, MODE_SEND | (MODE_ROUND if (self.fuse and it > 1) else 0))
            if self.fuse and it > 1:
                energy = e
                if (energy - lb) / energy < max_relgap:
                    return energy, lb, float(it - 1)
            _, lb = self._pass(1, 0)
            if (not self.fuse) or it >= iter_max:
                energy, _ = self._pass(0, MODE_ROUND)
                if it >= iter_max or (energy - lb) / energy < max_relgap:
                    return energy, lb, float(it)
            it += 1

    def labels(self, gather=True):
        out = np.zeros(self.N, dtype=np.float64)
        check(lib().sb_trws_grid_get_labels(self._h, out.ctypes.data_as(_dp)))
        if self.world > 1 and gather and self.dist is not None:
            t = self.torch.from_numpy(out)
            dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
            t = t.to(dev)
            self.dist.all_reduce(t, group=self.group)   # rows a rank does not sweep are 0
            out = t.cpu().numpy()
        return out

    def close(self):
        if getattr(self, "_h", None):
            lib().sb_trws_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TrwsGridLocalGroup:
    """All `world` ranks of a banded sweep as solvers of THIS process on the current device, their sweeps running
    concurrently on separate streams (sb_trws_grid_attach_local).  The ranks exchange their boundary messages
    exactly as over NVLink -- sender writes into the receiver's arrays, receiver polls self-validating words --
    so a one-GPU box exercises the multi-GPU protocol, cross-band dependencies included."""

    def __init__(self, kernel, H, W, L, tol, world, options=None):
        self.world = int(world)
        self.ranks = [TrwsGrid(kernel, H, W, L, tol, options, rank=r, world=self.world, local=True) for r in range(self.world)]
        self.fuse = self.ranks[0].fuse
        self.N = self.ranks[0].N

    def each(self, fn):
        return [fn(g) for g in self.ranks]

    def finalize(self):
        for g in self.ranks:
            g.finalize()
        for r, g in enumerate(self.ranks):
            up = self.ranks[(r - 1) % self.world]._h
            down = self.ranks[(r + 1) % self.world]._h
            check(lib().sb_trws_grid_attach_local(g._h, up, down, self.world))
            g._attached = True

    def minimize(self, iters):
        seq = TrwsGrid.pass_list(self.fuse, int(iters))
        tot = np.zeros((len(seq), 2))
        for i0 in range(0, len(seq), 200):
            part = seq[i0:i0 + 200]
            for which, mode in part:
                for g in self.ranks:
                    g.launch(which, mode)
            for g in self.ranks:
                tot[i0:i0 + len(part)] += np.asarray(g.wait()).reshape(len(part), 2)
        energy = lb = 0.0
        for (which, mode), row in zip(seq, tot):
            if which == 0 and (mode & MODE_ROUND):
                energy = row[0]
            if which == 1:
                lb = row[1]
        return float(energy), float(lb), float(iters)

    def labels(self):
        out = np.zeros(self.N)
        for g in self.ranks:
            out += g.labels(gather=False)     # nodes a rank does not sweep are 0
        return out

    def close(self):
        for g in self.ranks:
            g.close()


def trws_grid(kernel, unary, proposals, weights, tol, H, W, options=None, d_min=0.0, d_step=1.0):
    """One-shot form: [solution, energy, lower_bound, iterations] like ``trws`` (trws.m:2-33), from
    proposals (L x 4 x N), unary (L x N) and weights (E)."""
    unary = np.asarray(unary, dtype=np.float64)
    L = unary.shape[0]
    g = TrwsGrid(kernel, H, W, L, tol, options)
    try:
        g.set_labels(0, proposals, unary, d_min, d_step)
        g.set_weights(weights)
        g.finalize()
        maxiter = 1000 if options is None else options.get("maxiter", 1000)
        relgap = 0.0 if options is None else options.get("max_relgap", 0.0)
        e, lb, it = g.minimize(maxiter, relgap)
        sol = g.labels()
        trws_grid.last_timing = dict(g.timing)
        return sol, e, lb, it
    finally:
        g.close()


def positions_from_labels(H, W, own, gx, gy, dtype=np.float32):
    """q, qprim (L x E, float64 holding `dtype`-rounded values) from the stored label planes, with the
    arithmetic of the sweep kernel: q = own[head], qprim = own[tail] + step towards the head."""
    own = np.asarray(own, dtype=dtype)
    gx = np.asarray(gx, dtype=dtype)
    gy = np.asarray(gy, dtype=dtype)
    ind1, ind2 = construct_neighborhood(H, W)   # tail, head (1-based)
    t, h = ind1 - 1, ind2 - 1
    dr = (h % H) - (t % H)      # head row - tail row
    dc = (h // H) - (t // H)
    q = own[:, h]
    step = np.where(dr != 0, gy[:, t] * dr.astype(dtype), gx[:, t] * dc.astype(dtype)).astype(dtype)
    qprim = (own[:, t] + step).astype(dtype)
    return q.astype(np.float64), qprim.astype(np.float64)
