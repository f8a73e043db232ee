"""Row-banded TRW-S over the GPUs of one box (one process per GPU, torch.distributed plumbing).

``TrwsBandedSolver`` is the multi-rank form of ``TrwsSolver``: every rank passes the SAME problem
(static data is replicated), sweeps the strips of its own row band and pushes the messages that
cross a band boundary into the neighbouring GPU's memory (include/stereo_b200.h, "several GPUs").
torch.distributed is used for three things only: exchanging the CUDA IPC handles once, the
all-reduce of (energy, lower bound) after every pass -- which the stop rule of
minimize.cpp:97-112 needs and which also separates consecutive passes -- and gathering labels.
"""
from __future__ import annotations

import ctypes

import numpy as np

from . import _lib
from ._lib import _dp, _up, check, lib
from .solvers import _trws_args, _trws_options, _raise_nan

MODE_SEND, MODE_ROUND = 1, 2


def band_of_row(r, H, world):
    """Owner of image row r (trws_order.h: band_of_row)."""
    return (r * world) // H


class TrwsBandedSolver:
    def __init__(self, kernel, unary, connectivity, q, qprim, alphas, tol, options=None, group=None):
        import torch
        import torch.distributed as dist
        self.dist, self.torch, self.group = dist, torch, group
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        kernel, L, N, E, unary, conn0, q, qprim, alphas, tol = _trws_args(kernel, unary, connectivity, q, qprim,
                                                                          alphas, tol)
        self.N, self.L, self.E = N, L, E
        opt = _trws_options(options)
        self.fuse = bool(opt.fuse_rounding)
        self._h = ctypes.c_void_p()
        rc = lib().sb_trws_create_banded(kernel, L, N, E, unary.ctypes.data_as(_dp), conn0.ctypes.data_as(_up),
                                         q.ctypes.data_as(_dp), qprim.ctypes.data_as(_dp), alphas.ctypes.data_as(_dp),
                                         tol, ctypes.byref(opt), self.rank, self.world, ctypes.byref(self._h))
        _raise_nan(rc)
        check(rc)
        if self.world > 1:
            mine = ctypes.create_string_buffer(192)
            check(lib().sb_trws_ipc_export(self._h, mine))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, bytes(mine.raw), group=group)
            up = gathered[self.rank - 1] if self.rank > 0 else None
            down = gathered[self.rank + 1] if self.rank + 1 < self.world else None
            check(lib().sb_trws_ipc_attach(self._h, up, down))
            dist.barrier(group=group)
        self.kernel_ms = 0.0

    def _pass(self, which, mode):
        acc = (ctypes.c_double * 2)()
        check(lib().sb_trws_pass(self._h, which, mode, acc))
        t = self.torch.tensor([acc[0], acc[1]], dtype=self.torch.float64)
        if self.world > 1:
            dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
            t = t.to(dev)
            self.dist.all_reduce(t, group=self.group)   # also the barrier between passes
            t = t.cpu()
        return float(t[0]), float(t[1])

    def reset(self):
        check(lib().sb_trws_reset(self._h))
        if self.world > 1:
            self.dist.barrier(group=self.group)

    def minimize(self, maxiter=1000, max_relgap=0.0):
        """Minimize_TRW_S (minimize.cpp:7-116) with the primal rounding of iteration t fused into
        the forward sweep of t+1 (same control flow as the single-GPU driver in trws_solve.cu)."""
        iter_max = int(maxiter)
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

    def labels(self):
        """1-based labels of all pixels on every rank (each rank rounds its own band)."""
        out = np.zeros(self.N, dtype=np.float64)
        check(lib().sb_trws_get_labels(self._h, out.ctypes.data_as(_dp)))
        if self.world > 1:
            # labels of rows this rank does not own come back as 1 (label 0 + 1): keep own rows only
            t = self.torch.from_numpy(out - 1.0)
            dev = "cuda" if self.dist.get_backend(self.group) == "nccl" else "cpu"
            t = t.to(dev)
            self.dist.all_reduce(t, group=self.group)
            out = t.cpu().numpy() + 1.0
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
