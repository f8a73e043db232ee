"""Grid helpers shared by the dispmap classes (dispmap_super.m:275-302)."""
from __future__ import annotations

import numpy as np


def construct_neighborhood(H, W):
    """ind1, ind2 (1-based node numbers, MATLAB column-major) of the 4-connected grid in the
    order dispmap_super.construct_neighborhood builds them (dispmap_super.m:279-302):
    vertical start->finish, finish->start, horizontal start->finish, finish->start."""
    nodenr = np.arange(1, H * W + 1, dtype=np.int64).reshape(W, H).T  # nodenr(:) = 1:N
    vs = nodenr[:-1, :].T.reshape(-1)  # start(:) column-major
    vf = nodenr[1:, :].T.reshape(-1)
    hs = nodenr[:, :-1].T.reshape(-1)
    hf = nodenr[:, 1:].T.reshape(-1)
    ind1 = np.concatenate([vs, vf, hs, hf])
    ind2 = np.concatenate([vf, vs, hf, hs])
    return ind1, ind2


def get_points(H, W):
    """2 x N array [x; y] = [column; row], 1-based (dispmap_super.m:275-278)."""
    xx, yy = np.meshgrid(np.arange(1, W + 1, dtype=np.float64), np.arange(1, H + 1, dtype=np.float64))
    return np.stack([xx.T.reshape(-1), yy.T.reshape(-1)])
