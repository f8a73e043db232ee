"""Generates the committed golden vectors from the UNMODIFIED reference (oracle/_ref, built
from /root/reference by oracle/Makefile).  Run in the build container only:

    python tests/golden/make_golden.py

The fixtures pin the C port (oracle/*.c) and the product where /root/reference and
oracle/_ref are absent.  Inputs are regenerated from seeds by stereo_b200.synth, so only
seeds + reference outputs are stored.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle  # noqa: E402
from stereo_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

ORDER_SHAPES = [(1, 6), (6, 1), (2, 6), (6, 2), (3, 9), (9, 3), (3, 3), (4, 4), (4, 9), (9, 4), (5, 5), (10, 12),
                (17, 33), (40, 23), (64, 48)]

TRWS_CASES = [  # (H, W, L, kernel, seed, maxiter, max_relgap)
    (6, 8, 5, 1, 1, 5, 0.0), (12, 17, 15, 1, 2, 10, 0.0), (12, 17, 15, 2, 2, 10, 0.0),
    (24, 31, 16, 1, 7, 8, 0.0), (20, 25, 33, 1, 3, 7, 0.0), (20, 25, 33, 2, 3, 7, 0.0),
    (48, 64, 40, 1, 1, 20, 0.0), (48, 64, 40, 2, 1, 20, 0.0), (31, 45, 64, 1, 1, 10, 0.0),
    (31, 45, 100, 2, 1, 6, 0.0), (17, 19, 192, 1, 4, 5, 0.0), (13, 16, 256, 1, 5, 5, 0.0),
    (13, 16, 256, 2, 5, 5, 0.0), (3, 9, 7, 1, 6, 10, 0.0), (1, 9, 7, 1, 6, 10, 0.0), (9, 1, 7, 2, 6, 10, 0.0),
    (48, 64, 8, 1, 1, 200, 1e-3), (30, 30, 24, 1, 9, 100, 1e-2),
]


RD_CASES = [  # (H, W, seed, kernel, frustrated, improve)
    (6, 8, 1, 1, 0, 0), (6, 8, 1, 1, 1, 0), (17, 23, 2, 1, 0, 0), (17, 23, 2, 2, 0, 0), (17, 23, 3, 1, 1, 0),
    (40, 50, 3, 1, 0, 0), (40, 50, 3, 1, 1, 0), (64, 48, 4, 2, 0, 0), (1, 9, 5, 1, 0, 0), (9, 1, 5, 1, 1, 0),
    (2, 2, 6, 1, 1, 0), (120, 90, 7, 1, 0, 0), (12, 14, 8, 1, 1, 1), (20, 25, 9, 1, 1, 1), (40, 50, 3, 1, 0, 1),
]


def main():
    assert oracle.have_ref("trws"), "build oracle/_ref first (make -C oracle ref)"
    # 1. orderings
    d = {}
    for H, W in ORDER_SHAPES:
        d[f"{H}x{W}"] = oracle.trws_ordering(H, W)
    np.savez_compressed(os.path.join(HERE, "ordering.npz"), **d)
    # 2. full solves
    out = {}
    for i, (H, W, L, k, seed, it, gap) in enumerate(TRWS_CASES):
        pr = synth.trws_problem(H, W, L, seed=seed, kernel=k)
        lab, e, lb, n = oracle.trws_solve(k, pr["unary"].T, (pr["connectivity"] - 1).T, pr["q"].T, pr["qprim"].T,
                                          pr["alphas"], pr["tol"], it, gap)
        out[f"case{i}_params"] = np.array([H, W, L, k, seed, it, gap], dtype=np.float64)
        out[f"case{i}_labels"] = lab.astype(np.int16)
        out[f"case{i}_scalars"] = np.array([e, lb, n], dtype=np.float64)
        # checksum of the regenerated inputs so a drifting generator is detected
        out[f"case{i}_inputsum"] = np.array([pr["unary"].sum(), pr["q"].sum(), pr["qprim"].sum(), pr["alphas"].sum()])
    np.savez_compressed(os.path.join(HERE, "trws_solve.npz"), **out)
    # 3. single message updates (known-answer vectors incl. exact position ties)
    rng = np.random.Generator(np.random.PCG64(123))
    msgs = {}
    idx = 0
    for k in (1, 2):
        for L in (2, 15, 64, 256):
            for dir_ in (0, 1):
                for sw in (0, 1):
                    Di = rng.random(L) * 3
                    msg = rng.random(L) * 0.5
                    s0 = rng.random(L)
                    s1 = rng.random(L)
                    s1[: L // 3] = s0[: L // 3]
                    if L > 4:
                        s0[3] = s0[4]
                    alpha = float(rng.choice([0.0, 2.0, 18.0, 216.0])) if idx % 5 == 0 else float(rng.choice([2.0, 18.0]))
                    lam = 0.05 if k == 1 else 0.05 ** 2
                    gamma = float(rng.choice([0.25, 1 / 6, 0.125, 0.5]))
                    m2, vmin = oracle.trws_update_message(k, Di, msg, s0, s1, alpha, lam, gamma, dir_, sw)
                    msgs[f"u{idx}_in"] = np.stack([Di, msg, s0, s1])
                    msgs[f"u{idx}_par"] = np.array([k, alpha, lam, gamma, dir_, sw])
                    msgs[f"u{idx}_out"] = np.concatenate([m2, [vmin]])
                    idx += 1
    np.savez_compressed(os.path.join(HERE, "update_message.npz"), **msgs)
    # 4. vgg_interp2 linear
    A = rng.random((7, 9, 3)) * 255
    X = rng.random(300) * 11 - 1
    Y = rng.random(300) * 9 - 1
    X[:6] = [1, 9, 9, 3.5, 9, 1]
    Y[:6] = [1, 7, 3.2, 7, 1, 7]
    B = oracle.interp2_linear(A, X, Y, -1000.0)
    np.savez_compressed(os.path.join(HERE, "interp2.npz"), A=A, X=X, Y=Y, B=B)
    # 5. QPBO fusions (rd_mex.cpp through the shim); libc rand() reseeded like the tests do
    import ctypes
    libc = ctypes.CDLL(None)
    out = {}
    for i, (H, W, seed, kernel, mode, improve) in enumerate(RD_CASES):
        pr = synth.rd_problem(H, W, seed=seed, kernel=kernel, mode="frustrated" if mode else "stereo")
        libc.srand(1)
        lab, e, lb, nu = oracle.rd_solve(pr["U0"], pr["U1"], pr["E00"], pr["E01"], pr["E10"], pr["E11"],
                                         (pr["connectivity"] - 1).T, improve=bool(improve))
        out[f"case{i}_params"] = np.array([H, W, seed, kernel, mode, improve], dtype=np.float64)
        out[f"case{i}_labels"] = lab.astype(np.int8)
        out[f"case{i}_scalars"] = np.array([e, lb, nu], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "rd_solve.npz"), **out)
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
