"""Time the MATLAB-layout TRW-S entry (sb_trws_solve behind trws.m / trws_mex.cpp) at BASELINE configs[1]'s shape.
usage: python scripts/mlayout_probe.py [H W L iters]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import stereo_b200 as sb  # noqa: E402
from stereo_b200 import solvers, synth  # noqa: E402

a = [int(x) for x in sys.argv[1:5]]
H, W, L, it = a + [375, 450, 64, 20][len(a):]
pr = synth.trws_problem(H, W, L, seed=1, kernel=1)
for _ in range(2):
    sol, e, lb, n = sb.trws(1, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"],
                            dict(maxiter=it, max_relgap=0))
    tm = solvers.last_timing
    print(f"trws {H}x{W}x{L}, {it} it: E={e:.4f} LB={lb:.4f}; setup {tm['setup_ms']:.1f} ms, solve {tm['solve_ms']:.1f} ms, "
          f"{tm['sweep_kernel_launches']} sweep launches avg {tm['sweep_kernel_ms'] / max(1, tm['sweep_kernel_launches']):.3f} ms", flush=True)
