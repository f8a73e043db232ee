import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import stereo_b200 as sb
from stereo_b200 import synth, solvers
pr = synth.trws_problem(375, 450, 64, seed=1, kernel=1)
def pinned(a):
    a = np.asfortranarray(a, dtype=np.float64)
    t = torch.empty(a.size, dtype=torch.float64, pin_memory=True)
    v = t.numpy().reshape(a.shape, order="F"); v[...] = a
    return t, v
keep = [pinned(pr[k]) for k in ("unary", "q", "qprim", "alphas")]
u, q, qp, al = (k[1] for k in keep)
for i in range(4):
    t0 = time.perf_counter()
    r = sb.trws(1, u, pr["connectivity"], q, qp, al, pr["tol"], dict(maxiter=20))
    dt = time.perf_counter() - t0
    print(f"call {i}: wall {dt*1e3:.1f} ms  lib timing {solvers.last_timing}", flush=True)
