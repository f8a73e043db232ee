import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np
import stereo_b200 as sb
from stereo_b200 import synth
from oracle import oracle

def run(H, W, L, kernel, maxiter, prec, fuse=True, relgap=0.0):
    pr = synth.trws_problem(H, W, L, seed=1, kernel=kernel)
    conn0 = (pr['connectivity'] - 1).T.astype(np.uint32)
    t0 = time.time()
    rl, re, rlb, rit = oracle.trws_solve(kernel, pr['unary'].T, conn0, pr['q'].T, pr['qprim'].T, pr['alphas'], pr['tol'], maxiter, relgap)
    t1 = time.time()
    sol, e, lb, it = sb.trws(kernel, pr['unary'], pr['connectivity'], pr['q'], pr['qprim'], pr['alphas'], pr['tol'],
                             dict(maxiter=maxiter, max_relgap=relgap, precision=prec, fuse_rounding=fuse))
    t2 = time.time()
    from stereo_b200 import solvers
    print(f"{H}x{W} L={L} k={kernel} it={maxiter} {prec} fuse={fuse}: ref E={re:.6f} LB={rlb:.6f} it={rit} ({t1-t0:.2f}s) | gpu E={e:.6f} LB={lb:.6f} it={it} ({t2-t1:.2f}s, sweep {solvers.last_timing['sweep_ms_avg']:.3f} ms) | relE={abs(e-re)/abs(re):.2e} relLB={abs(lb-rlb)/abs(rlb):.2e} labels_equal={np.mean(sol==rl):.4f}", flush=True)

for prec in ['f64', 'f32']:
    for fuse in [False, True]:
        run(6, 8, 5, 1, 5, prec, fuse)
        run(12, 17, 15, 1, 10, prec, fuse)
run(12, 17, 15, 2, 10, 'f64')
run(12, 17, 15, 2, 10, 'f32')
run(48, 64, 8, 1, 20, 'f64')
run(48, 64, 8, 1, 20, 'f32')
run(48, 64, 40, 1, 20, 'f32')
run(48, 64, 40, 2, 20, 'f32')
run(31, 45, 64, 1, 10, 'f32')
run(31, 45, 100, 1, 10, 'f32')
run(31, 45, 192, 1, 5, 'f32')
run(31, 45, 256, 1, 5, 'f32')
run(31, 45, 256, 2, 5, 'f32')
run(3, 9, 7, 1, 10, 'f64')
run(1, 9, 7, 1, 10, 'f64')
run(48, 64, 8, 1, 200, 'f32', True, 1e-3)
