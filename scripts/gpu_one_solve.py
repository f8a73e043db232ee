"""One small resident solve (for ncu / sanitizer captures): python scripts/gpu_one_solve.py H W L iters [kernel] [seed]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stereo_b200 as sb
from stereo_b200 import synth
H, W, L, it = (int(x) for x in sys.argv[1:5])
k = int(sys.argv[5]) if len(sys.argv) > 5 else 1
seed = int(sys.argv[6]) if len(sys.argv) > 6 else 1
pr = synth.trws_problem(H, W, L, seed=seed, kernel=k)
s = sb.TrwsSolver(k, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"])
e, lb, n = s.minimize(it, 0.0)
print(H, W, L, it, e, lb, n, s.timing)
