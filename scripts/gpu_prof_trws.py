"""Per-phase cycle breakdown of the sweep kernel (SB_TRWS_PROFILE=1) on a few shapes."""
import os, sys
os.environ["SB_TRWS_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import stereo_b200 as sb
from stereo_b200 import synth
for (H, W, L, it) in [(128, 160, 64, 6), (375, 450, 64, 5), (256, 256, 256, 3)]:
    pr = synth.trws_problem(H, W, L, seed=1, kernel=1)
    s = sb.TrwsSolver(1, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"])
    s.minimize(2, 0.0)
    s.reset()
    print(f"== {H}x{W} L={L}", flush=True)
    e, lb, n = s.minimize(it, 0.0)
    print(f"   sweep {s.timing['sweep_ms_avg']:.3f} ms/iter  kernel {s.timing['sweep_kernel_ms']/s.timing['sweep_kernel_launches']:.3f} ms/launch", flush=True)
    s.close()
