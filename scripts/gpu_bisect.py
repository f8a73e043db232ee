"""Timing bisection of the sweep kernel on a ring-dominated grid (SB_TRWS_DEBUG switches)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    os.environ.pop("SB_TRWS_PROFILE", None)
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import stereo_b200 as sb
    from stereo_b200 import synth
    H, W, L = (int(x) for x in sys.argv[2:5])
    pr = synth.trws_problem(H, W, L, seed=1, kernel=1)
    s = sb.TrwsSolver(1, pr["unary"], pr["connectivity"], pr["q"], pr["qprim"], pr["alphas"], pr["tol"])
    s.minimize(2, 0.0)
    e, lb, n = s.minimize(6, 0.0)
    t = s.timing
    print(f"debug={os.environ.get('SB_TRWS_DEBUG')} kernel {t['sweep_kernel_ms']/t['sweep_kernel_launches']*1e3:.1f} us/launch -> {t['sweep_kernel_ms']/t['sweep_kernel_launches']*1e3/(3*H+2*W)*1965:.0f} cycles per critical-path step (3H+2W)", flush=True)
else:
    shapes = [tuple(int(x) for x in a.split("x")) for a in os.environ.get("SHAPES", "4x1500x64,375x450x8,128x160x64,375x450x64").split(",")]
    for shape in shapes:
        for dbg in (0, 8, 24):
            env = dict(os.environ, SB_TRWS_DEBUG=str(dbg))
            print(shape, end=" ", flush=True)
            subprocess.run([sys.executable, __file__, "x"] + [str(x) for x in shape], env=env)
