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
    print(f"debug={os.environ.get('SB_TRWS_DEBUG')} kernel {t['sweep_kernel_ms']/t['sweep_kernel_launches']*1e3:.1f} us/launch -> {t['sweep_kernel_ms']/t['sweep_kernel_launches']*1e3/(2*H+2*W-4)*1965:.0f} cycles per ring node", flush=True)
else:
    for shape in ((4, 1500, 64), (4, 1500, 8), (375, 450, 64)):
        for dbg in (0, 2, 4, 7):
            env = dict(os.environ, SB_TRWS_DEBUG=str(dbg))
            print(shape, end=" ", flush=True)
            subprocess.run([sys.executable, __file__, "x"] + [str(x) for x in shape], env=env)
