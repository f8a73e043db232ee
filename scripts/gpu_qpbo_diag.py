import os, sys, time
os.environ["SB_QPBO_PROFILE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stereo_b200 as sb
from stereo_b200 import synth
for (H, W, mode) in [(375, 450, "stereo"), (1080, 1920, "stereo"), (375, 450, "frustrated")]:
    rp = synth.rd_problem(H, W, seed=5, mode=mode)
    a = (rp["U0"], rp["U1"], rp["E00"], rp["E01"], rp["E10"], rp["E11"], rp["connectivity"])
    sb.rd(*a, {})
    t0 = time.perf_counter(); r = sb.rd(*a, {}); dt = time.perf_counter() - t0
    print(f"{H}x{W} {mode}: {dt*1e3:.1f} ms, unlabelled {r[3]}", flush=True)
