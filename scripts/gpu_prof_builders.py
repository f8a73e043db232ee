"""Small driver for ncu captures of the non-TRW-S kernels of the hot path at BASELINE configs[2]'s shape:
one QPBO fusion (K4) and an NCC cost volume (K1).   python scripts/gpu_prof_builders.py [qpbo|ncc] [levels]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stereo_b200 as sb
from stereo_b200 import builders, synth
what = sys.argv[1] if len(sys.argv) > 1 else "qpbo"
H, W = 1080, 1920
if what == "qpbo":
    rp = synth.rd_problem(H, W, seed=0xB203, mode="stereo")
    a = tuple(rp[k] for k in ("U0", "U1", "E00", "E01", "E10", "E11", "connectivity"))
    sb.rd(*a, {})
    t0 = time.perf_counter()
    lab, e, lb, nu = sb.rd(*a, {})
    print("qpbo", H, W, f"{(time.perf_counter()-t0)*1e3:.1f} ms", e, lb, nu, "launches", sb._lib.lib().sb_kernel_launches())
else:
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    im0, im1, _ = synth.stereo_pair(H, W, 127, seed=0xB203)
    d = np.arange(D, dtype=np.float64)
    builders.ncc_volume(im0, im1, d[:2], 4)
    t0 = time.perf_counter()
    v = builders.ncc_volume(im0, im1, d, 4)
    print("ncc", H, W, D, f"{(time.perf_counter()-t0)*1e3:.1f} ms", float(v[:, :, 0].sum()))
