"""Time the NCC volume kernels at BASELINE configs[2]'s shape (1080 x 1920, 128 levels, 9 x 9 window).
usage: python scripts/ncc_probe.py [H W D p]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from stereo_b200 import builders, synth  # noqa: E402

H, W, D, p = (int(x) for x in sys.argv[1:5]) if len(sys.argv) > 4 else (1080, 1920, 128, 4)
im0, im1, _ = synth.stereo_pair(H, W, D - 1, seed=3)
for _ in range(2):
    v = builders.NccVolume(im0, im1, np.arange(float(D)), p)
    i = v.info()
    algo = 4.0 * H * W * D + 6.0 * H * W
    print(f"{H}x{W}x{D} p={p}: {i}, {algo / (i['kernel_ms'] * 1e-3) / 1e9:.0f} GB/s algorithmic", flush=True)
    v.close()
