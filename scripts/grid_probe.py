"""Time the grid-native TRW-S sweep on the on-device synthetic problem.
usage: python scripts/grid_probe.py H W L [iters] [kernel]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stereo_b200.gridsolver import TrwsGrid  # noqa: E402

H, W, L = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
kernel = int(sys.argv[5]) if len(sys.argv) > 5 else 1
t0 = time.time()
g = TrwsGrid(kernel, H, W, L, 0.02 if kernel == 1 else 0.02 ** 2)
g.synth(0xB200)
t1 = time.time()
g.finalize()
t2 = time.time()
info = g.info()
e, lb, it = g.minimize(iters, 0.0)
tm = g.timing
ms = tm["sweep_kernel_ms"] / max(1, tm["sweep_kernel_launches"])
N = H * W
gbs = 64.0 * L * N / (ms * 1e-3) / 1e9
print(f"grid {H}x{W}x{L} k={kernel}: hbm {info['hbm_bytes'] / 2**30:.2f} GiB, ctas {info['ctas_fwd']}/{info['ctas_bwd']}, "
      f"smem/cta {info['smem_per_cta']}, create+synth {t1 - t0:.2f}s tables {t2 - t1:.2f}s; {iters} it: E={e:.4f} LB={lb:.4f} "
      f"solve {tm['solve_ms']:.1f} ms, {tm['sweep_kernel_launches']} launches avg {ms:.3f} ms -> {gbs:.0f} GB/s algorithmic "
      f"= {gbs / 6551:.3f} of 6551", flush=True)
g.close()
