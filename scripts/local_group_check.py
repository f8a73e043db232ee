"""Column-banded grid-native TRW-S with all ranks in ONE process on one GPU vs the single-rank sweep.
usage: python scripts/local_group_check.py H W L iters world [kernel]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
from stereo_b200.gridsolver import TrwsGrid, TrwsGridLocalGroup  # noqa: E402

H, W, L, it, world = (int(x) for x in sys.argv[1:6])
kernel = int(sys.argv[6]) if len(sys.argv) > 6 else 1
tol = 0.02 if kernel == 1 else 0.02 ** 2
ref = TrwsGrid(kernel, H, W, L, tol)
ref.synth(77)
ref.finalize()
t0 = time.perf_counter()
e1, lb1, _ = ref.minimize(it, 0.0)
t1 = time.perf_counter() - t0
lab1 = ref.labels()
ref.close()
grp = TrwsGridLocalGroup(kernel, H, W, L, tol, world)
grp.each(lambda g: g.synth(77))
grp.finalize()
t0 = time.perf_counter()
e, lb, _ = grp.minimize(it)
t2 = time.perf_counter() - t0
lab = grp.labels()
grp.close()
print(f"{H}x{W} L={L} k={kernel} {it} it, {world} ranks on one GPU: E={e:.6f} (1 rank {e1:.6f}) LB={lb:.6f} ({lb1:.6f}) "
      f"labels equal {np.mean(lab == lab1):.6f}; {t2 * 1e3 / it:.2f} ms/it vs {t1 * 1e3 / it:.2f} ms/it", flush=True)
assert abs(e - e1) <= 1e-5 * abs(e1) and abs(lb - lb1) <= 1e-5 * abs(lb1) and np.mean(lab == lab1) >= 0.999
