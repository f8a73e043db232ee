import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stereo_b200 import builders, synth
H, W, D, p = (int(x) for x in sys.argv[1:5])
im0, im1, _ = synth.stereo_pair(H, W, D, seed=H + W)
d = np.arange(D, dtype=np.float64)
v = builders.NccVolume(im0, im1, d, p)
print(v.info())
fast = v.get()
os.environ["SB_NCC_GENERAL"] = "1"
gen = builders.ncc_volume(im0, im1, d, p)
diff = np.abs(fast - gen)
print("max diff", diff.max(), "at", np.unravel_index(diff.argmax(), diff.shape), "frac>1e-4", (diff > 1e-4).mean())
for lv in range(min(D, 3)):
    bad = np.argwhere(diff[:, :, lv] > 1e-4)
    print("level", lv, "bad count", len(bad), "rows", np.unique(bad[:, 0])[:12], "cols", np.unique(bad[:, 1])[:12])
