"""tests/golden/teddy.npz: the Middlebury "teddy" pair the reference's examples run on (data/teddy/im2.png, im6.png;
example_ncc.m:9-10, example_global.m), as uint8 arrays, so BASELINE configs[0]/[1] plumbing runs on the real pair on
the GPU box (where /root/reference does not exist).  Run in the build container:
    python tests/golden/make_teddy.py"""
import os

import numpy as np
from PIL import Image

REF = "/root/reference/data/teddy"
HERE = os.path.dirname(os.path.abspath(__file__))
im2 = np.asarray(Image.open(os.path.join(REF, "im2.png")).convert("RGB"), dtype=np.uint8)
im6 = np.asarray(Image.open(os.path.join(REF, "im6.png")).convert("RGB"), dtype=np.uint8)
assert im2.shape == im6.shape == (375, 450, 3)
np.savez_compressed(os.path.join(HERE, "teddy.npz"), im2=im2, im6=im6)
print("teddy.npz", im2.shape, int(im2.astype(np.int64).sum()), int(im6.astype(np.int64).sum()))
