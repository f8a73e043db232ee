"""Throughput regime probe: a large TRW-S problem built directly in device-friendly chunks.
python scripts/gpu_large.py H W L iters"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stereo_b200 as sb
from stereo_b200 import synth
H, W, L, it = (int(x) for x in sys.argv[1:5])
t0 = time.time()
rng = np.random.Generator(np.random.PCG64(1))
N = H * W
i1, i2 = sb.construct_neighborhood(H, W)
E = i1.size
# cheap synthetic positions: per-label fronto-parallel levels + small per-pixel slopes
base = (np.arange(L)[:, None] + 0.5) / L
q = (base + 0.01 * rng.standard_normal((1, E))).astype(np.float64) * np.ones((L, 1))
qprim = q + 0.002 * rng.standard_normal((L, 1)) + 0.001 * rng.standard_normal((1, E))
unary = rng.random((L, N))
alphas = np.where(rng.random(E) < 0.8, 216.0, 18.0)
print(f"generated in {time.time()-t0:.1f}s: unary {unary.nbytes/1e9:.2f} GB, q+qprim {2*q.nbytes/1e9:.2f} GB", flush=True)
t0 = time.time()
s = sb.TrwsSolver(1, unary, np.stack([i1, i2]), q, qprim, alphas, 0.02)
print(f"setup {time.time()-t0:.1f}s", flush=True)
s.minimize(1, 0.0)
e, lb, n = s.minimize(it, 0.0)
t = s.timing
ms = t['sweep_kernel_ms'] / t['sweep_kernel_launches']
print(f"{H}x{W} L={L}: {ms:.2f} ms/launch, {64.0*L*N/ms/1e6:.1f} GB/s algorithmic ({64.0*L*N/ms/1e6/6549.4*100:.1f}% of 6549 GB/s), E={e:.3f} LB={lb:.3f}", flush=True)
