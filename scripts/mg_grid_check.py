"""Column-banded grid-native TRW-S on N GPUs vs the single-GPU sweep (run under torchrun):
   python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/mg_grid_check.py H W L iters [kernel]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from stereo_b200 import _lib  # noqa: E402
from stereo_b200.gridsolver import TrwsGrid  # noqa: E402

rank = int(os.environ["RANK"])
local = int(os.environ["LOCAL_RANK"])
world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
_lib.check(_lib.lib().sb_set_device(local))
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
H, W, L, it = (int(x) for x in sys.argv[1:5])
kernel = int(sys.argv[5]) if len(sys.argv) > 5 else 1
tol = 0.02 if kernel == 1 else 0.02 ** 2
ref_res = None
if rank == 0:
    ref = TrwsGrid(kernel, H, W, L, tol)
    ref.synth(77)
    ref.finalize()
    ref.minimize(1, 0.0)
    ref.reset()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    re, rlb, rn = ref.minimize(it, 0.0)
    t_ref = time.perf_counter() - t0
    rlab = ref.labels()
    mem1 = ref.info()["hbm_bytes"]
    ref.close()
dist.barrier()
s = TrwsGrid(kernel, H, W, L, tol, group=dist.group.WORLD)
s.synth(77)
s.finalize()
s.minimize(1, 0.0)
s.reset()
dist.barrier()
t0 = time.perf_counter()
e, lb, n = s.minimize(it, 0.0)
torch.cuda.synchronize()
dist.barrier()
t_mg = time.perf_counter() - t0
lab = s.labels()
info = s.info()
if rank == 0:
    print(f"{H}x{W} L={L} k={kernel} {it} it on {world} GPUs: E={e:.6f} (1 GPU {re:.6f}) LB={lb:.6f} ({rlb:.6f}) n={n} ({rn}) "
          f"labels equal {np.mean(lab == rlab):.6f} | {t_mg * 1e3 / it:.2f} ms/it banded vs {t_ref * 1e3 / it:.2f} ms/it single "
          f"= {t_ref / t_mg:.2f}x | HBM per rank {info['hbm_bytes'] / 2**30:.2f} GiB vs {mem1 / 2**30:.2f} GiB", flush=True)
    assert abs(e - re) <= 1e-5 * abs(re) and abs(lb - rlb) <= 1e-5 * abs(rlb) and np.mean(lab == rlab) >= 0.999
dist.barrier()
s.close()
dist.destroy_process_group()
